set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests15.log 2>&1; echo "rc=$?" >> gpurun_out/r2_tests15.log; tail -4 gpurun_out/r2_tests15.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r2g_c2_n1.json 2> gpurun_out/r2g_c2_n1.err
python profiles/stress_parity.py 2000 601 > gpurun_out/r2_stress_parity_g.log 2>&1; tail -n 1 gpurun_out/r2_stress_parity_g.log
FLEXGSEA_B200_LIB=$PWD/profiles/variants/libflexgsea_b200_xorcmp.so python profiles/es_variants.py > gpurun_out/r2_e1_plain.json 2> gpurun_out/r2_e1_plain.err &&
FLEXGSEA_B200_LIB=$PWD/profiles/variants/libflexgsea_b200_xorcmp.so ncu --set full --clock-control none -k regex:'es_wks_kernel' -s 12 -c 2 -o gpurun_out/r2_prof_e1 python profiles/es_variants.py > gpurun_out/r2_ncu_e1.log 2>&1
tail -2 gpurun_out/r2_ncu_e1.log
