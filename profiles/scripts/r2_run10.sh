set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests10.log 2>&1; echo "rc=$?" >> gpurun_out/r2_tests10.log; tail -5 gpurun_out/r2_tests10.log
python profiles/stress_pipeline.py 1000 408 > gpurun_out/r2_stress_pipeline.log 2>&1; tail -n 2 gpurun_out/r2_stress_pipeline.log
python profiles/ncu_targets.py SIG > gpurun_out/r2_sig_plain.json 2> gpurun_out/r2_sig_plain.err &&
ncu --set full --clock-control none --import-source on -k regex:'fdr_count_kernel' -c 2 -o gpurun_out/r2_prof_sig3 python profiles/ncu_targets.py SIG > gpurun_out/r2_ncu_sig3.log 2>&1
tail -2 gpurun_out/r2_sig_plain.json
