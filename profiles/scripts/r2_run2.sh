set -x
mkdir -p gpurun_out
python -m pytest tests/test_r_shim.py -m gpu -x -q > gpurun_out/r2_tests_shim.log 2>&1; echo "rc=$?" >> gpurun_out/r2_tests_shim.log; tail -15 gpurun_out/r2_tests_shim.log
python profiles/ncu_targets.py > gpurun_out/r2_targets_plain.json 2> gpurun_out/r2_targets_plain.err &&
ncu --set full --clock-control none --import-source on -k regex:'lm_gemm_kernel|es_mean_smem_kernel|sort_rank_big_kernel|fdr_count_kernel|sig_partial_kernel' -c 12 -o gpurun_out/r2_prof_gaps python profiles/ncu_targets.py > gpurun_out/r2_ncu_gaps.log 2>&1
tail -5 gpurun_out/r2_ncu_gaps.log
python bench.py --perms 1184 --steps 1 --warmup 3 --no-cpu > gpurun_out/r2_launch_plain.json 2> gpurun_out/r2_launch_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches.csv python bench.py --perms 1184 --steps 1 --warmup 3 --no-cpu > gpurun_out/r2_launch_ncu.log 2>&1
ls -la gpurun_out | grep r2_
