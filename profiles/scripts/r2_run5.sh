set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests5.log 2>&1; echo "rc=$?" >> gpurun_out/r2_tests5.log; tail -6 gpurun_out/r2_tests5.log
python profiles/dup_genes_timing.py > gpurun_out/r2_dup_genes.json 2> gpurun_out/r2_dup_genes.err; cat gpurun_out/r2_dup_genes.json
python profiles/stress_parity.py 1500 404 > gpurun_out/r2_stress_parity.log 2>&1; tail -n 2 gpurun_out/r2_stress_parity.log
python profiles/stress_pipeline.py 1500 405 > gpurun_out/r2_stress_pipeline.log 2>&1; tail -n 2 gpurun_out/r2_stress_pipeline.log
EXACT_SMALL=0 python profiles/stress_pipeline.py 1500 406 > gpurun_out/r2_stress_pipeline0.log 2>&1; tail -n 2 gpurun_out/r2_stress_pipeline0.log
python profiles/ncu_targets.py SIG > gpurun_out/r2_sig_plain.json 2> gpurun_out/r2_sig_plain.err &&
ncu --set full --clock-control none --import-source on -k regex:'fdr_count_kernel|sig_partial_kernel' -c 4 -o gpurun_out/r2_prof_sig2 python profiles/ncu_targets.py SIG > gpurun_out/r2_ncu_sig2.log 2>&1
tail -3 gpurun_out/r2_ncu_sig2.log
