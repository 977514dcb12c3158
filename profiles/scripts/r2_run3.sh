set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests3.log 2>&1; echo "rc=$?" >> gpurun_out/r2_tests3.log; tail -15 gpurun_out/r2_tests3.log
python profiles/dup_genes_timing.py > gpurun_out/r2_dup_genes.json 2> gpurun_out/r2_dup_genes.err; cat gpurun_out/r2_dup_genes.json
python profiles/ncu_targets.py SIG > gpurun_out/r2_sig_plain.json 2>&1; cat gpurun_out/r2_sig_plain.json
python profiles/stress_pipeline.py 600 401 > gpurun_out/r2_stress_pipeline.log 2>&1; tail -3 gpurun_out/r2_stress_pipeline.log
python profiles/stress_parity.py 800 402 > gpurun_out/r2_stress_parity.log 2>&1; tail -3 gpurun_out/r2_stress_parity.log
