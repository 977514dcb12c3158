set -x
mkdir -p gpurun_out
python profiles/stress_parity.py 10000 501 > gpurun_out/r2_stress_parity_final.log 2>&1; tail -n 2 gpurun_out/r2_stress_parity_final.log
python profiles/stress_pipeline.py 5000 502 > gpurun_out/r2_stress_pipeline_final.log 2>&1; tail -n 2 gpurun_out/r2_stress_pipeline_final.log
python profiles/stress_lm.py 1500 503 > gpurun_out/r2_stress_lm_final.log 2>&1; tail -n 3 gpurun_out/r2_stress_lm_final.log
EXACT_SMALL=0 python profiles/stress_parity.py 6000 504 > gpurun_out/r2_stress_parity0_final.log 2>&1; tail -n 2 gpurun_out/r2_stress_parity0_final.log
