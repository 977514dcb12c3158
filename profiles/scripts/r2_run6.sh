set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests6.log 2>&1; echo "rc=$?" >> gpurun_out/r2_tests6.log; tail -25 gpurun_out/r2_tests6.log
python profiles/stress_pipeline.py 800 407 > gpurun_out/r2_stress_pipeline.log 2>&1; tail -n 2 gpurun_out/r2_stress_pipeline.log
