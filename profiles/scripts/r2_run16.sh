set -x
mkdir -p gpurun_out
python -m pytest tests/test_multi_gpu.py -m gpu -x -q > gpurun_out/r2_tests16.log 2>&1; echo "rc=$?" >> gpurun_out/r2_tests16.log; tail -30 gpurun_out/r2_tests16.log
