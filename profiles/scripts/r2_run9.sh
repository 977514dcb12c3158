set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mean or maxmean or pipeline or two_responses or few_samples" > gpurun_out/r2_tests9.log 2>&1; echo "rc=$?" >> gpurun_out/r2_tests9.log; tail -5 gpurun_out/r2_tests9.log
python profiles/arrange_timing.py > gpurun_out/r2_arrange.json 2> gpurun_out/r2_arrange.err; cat gpurun_out/r2_arrange.json; tail -3 gpurun_out/r2_arrange.err
