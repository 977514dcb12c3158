set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "bulk_copy or c2_shape or C4 or tail or wks or mean" > gpurun_out/r2y_tests.log 2>&1; tail -3 gpurun_out/r2y_tests.log
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r2y_c2_n1.json 2> gpurun_out/r2y_c2_n1.err; tail -c 300 gpurun_out/r2y_c2_n1.err
timeout 300 python bench.py --config C4-mean --steps 5 --warmup 3 --no-cpu > gpurun_out/r2y_C4-mean_n1.json 2> gpurun_out/r2y_C4-mean_n1.err
timeout 300 python bench.py --config C4-maxmean --steps 5 --warmup 3 --no-cpu > gpurun_out/r2y_C4-maxmean_n1.json 2> gpurun_out/r2y_C4-maxmean_n1.err
python - <<'P'
import json
for f in ("r2y_c2_n1","r2y_C4-mean_n1","r2y_C4-maxmean_n1"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1]); print(f, d["ms_per_step"], d["stages"]["es"], d["clocks"])
    except Exception as e: print(f, "ERR", e)
P
