set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multi_gpu.py -x -q -m gpu > gpurun_out/r2z_multi.log 2>&1; tail -3 gpurun_out/r2z_multi.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29912 bench.py --gpus 2 --steps 10 --warmup 3 2> gpurun_out/r2z_c2_n2.err | grep -v "^NCCL version" > gpurun_out/r2z_c2_n2.json
timeout 300 python bench.py --gpus 2 --steps 10 --warmup 3 2> gpurun_out/r2z_c2_group2.err | grep -v "^NCCL version" > gpurun_out/r2z_c2_group2.json
python - <<'P'
import json
for f in ("r2z_c2_n2","r2z_c2_group2"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1]); print(f, d["ms_per_step"], d["value"], d.get("clocks"), d.get("collective_ms_per_step"), d.get("per_rank"))
    except Exception as e: print(f, "ERR", e)
P
