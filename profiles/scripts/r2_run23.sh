set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2z_tests.log 2>&1; tail -3 gpurun_out/r2z_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2z_smoke.log 2>&1; tail -2 gpurun_out/r2z_smoke.log
timeout 300 python bench.py > gpurun_out/r2z_c2_n1.json 2> gpurun_out/r2z_c2_n1.err; tail -c 200 gpurun_out/r2z_c2_n1.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2z_c2_ref.json 2> gpurun_out/r2z_c2_ref.err
timeout 300 python bench.py --config C3 --steps 3 --warmup 3 --no-cpu > gpurun_out/r2z_C3_n1.json 2> gpurun_out/r2z_C3_n1.err
python - <<'P'
import json
for f in ("r2z_c2_n1","r2z_c2_ref","r2z_C3_n1"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1]); print(f, d["ms_per_step"], d["value"], d.get("clocks"), d.get("e2e",{}).get("value"))
    except Exception as e: print(f, "ERR", e)
P
