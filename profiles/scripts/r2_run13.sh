set -x
mkdir -p gpurun_out
python profiles/e2e_breakdown.py > gpurun_out/r2_e2e_breakdown.txt 2>&1; cat gpurun_out/r2_e2e_breakdown.txt | tail -9
python bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_bench_c2_group2.json 2> gpurun_out/r2_bench_c2_group2.err; tail -2 gpurun_out/r2_bench_c2_group2.err; cut -c1-400 gpurun_out/r2_bench_c2_group2.json
