set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests12.log 2>&1; echo "rc=$?" >> gpurun_out/r2_tests12.log; tail -4 gpurun_out/r2_tests12.log
python bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_bench_c2_group2.json 2> gpurun_out/r2_bench_c2_group2.err; tail -2 gpurun_out/r2_bench_c2_group2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2f_c2_n2.json 2> gpurun_out/r2f_c2_n2.err
python bench.py --steps 5 --warmup 3 > gpurun_out/r2f_c2_n1.json 2> gpurun_out/r2f_c2_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2f_c2_ref.json 2>&1
for cfg in C4-maxmean C4-mean; do python bench.py --config $cfg --steps 2 --warmup 3 > gpurun_out/r2f_${cfg}_n1.json 2> gpurun_out/r2f_${cfg}_n1.err; done
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; tail -1 gpurun_out/r2_smoke.log
