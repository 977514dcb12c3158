set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2_gpus.txt
python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2_tests.log
tail -5 gpurun_out/r2_tests.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_c2_n1.json 2> gpurun_out/r2_bench_c2_n1.err; tail -2 gpurun_out/r2_bench_c2_n1.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_bench_c2_n2.json 2> gpurun_out/r2_bench_c2_n2.err; tail -3 gpurun_out/r2_bench_c2_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 --exchange 1 > gpurun_out/r2_bench_c2_n2_x1.json 2> gpurun_out/r2_bench_c2_n2_x1.err
./profiles/microbench/dmma_peak.bin > gpurun_out/r2_dmma_peak.json 2>&1
for cfg in C3 C4-maxmean C4-mean C5 C1; do
  python bench.py --config $cfg --steps 2 --warmup 3 > gpurun_out/r2_bench_${cfg}_n1.json 2> gpurun_out/r2_bench_${cfg}_n1.err; tail -2 gpurun_out/r2_bench_${cfg}_n1.err
done
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_c2_ref.json 2>&1
ls -la gpurun_out | head -40
