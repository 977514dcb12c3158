# 8-GPU box: strong-scaling curve of C2, and C4 / C5 at N = 8 (BASELINE configs[3], [4])
set -x
mkdir -p gpurun_out
nvidia-smi -L | wc -l > gpurun_out/r2_n8_gpus.txt
python -m pytest tests/test_multi_gpu.py -m gpu -x -q > gpurun_out/r2_tests_multi8.log 2>&1; echo "rc=$?" >> gpurun_out/r2_tests_multi8.log; tail -3 gpurun_out/r2_tests_multi8.log
python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/r2s_c2_n1.json 2> gpurun_out/r2s_c2_n1.err
for n in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29520+n)) bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/r2s_c2_n$n.json 2> gpurun_out/r2s_c2_n$n.err; tail -1 gpurun_out/r2s_c2_n$n.err
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 3 --warmup 3 --exchange 1 > gpurun_out/r2s_c2_n8_x1.json 2> gpurun_out/r2s_c2_n8_x1.err
for cfg in C4-maxmean C5; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --config $cfg --steps 2 --warmup 3 > gpurun_out/r2s_${cfg}_n8.json 2> gpurun_out/r2s_${cfg}_n8.err; tail -1 gpurun_out/r2s_${cfg}_n8.err
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --config C5 --steps 2 --warmup 3 --exchange 1 > gpurun_out/r2s_C5_n8_x1.json 2> gpurun_out/r2s_C5_n8_x1.err
ls -la gpurun_out | grep r2s_
