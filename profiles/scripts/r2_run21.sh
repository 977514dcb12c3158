set -x
mkdir -p gpurun_out
for n in 2 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29900+n)) bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/r2x_c2_n$n.json 2> gpurun_out/r2x_c2_n$n.err; tail -1 gpurun_out/r2x_c2_n$n.err
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29904 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/r2x_c2_n4.json 2> gpurun_out/r2x_c2_n4.err
