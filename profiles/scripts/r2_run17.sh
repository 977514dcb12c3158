# 8-GPU box, final build: strong-scaling curve of C2 again, C3 / C4 at N = 8
set -x
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/r2t_c2_n1.json 2> gpurun_out/r2t_c2_n1.err
for n in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29620+n)) bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/r2t_c2_n$n.json 2> gpurun_out/r2t_c2_n$n.err; tail -1 gpurun_out/r2t_c2_n$n.err
done
for cfg in C3 C4-maxmean C4-mean; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29641 bench.py --gpus 8 --config $cfg --steps 2 --warmup 3 > gpurun_out/r2t_${cfg}_n8.json 2> gpurun_out/r2t_${cfg}_n8.err; tail -1 gpurun_out/r2t_${cfg}_n8.err
done
python bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2t_c2_group8.json 2> gpurun_out/r2t_c2_group8.err; tail -1 gpurun_out/r2t_c2_group8.err
ls gpurun_out | grep r2t_
