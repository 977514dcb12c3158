# why is N = 2 slow on the 8-GPU box (3.5 ms in the collectives) when a 2-GPU box takes 0.3 ms?
set -x
mkdir -p gpurun_out
NCCL_DEBUG=INFO NCCL_DEBUG_FILE=gpurun_out/r2_nccl_n2_%h_%p.log python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2u_c2_n2.json 2> gpurun_out/r2u_c2_n2.err
CUDA_VISIBLE_DEVICES=0,1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29712 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2u_c2_n2_vis01.json 2> gpurun_out/r2u_c2_n2_vis01.err
CUDA_VISIBLE_DEVICES=2,5 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29713 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2u_c2_n2_vis25.json 2> gpurun_out/r2u_c2_n2_vis25.err
python bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2u_c2_group2.json 2> gpurun_out/r2u_c2_group2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29714 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2u_c2_n8.json 2> gpurun_out/r2u_c2_n8.err
grep -h "NCCL INFO" gpurun_out/r2_nccl_n2_*.log | grep -i "channel\|algo\|proto\|P2P\|NVLS\|via\|Connected" | head -40 > gpurun_out/r2_nccl_n2_summary.txt
rm -f gpurun_out/r2_nccl_n2_*_*.log
