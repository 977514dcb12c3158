mkdir -p gpurun_out
timeout 60 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29918 bench.py --gpus 8 --steps 10 --warmup 3 2> gpurun_out/r2z_c2_n8.err | grep -v "^NCCL version" > gpurun_out/r2z_c2_n8.json
timeout 40 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r2z_c2_n1_box8.json 2> gpurun_out/r2z_c2_n1_box8.err
timeout 40 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29914 bench.py --gpus 4 --steps 10 --warmup 3 2> gpurun_out/r2z_c2_n4.err | grep -v "^NCCL version" > gpurun_out/r2z_c2_n4.json
