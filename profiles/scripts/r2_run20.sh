# N = 2 on the 8-GPU box: what makes rank 0 late by 3.7 ms per step?
# (ran against commit d94a12d plus two diagnostic switches in bench.py, since removed: FGS_BENCH_NO_CVD = do not
# restrict CUDA_VISIBLE_DEVICES inside the process, FGS_BENCH_NO_SAMPLER = no nvidia-smi clock sampler)
set -x
mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29800 + RANDOM % 100)) bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2w_$tag.json 2> gpurun_out/r2w_$tag.err; }
run A_inproc_cvd FOO=1
run B_shell_cvd01 CUDA_VISIBLE_DEVICES=0,1
run C_all_visible_nonvls FGS_BENCH_NO_CVD=1 NCCL_NVLS_ENABLE=0
run D_all_visible_nosampler FGS_BENCH_NO_CVD=1 FGS_BENCH_NO_SAMPLER=1
run E_inproc_cvd_nosampler FGS_BENCH_NO_SAMPLER=1
run F_shell_cvd_all CUDA_VISIBLE_DEVICES=0,1,2,3,4,5,6,7
