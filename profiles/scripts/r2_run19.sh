set -x
mkdir -p gpurun_out
python -m pytest tests/test_multi_gpu.py tests/test_gpu_parity.py -m gpu -x -q -k "multi or group or comm or torch_distributed or two_gpus or calc_sig or staged" > gpurun_out/r2_tests19.log 2>&1; echo "rc=$?" >> gpurun_out/r2_tests19.log; tail -4 gpurun_out/r2_tests19.log
python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/r2v_c2_n1.json 2> gpurun_out/r2v_c2_n1.err
for n in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29720+n)) bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/r2v_c2_n$n.json 2> gpurun_out/r2v_c2_n$n.err; tail -1 gpurun_out/r2v_c2_n$n.err
done
