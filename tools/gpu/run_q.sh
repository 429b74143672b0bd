set -x
nvidia-smi topo -m 2>&1 | head -12
nproc
for N in 8 4; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$N bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_r02_n$N.json 2> gpurun_out/bench_r02_n$N.err; tail -c 2600 gpurun_out/bench_r02_n$N.json; tail -3 gpurun_out/bench_r02_n$N.err
done
timeout 300 python -m pytest tests/test_gpu_multi.py -x -q -m gpu -k device_mask 2>&1 | tail -3
