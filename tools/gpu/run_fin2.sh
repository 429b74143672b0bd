# end-of-round check on two GPUs: the multi-GPU tests, then the bench as the driver launches it
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_r02d_n2.json 2> gpurun_out/bench_r02d_n2.err
tail -c 1600 gpurun_out/bench_r02d_n2.json; tail -3 gpurun_out/bench_r02d_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 2>/dev/null | tail -c 700
