set -x
nvidia-smi topo -m 2>&1 | head -12
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -4
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_r02_n2.json 2> gpurun_out/bench_r02_n2.err; tail -c 2500 gpurun_out/bench_r02_n2.json; tail -5 gpurun_out/bench_r02_n2.err
