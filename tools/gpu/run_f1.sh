timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 --config4-streams 200000 > gpurun_out/bench_r02b_n2.json 2> gpurun_out/bench_r02b_n2.err
tail -c 1500 gpurun_out/bench_r02b_n2.json; tail -3 gpurun_out/bench_r02b_n2.err
