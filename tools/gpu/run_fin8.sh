# end-of-round record on eight GPUs, launched as the driver does
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/bench_r02e_n8.json 2> gpurun_out/bench_r02e_n8.err
tail -c 1500 gpurun_out/bench_r02e_n8.json; tail -2 gpurun_out/bench_r02e_n8.err
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
