timeout 900 python -m pytest tests/test_gpu_stream.py -x -q -m gpu 2>&1 | tail -3
python tools/stream_time.py 2>&1 | tail -1
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:"acm_walk1" -c 3 python tools/dbg_stream.py 2>&1 | grep -E "gpu__time|inst_exec" | head -6
