python tools/stream_time.py 2>&1 | tail -1
python tools/stream_time.py 2>&1 | tail -1
