timeout 900 python -m pytest tests/test_gpu_stream.py tests/test_gpu_acmtool.py -x -q -m gpu 2>&1 | tail -3
