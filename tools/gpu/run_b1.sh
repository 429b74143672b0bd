timeout 1200 python -m pytest tests/test_gpu_batch.py tests/test_gpu_multi.py tests/test_gpu_gen.py -x -q -m gpu 2>&1 | tail -5
bash tools/gpu/run_b4.sh
