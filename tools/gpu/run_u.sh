set -x
timeout 1500 python -m pytest tests/test_gpu_batch.py tests/test_gpu_gen.py -x -q -m gpu 2>&1 | tail -5
timeout 300 python tools/prof_aux.py --streams 20000 2>&1 | tail -2
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"acm_scan|acm_blocks|acm_finish" python tools/prof_aux.py --streams 20000 2>&1 | grep -E "gpu__time" | head -4
