# scan step without branches: parity of the general path, then config 3 and the forced-general config-4 shape
timeout 900 python -m pytest tests/test_gpu_batch.py -x -q -m gpu -k "general_path_stream_groups or routes or levels_11 or stress_corpus or truncations or single_fillers or garbage or unaligned or decode_twice" 2>&1 | tail -3
timeout 300 python tools/profile_run.py --streams 10000 --runs 4 --workload config3 2>&1 | tail -3
timeout 600 python tools/prof_aux.py --streams 20000 --kernel 1 2>&1 | tail -2
