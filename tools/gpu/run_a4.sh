LD_PRELOAD=$PWD/libacm_b200/_lib/var/w1dbg/libacm_b200.so python tools/dbg_stream.py 2>&1 | tail -16
