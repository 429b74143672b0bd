ncu --set full --clock-control none --import-source on -k regex:acm_walk1 -s 2 -c 1 -o gpurun_out/prof_r02_walk1 -f python tools/dbg_stream.py 2>&1 | tail -2
