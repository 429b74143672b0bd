ncu --set full --clock-control none --import-source on -k regex:acm_scan -c 1 -o gpurun_out/prof_r02_g3_scan -f python tools/prof_aux.py --streams 20000 --kernel 1 2>&1 | tail -2
