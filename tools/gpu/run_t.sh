set -x
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"acm_scan|acm_blocks|acm_finish" python tools/prof_aux.py --streams 20000 2>&1 | grep -E "acm_scan|acm_blocks|acm_finish|gpu__time" | head -12
