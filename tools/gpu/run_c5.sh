ACM_KERNEL=1 timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_run.py 2>&1 | grep -v '^=========     Host Frame\|^=========         in ' | tail -4
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_run.py 2>&1 | grep -v '^=========     Host Frame\|^=========         in ' | tail -3
ACM_KERNEL=1 timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_run.py 2>&1 | grep -v '^=========     Host Frame\|^=========         in ' | tail -3
ACM_KERNEL=1 timeout 900 compute-sanitizer --tool synccheck python tools/sanitize_run.py 2>&1 | grep -v '^=========     Host Frame\|^=========         in ' | tail -3
