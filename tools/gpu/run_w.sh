set -x
ACM_KERNEL=2 timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_run.py 2>&1 | grep -v "^=========     Host Frame\|^=========         in " | head -60
