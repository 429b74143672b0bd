set -x
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_run.py 2>&1 | grep -v '^=========     Host Frame\|^=========         in ' | tail -12
ACM_KERNEL=1 timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_run.py 2>&1 | grep -v '^=========     Host Frame\|^=========         in ' | tail -12
ACM_KERNEL=2 timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_run.py 2>&1 | grep -v '^=========     Host Frame\|^=========         in ' | tail -12
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_stream.py 2>&1 | grep -v '^=========     Host Frame\|^=========         in ' | tail -12
ACM_KERNEL=1 timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_run.py 2>&1 | grep -v '^=========     Host Frame\|^=========         in ' | tail -8
timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_stream.py 2>&1 | grep -v '^=========     Host Frame\|^=========         in ' | tail -8
