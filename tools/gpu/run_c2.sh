timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_stream.py 2>&1 | grep -v '^=========     Host Frame\|^=========         in ' | grep -E "Race reported|and .* access|SUMMARY|read " | sed 's/+0x[0-9a-f]*//' | sort | uniq -c | sort -rn | head -20
python tools/stream_time.py 2>&1 | tail -1
