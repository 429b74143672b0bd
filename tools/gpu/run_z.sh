for pct in 18 20 22 24 26 28; do
echo "== pct $pct"
ACM_B200_SCAN_PCT_WALK=$pct timeout 300 python tools/profile_run.py --streams 10000 --runs 4 2>&1 | tail -2 | head -1
done
