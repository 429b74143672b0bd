for a in 0 1 0 1; do
echo "== scan alone $a"
ACM_B200_SCAN_ALONE=$a timeout 300 python tools/profile_run.py --streams 10000 --runs 4 --workload config3 2>&1 | tail -3 | head -2
done
