for r in 0 1; do
echo "== seg ramp $r"
ACM_B200_TRACE=1 ACM_B200_SEG_RAMP=$r timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu --no-config4 --no-streaming 2>&1 >/dev/null | grep "acm trace" | tail -14
done
