ACM_B200_TRACE=1 python bench.py --steps 3 --warmup 3 --no-cpu --no-config4 --no-streaming --e2e-steps 3 2>&1 | grep "acm trace" | tail -24
