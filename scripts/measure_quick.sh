timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py -x -q 2>&1 | tail -3
timeout 120 python scripts/profile_sweep.py 20 500000 10 2>&1 | tail -2
for i in 1 2; do timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['launch_ms'])"; done
