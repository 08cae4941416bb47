for G in 3 4; do for P in 1 2 3; do
  export RCLC_SWEEP_PARTS=$P
  RCLC_SOLVE_GROUPS=$G timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('G=$G P=$P', round(d['ms_per_step'],2), round(d['e2e']['ms_per_step'],2))"
done; done
