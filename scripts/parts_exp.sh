for D in 0 1; do for P in 0 1 2; do
  if [ $P -eq 0 ]; then unset RCLC_SWEEP_PARTS; else export RCLC_SWEEP_PARTS=$P; fi
  echo "dense=$D parts=$P"; RCLC_MS_DENSE=$D timeout 300 python scripts/bench_multistart.py 2>&1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_search'], d['Mpts_poses_per_s'], d['best_pose'])"
done; done
