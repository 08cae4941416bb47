# end-of-round measurement pass (run under gpurun): tests, smoke, both bench arms, launch list, one full capture of k_sweep
set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_ref.json 2> gpurun_out/final_ref.err
timeout 900 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
RCLC_NO_GRAPHS=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 4000 -c 500 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
tail -c 600 gpurun_out/final_bench.json
