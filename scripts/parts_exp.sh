timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 120 python scripts/profile_sweep.py 20 500000 6 > gpurun_out/p4.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sweep -s 3 -c 2 -o gpurun_out/prof_sweep_r01_final2 -f python scripts/profile_sweep.py 20 500000 6 > gpurun_out/ncu_p4.log 2>&1; tail -2 gpurun_out/p4.log
