#!/bin/bash
# End-of-round measurement pass (run under gpurun): tests, smoke, both bench arms, launch lists, ncu captures of the kernels the
# round's claims rest on. Everything lands in gpurun_out/; the summaries that are to be judged are copied to profiles/ by hand.
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r02_pytest.log; cat gpurun_out/r02_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
# k_sweep: full capture, DRAM bytes per launch (before the bench, which quotes it only for the sources it was captured from)
bash scripts/capture_sweep_traffic.sh > gpurun_out/capture.log 2>&1; tail -2 gpurun_out/capture.log
cp gpurun_out/sweep_traffic.json profiles/sweep_traffic.json
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_ref.json 2> gpurun_out/r02_ref.err
timeout 900 python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err
# launch list of a bench step (per-launch times are serialised and cold: compare shares)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 600 --csv --log-file gpurun_out/r02_launches_step.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-calib --no-config4 --no-big-roofline > gpurun_out/ncu_l1.log 2>&1
# launch list of one calibration (stage 1 + stage 2 + extrinsic), 20 scans
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02_launches_calib.csv python scripts/chain_timing.py ncu > gpurun_out/ncu_l2.log 2>&1
# k_score (FP32 pipe), k_gmm_iter (fp64 pipe): full captures, raw pages
timeout 600 ncu --set full --clock-control none -k regex:"k_score|k_gmm_iter|k_select" -c 4 -f -o gpurun_out/r02_other python scripts/bench_other_kernels.py > gpurun_out/ncu_other.log 2>&1
ncu -i gpurun_out/r02_other.ncu-rep --page raw --csv > gpurun_out/r02_other_raw.csv
tail -c 400 gpurun_out/r02_bench.json
