#!/bin/bash
# Run on the GPU box (gpurun): one ncu --set full capture of k_sweep on the bench's batch (20 x 500 k points, cold L2), DRAM bytes
# per launch written to profiles/sweep_traffic.json together with the fingerprint of the kernel sources, the raw metrics of
# the capture to gpurun_out/ (the .ncu-rep stays scratch).
set -e
mkdir -p gpurun_out
python scripts/profile_sweep.py 128 500000 5 > gpurun_out/sweep_plain.log 2>&1          # the command exits 0 without ncu first
ncu --set full --clock-control none --import-source on -k regex:k_sweep -s 3 -c 2 -f -o gpurun_out/r02_sweep python scripts/profile_sweep.py 128 500000 5 > gpurun_out/ncu_sweep.log 2>&1
ncu -i gpurun_out/r02_sweep.ncu-rep --page raw --csv > gpurun_out/r02_sweep_raw.csv
python - <<'PY'
import csv, json, sys, os, datetime
sys.path.insert(0, os.getcwd())
import bench
rows = list(csv.reader(open("gpurun_out/r02_sweep_raw.csv")))
h = rows[0]
K = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
def col(name):                       # (every column has its own unit)
    i = h.index(name)
    return [float(r[i].replace(",", "")) * K[rows[1][i]] for r in rows[2:]]
rd, wr = col("dram__bytes_read.sum"), col("dram__bytes_write.sum")
out = {"dram_bytes_per_launch": int((sum(rd) + sum(wr)) / len(rd)), "dram_bytes_read": int(sum(rd) / len(rd)),
       "dram_bytes_write": int(sum(wr) / len(wr)), "kernel_source_sha": bench.kernel_source_sha(),
       "source": "ncu --set full --clock-control none, %d launches of k_sweep on 128 x 500k points (the bench's roofline batch), L2 flushed before each (scripts/capture_sweep_traffic.sh, %s)" % (len(rd), datetime.date.today().isoformat())}
json.dump(out, open("gpurun_out/sweep_traffic.json", "w"))
print(out)
PY
