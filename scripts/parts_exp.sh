for F in 20; do for P in 1 2; do echo "F=$F parts=$P"; RCLC_SWEEP_PARTS=$P python scripts/sweep_timeline.py $F 2>&1 | grep -v "late task"; done; done
