P='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["ms_per_step"], d["e2e"]["ms_per_step"])'
echo plain; python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "$P"
echo OMP1; OMP_NUM_THREADS=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "$P"
echo torchrun1; python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "$P"
