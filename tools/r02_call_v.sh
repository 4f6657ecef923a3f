#!/bin/bash
# Round-2 call V (2 GPUs): the N=2 bench with the e2e steps timed one by one (as at N=1).
set -u
out=gpurun_out/r02v
mkdir -p "$out"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 2 > "$out/bench_n2.json" 2> "$out/bench_n2.err"; echo "N=2 exit $?"
tail -1 "$out/bench_n2.json" | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("N=2", d["value"], d["details"]["pcg_iterations"], d["e2e"], d["gpu_launches"], d["details"]["true_rel_residual"])'
tail -3 "$out/bench_n2.err"
