#!/bin/bash
# Round-2 call T (2 GPUs): the multi-GPU parity cases and the C++ multi-GPU caller with the slab allocator, the N=2 bench,
# and what the box's host side looks like (CPU quota, NUMA) for the upload numbers.
set -u
out=gpurun_out/r02t
mkdir -p "$out"
{ echo "nproc $(nproc)"; echo "cpu.max $(cat /sys/fs/cgroup/cpu.max 2>/dev/null)"; echo "affinity $(python -c 'import os; print(len(os.sched_getaffinity(0)))')"; lscpu | grep -E "Model name|Socket|NUMA|^CPU\(s\)"; free -g | head -2; cat /proc/loadavg; } > "$out/host.txt" 2>&1; cat "$out/host.txt"
timeout 600 python -m pytest tests/test_gpu_dist.py tests/test_cpp_dropin.py -m gpu -x -q > "$out/tests.log" 2>&1; echo "tests exit $?" | tee -a "$out/tests.log"; tail -4 "$out/tests.log"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 2 --steps 3 --warmup 3 > "$out/bench_n2.json" 2> "$out/bench_n2.err"; echo "N=2 exit $?"
tail -1 "$out/bench_n2.json" | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("N=2", d["value"], d["details"]["pcg_iterations"], d["e2e"]["value"], d["gpu_launches"], d["details"].get("upload_seconds"))'
cat /proc/loadavg
