#!/bin/bash
# Round-2 call O (2 GPUs): the small levels of the cycle as one cooperative kernel (csrc/tail.cu): parity, then solves with
# and without it, thresholds, N=1 and N=2.
set -u
out=gpurun_out/r02o
mkdir -p "$out"
timeout 1500 python -m pytest tests -m gpu -x -q > "$out/tests.log" 2>&1; echo "tests exit $?" | tee -a "$out/tests.log"; tail -6 "$out/tests.log"
show() { python -c 'import sys,json; d=json.loads(sys.stdin.read().splitlines()[-1]); print(sys.argv[1], d["value"], d["details"]["pcg_iterations"], d["e2e"]["value"], d["gpu_launches"], "upload", round(d["details"].get("upload_seconds", 0),3))' "$1"; }
for rows in 131072 0 32768 524288 1100000; do
  CUDA_VISIBLE_DEVICES=0 SPARSH_TAIL_ROWS=$rows timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > "$out/n1_tail$rows.json" 2> "$out/n1_tail$rows.err"; show "N=1 tail_rows=$rows" < "$out/n1_tail$rows.json"
done
for rows in 131072 0 524288; do
  SPARSH_TAIL_ROWS=$rows timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29523 \
    bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > "$out/n2_tail$rows.json" 2> "$out/n2_tail$rows.err"; show "N=2 tail_rows=$rows" < "$out/n2_tail$rows.json"
done
