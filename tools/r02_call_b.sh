#!/bin/bash
# Round-2 second GPU call (8 GPUs): the C5 measurement, 512^3 AMG-PCG row-sharded over 8 B200 (BASELINE.json configs[4]).
set -u
out=gpurun_out/r02b
mkdir -p "$out"
SPARSH_SETUP_TIMING=1 timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus 8 --grid 512 --steps 3 --warmup 3 --share-hierarchy --no-cpu-baseline \
  > "$out/bench512_n8.json" 2> "$out/bench512_n8.err"
echo "N=8 512^3 exit $?"; tail -1 "$out/bench512_n8.json" | cut -c1-1500; tail -20 "$out/bench512_n8.err"
nproc; free -g | head -2
