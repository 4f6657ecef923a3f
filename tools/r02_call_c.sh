#!/bin/bash
# Round-2 third GPU call (4 GPUs): the N=4 and N=2 legs of C5 (512^3) and of the 256^3 headline curve.
set -u
out=gpurun_out/r02c
mkdir -p "$out"
run() { # N grid extra...
  local N=$1 G=$2; shift 2
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 295$N$N \
    bench.py --gpus $N --grid $G --steps 3 --warmup 3 --no-cpu-baseline "$@" > "$out/bench${G}_n$N.json" 2> "$out/bench${G}_n$N.err"
  echo "N=$N $G^3 exit $?"; tail -1 "$out/bench${G}_n$N.json" | cut -c1-1200; tail -3 "$out/bench${G}_n$N.err"
}
run 4 512 --share-hierarchy
run 2 512 --share-hierarchy
run 4 256
run 2 256
