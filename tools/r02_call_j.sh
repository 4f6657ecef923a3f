#!/bin/bash
# Round-2 call J (8 GPUs): 256^3 at N=8 and N=4, one launch per operator (default) against strips on a second stream;
# 512^3 at N=8 with the round's final kernels.
set -u
out=gpurun_out/r02j
mkdir -p "$out"
run() { # tag N grid env...
  local tag=$1 N=$2 G=$3; shift 3
  env "$@" timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $N --grid $G --steps 3 --warmup 3 --no-cpu-baseline $EXTRA > "$out/${tag}.json" 2> "$out/${tag}.err"
  echo "$tag: exit $? $(tail -1 "$out/${tag}.json" | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["value"], d["details"]["pcg_iterations"], d["details"]["final_rel_residual"], d["e2e"]["value"], d["gpu_launches"])' 2>&1 | tail -1)"
}
EXTRA=""
run n8_default 8 256 X=1
run n8_nomerge 8 256 SPARSH_DIST_MERGE=0
run n4_default 4 256 X=1
run n4_nomerge 4 256 SPARSH_DIST_MERGE=0
EXTRA="--share-hierarchy"
run n8_512 8 512 X=1
