#!/bin/bash
# Round-2 call L (8 GPUs): the N=4 / N=8 parity cases, 256^3 and 512^3 at N=8 with the round's final defaults.
set -u
out=gpurun_out/r02l
mkdir -p "$out"
timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -q -k "4-64 or 8-96 or 2-96" > "$out/tests_dist.log" 2>&1; echo "dist tests exit $?" | tee -a "$out/tests_dist.log"; tail -4 "$out/tests_dist.log"
run() { # tag N grid extra
  local tag=$1 N=$2 G=$3; shift 3
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $N --grid $G --steps 3 --warmup 3 --no-cpu-baseline "$@" > "$out/${tag}.json" 2> "$out/${tag}.err"
  echo "$tag: exit $? $(tail -1 "$out/${tag}.json" | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["value"], d["details"]["pcg_iterations"], d["details"]["final_rel_residual"], d["details"]["true_rel_residual"], d["e2e"]["value"], d["gpu_launches"])' 2>&1 | tail -1)"
}
run n8_256 8 256
run n8_512 8 512 --share-hierarchy
