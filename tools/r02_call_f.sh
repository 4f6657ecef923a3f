#!/bin/bash
# Round-2 call F (2 GPUs): A/B of the multi-GPU changes at 256^3, N=2 (dict kernels unless stated).
set -u
out=gpurun_out/r02f
mkdir -p "$out"
timeout 900 python -m pytest tests/test_gpu_dist.py -m gpu -x -q > "$out/tests.log" 2>&1; echo "dist tests exit $?"; tail -3 "$out/tests.log"
run() { # tag, env...
  local tag=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29523 \
    bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > "$out/n2_$tag.json" 2> "$out/n2_$tag.err"
  echo "$tag: exit $? $(tail -1 "$out/n2_$tag.json" | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["value"], d["details"]["pcg_iterations"], d["e2e"]["value"], d["gpu_launches"])' 2>&1)"
}
run default X=1
run nomerge SPARSH_DIST_MERGE=0
run nccl_coll SPARSH_PEER_COLL=0
run old_style SPARSH_DIST_MERGE=0 SPARSH_PEER_COLL=0
run pattern SPARSH_PATTERN=1
run pattern_nomerge SPARSH_PATTERN=1 SPARSH_DIST_MERGE=0
run plain_csr SPARSH_DICT=0
