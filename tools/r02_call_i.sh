#!/bin/bash
# Round-2 call I (2 GPUs): after the push/fused-push ordering fix: full GPU suite, N=2 A/B.
set -u
out=gpurun_out/r02i
mkdir -p "$out"
timeout 1500 python -m pytest tests -m gpu -x -q > "$out/tests.log" 2>&1; echo "tests exit $?" | tee -a "$out/tests.log"; tail -6 "$out/tests.log"
run() { # tag, env...
  local tag=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29523 \
    bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > "$out/n2_$tag.json" 2> "$out/n2_$tag.err"
  echo "$tag: exit $? $(tail -1 "$out/n2_$tag.json" | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["value"], d["details"]["pcg_iterations"], d["e2e"]["value"], d["gpu_launches"], d["roofline"].get("kernel"), d["roofline"].get("ms_per_launch"))' 2>&1 | tail -1)"
}
run default X=1
run nomerge SPARSH_DIST_MERGE=0
run dict SPARSH_PATTERN=0
run nccl_coll SPARSH_PEER_COLL=0
