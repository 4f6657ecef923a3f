#!/bin/bash
# Round-2 call K (2 GPUs): device-side csr-pattern8 encoder, two-stream strips as the default: GPU suite, upload timing,
# N=1 and N=2 benches.
set -u
out=gpurun_out/r02k
mkdir -p "$out"
timeout 1500 python -m pytest tests -m gpu -x -q > "$out/tests.log" 2>&1; echo "tests exit $?" | tee -a "$out/tests.log"; tail -6 "$out/tests.log"
CUDA_VISIBLE_DEVICES=0 SPARSH_UPLOAD_TIMING=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > "$out/bench_n1.json" 2> "$out/bench_n1.err"
tail -1 "$out/bench_n1.json" | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("N=1", d["value"], d["e2e"]["value"], d["details"]["upload_seconds"], d["details"]["host_setup_seconds"], d["roofline"]["ms_per_launch"])'
grep "upload 16777216 x 16777216\|upload 8388608 x" "$out/bench_n1.err" | head -12
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29523 \
    bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > "$out/n2_default.json" 2> "$out/n2_default.err"
tail -1 "$out/n2_default.json" | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("N=2", d["value"], d["details"]["pcg_iterations"], d["e2e"]["value"], d["gpu_launches"])'
