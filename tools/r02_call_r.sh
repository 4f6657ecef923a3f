#!/bin/bash
# Round-2 call R (1 GPU): Galerkin products on the device with the transfers staged through pinned buffers, phase times.
set -u
out=gpurun_out/r02r
mkdir -p "$out"
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "galerkin or encoder" > "$out/tests.log" 2>&1; echo "tests exit $?" | tee -a "$out/tests.log"; tail -3 "$out/tests.log"
show() { python -c 'import sys,json; d=json.loads(sys.stdin.read().splitlines()[-1]); print(sys.argv[1], d["value"], d["e2e"]["value"], "upload", round(d["details"]["upload_seconds"],3), "setup", round(d["details"]["host_setup_seconds"],3), d["details"]["galerkin_products"])' "$1"; }
SPARSH_SETUP_TIMING=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --gpu-rap > "$out/bench_n1_gpurap.json" 2> "$out/bench_n1_gpurap.err"; show "N=1 device RAP" < "$out/bench_n1_gpurap.json"
grep -E "^\[rap|RAP on" "$out/bench_n1_gpurap.err" "$out/bench_n1_gpurap.json" | head -30
