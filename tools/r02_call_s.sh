#!/bin/bash
# Round-2 call S (1 GPU): one device slab per hierarchy + per-thread scratch for the encoder (instead of ~450 device
# allocations): full GPU suite, then the N=1 bench with and without the slab (upload_seconds).
set -u
out=gpurun_out/r02s
mkdir -p "$out"
timeout 900 python -m pytest tests -m gpu -x -q > "$out/tests.log" 2>&1; echo "tests exit $?" | tee -a "$out/tests.log"; tail -4 "$out/tests.log"
show() { python -c 'import sys,json; d=json.loads(sys.stdin.read().splitlines()[-1]); print(sys.argv[1], d["value"], d["e2e"]["value"], "upload", round(d["details"]["upload_seconds"],3), "setup", round(d["details"]["host_setup_seconds"],3), "iters", d["details"]["pcg_iterations"], d["details"]["final_rel_residual"])' "$1"; }
SPARSH_UPLOAD_TIMING=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > "$out/bench_n1_slab.json" 2> "$out/bench_n1_slab.err"; show "N=1 slab" < "$out/bench_n1_slab.json"; grep "^\[upload\]" "$out/bench_n1_slab.err" | tail -2
SPARSH_SLAB=0 SPARSH_UPLOAD_TIMING=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > "$out/bench_n1_noslab.json" 2> "$out/bench_n1_noslab.err"; show "N=1 no slab" < "$out/bench_n1_noslab.json"; grep "^\[upload\]" "$out/bench_n1_noslab.err" | tail -2
