#!/bin/bash
# Round-2 call U (1 GPU): smoke() and the default bench of the final code (e2e steps timed one by one).
set -u
out=gpurun_out/r02u
mkdir -p "$out"
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > "$out/smoke.log" 2>&1; echo "smoke exit $?"; tail -2 "$out/smoke.log"
timeout 600 python bench.py > "$out/bench_n1.json" 2> "$out/bench_n1.err"; echo "bench exit $?"
tail -1 "$out/bench_n1.json" | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("N=1", d["value"], d["e2e"], d["gpu_launches"], round(d["roofline"]["frac"],3), d["cpu_baseline"]["value"], d["clocks"], "upload", d["details"]["upload_seconds"], "setup", d["details"]["host_setup_seconds"])'
