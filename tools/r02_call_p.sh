#!/bin/bash
# Round-2 call P (1 GPU): the fused tail as one 16-CTA cluster (SPARSH_TAIL_MODE=2) against per-kernel launches.
set -u
out=gpurun_out/r02p
mkdir -p "$out"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fused_tail or graph_and_direct or vcycle_amg_pcg" > "$out/tests.log" 2>&1; echo "tests exit $?"; tail -4 "$out/tests.log"
show() { python -c 'import sys,json; d=json.loads(sys.stdin.read().splitlines()[-1]); print(sys.argv[1], d["value"], d["details"]["pcg_iterations"], d["e2e"]["value"], d["gpu_launches"])' "$1"; }
for cfg in "0 0" "2 32768" "2 65536" "2 16384" "1 32768"; do
  set -- $cfg
  SPARSH_TAIL_MODE=$1 SPARSH_TAIL_ROWS=$2 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > "$out/n1_mode$1_rows$2.json" 2> "$out/n1_mode$1_rows$2.err"; show "N=1 mode=$1 rows=$2" < "$out/n1_mode$1_rows$2.json"
done
