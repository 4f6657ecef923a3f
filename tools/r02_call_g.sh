#!/bin/bash
# Round-2 call G (1 GPU): lean pattern kernel v3 (successor-tile L2 prefetch, table from global memory): parity + launch
# shapes, with and without the prefetch; the same kernels compiled with plain instead of non-coherent gathers.
set -u
out=gpurun_out/r02g
mkdir -p "$out"
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pattern or level_ops or coarse or edge" > "$out/tests.log" 2>&1; echo "tests exit $?"; tail -3 "$out/tests.log"
probe() { # tag env...
  local tag=$1; shift
  env SPARSH_PATTERN=2 "$@" timeout 300 python tools/perf_probe.py --n 256 --reps 20 --families pattern 2>&1 \
    | grep -E "^pattern|^default|not run" | sed "s/^/$tag  /" | tee -a "$out/sweep.log"
}
probe "pf=on  rpt=2" SPARSH_PAT2_RPT=2
probe "pf=off rpt=2" SPARSH_PAT2_RPT=2 SPARSH_PAT2_PF=0
probe "pf=on  rpt=4" SPARSH_PAT2_RPT=4
probe "pf=on  rpt=1" SPARSH_PAT2_RPT=1
probe "pf=2x  rpt=2" SPARSH_PAT2_RPT=2 SPARSH_PAT2_PF=2368
probe "pf=.5x rpt=2" SPARSH_PAT2_RPT=2 SPARSH_PAT2_PF=592
probe "PLAIN pf=on rpt=2" SPARSH_PAT2_RPT=2 SPARSH_LIB_OVERRIDE=$PWD/sparsh_amg_b200/lib/libsparsh_b200_plain.so
probe "PLAIN pf=off rpt=2" SPARSH_PAT2_RPT=2 SPARSH_PAT2_PF=0 SPARSH_LIB_OVERRIDE=$PWD/sparsh_amg_b200/lib/libsparsh_b200_plain.so
SPARSH_PATTERN=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > "$out/bench_pattern.json" 2> "$out/bench_pattern.err"; tail -1 "$out/bench_pattern.json" | cut -c1-300; tail -3 "$out/bench_pattern.err"
