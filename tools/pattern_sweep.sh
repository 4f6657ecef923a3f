#!/bin/bash
# Developer tool (run under gpurun, one GPU): validate and time the opt-in csr-pattern8 kernel, and give the other
# not-yet-run additions (smoothed aggregation, GMRES) their first GPU run.
#   gpurun --timeout 1500 -- 'bash tools/pattern_sweep.sh'
# 1. opt-in parity tests  2. launch-shape sweep on the 256^3 Jacobi sweep  3. whole-solve bench with the kernel selected
# 4. one ncu --set full capture of the pattern Jacobi kernel.  Everything lands in gpurun_out/pattern/.
set -u
out=gpurun_out/pattern
mkdir -p "$out"
SPARSH_TEST_PATTERN=1 timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k pattern > "$out/tests.log" 2>&1
echo "tests exit $?" | tee -a "$out/tests.log"
grep -q "passed" "$out/tests.log" || { tail -30 "$out/tests.log"; exit 1; }
# same tests through the TMA-staged variant of the kernel (the switch is read once per process)
SPARSH_TEST_PATTERN=1 SPARSH_PATTERN_TMA=1 timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k pattern \
  > "$out/tests_tma.log" 2>&1
echo "tests (TMA variant) exit $?" | tee -a "$out/tests_tma.log"
tail -3 "$out/tests_tma.log"
# the other two additions that have not had a GPU run yet: smoothed-aggregation hierarchies and GMRES(m)
SPARSH_TEST_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "smoothed_aggregation or gmres" \
  > "$out/tests_experimental.log" 2>&1
echo "experimental tests exit $?" | tee -a "$out/tests_experimental.log"
tail -5 "$out/tests_experimental.log"
for rpt in 2 4 8; do
  for jb in 2 4; do
    [ "$rpt" = 8 ] && [ "$jb" = 4 ] && continue
    SPARSH_PATTERN=2 SPARSH_PATTERN_RPT=$rpt SPARSH_PATTERN_JB=$jb timeout 300 python tools/perf_probe.py --n 256 --reps 20 \
      --families all 2>&1 | grep -E "^pattern|^dict128 +jacobi|^# default" | sed "s/^/rpt=$rpt jb=$jb  /" | tee -a "$out/sweep.log"
  done
done
SPARSH_PATTERN=2 SPARSH_PATTERN_TMA=1 timeout 300 python tools/perf_probe.py --n 256 --reps 20 --families all 2>&1 \
  | grep -E "^pattern" | sed "s/^/tma          /" | tee -a "$out/sweep.log"
SPARSH_PATTERN=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > "$out/bench_pattern.json" 2> "$out/bench_pattern.err"
tail -1 "$out/bench_pattern.json"
SPARSH_PATTERN=1 SPARSH_PATTERN_TMA=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > "$out/bench_pattern_tma.json" 2> "$out/bench_pattern_tma.err"
tail -1 "$out/bench_pattern_tma.json"
SPARSH_PATTERN=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:csr_pattern_kernel -s 40 -c 1 \
  -o "$out/pattern_jacobi" python bench.py --steps 1 --warmup 1 --no-cpu-baseline --grid 256 > "$out/ncu.log" 2>&1
echo "ncu exit $?"
