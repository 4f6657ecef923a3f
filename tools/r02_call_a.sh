#!/bin/bash
# Round-2 first GPU call (1 GPU): full GPU test suite (pattern8 / SA / GMRES tests no longer gated), the csr-pattern8
# launch-shape sweep, whole-solve benches with the pattern kernels, and the 512^3 single-GPU leg of C5.
set -u
out=gpurun_out/r02a
mkdir -p "$out"
timeout 1500 python -m pytest tests -m gpu -x -q > "$out/tests.log" 2>&1; echo "tests exit $?" | tee -a "$out/tests.log"; tail -5 "$out/tests.log"
SPARSH_PATTERN_TMA=1 timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k pattern > "$out/tests_tma.log" 2>&1
echo "tests (TMA variant) exit $?" | tee -a "$out/tests_tma.log"; tail -3 "$out/tests_tma.log"
for rpt in 2 4 8; do
  for jb in 2 4; do
    [ "$rpt" = 8 ] && [ "$jb" = 4 ] && continue
    SPARSH_PATTERN=2 SPARSH_PATTERN_RPT=$rpt SPARSH_PATTERN_JB=$jb timeout 300 python tools/perf_probe.py --n 256 --reps 20 \
      --families pattern 2>&1 | grep -E "^pattern|^default|^# default|not run" | sed "s/^/rpt=$rpt jb=$jb  /" | tee -a "$out/sweep.log"
  done
done
SPARSH_PATTERN=2 SPARSH_PATTERN_TMA=1 timeout 300 python tools/perf_probe.py --n 256 --reps 20 --families all 2>&1 \
  | grep -E "^pattern|not run|blas1|torch" | sed "s/^/tma          /" | tee -a "$out/sweep.log"
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --dump-hist "$out/hist256.json" > "$out/bench_dict.json" 2> "$out/bench_dict.err"; tail -1 "$out/bench_dict.json" | cut -c1-600
SPARSH_PATTERN=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > "$out/bench_pattern.json" 2> "$out/bench_pattern.err"; tail -1 "$out/bench_pattern.json" | cut -c1-600
SPARSH_PATTERN=1 SPARSH_PATTERN_TMA=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > "$out/bench_pattern_tma.json" 2> "$out/bench_pattern_tma.err"; tail -1 "$out/bench_pattern_tma.json" | cut -c1-600
SPARSH_SETUP_TIMING=1 timeout 1200 python bench.py --gpus 1 --grid 512 --steps 2 --warmup 3 --no-cpu-baseline --dump-hist "$out/hist512.json" \
  > "$out/bench512_n1.json" 2> "$out/bench512_n1.err"; echo "512 exit $?"; tail -1 "$out/bench512_n1.json" | cut -c1-900; tail -5 "$out/bench512_n1.err"
nvidia-smi --query-gpu=name,memory.used,memory.total --format=csv
nproc; free -g | head -2
