#!/bin/bash
# Round-2 call D (2 GPUs): merged strip+interior launch, peer-memory collectives (no NCCL on the iteration path), the lean
# csr-pattern8 kernel.  GPU suite incl. the 2-GPU parity cases, launch-shape sweep of the lean kernel, benches.
set -u
out=gpurun_out/r02d
mkdir -p "$out"
timeout 1500 python -m pytest tests -m gpu -x -q > "$out/tests.log" 2>&1; echo "tests exit $?" | tee -a "$out/tests.log"; tail -8 "$out/tests.log"
export CUDA_VISIBLE_DEVICES=0
for thr in 256 128; do for rpt in 1 2 4; do
  SPARSH_PATTERN=2 SPARSH_PAT2_RPT=$rpt timeout 300 python tools/perf_probe.py --n 256 --reps 20 --families pattern 2>&1 \
    | grep -E "^pattern$thr|not run" | sed "s/^/lean rpt=$rpt  /" | tee -a "$out/sweep.log"
done; done
SPARSH_PATTERN=2 SPARSH_PAT2=0 timeout 300 python tools/perf_probe.py --n 256 --reps 20 --families pattern 2>&1 | grep -E "^pattern|^default" | sed "s/^/old          /" | tee -a "$out/sweep.log"
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > "$out/bench_dict.json" 2> "$out/bench_dict.err"; tail -1 "$out/bench_dict.json" | cut -c1-400
SPARSH_PATTERN=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > "$out/bench_pattern.json" 2> "$out/bench_pattern.err"; tail -1 "$out/bench_pattern.json" | cut -c1-400; tail -3 "$out/bench_pattern.err"
unset CUDA_VISIBLE_DEVICES
for pat in 0 1; do
SPARSH_PATTERN=$pat timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 \
  bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > "$out/bench256_n2_pat$pat.json" 2> "$out/bench256_n2_pat$pat.err"
echo "N=2 pattern=$pat exit $?"; tail -1 "$out/bench256_n2_pat$pat.json" | cut -c1-700; tail -3 "$out/bench256_n2_pat$pat.err"
done
