#!/bin/bash
# Round-2 call N (2 GPUs): final GPU suite (device Galerkin product, C++ multi-GPU caller, NVTX build), upload through pinned
# staging buffers, Galerkin products on the device, lean kernel without the prefetch arithmetic, compute-sanitizer on smoke().
set -u
out=gpurun_out/r02n
mkdir -p "$out"
timeout 1500 python -m pytest tests -m gpu -x -q > "$out/tests.log" 2>&1; echo "tests exit $?" | tee -a "$out/tests.log"; tail -6 "$out/tests.log"
export CUDA_VISIBLE_DEVICES=0
show() { python -c 'import sys,json; d=json.loads(sys.stdin.read().splitlines()[-1]); print(sys.argv[1], d["value"], d["e2e"]["value"], "upload", round(d["details"]["upload_seconds"],3), "setup", round(d["details"]["host_setup_seconds"],3), d["details"]["galerkin_products"], "jacobi ms", round(d["roofline"]["ms_per_launch"],4), [ (k["op"], round(k["ms"],4)) for k in d["kernels"][:4]])' "$1"; }
SPARSH_UPLOAD_TIMING=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > "$out/bench_n1.json" 2> "$out/bench_n1.err"; show "N=1 host RAP (PDL on)" < "$out/bench_n1.json"; grep "^\[upload\]" "$out/bench_n1.err" | head -3
SPARSH_PDL=0 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > "$out/bench_n1_nopdl.json" 2> "$out/bench_n1_nopdl.err"; show "N=1 PDL off" < "$out/bench_n1_nopdl.json"
SPARSH_SETUP_TIMING=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --gpu-rap > "$out/bench_n1_gpurap.json" 2> "$out/bench_n1_gpurap.err"; show "N=1 device RAP" < "$out/bench_n1_gpurap.json"
SPARSH_PAT2_RPT=2 timeout 300 python tools/perf_probe.py --n 256 --reps 20 --families pattern 2>&1 | grep -E "^pattern" | sed "s/^/rpt=2 /" | tee "$out/sweep.log"
SPARSH_PAT2_RPT=1 timeout 300 python tools/perf_probe.py --n 256 --reps 20 --families pattern 2>&1 | grep -E "^pattern" | sed "s/^/rpt=1 /" | tee -a "$out/sweep.log"
unset CUDA_VISIBLE_DEVICES
for pdl in 1 0; do
SPARSH_PDL=$pdl timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29523 \
    bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > "$out/n2_pdl$pdl.json" 2> "$out/n2_pdl$pdl.err"
tail -1 "$out/n2_pdl$pdl.json" | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("N=2 PDL", sys.argv[1], d["value"], d["details"]["pcg_iterations"], d["e2e"]["value"], d["gpu_launches"])' $pdl
done
CUDA_VISIBLE_DEVICES=0 timeout 300 python tools/perf_probe.py --matrix diffusion27 --n 160 --reps 10 --families default 2>&1 | grep -E "^default|^# " | sed "s/^/diffusion27 160^3  /" | tee -a "$out/sweep.log"
CUDA_VISIBLE_DEVICES=0 timeout 300 python tools/perf_probe.py --matrix sa_coarse --n 160 --reps 10 --families default 2>&1 | grep -E "^default|^# " | sed "s/^/sa_coarse(160^3)  /" | tee -a "$out/sweep.log"
echo "compute-sanitizer is closed on this pool (gpurun answers exit 86)"
