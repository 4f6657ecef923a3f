#!/bin/bash
# Round-2 call M (1 GPU): threaded upload; 512^3 on one GPU with the final kernels; ncu launch list and --set full of the
# final Jacobi kernel.
set -u
out=gpurun_out/r02m
mkdir -p "$out"
timeout 1200 python -m pytest tests -m gpu -x -q > "$out/tests.log" 2>&1; echo "tests exit $?" | tee -a "$out/tests.log"; tail -5 "$out/tests.log"
SPARSH_UPLOAD_TIMING=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > "$out/bench_n1.json" 2> "$out/bench_n1.err"
tail -1 "$out/bench_n1.json" | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("N=1 256", d["value"], d["e2e"]["value"], "upload", d["details"]["upload_seconds"], "setup", d["details"]["host_setup_seconds"], d["roofline"]["ms_per_launch"])'
grep "^\[upload\]" "$out/bench_n1.err" | head
SPARSH_UPLOAD_THREADS=1 timeout 600 python bench.py --steps 1 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c 'import sys,json; d=json.loads(sys.stdin.read().splitlines()[-1]); print("sequential upload", d["details"]["upload_seconds"])'
timeout 1200 python bench.py --gpus 1 --grid 512 --steps 2 --warmup 3 --no-cpu-baseline > "$out/bench512_n1.json" 2> "$out/bench512_n1.err"
tail -1 "$out/bench512_n1.json" | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("N=1 512", d["value"], d["details"]["pcg_iterations"], d["details"]["final_rel_residual"], d["details"]["true_rel_residual"], d["e2e"]["value"], "upload", d["details"]["upload_seconds"], "setup", d["details"]["host_setup_seconds"])'
# launch list of a short solve (every launch with its device time), then the top kernel in full
python bench.py --steps 1 --warmup 1 --max-iter 2 --profile --no-cpu-baseline > "$out/plain_profile.log" 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file "$out/launches.csv" \
  python bench.py --steps 1 --warmup 1 --max-iter 2 --profile --no-cpu-baseline > "$out/ncu_list.log" 2>&1
echo "ncu list exit $?"
python tools/prof_jacobi.py > "$out/plain_jacobi.log" 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:csr_pat2_kernel -s 2 -c 2 -o "$out/pat2v3_jacobi" \
  python tools/prof_jacobi.py > "$out/ncu_full.log" 2>&1
echo "ncu full exit $?"; ls -la "$out" | head -20
