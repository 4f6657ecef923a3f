#!/bin/bash
# Round-2 call E (1 GPU): ncu --set full of the csr-pattern8 Jacobi kernels (lean and first variant) at 256^3.
set -u
out=gpurun_out/r02e
mkdir -p "$out"
export SPARSH_PATTERN=1
python tools/prof_jacobi.py > "$out/plain_lean.log" 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:csr_pat2_kernel -s 2 -c 2 -o "$out/pat2_jacobi" \
  python tools/prof_jacobi.py > "$out/ncu_lean.log" 2>&1
echo "ncu lean exit $?"; tail -3 "$out/ncu_lean.log"
SPARSH_PAT2=0 ncu --set full --clock-control none --import-source on -k regex:csr_pattern_kernel -s 2 -c 2 -o "$out/pat1_jacobi" \
  python tools/prof_jacobi.py > "$out/ncu_old.log" 2>&1
echo "ncu old exit $?"; tail -3 "$out/ncu_old.log"
ls -la "$out"
