#!/bin/bash
# Round-2 call Q (2 GPUs): final confirmation of the round's default configuration: full GPU suite (incl. the N=2 parity and
# C++ multi-GPU cases), smoke(), the N=1 bench with its CPU baseline, the reference arm, the N=2 bench.
set -u
out=gpurun_out/r02q
mkdir -p "$out"
timeout 1500 python -m pytest tests -m gpu -x -q > "$out/tests.log" 2>&1; echo "tests exit $?" | tee -a "$out/tests.log"; tail -5 "$out/tests.log"
CUDA_VISIBLE_DEVICES=0 timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > "$out/smoke.log" 2>&1; echo "smoke exit $?"; tail -2 "$out/smoke.log"
CUDA_VISIBLE_DEVICES=0 timeout 900 python bench.py > "$out/bench_n1.json" 2> "$out/bench_n1.err"; echo "bench exit $?"
tail -1 "$out/bench_n1.json" | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("N=1", d["value"], d["e2e"]["value"], d["gpu_launches"], d["roofline"]["kernel"], round(d["roofline"]["frac"],3), d["roofline"]["traffic"], round(d["roofline_csr"]["frac"],3), round(d["roofline_csr"]["spmv"]["frac_of_8TBs_nominal"],3), d["cpu_baseline"]["value"], d["clocks"], "upload", d["details"]["upload_seconds"])'
CUDA_VISIBLE_DEVICES=0 timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > "$out/bench_reference.json" 2> "$out/bench_reference.err"; echo "reference exit $?"
tail -1 "$out/bench_reference.json" | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("REF", d["value"], d["details"]["converged_solve"], d["cpu_baseline"]["cores"], d["config"])'
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 2 > "$out/bench_n2.json" 2> "$out/bench_n2.err"; echo "N=2 exit $?"
tail -1 "$out/bench_n2.json" | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("N=2", d["value"], d["details"]["pcg_iterations"], d["e2e"]["value"], d["gpu_launches"], d["roofline"].get("kernel"), d["config"])'
