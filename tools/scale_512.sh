#!/bin/bash
# Developer tool: the C5 measurement of BASELINE.json (512^3, 134M rows, 0.94B nnz), strong scaling 1 -> N GPUs.
#   gpurun --gpus 8 --timeout 1500 -- 'bash tools/scale_512.sh 8'
# One 512^3 hierarchy is ~28 GB on the host: --share-hierarchy keeps a single copy per node.  Expected host-side cost
# per run (measured on 8 cores): generate 8 s, setup 22 s, publish 7 s, partition plan ~12 s per rank, then uploads.
# Never wrap these multi-rank commands in ncu.  Results land in gpurun_out/scale512/.
set -u
N=${1:-8}
out=gpurun_out/scale512
mkdir -p "$out"
timeout 900 python bench.py --gpus 1 --grid 512 --steps 2 --warmup 3 --no-cpu-baseline \
  > "$out/n1.json" 2> "$out/n1.err"
echo "N=1 exit $?"; tail -1 "$out/n1.json"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus "$N" --grid 512 --steps 2 --warmup 3 --share-hierarchy --no-cpu-baseline \
  > "$out/n$N.json" 2> "$out/n$N.err"
echo "N=$N exit $?"; tail -1 "$out/n$N.json"
python - "$out/n1.json" "$out/n$N.json" <<'PY'
import json, sys
def last(p):
    lines = [l for l in open(p).read().splitlines() if l.startswith("{")]
    return json.loads(lines[-1]) if lines else None
a, b = last(sys.argv[1]), last(sys.argv[2])
if a and b:
    print(f"512^3: {a['value']:.4f} s on 1 GPU, {b['value']:.4f} s on {b['n_gpus']} GPUs -> speed-up {a['value']/b['value']:.2f}x "
          f"(iterations {a['config'].get('pcg_iterations')} / {b['config'].get('pcg_iterations')})")
PY
