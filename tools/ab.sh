#!/bin/bash
# A/B the variant libraries on the GPU box: tools/ab.sh TREES v1 v2 ...  -> gpurun_out/ab_<v>.json, one summary line each
TREES=$1; shift
for v in "$@"; do
  PDE_B200_LIB=variants/libpde_$v.so python bench.py --trees $TREES --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
  python - "$v" <<'PY'
import json,sys
v=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/ab_{v}.json").read().strip().splitlines()[-1])
    d4 = d.get("depth4_validation") or {}
    print(f"{v:24s} ms/step {d['ms_per_step']:9.2f}  evals/s {d['value']:.4g}  frac {d['roofline']['frac']:.4f}  depth4 kernel {d4.get('kernel_ms', float('nan')):.2f} ms wall {d4.get('wall_ms_host_strings_to_survivor_bits', float('nan')):.1f} ms")
except Exception as e:
    print(v, "FAILED", e)
PY
done
