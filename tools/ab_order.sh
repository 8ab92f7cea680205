# A/B of the evaluation order (order.cu) on ONE box: bash tools/ab_order.sh [trees]
T=${1:-1000000}
for rep in 1 2; do
  for v in "" 1; do
    if [ -n "$v" ]; then export PDE_B200_NO_ORDER=1; tag="as given"; else unset PDE_B200_NO_ORDER; tag="sorted  "; fi
    python bench.py --trees $T --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --ref-wall 0 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$tag', d['ms_per_step'], 'ms/step', d['value'], 'launches', d['gpu_launches'], 'depth4 kernel', d['depth4_validation']['kernel_ms'], 'wall', d['depth4_validation']['wall_ms_host_strings_to_survivor_bits'], 'd4dev', d['depth4_device_resident']['wall_ms_operand_strings_to_survivor_flags'])"
  done
done
