#!/bin/bash
# Same-box A/B of two builds of libaclip_b200.so through the full bench step (box-to-box clock spread
# under the power cap is larger than most kernel changes): scripts/ab_bench.sh <other.so> [rounds]
# Prints frames/s and the per-kind kernel times of each run, alternating OTHER / CURRENT.
other=${1:?path of the other library build}; rounds=${2:-2}
for r in $(seq 1 "$rounds"); do
  for which in OTHER CURRENT; do
    if [ "$which" = OTHER ]; then export ACLIP_LIB="$other"; else unset ACLIP_LIB; fi
    python bench.py --no-cpu-baseline --steps 20 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); k=d['kernels']
print('$which', round(d['value']), d['clocks']['sm_mhz'], {n: k[n]['ms'] for n in ('gemm_tcgen05','vit_attention','layernorm')})"
  done
done
