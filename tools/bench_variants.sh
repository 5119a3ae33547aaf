#!/bin/bash
# Bench every library variant under bias.jl_b200/csrc/_exp (see tools/build_variant.sh) with the headline workload.
for so in "" bias.jl_b200/csrc/_exp/libbias_*.so; do
  name=${so:-default}
  BIAS_CUDA_LIB=$so timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-hbm-regime --e2e-steps 1 2>/dev/null | tail -1 | \
    python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$name', round(d['ms_per_step'],2), d['sweep_stats']['n_moved_last'])"
done
