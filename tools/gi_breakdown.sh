#!/bin/bash
# grad-input timing with parts of the kernel switched off (SINC_DEBUG_GI: 1 col2im, 2 MMAs, 4 producers)
for d in 0 1 2 4 3 5 6 7; do
  SINC_DEBUG_GI=$d timeout 120 python bench.py --no-cpu --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('debug=$d bwd_input ms', round(d['kernels']['bwd_input']['ms'],4))"
done
