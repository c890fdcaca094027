#!/bin/bash
for ns in 0 200 400 600 800 1000 1200; do
  SA_B200_CORR_PACK_PACE_NS=$ns python bench.py --extras 0 --no-cpu-baseline --steps 20 2>/dev/null | \
    python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('pace_ns=$ns', 'corr_pack_us', d['kernels']['corr_pack_tf32']['us'], 'ms_step', d['ms_per_step'])"
done
