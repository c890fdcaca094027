#!/bin/bash
# lookup L2-prefetch distance sweep (SA_B200_LOOKUP_PF, CTAs ahead): prints launch_us / ms_per_step per setting
for pf in 0 740 1480 2960 5920; do
  for mono in factored packed; do
    SA_B200_LOOKUP_PF=$pf python bench.py --extras 0 --no-cpu-baseline --steps 20 --mono $mono 2>/dev/null | \
      python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('pf=$pf mono=$mono', 'launch_us', d['roofline']['launch_us'], 'ms_step', d['ms_per_step'])"
  done
done
