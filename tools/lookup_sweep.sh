#!/bin/bash
# lookup kernel variants at the KITTI-size workload: output path (TMA / 128-bit stores) x pixels per CTA x mono form
for mono in factored packed; do
  for tma in 1 0; do
    for tile in 32 64; do
      SA_B200_LOOKUP_TMA=$tma SA_B200_LOOKUP_TILE=$tile python bench.py --extras 0 --no-cpu-baseline --steps 20 --mono $mono 2>/dev/null | \
        python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('mono=$mono tma=$tma tile=$tile', 'launch_us', d['roofline']['launch_us'], 'ms_step', d['ms_per_step'])"
    done
  done
done
