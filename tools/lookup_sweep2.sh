#!/bin/bash
for lib in "" t56v56 t64v64 t48v40; do
  for tma in 1 0; do
    for tile in 32 64; do
      L=""; [ -n "$lib" ] && L=/root/repo/stereoanywhere_b200/lib/variants/libsa_b200_$lib.so
      SA_B200_LIB=$L SA_B200_LOOKUP_TMA=$tma SA_B200_LOOKUP_TILE=$tile python bench.py --extras 0 --no-cpu-baseline --steps 20 2>/dev/null | \
        python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('lib=${lib:-default(t56v48)} tma=$tma tile=$tile', 'launch_us', d['roofline']['launch_us'], 'ms_step', d['ms_per_step'])"
    done
  done
done
