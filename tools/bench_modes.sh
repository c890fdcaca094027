#!/bin/bash
# every bench mode at KITTI size with the current library: name, pairs/s, ms per step, us per lookup, corr_pack us
p() { python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('$1', d['value'], d['ms_per_step'], d['roofline']['launch_us'], d['roofline']['kernels']['corr_pack_tf32']['us'] if 'corr_pack_tf32' in d['roofline'].get('kernels',{}) else '')"; }
python bench.py --extras 0 --no-cpu-baseline --steps 10 2>/dev/null | p default
SA_B200_LOOKUP_TMA=0 python bench.py --extras 0 --no-cpu-baseline --steps 10 --mono packed 2>/dev/null | p packed_tma0
SA_B200_LOOKUP_TMA=1 python bench.py --extras 0 --no-cpu-baseline --steps 10 --mono packed 2>/dev/null | p packed_tma1
python bench.py --extras 0 --no-cpu-baseline --steps 10 --mono aggregated 2>/dev/null | p aggregated
python bench.py --extras 0 --no-cpu-baseline --steps 10 --mono otf 2>/dev/null | p otf
python bench.py --extras 0 --no-cpu-baseline --steps 10 --variant protocol 2>/dev/null | p protocol
python bench.py --extras 0 --no-cpu-baseline --steps 10 --storage fp16 2>/dev/null | p fp16
python bench.py --workload c4_middlebury_1984x2872_tiled --steps 10 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('tiled', d['value'], d['ms_per_step'])"
python bench.py --extras 0 --no-cpu-baseline --steps 10 --workload c3_sceneflow_540x960_b8 2>/dev/null | p c3
