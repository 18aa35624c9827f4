#!/bin/bash
# Same-box A/B of two builds or two routings (run through gpurun; box-to-box clocks differ by +-3 %, so only runs on ONE
# box compare).  usage: bash scripts/gpu_ab.sh <out-subdir> [ENV=VALUE for arm B, e.g. WJ_LIB=wavjepa_b200/libwavjepa_prev.so
# or WJ_GEMM_PAIR_GRADS=0]
set -u
O=gpurun_out/${1:-ab}
B_ENV=${2:-WJ_GEMM_PAIR_GRADS=0}
mkdir -p $O
for i in 1 2; do
  env $B_ENV timeout -s KILL 300 python bench.py --steps 20 --no-cpu-baseline --no-gpu-baseline > $O/bench_B_$i.json 2> $O/bench_B_$i.err
  timeout -s KILL 300 python bench.py --steps 20 --no-cpu-baseline --no-gpu-baseline > $O/bench_A_$i.json 2> $O/bench_A_$i.err
done
for f in $O/bench_*.json; do echo $f; python -c "
import json
l=[x for x in open('$f').read().splitlines() if x.startswith('{')]
d=json.loads(l[-1]) if l else {}
print(d.get('value'), d.get('ms_per_step'), d.get('loss'), d.get('clocks',{}).get('sm_mhz'), (d.get('roofline') or {}).get('frac'))"; done
