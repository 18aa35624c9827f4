#!/bin/bash
# Same-box A/B of two builds: WJ_LIB selects the library.   bash scripts/gpu_ab.sh <out-subdir>
set -u
O=gpurun_out/${1:-ab}
mkdir -p $O
PREV=wavjepa_b200/libwavjepa_prev.so
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "attention or layernorm" > $O/pytest_kernels.log 2>&1; echo "rc=$?" >> $O/pytest_kernels.log
tail -3 $O/pytest_kernels.log
timeout 200 python scripts/bench_attn.py > $O/attn_new.txt 2>&1
WJ_LIB=$PREV timeout 200 python scripts/bench_attn.py > $O/attn_prev.txt 2>&1
timeout 100 python scripts/bench_ln.py > $O/ln_new.txt 2>&1
WJ_LIB=$PREV timeout 100 python scripts/bench_ln.py > $O/ln_prev.txt 2>&1
for i in 1 2; do
  WJ_LIB=$PREV timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-gpu-baseline > $O/bench_prev_$i.json 2> $O/bench_prev_$i.err
  timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-gpu-baseline > $O/bench_new_$i.json 2> $O/bench_new_$i.err
done
WJ_LIB=$PREV timeout 300 python bench.py --config nat --no-cpu-baseline --no-gpu-baseline > $O/nat_prev.json 2> $O/nat_prev.err
timeout 300 python bench.py --config nat --no-cpu-baseline --no-gpu-baseline > $O/nat_new.json 2> $O/nat_new.err
for f in $O/bench_*.json $O/nat_*.json; do echo $f; python -c "
import json,sys
l=[x for x in open('$f').read().splitlines() if x.startswith('{')]
d=json.loads(l[-1]) if l else {}
print(d.get('value'), d.get('ms_per_step'), d.get('clocks',{}).get('sm_mhz'), {k:v for k,v in (d.get('kernel_ms_per_step') or {}).items() if 'attn' in k or 'layernorm' in k})
"; done
