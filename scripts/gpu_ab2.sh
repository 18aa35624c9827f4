#!/bin/bash
set -u
O=gpurun_out/${1:-ab2}
mkdir -p $O
PREV=wavjepa_b200/libwavjepa_prev.so
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "attention or layernorm" > $O/pytest_kernels.log 2>&1; echo "rc=$?" >> $O/pytest_kernels.log
tail -3 $O/pytest_kernels.log
timeout 200 python scripts/bench_attn.py > $O/attn_new.txt 2>&1
WJ_LIB=$PREV timeout 200 python scripts/bench_attn.py > $O/attn_prev.txt 2>&1
cat $O/attn_new.txt $O/attn_prev.txt | grep -v "^  \|Traceback"
WJ_LIB=$PREV timeout 300 python bench.py --config nat --no-cpu-baseline --no-gpu-baseline > $O/nat_prev.json 2> $O/nat_prev.err
timeout 300 python bench.py --config nat --no-cpu-baseline --no-gpu-baseline > $O/nat_new.json 2> $O/nat_new.err
timeout 300 python -m pytest tests/test_gpu_model.py -x -q -m gpu -k "nat or Nat or binaural" > $O/pytest_nat.log 2>&1; tail -2 $O/pytest_nat.log
for f in $O/nat_*.json; do echo $f; python -c "
import json,sys
l=[x for x in open('$f').read().splitlines() if x.startswith('{')]
d=json.loads(l[-1]) if l else {}
print(d.get('value'), d.get('ms_per_step'), d.get('loss'), d.get('clocks',{}).get('sm_mhz'), {k:v for k,v in (d.get('kernel_ms_per_step') or {}).items() if 'attn' in k or 'layernorm' in k})
"; done
