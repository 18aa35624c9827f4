#!/bin/bash
set -u
O=gpurun_out/${1:-ab4}
mkdir -p $O
timeout -s KILL 200 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "cta_pairs" > $O/pytest_pairs.log 2>&1; echo "rc=$?" >> $O/pytest_pairs.log
tail -4 $O/pytest_pairs.log
if ! grep -q "rc=0" $O/pytest_pairs.log; then grep -E "^E |FAILED" $O/pytest_pairs.log | head -20; fi
timeout -s KILL 300 python scripts/bench_gemm_shapes.py --grads > $O/gemm_grads.txt 2>&1; cat $O/gemm_grads.txt
timeout -s KILL 900 python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log
tail -3 $O/pytest_gpu.log
for i in 1 2; do
  WJ_GEMM_PAIR_GRADS=0 timeout -s KILL 300 python bench.py --steps 20 --no-cpu-baseline --no-gpu-baseline > $O/bench_single_$i.json 2> $O/bench_single_$i.err
  timeout -s KILL 300 python bench.py --steps 20 --no-cpu-baseline --no-gpu-baseline > $O/bench_pair_$i.json 2> $O/bench_pair_$i.err
done
for f in $O/bench_*.json; do echo $f; python -c "
import json,sys
l=[x for x in open('$f').read().splitlines() if x.startswith('{')]
d=json.loads(l[-1]) if l else {}
print(d.get('value'), d.get('ms_per_step'), d.get('loss'), d.get('clocks',{}).get('sm_mhz'), (d.get('roofline') or {}).get('frac'), {k:v for k,v in (d.get('kernel_ms_per_step') or {}).items() if 'gemm' in k})
"; done
