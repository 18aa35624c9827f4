#!/bin/bash
set -u
O=gpurun_out/${1:-det2}
mkdir -p $O
timeout -s KILL 600 python -m pytest tests/test_gpu_model.py -x -q -m gpu -k "deterministic or run_to_run" > $O/pytest_det.log 2>&1; echo "rc=$?" >> $O/pytest_det.log; tail -4 $O/pytest_det.log
if ! grep -q "rc=0" $O/pytest_det.log; then grep -E "^E |FAILED|Error" $O/pytest_det.log | head -20; fi
timeout -s KILL 300 python scripts/determinism_probe.py 16 --det > $O/determinism_det.json 2> $O/determinism_det.err; cat $O/determinism_det.json; tail -3 $O/determinism_det.err
timeout -s KILL 900 python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log; tail -3 $O/pytest_gpu.log
timeout -s KILL 300 python bench.py --steps 20 --no-cpu-baseline --no-gpu-baseline > $O/bench_default.json 2> $O/bench_default.err
timeout -s KILL 300 python bench.py --steps 20 --no-cpu-baseline --no-gpu-baseline --deterministic > $O/bench_det.json 2> $O/bench_det.err
for f in $O/bench_default.json $O/bench_det.json; do python -c "
import json
l=[x for x in open('$f').read().splitlines() if x.startswith('{')]
d=json.loads(l[-1]) if l else {}
print('$f', d.get('value'), d.get('ms_per_step'), d.get('loss'), d.get('clocks',{}).get('sm_mhz'), (d.get('roofline') or {}).get('frac'), d.get('gpu_launches'))"; done
tail -3 $O/bench_det.err
