#!/bin/bash
set -u
O=gpurun_out/${1:-det}
mkdir -p $O
PREV=wavjepa_b200/libwavjepa_prev.so
timeout -s KILL 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "predictor_ctx_grad or conv0 or gather" > $O/pytest_k.log 2>&1; echo "rc=$?" >> $O/pytest_k.log; tail -3 $O/pytest_k.log
timeout -s KILL 300 python scripts/determinism_probe.py > $O/determinism_new.json 2> $O/determinism_new.err; cat $O/determinism_new.json; tail -3 $O/determinism_new.err
timeout -s KILL 900 python -m pytest tests/test_gpu_model.py -x -q -m gpu > $O/pytest_model.log 2>&1; echo "rc=$?" >> $O/pytest_model.log; tail -3 $O/pytest_model.log
timeout -s KILL 300 python bench.py --steps 20 --no-cpu-baseline --no-gpu-baseline > $O/bench_new.json 2> $O/bench_new.err
python -c "
import json
l=[x for x in open('$O/bench_new.json').read().splitlines() if x.startswith('{')]
d=json.loads(l[-1]); print(d['value'], d['ms_per_step'], d['loss'], d['clocks']['sm_mhz'], d['roofline']['frac'])"
