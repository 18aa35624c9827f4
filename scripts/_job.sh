set -u
O=gpurun_out/r2z
mkdir -p $O
timeout -s KILL 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_kernels.py -x -q -m gpu -k "hear or host_known or conv0 or hf or arch or interop" > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log; tail -3 $O/pytest.log
for i in 1 2; do timeout -s KILL 300 python bench.py --config hear --no-gpu-baseline $( [ $i = 2 ] && echo --no-cpu-baseline ) > $O/bench_hear_$i.json 2> $O/bench_hear_$i.err; done
timeout -s KILL 300 python bench.py --config hear --no-gpu-baseline --no-cpu-baseline --steps 30 > $O/bench_hear_3.json 2> $O/bench_hear_3.err
timeout -s KILL 300 python scripts/profile_step.py > $O/step_profile.txt 2>&1; grep -E "conv0|step \(unprofiled\)" $O/step_profile.txt
for f in $O/bench_hear_*.json; do python -c "
import json
l=[x for x in open('$f').read().splitlines() if x.startswith('{')]
d=json.loads(l[-1]) if l else {}
print('$f', d.get('value'), d.get('ms_per_step'), (d.get('e2e') or {}).get('value'), d.get('clocks',{}).get('sm_mhz'), (d.get('roofline') or {}).get('frac'), d.get('gpu_launches'))"; done
