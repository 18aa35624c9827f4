#!/bin/bash
# Evidence refresh on one B200 (run through gpurun): GPU tests, the three bench configs, the reference arm, the
# per-entry step profile, the ncu launch list with DRAM traffic, and ncu --set full captures of the HBM-bound kernels.
# usage: bash scripts/gpu_refresh.sh <out-subdir>
set -u
O=gpurun_out/${1:-refresh}
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log
timeout 600 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err
timeout 300 python bench.py --config hear --no-gpu-baseline > $O/bench_hear.json 2> $O/bench_hear.err
timeout 300 python bench.py --config nat --no-cpu-baseline --no-gpu-baseline > $O/bench_nat.json 2> $O/bench_nat.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err
timeout 300 python scripts/profile_step.py > $O/step_profile.txt 2>&1
timeout 300 python scripts/bench_denoiser.py > $O/bench_denoiser.json 2> $O/bench_denoiser.err
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file $O/traffic.csv python scripts/profile_step.py --ncu > $O/ncu_traffic.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv0_bwd_kernel -c 1 \
    -o $O/conv0_bwd python scripts/profile_step.py --ncu > $O/ncu_conv0_bwd.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:layernorm_bwd_kernel -c 4 \
    -o $O/ln_bwd python scripts/profile_step.py --ncu > $O/ncu_ln_bwd.log 2>&1
timeout 300 python scripts/determinism_probe.py > $O/determinism.json 2> $O/determinism.err
timeout 300 python scripts/determinism_probe.py 16 --det > $O/determinism_det.json 2> $O/determinism_det.err
timeout 300 python bench.py --deterministic --no-cpu-baseline --no-gpu-baseline > $O/bench_det.json 2> $O/bench_det.err
timeout 900 compute-sanitizer --tool memcheck python scripts/sanitize_small.py > $O/sanitizer_memcheck.log 2>&1
tail -3 $O/sanitizer_memcheck.log
tail -3 $O/pytest_gpu.log; tail -2 $O/smoke.log; cat $O/bench_n1.json | cut -c1-400
