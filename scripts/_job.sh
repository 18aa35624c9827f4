set -u
O=gpurun_out/r2x
mkdir -p $O
PREV=wavjepa_b200/libwavjepa_prev.so
timeout -s KILL 300 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "layernorm" > $O/pytest_k.log 2>&1; echo "rc=$?" >> $O/pytest_k.log; tail -2 $O/pytest_k.log
for i in 1 2; do
timeout -s KILL 100 python scripts/bench_ln.py > $O/ln_new.txt 2>&1; cat $O/ln_new.txt
WJ_LIB=$PREV timeout -s KILL 100 python scripts/bench_ln.py > $O/ln_prev.txt 2>&1; cat $O/ln_prev.txt
done
