"""Isolated timing of the residual-add + LayerNorm kernels on the step's shapes (CUDA events; buffers >> L2).
    python scripts/bench_ln.py          (WJ_LIB=<other build> for an A/B on the same box)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from wavjepa_b200 import ops  # noqa: E402

dev = "cuda"
torch.manual_seed(0)
for name, M, D in (("student", 19906, 768), ("predictor", 172953, 384), ("teacher", 102400, 768)):
    x, a = torch.randn(M, D, device=dev), torch.randn(M, D, device=dev).bfloat16()
    g, b = torch.ones(D, device=dev), torch.zeros(D, device=dev)
    of, ob = torch.empty(M, D, device=dev), torch.empty(M, D, device=dev, dtype=torch.bfloat16)
    st = torch.empty(M, 2, device=dev)
    dy, dyb = torch.randn(M, D, device=dev), torch.randn(M, D, device=dev).bfloat16()
    dg, db, cs = (torch.zeros(D, device=dev) for _ in range(3))
    fwd = lambda: ops.add_layernorm_fwd(x, a, g, b, 1e-6, of, ob, st, None)
    bwd = lambda: ops.add_layernorm_bwd(dy, dyb, x, a, st, g, of, ob, dg, db, cs)
    res = []
    for fn in (fwd, bwd):
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            fn()
        e1.record()
        torch.cuda.synchronize()
        res.append(e0.elapsed_time(e1) / 20)
    print(f"{name:10s} [{M} x {D}]: fwd {res[0]*1e3:7.1f} us ({M*D*14/res[0]/1e6:6.0f} GB/s)   bwd {res[1]*1e3:7.1f} us ({M*D*18/res[1]/1e6:6.0f} GB/s)")
