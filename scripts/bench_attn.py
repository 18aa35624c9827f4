"""Isolated timing of the attention shapes of the training step (CUDA events, L2-cold by size).
    PYTHONPATH=. python scripts/bench_attn.py [--ncu]     (--ncu: one launch per shape, for an ncu capture)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from wavjepa_b200 import ops  # noqa: E402

ncu = "--ncu" in sys.argv
torch.manual_seed(0)
dev = "cuda"
for name, S, D, H, lo, hi in (("student (visible context)", 512, 768, 12, 20, 73), ("predictor (context + targets)", 2048, 384, 12, 45, 123),
                              ("teacher (full sequences)", 512, 768, 12, 200, 201),
                              # binaural (configs[4]): 400 tokens per instance
                              ("Nat teacher (full 400-token sequences)", 512, 768, 12, 400, 401),
                              ("Nat predictor (context + targets)", 2048, 384, 12, 90, 246),
                              ("Nat student (visible context)", 512, 768, 12, 40, 146)):
    lens = torch.randint(lo, hi, (S,))
    cu = torch.zeros(S + 1, dtype=torch.int32)
    cu[1:] = lens.cumsum(0)
    tot, ml = int(cu[-1]), int(lens.max())
    cu = cu.to(dev)
    qkv = torch.randn(tot, 3 * D, device=dev).bfloat16()
    out = torch.empty(tot, D, device=dev, dtype=torch.bfloat16)
    lse = torch.empty(tot, H, device=dev)
    do = torch.randn(tot, D, device=dev).bfloat16()
    dqkv = torch.empty_like(qkv)
    fwd = lambda: ops.attn_fwd(qkv, cu, S, ml, D, H, out, lse)
    bwd = lambda: ops.attn_bwd(qkv, out, do, lse, cu, S, ml, D, H, dqkv)
    dbias = torch.zeros(3 * D, device=dev)
    bwd_b = lambda: ops.attn_bwd(qkv, out, do, lse, cu, S, ml, D, H, dqkv, dbias=dbias) if hasattr(ops, "add_bf16") else bwd()
    fwd()
    if ncu:
        bwd()
        torch.cuda.synchronize()
        continue
    res = []
    for fn in (fwd, bwd, bwd_b):
        if fn is not fwd and name.startswith("Nat teacher"):   # the teacher has no backward (and 400 x dh 64 exceeds it)
            res.append(float("nan"))
            continue
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            fn()
        e1.record()
        torch.cuda.synchronize()
        res.append(e0.elapsed_time(e1) / 20)
    dh = D // H
    flops = 4.0 * float((lens.double() ** 2).sum()) * dh * H     # QK^T + PV
    print(f"{name:32s} seqs {S} tokens {tot} max {ml} dh {dh}: fwd {res[0]*1e3:7.1f} us ({flops/res[0]/1e9:6.1f} TFLOP/s)   "
          f"bwd {res[1]*1e3:7.1f} us ({2.5*flops/res[1]/1e9:6.1f} TFLOP/s)   bwd + in_proj bias gradient {res[2]*1e3:7.1f} us")
