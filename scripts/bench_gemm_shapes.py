"""Isolated timing of the GEMM shapes that dominate the training step (CUDA events, L2-cold by rotating buffers).
    python scripts/bench_gemm_shapes.py [--ncu]     (--ncu: one launch per shape, for an ncu capture)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from wavjepa_b200 import ops  # noqa: E402

dev = "cuda"
ncu = "--ncu" in sys.argv
torch.manual_seed(0)


def run(name, M, N, K, *, act=0, f32=False, resid=False, out2=False, aux=False, bias=True, mode="fwd", reps=20):
    a = torch.randn(M, K, device=dev).bfloat16()
    w = (torch.randn(N, K, device=dev) * 0.05).bfloat16() if mode == "fwd" else (torch.randn(K, N, device=dev) * 0.05).bfloat16()
    outs = [torch.empty(M, N, device=dev, dtype=torch.float32 if f32 else torch.bfloat16) for _ in range(2)]
    b = torch.randn(N, device=dev) if bias else None
    r = torch.randn(M, N, device=dev) if resid else None
    o2 = torch.empty(M, N, device=dev, dtype=torch.bfloat16) if out2 else None
    ax = torch.randn(M, N, device=dev).bfloat16() if aux else None

    def call(i):
        if mode == "fwd":
            ops.gemm(ops.plain_operand(a), w, M, 1, outs[i % 2], bias=b, act=act, resid=r, out2=o2, aux=ax)
        else:
            ops.gemm_dgrad(ops.plain_operand(a), w, M, 1, outs[i % 2], K=K, N=N, act=act, aux=ax, resid=r)

    if ncu:
        call(0)
        torch.cuda.synchronize()
        return
    for i in range(3):
        call(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        call(i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{name:34s} M{M} N{N} K{K} act{act} f32{int(f32)} resid{int(resid)}: {ms:.3f} ms  {2.0 * M * N * K / ms / 1e9:.0f} TFLOP/s")


run("pred qkv plain", 172433, 1152, 384)
run("pred fc1 gelu+save", 172433, 1536, 384, act=1, out2=True)
run("pred fc1 gelu", 172433, 1536, 384, act=1)
run("pred fc1 plain", 172433, 1536, 384)
run("pred fc2 resid f32", 172433, 384, 1536, act=3, f32=True, resid=True)
run("pred outproj resid f32", 172433, 384, 384, act=3, f32=True, resid=True)
run("pred dgrad fc2 x aux", 172433, 1536, 384, act=2, aux=True, bias=False, mode="dgrad")
run("pred dgrad plain", 172433, 1536, 384, bias=False, mode="dgrad")
run("teacher fc1 gelu", 102400, 3072, 768, act=1)
run("teacher fc1 plain", 102400, 3072, 768)
run("teacher qkv plain", 102400, 2304, 768)
run("teacher outproj resid f32", 102400, 768, 768, act=3, f32=True, resid=True)
run("teacher fc2 resid f32", 102400, 768, 3072, act=3, f32=True, resid=True)
run("teacher fc2 plain bf16", 102400, 768, 3072)
