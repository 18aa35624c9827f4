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


def run(name, M, N, K, *, act=0, f32=False, resid=False, out2=False, aux=False, bias=True, mode="fwd", reps=20, bn=0, colsum=False):
    a = torch.randn(M, K, device=dev).bfloat16()
    w = (torch.randn(N, K, device=dev) * 0.05).bfloat16() if mode == "fwd" else (torch.randn(K, N, device=dev) * 0.05).bfloat16()
    outs = [torch.empty(M, N, device=dev, dtype=torch.float32 if f32 else torch.bfloat16) for _ in range(2)]
    b = torch.randn(N, device=dev) if bias else None
    r = torch.randn(M, N, device=dev) if resid else None
    o2 = torch.empty(M, N, device=dev, dtype=torch.bfloat16) if out2 else None
    ax = torch.randn(M, N, device=dev).bfloat16() if aux else None
    cs = torch.zeros(N, device=dev) if colsum else None

    def call(i):
        if mode == "fwd":
            ops.gemm(ops.plain_operand(a), w, M, 1, outs[i % 2], bias=b, act=act, resid=r, out2=o2, aux=ax, block_n=bn)
        else:
            ops.gemm_dgrad(ops.plain_operand(a), w, M, 1, outs[i % 2], K=K, N=N, act=act, aux=ax, resid=r, colsum=cs)

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


def run_conv_wgrad(B=512, L_in=6430, C=512, k=3, reps=5):
    L_out = (L_in - k) // 2 + 1
    x = torch.randn(B, L_in, C, device=dev).bfloat16()
    dy = torch.randn(B, L_out, C, device=dev).bfloat16()
    out = torch.empty(C, k * C, device=dev)

    def call():
        ops.gemm_wgrad(ops.make_operand(dy, C, L_out, B), ops.conv_operand(x, k), L_out, B, out)

    def call_plain():   # same FLOPs, B operand as a plain [B*L_out, k*C] matrix (no segments): isolates the addressing
        ops.gemm_wgrad(ops.make_operand(dy, C, L_out, B), ops.make_operand(x, 1024, L_out, B, row_stride=1024, batch_stride=L_in * C),
                       L_out, B, out[:, :1024], N=1024, ld_out=k * C)

    if ncu:
        call()
        torch.cuda.synchronize()
        return
    for nm, fn, ncols in (("conv1 wgrad (segments)", call, k * C), ("conv1 wgrad (plain 1024 cols)", call_plain, 1024)):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        print(f"{nm:34s} B{B} L_out{L_out}: {ms:.3f} ms  {2.0 * B * L_out * C * ncols / ms / 1e9:.0f} TFLOP/s")


if "--pair" in sys.argv:
    for bn in (0, -256):
        tag = "pair" if bn < 0 else "1cta"
        run(f"{tag} teacher fc1 plain", 102400, 3072, 768, bn=bn)
        run(f"{tag} teacher fc1 gelu", 102400, 3072, 768, act=1, bn=bn)
        run(f"{tag} teacher qkv plain", 102400, 2304, 768, bn=bn)
        run(f"{tag} teacher fc2 resid f32", 102400, 768, 3072, act=3, f32=True, resid=True, bn=bn)
        run(f"{tag} teacher outproj resid f32", 102400, 768, 768, act=3, f32=True, resid=True, bn=bn)
        run(f"{tag} pred fc1 plain", 172433, 1536, 384, bn=bn)
        run(f"{tag} pred fc1 gelu+save", 172433, 1536, 384, act=1, out2=True, bn=bn)
        run(f"{tag} conv-like M1.6M N512 K1536", 1645568, 512, 1536, act=1, bias=False, bn=bn, reps=5)
        run(f"{tag} conv-like gelu+save", 1645568, 512, 1536, act=1, out2=True, bias=False, bn=bn, reps=5)
        run(f"{tag} conv2-like gelu+save", 822272, 512, 1536, act=1, out2=True, bias=False, bn=bn, reps=5)
        run(f"{tag} teacher fc2 plain bf16", 102400, 768, 3072, bn=bn)
        run(f"{tag} teacher outproj plain bf16", 102400, 768, 768, bn=bn)
        run(f"{tag} student fc1 gelu+save", 19906, 3072, 768, act=1, out2=True, bn=bn)
        run(f"{tag} student fc2 plain", 19906, 768, 3072, bn=bn)
        run(f"{tag} student qkv plain", 19906, 2304, 768, bn=bn)
    for bn in (0, -128):
        tag = "pair128" if bn < 0 else "1cta"
        run(f"{tag} pred qkv plain", 172433, 1152, 384, bn=bn)
        run(f"{tag} pred fc2 resid f32", 172433, 384, 1536, act=3, f32=True, resid=True, bn=bn)
        run(f"{tag} pred outproj resid f32", 172433, 384, 384, act=3, f32=True, resid=True, bn=bn)
    sys.exit(0)
if "--grads" in sys.argv:
    # weight / data gradients that may run on CTA pairs: same process, routing switch on / off
    def timed(fn, reps=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    for name, M, Nw, Kw in (("student fc1 wgrad", 19906, 3072, 768), ("student fc2 wgrad", 19906, 768, 3072),
                            ("student qkv wgrad", 19906, 2304, 768), ("student out_proj wgrad", 19906, 768, 768),
                            ("teacher-size fc1 wgrad", 102400, 3072, 768)):
        dy, x = torch.randn(M, Nw, device=dev).bfloat16(), torch.randn(M, Kw, device=dev).bfloat16()
        out = torch.empty(Nw, Kw, device=dev)
        res = []
        for pair in (1, 0):
            ops.gemm_option("pair_wgrad", pair)
            res.append(timed(lambda: ops.gemm_wgrad(ops.plain_operand(dy), ops.plain_operand(x), M, 1, out)))
        ops.gemm_option("pair_wgrad", 1)
        fl = 2.0 * M * Nw * Kw / 1e9
        print(f"{name:28s} M{M} [{Nw} x {Kw}]: pairs {res[0]:.3f} ms {fl/res[0]:.0f} TFLOP/s | single {res[1]:.3f} ms {fl/res[1]:.0f} TFLOP/s")
    B, L_in, C, k = 512, 6430, 512, 3
    L_out = (L_in - k) // 2 + 1
    x, dy = torch.randn(B, L_in, C, device=dev).bfloat16(), torch.randn(B, L_out, C, device=dev).bfloat16()
    out = torch.empty(C, k * C, device=dev)
    res = []
    for pair in (1, 0):
        ops.gemm_option("pair_wgrad", pair)
        res.append(timed(lambda: ops.gemm_wgrad(ops.make_operand(dy, C, L_out, B), ops.conv_operand(x, k), L_out, B, out), reps=5))
    ops.gemm_option("pair_wgrad", 1)
    fl = 2.0 * B * L_out * C * k * C / 1e9
    print(f"{'conv1 wgrad':28s} B{B} L_out{L_out}: pairs {res[0]:.3f} ms {fl/res[0]:.0f} TFLOP/s | single {res[1]:.3f} ms {fl/res[1]:.0f} TFLOP/s")
    del x, dy
    for name, M, N, K, gelu in (("student fc2 dgrad", 19906, 3072, 768, False), ("student fc1 dgrad", 19906, 768, 3072, False),
                                ("student qkv dgrad", 19906, 768, 2304, False), ("student out_proj dgrad", 19906, 768, 768, False),
                                ("conv-like dgrad", 1646080, 512, 1024, False), ("conv-like dgrad K512", 1646080, 512, 512, False),
                                ("student fc2 dgrad x GELU'", 19906, 3072, 768, True), ("conv-like dgrad x GELU'", 822784, 512, 1024, True),
                                ("conv-like dgrad K512 x GELU'", 822784, 512, 512, True)):
        a, w = torch.randn(M, K, device=dev).bfloat16(), (torch.randn(K, N, device=dev) * 0.05).bfloat16()
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        aux = torch.randn(M, N, device=dev).bfloat16() if gelu else None
        cs = torch.zeros(N, device=dev) if gelu else None
        res = []
        for bn in (-256, 256):
            res.append(timed(lambda: ops.gemm_dgrad(ops.plain_operand(a), w, M, 1, out, K=K, N=N, block_n=bn,
                                                    act=ops.ACT_DGELU if gelu else 0, aux=aux, colsum=cs), reps=10))
        fl = 2.0 * M * N * K / 1e9
        print(f"{name:28s} M{M} N{N} K{K}: pairs {res[0]:.3f} ms {fl/res[0]:.0f} TFLOP/s | single {res[1]:.3f} ms {fl/res[1]:.0f} TFLOP/s")
    sys.exit(0)
if "--conv-wgrad" in sys.argv:
    run_conv_wgrad()
    sys.exit(0)
run("pred qkv plain", 172433, 1152, 384)
run("pred qkv plain bn192", 172433, 1152, 384, bn=192)
run("pred qkv plain bn128", 172433, 1152, 384, bn=128)
run("pred fc2 plain bf16 (bn192)", 172433, 384, 1536)
run("pred fc2 plain bf16 bn128", 172433, 384, 1536, bn=128)
run("pred outproj plain bf16 (bn192)", 172433, 384, 384)
run("pred outproj plain bf16 bn128", 172433, 384, 384, bn=128)
run("pred fc1 gelu+save", 172433, 1536, 384, act=1, out2=True)
run("pred fc1 gelu", 172433, 1536, 384, act=1)
run("pred fc1 plain", 172433, 1536, 384)
run("pred fc2 resid f32", 172433, 384, 1536, act=3, f32=True, resid=True)
run("pred outproj resid f32", 172433, 384, 384, act=3, f32=True, resid=True)
run("pred dgrad fc2 x aux", 172433, 1536, 384, act=2, aux=True, bias=False, mode="dgrad")
run("pred dgrad fc2 x aux + colsum", 172433, 1536, 384, act=2, aux=True, bias=False, mode="dgrad", colsum=True)
run("pred dgrad plain", 172433, 1536, 384, bias=False, mode="dgrad")
run("teacher fc1 gelu", 102400, 3072, 768, act=1)
run("teacher fc1 plain", 102400, 3072, 768)
run("teacher qkv plain", 102400, 2304, 768)
run("teacher outproj resid f32", 102400, 768, 768, act=3, f32=True, resid=True)
run("teacher fc2 resid f32", 102400, 768, 3072, act=3, f32=True, resid=True)
run("teacher fc2 plain bf16", 102400, 768, 3072)
