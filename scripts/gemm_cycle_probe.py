"""Where does a short-K GEMM tile spend its cycles?  clock64 counters of the MMA thread, the TMA producer and one epilogue
warp (wj_gemm_debug) for the predictor shapes.   python scripts/gemm_cycle_probe.py > profiles/r02_gemm_cycle_counters.txt"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from wavjepa_b200 import _lib, ops  # noqa: E402

dev = "cuda"
lib = _lib.load()
cnt = torch.zeros(148 * 8, device=dev, dtype=torch.int64)


def probe(name, M, N, K, mode="fwd", **kw):
    a = torch.randn(M, K, device=dev).bfloat16()
    w = (torch.randn(N, K, device=dev) * 0.05).bfloat16() if mode == "fwd" else (torch.randn(K, N, device=dev) * 0.05).bfloat16()
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    extra = {}
    if kw.get("out2"):
        extra["out2"] = torch.empty_like(out)
    if kw.get("aux"):
        extra["aux"] = torch.randn(M, N, device=dev).bfloat16()
    if kw.get("colsum"):
        extra["colsum"] = torch.zeros(N, device=dev)
    act = kw.get("act", 0)

    def call():
        if mode == "fwd":
            ops.gemm(ops.plain_operand(a), w, M, 1, out, act=act, block_n=kw.get("bn", 0), **extra)
        else:
            ops.gemm_dgrad(ops.plain_operand(dy := a), w, M, 1, out, K=K, N=N, act=act, **extra)

    for _ in range(3):
        call()
    torch.cuda.synchronize()
    cnt.zero_()
    lib.wj_gemm_debug(C.c_void_p(cnt.data_ptr()))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    call()
    e1.record()
    torch.cuda.synchronize()
    lib.wj_gemm_debug(None)
    c = cnt.view(148, 8).double()
    tiles = c[:, 6].clamp(min=1)
    per = lambda i: (c[:, i] / tiles).mean().item()
    print(f"{name:34s} {e0.elapsed_time(e1)*1e3:7.1f} us | per tile (cycles): MMA thread total {per(2):7.0f} = wait operands {per(0):6.0f} + wait free accumulator "
          f"{per(1):6.0f} + issue/other {per(2)-per(0)-per(1):6.0f} | producer waits for a free stage {per(3):6.0f} | epilogue warp: total {per(5):7.0f}, "
          f"of which waiting for the accumulator {per(4):6.0f} | tiles/CTA {tiles.mean().item():.1f}")


probe("pred fc1 plain K384 N1536", 172433, 1536, 384)
probe("pred fc1 gelu K384", 172433, 1536, 384, act=1)
probe("pred fc1 gelu+save K384", 172433, 1536, 384, act=1, out2=True)
probe("pred qkv plain K384 N1152", 172433, 1152, 384)
probe("pred fc2 plain K1536 N384 (bn192)", 172433, 384, 1536)
probe("pred outproj K384 N384 (bn192)", 172433, 384, 384)
probe("pred dgrad plain K384 N1536", 172433, 1536, 384, mode="dgrad")
probe("pred dgrad x aux K384 N1536", 172433, 1536, 384, mode="dgrad", act=2, aux=True)
probe("pred dgrad x aux + colsum", 172433, 1536, 384, mode="dgrad", act=2, aux=True, colsum=True)
probe("teacher fc1 1cta K768 N3072", 102400, 3072, 768, bn=256)


def probe_wgrad(name, M, Nw, Kw):
    dy, x = torch.randn(M, Nw, device=dev).bfloat16(), torch.randn(M, Kw, device=dev).bfloat16()
    out = torch.zeros(Nw, Kw, device=dev)
    call = lambda: ops.gemm_wgrad(ops.plain_operand(dy), ops.plain_operand(x), M, 1, out, accumulate=True)
    for _ in range(3):
        call()
    torch.cuda.synchronize()
    cnt.zero_()
    lib.wj_gemm_debug(C.c_void_p(cnt.data_ptr()))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    call()
    e1.record()
    torch.cuda.synchronize()
    lib.wj_gemm_debug(None)
    c = cnt.view(148, 8).double()
    tiles = c[:, 6].clamp(min=1)
    tot, wf, we = c[:, 2].mean().item(), c[:, 0].mean().item(), c[:, 1].mean().item()
    print(f"{name:34s} {e0.elapsed_time(e1)*1e3:7.1f} us ({2.0*M*Nw*Kw/e0.elapsed_time(e1)/1e9:6.0f} TFLOP/s) | MMA thread per CTA (cycles): total {tot:9.0f} = "
          f"wait operands {wf:9.0f} ({100*wf/tot:4.1f} %) + wait free accumulator {we:7.0f} + issue/other {tot-wf-we:9.0f} | tiles/CTA {tiles.mean().item():.1f}")


probe_wgrad("wgrad pred [1536 x 384]", 172433, 1536, 384)
probe_wgrad("wgrad pred [384 x 1536]", 172433, 384, 1536)
probe_wgrad("wgrad pred [1152 x 384]", 172433, 1152, 384)
probe_wgrad("wgrad pred [384 x 384]", 172433, 384, 384)
probe_wgrad("wgrad student [3072 x 768]", 19906, 3072, 768)
probe_wgrad("wgrad student [768 x 768]", 19906, 768, 768)
probe("student dgrad plain K3072 N768", 19906, 768, 3072, mode="dgrad")
probe("student fc2 fwd 1cta K3072 N768", 19906, 768, 3072, bn=256)
probe("pred fc2 dgrad K1536 N384", 172433, 384, 1536, mode="dgrad")
