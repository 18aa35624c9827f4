"""Developer check of the tcgen05 GEMM core on a B200 (run under gpurun)."""
import sys, time, traceback
sys.path.insert(0, ".")
import torch
import torch.nn.functional as F
from wavjepa_b200 import ops, _lib

torch.manual_seed(0)
dev = "cuda"
_lib.require_device()

def rel(a, b):
    a = a.float(); b = b.float()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()

def run(name, fn):
    try:
        t0 = time.time()
        r = fn()
        torch.cuda.synchronize()
        print(f"[{name}] {r}  ({time.time()-t0:.2f}s)", flush=True)
    except Exception as e:
        print(f"[{name}] EXC {type(e).__name__}: {e}", flush=True)
        traceback.print_exc()

def t_plain(M, N, K, bn=0, out_dtype=torch.bfloat16):
    a = torch.randn(M, K, device=dev).bfloat16()
    w = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
    out = torch.full((M, N), float("nan"), device=dev, dtype=out_dtype)
    ops.gemm(ops.plain_operand(a), w, M, 1, out, block_n=bn)
    ref = a.float() @ w.float().t()
    return f"M{M} N{N} K{K} bn{bn} rel={rel(out, ref):.3e} nan={torch.isnan(out.float()).sum().item()}"

def t_epi():
    M, N, K = 1000, 1536, 384
    a = torch.randn(M, K, device=dev).bfloat16()
    w = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
    bias = torch.randn(N, device=dev)
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    out2 = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    ops.gemm(ops.plain_operand(a), w, M, 1, out, bias=bias, act=ops.ACT_GELU, out2=out2)
    h = (a.float() @ w.float().t() + bias).bfloat16()
    g = F.gelu(h.float())
    r1 = rel(out2, h); r2 = rel(out, g)
    # residual fp32 + fp32 out
    res = torch.randn(M, N, device=dev)
    o3 = torch.empty(M, N, device=dev)
    ops.gemm(ops.plain_operand(a), w, M, 1, o3, bias=bias, resid=res)
    r3 = rel(o3, a.float() @ w.float().t() + bias + res)
    # dgelu
    aux = torch.randn(M, N, device=dev).bfloat16()
    o4 = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    ops.gemm(ops.plain_operand(a), w, M, 1, o4, act=ops.ACT_DGELU, aux=aux)
    x = aux.float().requires_grad_(True)
    F.gelu(x).sum().backward()
    r4 = rel(o4, (a.float() @ w.float().t()) * x.grad)
    # pos table (resid_mod) and row scatter
    pos = torch.randn(200, N, device=dev)
    o5 = torch.empty(M, N, device=dev)
    ops.gemm(ops.plain_operand(a), w, M, 1, o5, resid=pos, resid_mod=200)
    idx = torch.arange(M, device=dev) % 200
    r5 = rel(o5, a.float() @ w.float().t() + pos[idx])
    perm = torch.randperm(M, device=dev).int()
    o6 = torch.zeros(M, N, device=dev)
    ops.gemm(ops.plain_operand(a), w, M, 1, o6, out_rows=perm)
    ref6 = torch.zeros(M, N, device=dev); ref6[perm.long()] = a.float() @ w.float().t()
    r6 = rel(o6, ref6)
    return f"gelu_pre={r1:.2e} gelu={r2:.2e} resid={r3:.2e} dgelu={r4:.2e} posmod={r5:.2e} scatter={r6:.2e}"

def conv_ref(x_nlc, w, stride):
    # x [B, L, C] ; w [O, C, k]
    y = F.conv1d(x_nlc.float().transpose(1, 2), w.float(), stride=stride)
    return y.transpose(1, 2).contiguous()

def t_conv_seg(Bn=3, L_in=402, Cc=512, k=3):
    x = torch.randn(Bn, L_in, Cc, device=dev).bfloat16()
    w = (torch.randn(Cc, Cc, k, device=dev) * 0.03).bfloat16()
    L_out = (L_in - k) // 2 + 1
    wk = w.permute(0, 2, 1).reshape(Cc, k * Cc).contiguous()  # [O, (j, c)]
    a = ops.make_operand(x, Cc, L_in // 2, Bn, nq=2, q_stride=Cc, row_stride=2 * Cc, batch_stride=L_in * Cc,
                         seg_width=Cc, seg_q=(0, 1, 0)[:k], seg_p=(0, 0, 1)[:k])
    out = torch.full((Bn, L_out, Cc), float("nan"), device=dev, dtype=torch.bfloat16)
    ops.gemm(a, wk, L_out, Bn, out.view(-1, Cc))
    ref = conv_ref(x, w, 2)
    return f"conv seg k{k} L_in{L_in} rel={rel(out, ref):.3e} nan={torch.isnan(out.float()).sum().item()}"

def t_conv_overlap(Bn=3, L_in=402, Cc=512, k=3):
    x = torch.randn(Bn, L_in, Cc, device=dev).bfloat16()
    w = (torch.randn(Cc, Cc, k, device=dev) * 0.03).bfloat16()
    L_out = (L_in - k) // 2 + 1
    wk = w.permute(0, 2, 1).reshape(Cc, k * Cc).contiguous()
    a = ops.make_operand(x, k * Cc, L_out, Bn, row_stride=2 * Cc, batch_stride=L_in * Cc)
    out = torch.full((Bn, L_out, Cc), float("nan"), device=dev, dtype=torch.bfloat16)
    ops.gemm(a, wk, L_out, Bn, out.view(-1, Cc))
    ref = conv_ref(x, w, 2)
    return f"conv overlap k{k} rel={rel(out, ref):.3e} nan={torch.isnan(out.float()).sum().item()}"

def t_wgrad(M, Nw, Kw, splits=0):
    # dW[Nw_out(m), Kw(n)] = dY[M, Nw]^T X[M, Kw]
    dy = torch.randn(M, Nw, device=dev).bfloat16()
    x = torch.randn(M, Kw, device=dev).bfloat16()
    out = torch.full((Nw, Kw), float("nan"), device=dev)
    ops.gemm_wgrad(ops.plain_operand(dy), ops.plain_operand(x), M, 1, out, splits=splits)
    ref = dy.float().t() @ x.float()
    return f"wgrad M{M} {Nw}x{Kw} splits{splits} rel={rel(out, ref):.3e} nan={torch.isnan(out).sum().item()}"

def t_wgrad_conv(Bn=3, L_in=402, Cc=512, k=3):
    x = torch.randn(Bn, L_in, Cc, device=dev).bfloat16()
    L_out = (L_in - k) // 2 + 1
    dy = torch.randn(Bn, L_out, Cc, device=dev).bfloat16()
    xo = ops.make_operand(x, Cc, L_in // 2, Bn, nq=2, q_stride=Cc, row_stride=2 * Cc, batch_stride=L_in * Cc,
                          seg_width=Cc, seg_q=(0, 1, 0)[:k], seg_p=(0, 0, 1)[:k])
    dyo = ops.make_operand(dy, Cc, L_out, Bn)
    out = torch.zeros(Cc, k * Cc, device=dev)
    ops.gemm_wgrad(dyo, xo, L_out, Bn, out)
    xf = x.float().transpose(1, 2).requires_grad_(False)
    w = torch.zeros(Cc, Cc, k, device=dev, requires_grad=True)
    y = F.conv1d(xf, w, stride=2)
    y.backward(dy.float().transpose(1, 2))
    ref = w.grad.permute(0, 2, 1).reshape(Cc, k * Cc)
    return f"wgrad conv k{k} rel={rel(out, ref):.3e}"

def t_time(M, N, K, bn=0, iters=20):
    a = torch.randn(M, K, device=dev).bfloat16()
    w = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    op = ops.plain_operand(a)
    for _ in range(3):
        ops.gemm(op, w, M, 1, out, block_n=bn)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        ops.gemm(op, w, M, 1, out, block_n=bn)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    e0.record()
    for _ in range(iters):
        torch.matmul(a, w.t(), out=out)
    e1.record(); torch.cuda.synchronize()
    ms2 = e0.elapsed_time(e1) / iters
    fl = 2.0 * M * N * K
    return f"M{M} N{N} K{K} bn{bn}: ours {ms:.3f} ms = {fl/ms/1e9:.1f} TF/s ; cublas {ms2:.3f} ms = {fl/ms2/1e9:.1f} TF/s"

run("plain-small", lambda: t_plain(128, 256, 64, 256))
run("plain-small128", lambda: t_plain(128, 128, 64, 128))
run("plain-k", lambda: t_plain(128, 256, 512, 256))
run("plain-tail", lambda: t_plain(300, 768, 512))
run("plain-tail-n", lambda: t_plain(300, 384, 768, 256))
run("plain-big", lambda: t_plain(20000, 2304, 768, 256))
run("plain-big128", lambda: t_plain(20000, 1152, 384, 128))
run("plain-f32", lambda: t_plain(5000, 768, 3072, 0, torch.float32))
run("epilogues", t_epi)
run("conv-seg3", lambda: t_conv_seg(k=3))
run("conv-seg2", lambda: t_conv_seg(L_in=400, k=2))
run("conv-seg3-big", lambda: t_conv_seg(Bn=4, L_in=6430, k=3))
run("conv-overlap", t_conv_overlap)
run("wgrad", lambda: t_wgrad(1000, 256, 384))
run("wgrad1", lambda: t_wgrad(64, 128, 128, 1))
run("wgrad-big", lambda: t_wgrad(50000, 768, 3072))
run("wgrad-conv", t_wgrad_conv)
run("time1", lambda: t_time(102400, 3072, 768, 256))
run("time2", lambda: t_time(102400, 768, 3072, 256))
run("time3", lambda: t_time(173000, 1536, 384, 256))
run("time4", lambda: t_time(173000, 1536, 384, 128))
run("time5", lambda: t_time(1645568, 512, 1536, 256, iters=5))
