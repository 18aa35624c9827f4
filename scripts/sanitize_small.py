"""One launch of every hand-written kernel family at small shapes, for compute-sanitizer:

    compute-sanitizer --tool memcheck  python scripts/sanitize_small.py > profiles/r02_sanitizer_memcheck.log
    compute-sanitizer --tool racecheck python scripts/sanitize_small.py > profiles/r02_sanitizer_racecheck.log

Covers: the tcgen05 GEMM in its three modes (fwd / dgrad / wgrad) on every epilogue path (plain TMA store, GELU,
GELU + saved GELU', 192-column tiles, GELU-backward with column sums, generic fp32 + residual, row scatter, transposed
wgrad, split-K), the CTA-pair GEMM in all three modes (incl. the GELU-backward factor pipeline), the implicit-GEMM conv
operands, both attention kernels (tcgen05 forward at head dim 64 / 32 with 128, 256 and 512 keys, tcgen05 backward with
fused column sums, mma.sync forward / backward), conv0 fwd / bwd, LayerNorm fwd / bwd with the fused residual add, the
elementwise / reduction kernels (incl. the context-gradient gather), the mask kernels, and a second pass over the
reductions in deterministic mode (workspace + fixed-order second kernels)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from wavjepa_b200 import ops  # noqa: E402

dev = "cuda"
torch.manual_seed(0)
bf = torch.bfloat16


def r(*shape, dtype=torch.float32, scale=1.0):
    return (torch.randn(*shape, device=dev) * scale).to(dtype)


# ---------------------------------------------------------------- GEMM forward paths
for (M, N, K, bn) in ((300, 256, 128, 0), (300, 384, 128, 0), (300, 1152, 384, 192), (300, 512, 768, 0), (600, 256, 768, -256), (300, 128, 64, 128)):
    a, w = r(M, K, dtype=bf), r(N, K, dtype=bf, scale=0.05)
    out = torch.empty(M, N, device=dev, dtype=bf)
    ops.gemm(ops.plain_operand(a), w, M, 1, out, bias=r(N), block_n=bn)
    ops.gemm(ops.plain_operand(a), w, M, 1, out, bias=r(N), act=ops.ACT_GELU, block_n=bn)
    if N % 256 == 0:
        ops.gemm(ops.plain_operand(a), w, M, 1, out, bias=r(N), act=ops.ACT_GELU, out2=torch.empty_like(out), block_n=bn)
    o32 = torch.empty(M, N, device=dev)
    ops.gemm(ops.plain_operand(a), w, M, 1, o32, bias=r(N), resid=r(M, N), act=ops.ACT_BF16, block_n=bn if bn < 0 else 0)
    ops.gemm(ops.plain_operand(a), w, M, 1, out, colsum=torch.zeros(N, device=dev), block_n=bn if bn < 0 else 0)
# ---------------------------------------------------------------- dgrad / wgrad
for (M, N, K) in ((300, 384, 256), (300, 768, 384), (300, 1536, 384)):
    dy, w = r(M, K, dtype=bf), r(K, N, dtype=bf, scale=0.05)
    ob = torch.empty(M, N, device=dev, dtype=bf)
    ops.gemm_dgrad(ops.plain_operand(dy), w, M, 1, ob, K=K, N=N)
    if N % 256 == 0:
        ops.gemm_dgrad(ops.plain_operand(dy), w, M, 1, ob, K=K, N=N, act=ops.ACT_DGELU, aux=r(M, N, dtype=bf), colsum=torch.zeros(N, device=dev))
    of = torch.zeros(M, N, device=dev)
    ops.gemm_dgrad(ops.plain_operand(dy), w, M, 1, of, K=K, N=N, resid=r(M, N))
    rows = torch.randperm(M, device=dev).int()
    ops.gemm_dgrad(ops.plain_operand(dy), w, M, 1, of, K=K, N=N, out_rows=rows)
for (M, Nw, Kw, sp) in ((1000, 256, 384, 0), (1000, 1536, 384, 0), (1000, 384, 384, 3), (700, 128, 128, 1)):
    dy, x = r(M, Nw, dtype=bf), r(M, Kw, dtype=bf)
    ops.gemm_wgrad(ops.plain_operand(dy), ops.plain_operand(x), M, 1, torch.zeros(Nw, Kw, device=dev), accumulate=True, splits=sp)
# CTA pairs: whole 256 x 256 weight tiles (wgrad), plain and GELU-backward data gradients
for (M, Nw, Kw, sp) in ((1000, 256, 256, 0), (700, 512, 768, 3)):
    dy, x = r(M, Nw, dtype=bf), r(M, Kw, dtype=bf)
    ops.gemm_wgrad(ops.plain_operand(dy), ops.plain_operand(x), M, 1, torch.zeros(Nw, Kw, device=dev), accumulate=True, splits=sp)
for (M, N, K) in ((600, 256, 768), (300, 512, 128)):
    dy, w = r(M, K, dtype=bf), r(K, N, dtype=bf, scale=0.05)
    ob = torch.empty(M, N, device=dev, dtype=bf)
    ops.gemm_dgrad(ops.plain_operand(dy), w, M, 1, ob, K=K, N=N, block_n=-256)
    ops.gemm_dgrad(ops.plain_operand(dy), w, M, 1, ob, K=K, N=N, act=ops.ACT_DGELU, aux=r(M, N, dtype=bf),
                   colsum=torch.zeros(N, device=dev), block_n=-256)
# ---------------------------------------------------------------- implicit-GEMM conv (k = 3 and 2), fwd / wgrad / dgrad
for k, L_in in ((3, 66), (2, 40)):
    Bn, Cc = 3, 512
    x = r(Bn, L_in, Cc, dtype=bf)
    L_out = (L_in - k) // 2 + 1
    wk = r(Cc, k * Cc, dtype=bf, scale=0.02)
    g = torch.empty(Bn, L_out, Cc, device=dev, dtype=bf)
    h = torch.empty_like(g)
    ops.gemm(ops.conv_operand(x, k), wk, L_out, Bn, g.view(-1, Cc), act=ops.ACT_GELU, out2=h.view(-1, Cc))
    ops.gemm_wgrad(ops.make_operand(g, Cc, L_out, Bn), ops.conv_operand(x, k), L_out, Bn, torch.empty(Cc, k * Cc, device=dev))
    dx = torch.empty(Bn, L_in, Cc, device=dev, dtype=bf)
    ops.conv_dgrad(g, wk, dx, k, act=ops.ACT_DGELU, aux=r(Bn, L_in, Cc, dtype=bf))
# ---------------------------------------------------------------- attention
for (D, H, lens) in ((768, 12, [39, 72, 1]), (384, 12, [85, 122, 128]), (768, 12, [200, 130]), (384, 12, [250, 3]), (384, 12, [300]),
                     (768, 12, [384, 257]), (384, 12, [512, 400, 9]), (384, 12, [600])):
    cu_l = [0]
    for n in lens:
        cu_l.append(cu_l[-1] + n)
    tot = cu_l[-1]
    cu = torch.tensor(cu_l, device=dev, dtype=torch.int32)
    qkv = r(tot, 3 * D, dtype=bf)
    out = torch.empty(tot, D, device=dev, dtype=bf)
    lse = torch.empty(tot, H, device=dev)
    ops.attn_fwd(qkv, cu, len(lens), max(lens), D, H, out, lse)
    dq = torch.empty(tot, 3 * D, device=dev, dtype=bf)
    ops.attn_bwd(qkv, out, r(tot, D, dtype=bf), lse, cu, len(lens), max(lens), D, H, dq, dbias=torch.zeros(3 * D, device=dev))
# ---------------------------------------------------------------- conv0, norms, elementwise, masks
for Cin, L in ((1, 2577), (2, 1300)):
    B, C = 2, 512
    x = r(B, Cin, L, dtype=bf)
    w0, ga, be = r(C, Cin, 10, scale=0.3), 1 + 0.1 * r(C), 0.1 * r(C)
    L_out = (L - 10) // 5 + 1
    out = torch.empty(B, L_out, C, device=dev, dtype=bf)
    mom, stats, red = ops.conv0_workspaces(B, Cin, C, dev, backward=True)
    ops.conv0_fwd(x, w0, ga, be, out, mom, stats)
    ops.conv0_bwd(x, w0, ga, be, mom, stats, r(B, L_out, C, dtype=bf), red, torch.zeros_like(w0), torch.zeros(C, device=dev), torch.zeros(C, device=dev))
for D in (384, 768):
    M = 333
    x, a = r(M, D), r(M, D, dtype=bf)
    g, b = 1 + 0.1 * r(D), 0.1 * r(D)
    of, ob, st, rs = torch.empty(M, D, device=dev), torch.empty(M, D, device=dev, dtype=bf), torch.empty(M, 2, device=dev), torch.empty(M, 2, device=dev)
    ops.add_layernorm_fwd(x, a, g, b, 1e-6, of, ob, st, rs)
    ops.add_layernorm_bwd(r(M, D), r(M, D, dtype=bf), x, a, st, g, of, ob, torch.zeros(D, device=dev), torch.zeros(D, device=dev), torch.zeros(D, device=dev))
n = 100_003
p_, g_, m_, v_ = r(n), r(n, scale=0.1), torch.zeros(n, device=dev), torch.zeros(n, device=dev)
ss = torch.zeros(1, device=dev, dtype=torch.float64)
ops.sumsq(g_, 1.0, ss)
ops.adamw_ema_step(p_, g_, m_, v_, 4e-4, 0.9, 0.98, 1e-6, 0.04, 1, 1.0, 5.0, ss, torch.empty(n, device=dev, dtype=bf), r(40000),
                   torch.empty(40000, device=dev, dtype=bf), 4096, 44096, 0.999)
ops.colsum(r(777, 1152, dtype=bf), torch.zeros(1152, device=dev))
ctx, tgt, vis, att, err = ops.masks_generate(0, 64, 200, 1, False, 4, 0.65, 10, 0.25, 10, 0.1, 0, 1234, 7, dev)
mi = ops.mask_indices(ctx, tgt, vis)
audio = r(2, 1, 50000)
ops.crop_norm(audio, torch.tensor([0, 100, 5, 7], device=dev, dtype=torch.int32), 2, 32159, torch.empty(4, 1, 32159, device=dev, dtype=bf), None)
Nc, G = mi.Nc, mi.G
dx0 = r(mi.Nv, 384)
ops.predictor_assemble_bwd(dx0, mi.vis_src, mi.Nv, 384, None, torch.zeros(384, device=dev))
ops.predictor_ctx_grad(dx0, mi.vis_src, mi.cu_v, mi.B * G, G, Nc, 384, torch.empty(Nc, 384, device=dev), torch.empty(Nc, 384, device=dev, dtype=bf))
# ---------------------------------------------------------------- the reductions once more in deterministic mode
ops.set_deterministic(True, 64 << 20)
dy, x = r(1000, 512, dtype=bf), r(1000, 768, dtype=bf)
ops.gemm_wgrad(ops.plain_operand(dy), ops.plain_operand(x), 1000, 1, torch.zeros(512, 768, device=dev), accumulate=True)
ops.gemm_wgrad(ops.plain_operand(dy), ops.plain_operand(r(1000, 384, dtype=bf)), 1000, 1, torch.zeros(512, 384, device=dev))
dy, w = r(300, 384, dtype=bf), r(384, 1536, dtype=bf, scale=0.05)
ops.gemm_dgrad(ops.plain_operand(dy), w, 300, 1, torch.empty(300, 1536, device=dev, dtype=bf), K=384, N=1536, act=ops.ACT_DGELU,
               aux=r(300, 1536, dtype=bf), colsum=torch.zeros(1536, device=dev))
M, D = 333, 768
x, a = r(M, D), r(M, D, dtype=bf)
g, st = 1 + 0.1 * r(D), torch.empty(M, 2, device=dev)
ops.add_layernorm_fwd(x, a, g, 0.1 * r(D), 1e-6, torch.empty(M, D, device=dev), torch.empty(M, D, device=dev, dtype=bf), st, torch.empty(M, 2, device=dev))
ops.add_layernorm_bwd(r(M, D), r(M, D, dtype=bf), x, a, st, g, torch.empty(M, D, device=dev), torch.empty(M, D, device=dev, dtype=bf),
                      torch.zeros(D, device=dev), torch.zeros(D, device=dev), torch.zeros(D, device=dev))
ops.colsum(r(777, 1152, dtype=bf), torch.zeros(1152, device=dev))
ops.sumsq(g_, 1.0, ss)
ops.predictor_assemble_bwd(dx0, mi.vis_src, mi.Nv, 384, None, torch.zeros(384, device=dev))
cu = torch.tensor([0, 85, 207], device=dev, dtype=torch.int32)
qkv = r(207, 3 * 384, dtype=bf)
out, lse = torch.empty(207, 384, device=dev, dtype=bf), torch.empty(207, 12, device=dev)
ops.attn_fwd(qkv, cu, 2, 122, 384, 12, out, lse)
ops.attn_bwd(qkv, out, r(207, 384, dtype=bf), lse, cu, 2, 122, 384, 12, torch.empty(207, 3 * 384, device=dev, dtype=bf),
             dbias=torch.zeros(3 * 384, device=dev))
B, C, L = 2, 512, 2577
x = r(B, 1, L, dtype=bf)
w0, ga, be = r(C, 1, 10, scale=0.3), 1 + 0.1 * r(C), 0.1 * r(C)
L_out = (L - 10) // 5 + 1
mom, stats, red = ops.conv0_workspaces(B, 1, C, dev, backward=True)
ops.conv0_fwd(x, w0, ga, be, torch.empty(B, L_out, C, device=dev, dtype=bf), mom, stats)
ops.conv0_bwd(x, w0, ga, be, mom, stats, r(B, L_out, C, dtype=bf), red, torch.zeros_like(w0), torch.zeros(C, device=dev), torch.zeros(C, device=dev))
ops.set_deterministic(False)
torch.cuda.synchronize()
print("sanitize_small: all launches completed")
