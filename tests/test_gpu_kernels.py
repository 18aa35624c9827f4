"""Kernel-level parity on the B200: every C-ABI entry point against a plain PyTorch fp32 restatement of the same op
(float kernels, tolerance stated per test) or against the numpy oracle (integer mask kernels, bit-exact)."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from wavjepa_b200 import _lib, ops  # noqa: E402
from oracle import jepa_oracle as jo  # noqa: E402  (the checker)

DEV = "cuda"
BF16_TOL = 4e-3   # rel-L2 of a bf16-rounded result against fp32 (bf16 eps = 3.9e-3, rel-L2 of rounding ~1.7e-3)


def rel(a, b):
    a = a.float()
    b = b.float()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


@pytest.fixture(autouse=True, scope="module")
def _device():
    _lib.require_device()
    torch.manual_seed(0)


# ------------------------------------------------------------------------------------------------------- GEMM
@pytest.mark.parametrize("M,N,K,bn", [(128, 256, 64, 256), (128, 128, 64, 128), (300, 768, 512, 0), (300, 384, 768, 256),
                                      (20000, 2304, 768, 256), (20000, 1152, 384, 128), (77, 3072, 768, 0),
                                      # 192-column tiles on the bf16 TMA-store epilogue (auto for N = 384 / 576)
                                      (777, 384, 384, 0), (5000, 384, 1536, 0), (20000, 1152, 384, 192),
                                      (300, 576, 128, 0), (1000, 768, 256, 192)])
def test_gemm_plain(M, N, K, bn):
    a = torch.randn(M, K, device=DEV).bfloat16()
    w = (torch.randn(N, K, device=DEV) * 0.05).bfloat16()
    out = torch.full((M, N), float("nan"), device=DEV, dtype=torch.bfloat16)
    ops.gemm(ops.plain_operand(a), w, M, 1, out, block_n=bn)
    assert rel(out, a.float() @ w.float().t()) < BF16_TOL


@pytest.mark.parametrize("M,N,K", [(5000, 768, 3072), (5000, 384, 1536), (777, 384, 384), (300, 576, 128)])
def test_gemm_f32_out(M, N, K):
    """fp32 outputs; N = 384 / 576 take the 192-column tiles."""
    a = torch.randn(M, K, device=DEV).bfloat16()
    w = (torch.randn(N, K, device=DEV) * 0.05).bfloat16()
    out = torch.empty(M, N, device=DEV)
    ops.gemm(ops.plain_operand(a), w, M, 1, out)
    base = a.float() @ w.float().t()
    assert rel(out, base) < 2e-5
    bias, res = torch.randn(N, device=DEV), torch.randn(M, N, device=DEV)
    ops.gemm(ops.plain_operand(a), w, M, 1, out, bias=bias, resid=res, act=ops.ACT_BF16)
    # (a few accumulators sit on a bf16 rounding boundary and land one ulp away from torch's: 1e-4 .. 2e-4 at K = 3072)
    assert rel(out, (base + bias).bfloat16().float() + res) < 3e-4


def test_gemm_epilogues():
    M, N, K = 1000, 1536, 384
    a = torch.randn(M, K, device=DEV).bfloat16()
    w = (torch.randn(N, K, device=DEV) * 0.05).bfloat16()
    bias = torch.randn(N, device=DEV)
    base = a.float() @ w.float().t()
    out = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
    out2 = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
    ops.gemm(ops.plain_operand(a), w, M, 1, out, bias=bias, act=ops.ACT_GELU, out2=out2)
    h = (base + bias).bfloat16()
    assert rel(out, F.gelu(h.float())) < BF16_TOL
    hx = h.float().requires_grad_(True)       # out2 = GELU'(h): all the backward needs of the pre-activation
    F.gelu(hx).sum().backward()
    assert rel(out2, hx.grad) < BF16_TOL
    cs = torch.zeros(N, device=DEV)           # fused column sums of the stored (bf16) output = bias gradient
    ops.gemm(ops.plain_operand(a), w, M, 1, out, bias=bias, colsum=cs)
    assert rel(cs, out.float().sum(0)) < 1e-5
    res = torch.randn(M, N, device=DEV)
    o3 = torch.empty(M, N, device=DEV)
    ops.gemm(ops.plain_operand(a), w, M, 1, o3, bias=bias, resid=res)
    assert rel(o3, base + bias + res) < 2e-5
    # act 3: the Linear output is rounded to bf16 before the fp32 residual
    ops.gemm(ops.plain_operand(a), w, M, 1, o3, bias=bias, resid=res, act=ops.ACT_BF16)
    assert rel(o3, (base + bias).bfloat16().float() + res) < 1e-4
    aux = torch.randn(M, N, device=DEV).bfloat16()
    o4 = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
    ops.gemm(ops.plain_operand(a), w, M, 1, o4, act=ops.ACT_DGELU, aux=aux)
    assert rel(o4, base * aux.float()) < BF16_TOL
    pos = torch.randn(200, N, device=DEV)
    o5 = torch.empty(M, N, device=DEV)
    ops.gemm(ops.plain_operand(a), w, M, 1, o5, resid=pos, resid_mod=200)
    assert rel(o5, base + pos[torch.arange(M, device=DEV) % 200]) < 2e-5
    perm = torch.randperm(M, device=DEV).int()
    o6 = torch.zeros(M, N, device=DEV)
    ops.gemm(ops.plain_operand(a), w, M, 1, o6, out_rows=perm)
    ref6 = torch.zeros(M, N, device=DEV)
    ref6[perm.long()] = base
    assert rel(o6, ref6) < 2e-5


def _conv_operand(x, L_in, k):
    Bn, _, Cc = x.shape
    return ops.make_operand(x, Cc, L_in // 2, Bn, nq=2, q_stride=Cc, row_stride=2 * Cc, batch_stride=L_in * Cc,
                            seg_width=Cc, seg_q=(0, 1, 0)[:k], seg_p=(0, 0, 1)[:k])


@pytest.mark.parametrize("Bn,L_in,k", [(3, 402, 3), (3, 400, 2), (2, 6430, 3)])
def test_conv_implicit_gemm(Bn, L_in, k):
    Cc = 512
    x = torch.randn(Bn, L_in, Cc, device=DEV).bfloat16()
    w = (torch.randn(Cc, Cc, k, device=DEV) * 0.03).bfloat16()
    L_out = (L_in - k) // 2 + 1
    wk = w.permute(0, 2, 1).reshape(Cc, k * Cc).contiguous()
    out = torch.full((Bn, L_out, Cc), float("nan"), device=DEV, dtype=torch.bfloat16)
    ops.gemm(_conv_operand(x, L_in, k), wk, L_out, Bn, out.view(-1, Cc))
    ref = F.conv1d(x.float().transpose(1, 2), w.float(), stride=2).transpose(1, 2)
    assert rel(out, ref) < BF16_TOL
    # overlapping-row view (row pitch 2C < row length kC)
    a = ops.make_operand(x, k * Cc, L_out, Bn, row_stride=2 * Cc, batch_stride=L_in * Cc)
    out.fill_(float("nan"))
    ops.gemm(a, wk, L_out, Bn, out.view(-1, Cc))
    assert rel(out, ref) < BF16_TOL


@pytest.mark.parametrize("M,Nw,Kw,splits", [(1000, 256, 384, 0), (64, 128, 128, 1), (50000, 768, 3072, 0), (333, 1152, 384, 0),
                                              (20000, 1536, 384, 0), (700, 512, 128, 3)])
def test_wgrad(M, Nw, Kw, splits):
    dy = torch.randn(M, Nw, device=DEV).bfloat16()
    x = torch.randn(M, Kw, device=DEV).bfloat16()
    out = torch.full((Nw, Kw), float("nan"), device=DEV)
    ops.gemm_wgrad(ops.plain_operand(dy), ops.plain_operand(x), M, 1, out, splits=splits)
    ref = dy.float().t() @ x.float()
    assert rel(out, ref) < 5e-5
    # accumulate semantics (out += ...), incl. the swapped-operand path taken when only the row count allows 256-wide tiles
    ops.gemm_wgrad(ops.plain_operand(dy), ops.plain_operand(x), M, 1, out, splits=splits, accumulate=True)
    assert rel(out, 2 * ref) < 5e-5


@pytest.mark.parametrize("L_in,k", [(402, 3), (400, 2)])
def test_conv_wgrad_dgrad(L_in, k):
    Bn, Cc = 3, 512
    x = torch.randn(Bn, L_in, Cc, device=DEV).bfloat16()
    w = (torch.randn(Cc, Cc, k, device=DEV) * 0.03).bfloat16()
    L_out = (L_in - k) // 2 + 1
    dy = torch.randn(Bn, L_out, Cc, device=DEV).bfloat16()
    xf = x.float().transpose(1, 2).requires_grad_(True)
    wf = w.float().requires_grad_(True)
    F.conv1d(xf, wf, stride=2).backward(dy.float().transpose(1, 2))
    # wgrad
    out = torch.zeros(Cc, k * Cc, device=DEV)
    ops.gemm_wgrad(ops.make_operand(dy, Cc, L_out, Bn), _conv_operand(x, L_in, k), L_out, Bn, out)
    assert rel(out, wf.grad.permute(0, 2, 1).reshape(Cc, k * Cc)) < 5e-5
    # dgrad: even / odd input positions are two GEMMs over shifted views of dY
    wk = w.permute(0, 2, 1).reshape(Cc, k * Cc).contiguous()
    dx = torch.full((Bn, L_in, Cc), float("nan"), device=DEV, dtype=torch.bfloat16)
    ops.conv_dgrad(dy, wk, dx, k)
    assert rel(dx, xf.grad.transpose(1, 2)) < BF16_TOL


@pytest.mark.parametrize("M,N,K", [(1000, 768, 2304), (515, 384, 1536), (4000, 3072, 768)])
def test_dgrad(M, N, K):
    dy = torch.randn(M, K, device=DEV).bfloat16()
    w = (torch.randn(K, N, device=DEV) * 0.05).bfloat16()   # y = x W^T, W [K(out), N(in)]
    res = torch.randn(M, N, device=DEV)
    out = torch.empty(M, N, device=DEV)
    ops.gemm_dgrad(ops.plain_operand(dy), w, M, 1, out, K=K, N=N, resid=res)
    assert rel(out, dy.float() @ w.float() + res) < 2e-5
    aux = torch.randn(M, N, device=DEV).bfloat16()
    o2 = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
    cs = torch.zeros(N, device=DEV)
    ops.gemm_dgrad(ops.plain_operand(dy), w, M, 1, o2, K=K, N=N, act=ops.ACT_DGELU, aux=aux, colsum=cs)
    assert rel(o2, (dy.float() @ w.float()) * aux.float()) < BF16_TOL
    assert rel(cs, o2.float().sum(0)) < 1e-5


@pytest.mark.parametrize("M,N,K", [(1000, 768, 2304), (515, 384, 1536), (4000, 384, 1152), (777, 384, 384)])
def test_dgrad_bf16_out(M, N, K):
    """bf16 data gradients (what autocast's Linear backward returns): TMA-store epilogue, 192-column tiles for N = 384."""
    dy = torch.randn(M, K, device=DEV).bfloat16()
    w = (torch.randn(K, N, device=DEV) * 0.05).bfloat16()
    out = torch.full((M, N), float("nan"), device=DEV, dtype=torch.bfloat16)
    ops.gemm_dgrad(ops.plain_operand(dy), w, M, 1, out, K=K, N=N)
    assert rel(out, dy.float() @ w.float()) < BF16_TOL


@pytest.mark.parametrize("M,Nw,Kw,splits", [(1000, 256, 256, 0), (50000, 768, 3072, 0), (19906, 2304, 768, 0), (700, 512, 768, 3),
                                              (300, 768, 768, 1)])
def test_wgrad_cta_pairs_match_single_ctas(M, Nw, Kw, splits):
    """Weight gradients with whole 256 x 256 tiles run on CTA pairs (tcgen05 cta_group::2) by default: same result as
    the single-CTA kernel (routing switch off) up to the fp32 order of the split-token reduction."""
    dy = torch.randn(M, Nw, device=DEV).bfloat16()
    x = torch.randn(M, Kw, device=DEV).bfloat16()
    ref = dy.float().t() @ x.float()
    outs = []
    try:
        for pair in (1, 0):
            ops.gemm_option("pair_wgrad", pair)
            out = torch.full((Nw, Kw), float("nan"), device=DEV)
            ops.gemm_wgrad(ops.plain_operand(dy), ops.plain_operand(x), M, 1, out, splits=splits)
            assert rel(out, ref) < 5e-5
            ops.gemm_wgrad(ops.plain_operand(dy), ops.plain_operand(x), M, 1, out, splits=splits, accumulate=True)
            assert rel(out, 2 * ref) < 5e-5
            outs.append(out)
    finally:
        ops.gemm_option("pair_wgrad", 1)
    assert rel(outs[0], outs[1]) < 1e-5


def test_conv_wgrad_cta_pairs():
    Bn, Cc, L_in, k = 3, 512, 402, 3
    x = torch.randn(Bn, L_in, Cc, device=DEV).bfloat16()
    L_out = (L_in - k) // 2 + 1
    dy = torch.randn(Bn, L_out, Cc, device=DEV).bfloat16()
    outs = []
    try:
        for pair in (1, 0):
            ops.gemm_option("pair_wgrad", pair)
            out = torch.zeros(Cc, k * Cc, device=DEV)
            ops.gemm_wgrad(ops.make_operand(dy, Cc, L_out, Bn), _conv_operand(x, L_in, k), L_out, Bn, out)
            outs.append(out)
    finally:
        ops.gemm_option("pair_wgrad", 1)
    assert rel(outs[0], outs[1]) < 1e-5


@pytest.mark.parametrize("M,N,K", [(256, 256, 768), (5000, 768, 2304), (19906, 768, 3072), (4099, 3072, 768)])
def test_dgrad_cta_pairs(M, N, K):
    """Plain bf16 data gradients with K >= 768 and N % 256 == 0: CTA pairs, bit-identical to single CTAs (same k order)."""
    dy = torch.randn(M, K, device=DEV).bfloat16()
    w = (torch.randn(K, N, device=DEV) * 0.05).bfloat16()
    o_pair = torch.full((M, N), float("nan"), device=DEV, dtype=torch.bfloat16)
    o_single = torch.full((M, N), float("nan"), device=DEV, dtype=torch.bfloat16)
    ops.gemm_dgrad(ops.plain_operand(dy), w, M, 1, o_pair, K=K, N=N, block_n=-256)
    ops.gemm_dgrad(ops.plain_operand(dy), w, M, 1, o_single, K=K, N=N, block_n=256)
    assert rel(o_pair, dy.float() @ w.float()) < BF16_TOL
    assert torch.equal(o_pair, o_single)
    # the GELU-backward form (out = acc * saved factor, fused bias-gradient column sums) on CTA pairs
    aux = torch.randn(M, N, device=DEV).bfloat16()
    cs_pair, cs_single = torch.zeros(N, device=DEV), torch.zeros(N, device=DEV)
    o_pair.fill_(float("nan"))
    o_single.fill_(float("nan"))
    ops.gemm_dgrad(ops.plain_operand(dy), w, M, 1, o_pair, K=K, N=N, act=ops.ACT_DGELU, aux=aux, colsum=cs_pair, block_n=-256)
    ops.gemm_dgrad(ops.plain_operand(dy), w, M, 1, o_single, K=K, N=N, act=ops.ACT_DGELU, aux=aux, colsum=cs_single, block_n=256)
    assert rel(o_pair, (dy.float() @ w.float()) * aux.float()) < BF16_TOL
    assert torch.equal(o_pair, o_single)
    assert rel(cs_pair, o_pair.float().sum(0)) < 1e-5 and rel(cs_pair, cs_single) < 1e-5
    ops.gemm_dgrad(ops.plain_operand(dy), w, M, 1, o_pair, K=K, N=N, act=ops.ACT_DGELU, aux=aux, block_n=-256)   # no column sums
    assert torch.equal(o_pair, o_single)


# ------------------------------------------------------------------------------------------------------- masks
def test_masks_time_inverse_bit_exact():
    from oracle import masks_oracle as mo

    B, T = 512, 200
    ctx, tgt, vis, att, err = ops.masks_generate(0, B, T, 1, False, 4, 0.65, 10, 0.25, 10, 0.1, 0, 1234, 7, DEV)
    assert err.item() == 0
    rc, rt, rv, ra = mo.time_inverse_masks(1234, 7, B, T)
    assert np.array_equal(ctx.cpu().numpy(), rc)
    assert np.array_equal(tgt.cpu().numpy(), rt)
    assert np.array_equal(vis.cpu().numpy(), rv)
    assert np.array_equal(att.cpu().numpy(), ra)


def test_masks_bit_exact_100k_rows():
    """>= 1e5 rows per masker, bit for bit against the numpy oracle (row blocks spread over the 32-bit row
    space, including the per-rank offsets bench.py uses)."""
    from oracle import masks_oracle as mo

    T, n = 200, 25000
    for row0 in (0, 1 << 24, 7 * (1 << 24) + 12345, (1 << 31) - n - 1):
        ctx, tgt, vis, att, err = ops.masks_generate(0, n, T, 1, False, 4, 0.65, 10, 0.25, 10, 0.1, 0, 99, row0, DEV)
        assert err.item() == 0
        rc, rt, rv, ra = mo.time_inverse_masks(99, row0, n, T)
        assert np.array_equal(ctx.cpu().numpy(), rc) and np.array_equal(tgt.cpu().numpy(), rt)
        assert np.array_equal(vis.cpu().numpy(), rv) and np.array_equal(att.cpu().numpy(), ra)
        ctx, tgt, vis, att, err = ops.masks_generate(1, n, T, 1, False, 4, 0.0, 1, 0.1, 10, 0.5, 5, 31, row0, DEV)
        assert err.item() == 0
        rc, rt, rv, ra = mo.speech_masks(31, row0, n, T, n_targets=4, tgt_prob=0.1, tgt_len=10, cutoff=0.5, min_context_len=5)
        assert np.array_equal(ctx.cpu().numpy(), rc) and np.array_equal(tgt.cpu().numpy(), rt)
        assert np.array_equal(vis.cpu().numpy(), rv) and np.array_equal(att.cpu().numpy(), ra)


def test_masks_speech_and_binaural_bit_exact():
    from oracle import masks_oracle as mo

    B, T = 256, 200
    ctx, tgt, vis, att, err = ops.masks_generate(1, B, T, 1, False, 4, 0.0, 1, 0.1, 10, 0.5, 5, 99, 0, DEV)
    assert err.item() == 0
    rc, rt, rv, ra = mo.speech_masks(99, 0, B, T, tgt_prob=0.1, tgt_len=10, cutoff=0.5, min_context_len=5)
    assert np.array_equal(ctx.cpu().numpy(), rc) and np.array_equal(tgt.cpu().numpy(), rt)
    assert np.array_equal(vis.cpu().numpy(), rv) and np.array_equal(att.cpu().numpy(), ra)
    # WavJEPA-Nat: n_times = 400 over 2 channels, "(S C)" interleave
    ctx, tgt, vis, att, err = ops.masks_generate(0, 64, 400, 2, True, 4, 0.65, 10, 0.25, 10, 0.1, 0, 5, 1000, DEV)
    rc, rt, rv, ra = mo.time_inverse_masks(5, 1000, 64, 400, in_channels=2, channel_based=True)
    assert ctx.shape == (64, 400)
    assert np.array_equal(ctx.cpu().numpy(), rc) and np.array_equal(tgt.cpu().numpy(), rt)
    assert np.array_equal(vis.cpu().numpy(), rv)


def test_mask_indices():
    B, G, T = 37, 4, 200
    ctx, tgt, vis, _, _ = ops.masks_generate(0, B, T, 1, False, G, 0.65, 10, 0.25, 10, 0.1, 0, 3, 0, DEV)
    mi = ops.mask_indices(ctx, tgt, vis)
    c, t, v = ctx.cpu(), tgt.cpu(), vis.cpu()
    ctx_rows = torch.nonzero(~c.reshape(-1)).squeeze(1)
    assert mi.Nc == ctx_rows.numel() and torch.equal(mi.ctx_rows[:mi.Nc].cpu().long(), ctx_rows)
    assert torch.equal(mi.n_c.cpu().long(), (~c).sum(1))
    assert torch.equal(mi.cu_c.cpu().long(), F.pad((~c).sum(1).cumsum(0), (1, 0)))
    visible = ~v.reshape(B * G, T)
    assert mi.Nv == int(visible.sum()) and mi.Nt == int(t.sum())
    assert torch.equal(mi.cu_v.cpu().long(), F.pad(visible.sum(1).cumsum(0), (1, 0)))
    # packed predictor rows: position and source
    pos = torch.nonzero(visible)[:, 1]
    assert torch.equal(mi.vis_pos[:mi.Nv].cpu().long(), pos)
    packed_ctx_index = torch.full((B * T,), -1, dtype=torch.long)
    packed_ctx_index[ctx_rows] = torch.arange(mi.Nc)
    bg = torch.nonzero(visible)[:, 0]
    src = packed_ctx_index[(bg // G) * T + pos]
    assert torch.equal(mi.vis_src[:mi.Nv].cpu().long(), src)
    # target rows
    tg = t.reshape(B * G, T)
    vrow_of = torch.full((B * G, T), -1, dtype=torch.long)
    vrow_of[visible] = torch.arange(mi.Nv)
    nz = torch.nonzero(tg)
    assert torch.equal(mi.tgt_vrow[:mi.Nt].cpu().long(), vrow_of[nz[:, 0], nz[:, 1]])
    assert torch.equal(mi.tgt_trow[:mi.Nt].cpu().long(), (nz[:, 0] // G) * T + nz[:, 1])
    assert mi.max_nc == int((~c).sum(1).max()) and mi.max_nv == int(visible.sum(1).max())


# ------------------------------------------------------------------------------------------------------- conv0
@pytest.mark.parametrize("Cin,L", [(1, 32159), (2, 4000), (1, 2577)])
def test_conv0_gn_gelu_fwd_bwd(Cin, L):
    B, C = 3, 512
    x = torch.randn(B, Cin, L, device=DEV).bfloat16()
    w = torch.randn(C, Cin, 10, device=DEV) * math.sqrt(2.0 / (Cin * 10))
    gamma = 1.0 + 0.1 * torch.randn(C, device=DEV)
    beta = 0.1 * torch.randn(C, device=DEV)
    L_out = (L - 10) // 5 + 1
    out = torch.empty(B, L_out, C, device=DEV, dtype=torch.bfloat16)
    mom, stats, red = ops.conv0_workspaces(B, Cin, C, DEV, backward=True)
    ops.conv0_fwd(x, w, gamma, beta, out, mom, stats)
    wr = w.bfloat16().float().requires_grad_(True)
    gr = gamma.clone().requires_grad_(True)
    br = beta.clone().requires_grad_(True)
    h = F.conv1d(x.float(), wr, stride=5)
    hq = h + (h.bfloat16().float() - h).detach()       # bf16-rounded conv output (autocast), straight-through
    y = F.gelu(F.group_norm(hq, C, gr, br, 1e-5))
    assert rel(out, y.transpose(1, 2)) < BF16_TOL
    dy = torch.randn(B, L_out, C, device=DEV).bfloat16()
    y.backward(dy.float().transpose(1, 2))
    dw = torch.zeros_like(w)
    dg = torch.zeros(C, device=DEV)
    db = torch.zeros(C, device=DEV)
    ops.conv0_bwd(x, w, gamma, beta, mom, stats, dy, red, dw, dg, db)     # GELU' recomputed from x: nothing saved
    # closed-form GroupNorm statistics == statistics of the conv output (to fp32 accuracy)
    hm = h.detach().mean(dim=2)
    assert (stats[..., 0] - hm).abs().max() < 1e-4
    assert rel(stats[..., 1], 1.0 / torch.sqrt(h.detach().var(dim=2, unbiased=False) + 1e-5)) < 1e-4
    # the recomputed GELU' (packed fp16) and dz (bf16, like autograd's bf16 conv weight gradient): ~2^-9 noise per term
    assert rel(dw, wr.grad) < 4e-3
    assert rel(dg, gr.grad) < 4e-3
    assert rel(db, br.grad) < 4e-3


# ------------------------------------------------------------------------------------------------------- norms
@pytest.mark.parametrize("D,dtype", [(768, torch.float32), (384, torch.float32), (512, torch.bfloat16)])
def test_layernorm_fwd_bwd(D, dtype):
    M = 1237
    x = (torch.randn(M, D, device=DEV) * 2 + 0.5).to(dtype)
    g = 1 + 0.1 * torch.randn(D, device=DEV)
    b = 0.1 * torch.randn(D, device=DEV)
    of = torch.empty(M, D, device=DEV)
    ob = torch.empty(M, D, device=DEV, dtype=torch.bfloat16)
    st = torch.empty(M, 2, device=DEV)
    rs = torch.empty(M, 2, device=DEV)
    ops.layernorm_fwd(x, g, b, 1e-6, of, ob, st, rs)
    xr = x.float().requires_grad_(True)
    gr = g.clone().requires_grad_(True)
    br = b.clone().requires_grad_(True)
    y = F.layer_norm(xr, (D,), gr, br, 1e-6)
    assert rel(of, y) < 1e-5 and rel(ob, y) < BF16_TOL
    assert rel(rs[:, 0], y.sum(1)) < 1e-3 and rel(rs[:, 1], (y * y).sum(1)) < 1e-5
    dy = torch.randn(M, D, device=DEV)
    y.backward(dy)
    dxf = torch.empty(M, D, device=DEV)
    dxb = torch.empty(M, D, device=DEV, dtype=torch.bfloat16)
    dg = torch.zeros(D, device=DEV)
    db = torch.zeros(D, device=DEV)
    cs = torch.zeros(D, device=DEV)
    ops.layernorm_bwd(dy, x, st, g, dxf, dxb, dg, db, cs)
    assert rel(dxf, xr.grad) < 1e-4 and rel(dxb, xr.grad) < BF16_TOL
    assert rel(dg, gr.grad) < 1e-4 and rel(db, br.grad) < 1e-4
    assert (cs - xr.grad.sum(0)).abs().max().item() < 1e-2


@pytest.mark.parametrize("D", [384, 768])
def test_add_layernorm_fwd_bwd(D):
    """LayerNorm(x + bf16 addend) with the sum formed in registers, and its backward with a two-part output gradient
    (fp32 residual branch + bf16 Linear data gradient) -- the post-norm layer of wavjepa/types/wavjepa_configs.py:29-47."""
    M = 1237
    x = torch.randn(M, D, device=DEV) * 2 + 0.5
    a = torch.randn(M, D, device=DEV).bfloat16()
    g = 1 + 0.1 * torch.randn(D, device=DEV)
    b = 0.1 * torch.randn(D, device=DEV)
    of = torch.empty(M, D, device=DEV)
    ob = torch.empty(M, D, device=DEV, dtype=torch.bfloat16)
    st = torch.empty(M, 2, device=DEV)
    rs = torch.empty(M, 2, device=DEV)
    ops.add_layernorm_fwd(x, a, g, b, 1e-6, of, ob, st, rs)
    xr = x.clone().requires_grad_(True)
    ar = a.float().requires_grad_(True)
    gr = g.clone().requires_grad_(True)
    br = b.clone().requires_grad_(True)
    y = F.layer_norm(xr + ar, (D,), gr, br, 1e-6)
    assert rel(of, y) < 1e-5 and rel(ob, y) < BF16_TOL
    assert rel(rs[:, 1], (y * y).sum(1)) < 1e-5
    d32 = torch.randn(M, D, device=DEV)
    d16 = torch.randn(M, D, device=DEV).bfloat16()
    y.backward(d32 + d16.float())
    for (da, db_) in ((d32, d16), (d32 + d16.float(), None), (None, None)):
        dxf = torch.empty(M, D, device=DEV)
        dxb = torch.empty(M, D, device=DEV, dtype=torch.bfloat16)
        dg = torch.zeros(D, device=DEV)
        db = torch.zeros(D, device=DEV)
        cs = torch.zeros(D, device=DEV)
        if da is None:
            with pytest.raises(_lib.WavJepaLibError):
                ops.add_layernorm_bwd(None, None, x, a, st, g, dxf, dxb, dg, db, cs)
            continue
        ops.add_layernorm_bwd(da, db_, x, a, st, g, dxf, dxb, dg, db, cs)
        assert rel(dxf, xr.grad) < 1e-4 and rel(dxb, xr.grad) < BF16_TOL
        assert rel(dg, gr.grad) < 1e-4 and rel(db, br.grad) < 1e-4
        assert (cs - xr.grad.sum(0)).abs().max().item() < 1e-2
    # bf16-only output gradient, no addend (degenerates to the plain LayerNorm backward)
    ops.add_layernorm_fwd(x, None, g, b, 1e-6, of, None, st, None)
    xr2 = x.clone().requires_grad_(True)
    F.layer_norm(xr2, (D,), g, b, 1e-6).backward(d16.float())
    dxf = torch.empty(M, D, device=DEV)
    ops.add_layernorm_bwd(None, d16, x, None, st, g, dxf, None, None, None, None)
    assert rel(dxf, xr2.grad) < 1e-4
    acc = d32.clone()
    ops.add_bf16(acc, d16)
    assert torch.equal(acc, d32 + d16.float())


def test_crop_norm():
    n_clips, S, Lc, Lfull = 4, 8, 32159, 160000
    audio = torch.randn(n_clips, 1, Lfull, device=DEV) * 0.3 + 0.01
    starts = torch.randint(0, Lfull - Lc + 1, (n_clips * S,), device=DEV, dtype=torch.int32)
    ob = torch.empty(n_clips * S, 1, Lc, device=DEV, dtype=torch.bfloat16)
    of = torch.empty(n_clips * S, 1, Lc, device=DEV)
    ops.crop_norm(audio, starts, S, Lc, ob, of)
    idx = starts.long().view(n_clips, S, 1) + torch.arange(Lc, device=DEV)
    crops = torch.gather(audio.expand(-1, S, -1), 2, idx).reshape(n_clips * S, 1, Lc)
    ref = (crops - crops.mean(dim=(-2, -1), keepdim=True)) / (crops.std(dim=(-2, -1), keepdim=True) + 1e-5)
    assert rel(of, ref) < 1e-5 and rel(ob, ref) < BF16_TOL
    # HEAR chunking: no starts table, fixed stride, zero padding past the clip end, 2 channels normalised jointly
    audio2 = torch.randn(3, 2, 70000, device=DEV)
    n_chunks = 3
    st2 = (torch.arange(n_chunks, device=DEV, dtype=torch.int32) * Lc).repeat(3)
    of2 = torch.empty(3 * n_chunks, 2, Lc, device=DEV)
    ops.crop_norm(audio2, st2, n_chunks, Lc, None, of2)
    padded = F.pad(audio2, (0, n_chunks * Lc - 70000))
    ch = padded.view(3, 2, n_chunks, Lc).permute(0, 2, 1, 3).reshape(3 * n_chunks, 2, Lc)
    ref2 = (ch - ch.mean(dim=(-2, -1), keepdim=True)) / (ch.std(dim=(-2, -1), keepdim=True) + 1e-5)
    assert rel(of2, ref2) < 1e-5


def test_target_accum():
    B, T, D, K = 5, 200, 768, 3
    layers = [torch.randn(B * T, D, device=DEV) * (i + 1) + 0.3 * i for i in range(K)]
    g = torch.ones(D, device=DEV)
    b = torch.zeros(D, device=DEV)
    targets = torch.empty(B * T, D, device=DEV)
    inst = torch.empty(B, 2, device=DEV)
    outs = []
    for i, x in enumerate(layers):
        of = torch.empty_like(x)
        rs = torch.empty(B * T, 2, device=DEV)
        ops.layernorm_fwd(x, g, b, 1e-6, of, None, None, rs)
        outs.append(of)
        ops.target_accum(of, rs, B, T, D, 1.0 / K, i == 0, inst, targets)
    stacked = torch.stack([o.view(B, T, D) for o in outs]).transpose(2, 3)
    ref = F.instance_norm(stacked).transpose(2, 3).mean(0)
    assert rel(targets.view(B, T, D), ref) < 1e-5
    # the one-pass form over all layers (what the teacher forward uses)
    sums = []
    for o in outs:
        rs = torch.empty(B * T, 2, device=DEV)
        rs[:, 0] = o.sum(1)
        rs[:, 1] = (o * o).sum(1)
        sums.append(rs)
    t2 = torch.full_like(targets, float("nan"))
    ops.target_combine(outs, sums, B, T, D, 1.0 / K, torch.empty(K, B, 2, device=DEV), t2)
    assert rel(t2.view(B, T, D), ref) < 1e-5


# ------------------------------------------------------------------------------------------------------- attention
def _attn_ref(qkv, cu, D, H):
    dh = D // H
    outs = []
    for s in range(len(cu) - 1):
        x = qkv[cu[s]:cu[s + 1]]
        n = x.shape[0]
        if n == 0:
            continue
        q, k, v = (x[:, i * D:(i + 1) * D].reshape(n, H, dh).transpose(0, 1) for i in range(3))
        o = F.scaled_dot_product_attention(q, k, v)
        outs.append(o.transpose(0, 1).reshape(n, D))
    return torch.cat(outs)


@pytest.mark.parametrize("D,H,lens", [(768, 12, [39, 72, 20, 1, 64, 65]), (384, 12, [85, 122, 128, 17]),
                                      (768, 12, [200, 200, 200]), (384, 12, [250, 3]), (384, 12, [5, 0, 128, 0, 33]),
                                      (768, 12, [128] * 40 + [7]), (768, 24, [129, 256, 1, 128]), (384, 12, [257, 5]),
                                      (768, 12, [256, 255, 130, 2]), (384, 12, [64] * 300 + [100] * 300),
                                      # 257..512 tokens: four key blocks / query tiles on the 512-column forward
                                      (768, 12, [400, 400, 37]), (384, 12, [400, 230, 512, 1]),
                                      (768, 12, [512, 385, 384, 129] * 3), (384, 12, [272, 7] * 40), (384, 12, [600, 5])])
def test_attention_fwd_bwd(D, H, lens):
    cu_l = [0]
    for n in lens:
        cu_l.append(cu_l[-1] + n)
    tot = cu_l[-1]
    cu = torch.tensor(cu_l, device=DEV, dtype=torch.int32)
    qkv = torch.randn(tot, 3 * D, device=DEV).bfloat16()
    out = torch.full((tot, D), float("nan"), device=DEV, dtype=torch.bfloat16)
    lse = torch.empty(tot, H, device=DEV)
    ops.attn_fwd(qkv, cu, len(lens), max(lens), D, H, out, lse)
    qr = qkv.float().requires_grad_(True)
    ref = _attn_ref(qr, cu_l, D, H)
    assert rel(out, ref) < 6e-3
    if max(lens) > (384 if D // H == 64 else 704):
        # forward-only length (the 400-token binaural teacher has no backward): the backward keeps a whole sequence in
        # shared memory and must refuse loudly rather than fall back
        with pytest.raises(_lib.WavJepaLibError):
            ops.attn_bwd(qkv, out, torch.zeros_like(out), lse, cu, len(lens), max(lens), D, H, torch.empty_like(qkv))
        return
    do = torch.randn(tot, D, device=DEV).bfloat16()
    ref.backward(do.float())
    dqkv = torch.full((tot, 3 * D), float("nan"), device=DEV, dtype=torch.bfloat16)
    ops.attn_bwd(qkv, out, do, lse, cu, len(lens), max(lens), D, H, dqkv)
    assert rel(dqkv, qr.grad) < 1.5e-2
    # the same call with the in_proj bias gradient (column sums of the stored dqkv) folded in: identical dqkv
    dq2 = torch.full((tot, 3 * D), float("nan"), device=DEV, dtype=torch.bfloat16)
    db = torch.ones(3 * D, device=DEV)
    ops.attn_bwd(qkv, out, do, lse, cu, len(lens), max(lens), D, H, dq2, dbias=db)
    assert torch.equal(dq2, dqkv)
    # against the fp32 autograd bias gradient (its key third is zero up to rounding, its value third is colsum(dO))
    cs = qr.grad.sum(0)
    assert (db - 1.0 - cs).abs().max().item() <= 1e-2 * cs.abs().max().item() + 1e-3
    assert rel(db - 1.0, cs) < 1e-2


# ------------------------------------------------------------------------------------------------------- elementwise
def test_gather_scatter_assemble():
    N, D, R = 700, 384, 1000
    src = torch.randn(R, D, device=DEV)
    idx = torch.randperm(R, device=DEV)[:N].int()
    of = torch.empty(N, D, device=DEV)
    ob = torch.empty(N, D, device=DEV, dtype=torch.bfloat16)
    ops.gather_rows(src, idx, N, of, ob)
    assert torch.equal(of, src[idx.long()]) and torch.equal(ob, src[idx.long()].bfloat16())
    ops.gather_rows(src.bfloat16(), None, R, of2 := torch.empty(R, D, device=DEV), None)
    assert torch.equal(of2, src.bfloat16().float())
    h = torch.randn(R, D, device=DEV).bfloat16()
    out = torch.zeros(R, D, device=DEV, dtype=torch.bfloat16)
    ops.scatter_dgelu(of, idx, h, N, out)
    ref = torch.zeros(R, D, device=DEV)
    ref[idx.long()] = of * h.float()[idx.long()]
    assert rel(out, ref) < BF16_TOL
    out.zero_()
    ops.scatter_dgelu(src, None, h, R, out)
    assert rel(out, src * h.float()) < BF16_TOL
    # predictor input assembly
    Nc, Nv, T = 300, 900, 200
    ctx = torch.randn(Nc, D, device=DEV).bfloat16()
    mt = torch.randn(D, device=DEV) * 0.02
    pos = torch.randn(T, D, device=DEV)
    vsrc = torch.randint(-1, Nc, (Nv,), device=DEV, dtype=torch.int32)
    vpos = torch.randint(0, T, (Nv,), device=DEV, dtype=torch.int32)
    xf = torch.empty(Nv, D, device=DEV)
    xb = torch.empty(Nv, D, device=DEV, dtype=torch.bfloat16)
    ops.predictor_assemble(ctx, mt, pos, vsrc, vpos, Nv, D, xf, xb)
    base = torch.where((vsrc >= 0)[:, None], ctx.float()[vsrc.clamp(min=0).long()], mt.bfloat16().float()[None])
    assert torch.equal(xf, base + pos[vpos.long()])
    dx0 = torch.randn(Nv, D, device=DEV)
    dctx = torch.zeros(Nc, D, device=DEV)
    dmt = torch.zeros(D, device=DEV)
    ops.predictor_assemble_bwd(dx0, vsrc, Nv, D, dctx, dmt)
    rctx = torch.zeros(Nc, D, device=DEV).index_add_(0, vsrc.clamp(min=0).long(), dx0 * (vsrc >= 0)[:, None])
    assert rel(dctx, rctx) < 1e-5 and rel(dmt, dx0[vsrc < 0].sum(0)) < 1e-4


def test_predictor_ctx_grad_is_the_scatter_without_atomics():
    """Context part of the predictor-input backward as a fixed-order gather: equals the fp32 index_add, is bit-identical
    from call to call, and the mask-token-only form of predictor_assemble_bwd leaves d_ctx alone."""
    B, G, T, D = 37, 4, 200, 384
    g = torch.Generator(device="cpu").manual_seed(3)
    ctx_vis = torch.rand(B, T, generator=g) < 0.35
    tgt = torch.rand(B, G, T, generator=g) < 0.2
    tgt &= ~ctx_vis[:, None]
    vis = ctx_vis[:, None] | tgt                              # predictor sequences: context + that group's targets
    n_c = ctx_vis.sum(1)
    cu_c = torch.zeros(B + 1, dtype=torch.int64); cu_c[1:] = n_c.cumsum(0)
    n_v = vis.view(B * G, T).sum(1)
    cu_v = torch.zeros(B * G + 1, dtype=torch.int64); cu_v[1:] = n_v.cumsum(0)
    Nc, Nv = int(cu_c[-1]), int(cu_v[-1])
    ctx_row = torch.full((B, T), -1, dtype=torch.int64)
    ctx_row[ctx_vis] = torch.arange(Nc)
    vsrc = ctx_row[:, None].expand(B, G, T)[vis].to(torch.int32).to(DEV)      # time order inside each (b, g) sequence
    dx0 = torch.randn(Nv, D, device=DEV)
    o32 = torch.full((Nc, D), float("nan"), device=DEV)
    o16 = torch.full((Nc, D), float("nan"), device=DEV, dtype=torch.bfloat16)
    cu_v_d = cu_v.to(torch.int32).to(DEV)
    ops.predictor_ctx_grad(dx0, vsrc, cu_v_d, B * G, G, Nc, D, o32, o16)
    ref = torch.zeros(Nc, D, device=DEV).index_add_(0, vsrc.clamp(min=0).long(), dx0 * (vsrc >= 0)[:, None])
    assert rel(o32, ref) < 1e-6 and torch.equal(o16, o32.bfloat16())
    o32b = torch.empty_like(o32)
    ops.predictor_ctx_grad(dx0, vsrc, cu_v_d, B * G, G, Nc, D, o32b, None)
    assert torch.equal(o32, o32b)
    dmt = torch.zeros(D, device=DEV)
    ops.predictor_assemble_bwd(dx0, vsrc, Nv, D, None, dmt)
    assert rel(dmt, dx0[vsrc < 0].sum(0)) < 1e-4


def test_masked_mse():
    Nt, D, R = 4321, 768, 2000
    pred = torch.randn(Nt, D, device=DEV).bfloat16()
    tg = torch.randn(R, D, device=DEV)
    rows = torch.randint(0, R, (Nt,), device=DEV, dtype=torch.int32)
    loss = torch.zeros(1, device=DEV)
    dp = torch.empty(Nt, D, device=DEV, dtype=torch.bfloat16)
    ops.masked_mse(pred, tg, rows, Nt, D, loss, dp)
    pr = pred.float().requires_grad_(True)
    ref = ((pr - tg[rows.long()]) ** 2).mean(-1).sum() / (Nt + 1e-8)
    ref.backward()
    assert abs(loss.item() - ref.item()) / ref.item() < 1e-5
    assert rel(dp, pr.grad) < BF16_TOL


def test_ema_bitwise_and_optimizer_tail():
    n = 1_000_003
    t = torch.randn(n, device=DEV)
    s = torch.randn(n, device=DEV)
    r = 0.999
    ref = t.clone().mul_(r).add_((1 - r) * s)
    ops.ema_update(t, s, r)
    assert torch.equal(t, ref)
    # AdamW + clip against torch.optim.AdamW / clip_grad_norm_
    p = torch.randn(n, device=DEV)
    g = torch.randn(n, device=DEV) * 0.1
    pr = torch.nn.Parameter(p.clone())
    opt = torch.optim.AdamW([pr], lr=4e-4, betas=(0.9, 0.98), eps=1e-6, weight_decay=0.04)
    m = torch.zeros(n, device=DEV)
    v = torch.zeros(n, device=DEV)
    pb = torch.empty(n, device=DEV, dtype=torch.bfloat16)
    for step in range(1, 4):
        pr.grad = g.clone() * step
        torch.nn.utils.clip_grad_norm_([pr], 5.0)
        opt.step()
        ss = torch.zeros(1, device=DEV, dtype=torch.float64)
        ops.sumsq(g * step, 1.0, ss)
        ops.adamw_step(p, g * step, m, v, 4e-4, 0.9, 0.98, 1e-6, 0.04, step, 1.0, 5.0, ss, pb)
        assert rel(p, pr.data) < 1e-6
    assert torch.equal(pb, p.bfloat16())
    # the same step with the EMA teacher update folded in (pre-step student values of p[lo:hi]): bitwise equal to the
    # two separate kernels
    lo, hi = 4096, 4096 + 500_000
    p2, m2, v2 = p.clone(), m.clone(), v.clone()
    tea = torch.randn(hi - lo, device=DEV)
    tea_ref = tea.clone()
    ops.ema_update(tea_ref, p[lo:hi], 0.9995)
    ss = torch.zeros(1, device=DEV, dtype=torch.float64)
    ops.sumsq(g, 1.0, ss)
    ops.adamw_step(p, g, m, v, 4e-4, 0.9, 0.98, 1e-6, 0.04, 4, 1.0, 5.0, ss, pb)
    pb2 = torch.empty_like(pb)
    tb = torch.empty(hi - lo, device=DEV, dtype=torch.bfloat16)
    ops.adamw_ema_step(p2, g, m2, v2, 4e-4, 0.9, 0.98, 1e-6, 0.04, 4, 1.0, 5.0, ss, pb2, tea, tb, lo, hi, 0.9995)
    assert torch.equal(p2, p) and torch.equal(m2, m) and torch.equal(v2, v) and torch.equal(pb2, pb)
    assert torch.equal(tea, tea_ref) and torch.equal(tb, tea_ref.bfloat16())
    y = torch.empty(n, device=DEV, dtype=torch.bfloat16)
    ops.cast_bf16(s, y)
    assert torch.equal(y, s.bfloat16())
    x = torch.randn(3000, 768, device=DEV)
    cs = torch.zeros(768, device=DEV)
    ops.colsum(x.bfloat16(), cs)
    assert rel(cs, x.bfloat16().float().sum(0)) < 1e-4
    # fp32 input, ragged row count, non-multiple-of-8 width (generic kernel), strided view (ld > N)
    for (M, N, dt) in ((1237, 1152, torch.bfloat16), (999, 384, torch.float32), (77, 1150, torch.bfloat16)):
        xx = torch.randn(M, N + 8, device=DEV).to(dt)[:, :N]
        cs = torch.ones(N, device=DEV)
        ops.colsum(xx, cs)
        assert rel(cs, 1.0 + xx.float().sum(0)) < 1e-4, (M, N, dt)


# ------------------------------------------------------------------------------------------------------- CTA-pair GEMM
@pytest.mark.parametrize("M,N,K,bn", [(256, 256, 64, -256), (300, 768, 512, -256), (20000, 2304, 768, -256),
                                      (777, 384, 384, -128), (5000, 1152, 384, -128), (131, 3072, 768, -256)])
def test_gemm_pair_plain(M, N, K, bn):
    """tcgen05 cta_group::2 kernel (256-row tiles shared by two CTAs of a cluster): same results as the 1-CTA path."""
    a = torch.randn(M, K, device=DEV).bfloat16()
    w = (torch.randn(N, K, device=DEV) * 0.05).bfloat16()
    out = torch.full((M, N), float("nan"), device=DEV, dtype=torch.bfloat16)
    ops.gemm(ops.plain_operand(a), w, M, 1, out, block_n=bn)
    assert rel(out, a.float() @ w.float().t()) < BF16_TOL
    ref = torch.empty_like(out)
    ops.gemm(ops.plain_operand(a), w, M, 1, ref, block_n=-bn)
    assert torch.equal(out, ref)


def test_gemm_pair_epilogues_and_conv():
    M, N, K = 1000, 1536, 384
    a = torch.randn(M, K, device=DEV).bfloat16()
    w = (torch.randn(N, K, device=DEV) * 0.05).bfloat16()
    bias = torch.randn(N, device=DEV)
    res = torch.randn(M, N, device=DEV)
    base = a.float() @ w.float().t()
    o = torch.empty(M, N, device=DEV)
    ops.gemm(ops.plain_operand(a), w, M, 1, o, bias=bias, resid=res, act=ops.ACT_BF16, block_n=-256)
    assert rel(o, (base + bias).bfloat16().float() + res) < 1e-4
    out = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
    out2 = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
    ops.gemm(ops.plain_operand(a), w, M, 1, out, bias=bias, act=ops.ACT_GELU, out2=out2, block_n=-256)
    assert rel(out, F.gelu((base + bias).bfloat16().float())) < BF16_TOL
    # implicit-GEMM conv (k3, stride 2) over channels-last activations, ragged L_out per batch entry
    B, L_in, C = 3, 802, 512
    x = torch.randn(B, L_in, C, device=DEV).bfloat16()
    wc = (torch.randn(C, C, 3, device=DEV) * 0.03)
    wk = wc.permute(0, 2, 1).reshape(C, 3 * C).bfloat16().contiguous()
    L_out = (L_in - 3) // 2 + 1
    y = torch.empty(B, L_out, C, device=DEV, dtype=torch.bfloat16)
    ops.gemm(ops.conv_operand(x, 3), wk, L_out, B, y.view(-1, C), block_n=-256)
    ref = F.conv1d(x.float().transpose(1, 2), wk.float().view(C, 3, C).permute(0, 2, 1), stride=2).transpose(1, 2)
    assert rel(y, ref) < BF16_TOL


# ------------------------------------------------------------------------------------------------------- denoiser stage
@pytest.mark.parametrize("M,alpha", [(4 * 200 * 768, 0.25), (1000, 0.0), (4096 * 33, 1.0)])
def test_mse_pair(M, alpha):
    pred = torch.randn(2, M, device=DEV)
    tgt = torch.randn(M, device=DEV)
    sums = torch.zeros(2, device=DEV, dtype=torch.float64)
    dp = torch.full((2, M), float("nan"), device=DEV)
    ops.mse_pair(pred, tgt, alpha, sums, dp)
    ref = ((pred.double() - tgt.double()) ** 2).sum(dim=1)
    assert torch.allclose(sums, ref, rtol=1e-6)
    p2 = pred.clone().requires_grad_(True)
    loss = alpha * F.mse_loss(p2[0], tgt) + (1 - alpha) * F.mse_loss(p2[1], tgt)
    loss.backward()
    assert rel(dp, p2.grad) < 1e-6
    ops.mse_pair(pred, tgt, alpha, sums, None)          # forward only: accumulates, no gradient buffer
    assert torch.allclose(sums, 2 * ref, rtol=1e-6)


def test_snr_mix_matches_reference_formula():
    B, T = 5, 70001
    g = torch.Generator().manual_seed(4)
    src = torch.randn(B, T, generator=g)
    noise = torch.randn(B, T, generator=g) * 0.2
    start = torch.tensor([0, 1000, 69000, 35000, 123])       # a window that runs past the end, an empty window
    length = torch.tensor([T, 20000, 5000, 0, 1])
    snr = torch.tensor([0.0, 10.0, -5.0, 20.0, 3.0])
    ref = jo.scene_add_noise(src[:, None], noise[:, None], snr, start, length)[:, 0]
    out = torch.full((B, T), float("nan"), device=DEV)
    ops.snr_mix(src.to(DEV), noise.to(DEV), start.int().to(DEV), length.int().to(DEV), snr.to(DEV), out)
    assert rel(out.cpu(), ref) < 1e-6
    assert torch.equal(out[3].cpu(), src[3])                 # empty window: a = 0, the source passes through
