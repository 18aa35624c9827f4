"""Python-side launchers for the C-ABI kernels (raw pointers + the current PyTorch CUDA stream).

PyTorch is used only for device memory and streams.  Every function here fails loudly if the CUDA library is
missing or a tensor is not on a CUDA device -- there is no eager fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import Epilogue, Operand, check

ACT_NONE, ACT_GELU, ACT_DGELU = 0, 1, 2


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    if t is None:
        return None
    if not t.is_cuda:
        raise _lib.WavJepaLibError("wavjepa_b200 ops need CUDA tensors (no CPU fallback)")
    return t.data_ptr()


def make_operand(t: torch.Tensor, cols: int, rows: int, batch: int = 1, *, row_stride: Optional[int] = None,
                 batch_stride: Optional[int] = None, nq: int = 1, q_stride: Optional[int] = None,
                 seg_width: int = 0, seg_q: Sequence[int] = (), seg_p: Sequence[int] = (),
                 offset: int = 0) -> Operand:
    """4-D strided bf16 view (see wj_operand_t).  Strides are in ELEMENTS here."""
    assert t.dtype == torch.bfloat16
    op = Operand()
    op.ptr = _ptr(t) + offset * 2
    row_stride = cols if row_stride is None else row_stride
    q_stride = row_stride if q_stride is None else q_stride
    batch_stride = rows * row_stride if batch_stride is None else batch_stride
    op.dim[0], op.dim[1], op.dim[2], op.dim[3] = cols, nq, rows, batch
    op.stride_bytes[0], op.stride_bytes[1], op.stride_bytes[2] = q_stride * 2, row_stride * 2, max(batch_stride, 8) * 2
    op.seg_width = seg_width
    for i, v in enumerate(seg_q):
        op.seg_q[i] = v
    for i, v in enumerate(seg_p):
        op.seg_p[i] = v
    return op


def plain_operand(t: torch.Tensor) -> Operand:
    """Row-major [M, K] matrix (last dim contiguous)."""
    assert t.dim() == 2 and t.stride(1) == 1
    return make_operand(t, t.shape[1], t.shape[0], 1, row_stride=t.stride(0))


def gemm(a: Operand, w: torch.Tensor, L: int, batch: int, out: torch.Tensor, *, K: Optional[int] = None,
         N: Optional[int] = None, ld_out: Optional[int] = None, out_offset: int = 0,
         bias: Optional[torch.Tensor] = None, act: int = ACT_NONE, out2: Optional[torch.Tensor] = None,
         ld_out2: Optional[int] = None, out2_offset: int = 0,
         resid: Optional[torch.Tensor] = None, ld_resid: Optional[int] = None, resid_mod: int = 0,
         resid_offset: int = 0, aux: Optional[torch.Tensor] = None, ld_aux: Optional[int] = None, aux_offset: int = 0,
         accumulate: bool = False, out_rows: Optional[torch.Tensor] = None, block_n: int = 0) -> None:
    """out[b*L+t, :] = epilogue(A(t,b) @ w.T); w is bf16 [N, K] (row-major)."""
    assert w.dtype == torch.bfloat16 and w.stride(-1) == 1
    N = w.shape[0] if N is None else N
    K = w.shape[1] if K is None else K
    e = Epilogue()
    esz = out.element_size()
    e.out = _ptr(out) + out_offset * esz
    e.ld_out = out.stride(-2) if ld_out is None else ld_out
    e.out_f32 = 1 if out.dtype == torch.float32 else 0
    assert out.dtype in (torch.float32, torch.bfloat16)
    e.accumulate = 1 if accumulate else 0
    if out2 is not None:
        assert out2.dtype == torch.bfloat16
        e.out2 = _ptr(out2) + out2_offset * 2
        e.ld_out2 = out2.stride(-2) if ld_out2 is None else ld_out2
    if bias is not None:
        assert bias.dtype == torch.float32
        e.bias = _ptr(bias)
    if resid is not None:
        e.resid = _ptr(resid) + resid_offset * resid.element_size()
        e.resid_f32 = 1 if resid.dtype == torch.float32 else 0
        e.ld_resid = resid.stride(-2) if ld_resid is None else ld_resid
        e.resid_mod = resid_mod
    if aux is not None:
        assert aux.dtype == torch.bfloat16
        e.aux = _ptr(aux) + aux_offset * 2
        e.ld_aux = aux.stride(-2) if ld_aux is None else ld_aux
    e.act = act
    if out_rows is not None:
        assert out_rows.dtype == torch.int32
        e.out_rows = _ptr(out_rows)
    lib = _lib.load()
    check(lib.wj_gemm_bf16(C.byref(a), C.c_void_p(_ptr(w)), C.c_int64(w.stride(0)), L, batch, N, K, C.byref(e),
                           block_n, _stream()))


def gemm_wgrad(dy: Operand, x: Operand, L: int, batch: int, out: torch.Tensor, *, M: Optional[int] = None,
               N: Optional[int] = None, ld_out: Optional[int] = None, out_offset: int = 0, accumulate: bool = False,
               splits: int = 0) -> None:
    """out[m, n] (+)= sum_{b,t} dY[b,t,m] * X(n; t, b); out fp32 [M, N]."""
    assert out.dtype == torch.float32
    M = out.shape[0] if M is None else M
    N = out.shape[1] if N is None else N
    ld = out.stride(0) if ld_out is None else ld_out
    lib = _lib.load()
    check(lib.wj_gemm_wgrad_bf16(C.byref(dy), C.byref(x), L, batch, M, N, C.c_void_p(_ptr(out) + out_offset * 4),
                                 C.c_int64(ld), 1 if accumulate else 0, splits, _stream()))
