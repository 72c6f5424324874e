"""Drop-in for the two ``torch_scatter`` functions the reference calls (``ASF/utils/soflow.py:13,474,481``):
``scatter_softmax(src[B,L,C], index[B,L], dim=1)`` and ``scatter_sum(src, index, dim=1)``.

Unlike torch_scatter's atomics, accumulation order is fixed (rows of a segment in ascending order), so the
result is reproducible run to run.  ``dim_size`` defaults to ``index.max() + 1`` like torch_scatter, which costs a
device sync; the fused model path passes the known size instead.
"""
import torch

from . import _native as nat


def _csr(index, n_seg):
    B, L = index.shape
    index = index.contiguous()
    ws = torch.empty(int(nat.lib().ssf_csr_workspace_ints(B, L, n_seg)), dtype=torch.int32, device=index.device)
    if index.dtype == torch.int64:
        nat.check(nat.lib().ssf_build_csr_i64(nat.ptr(index), B, L, n_seg, nat.ptr(ws), nat.stream()))
    elif index.dtype == torch.int32:
        nat.check(nat.lib().ssf_build_csr_i32(nat.ptr(index), B, L, n_seg, nat.ptr(ws), nat.stream()))
    else:
        raise nat.SsfError("scatter index must be int64 or int32")
    return ws


def _prep(src, index, dim):
    nat.require_device()
    if dim != 1 or src.dim() != 3 or index.dim() != 2:
        raise nat.SsfError("only the reference's call shape is supported: src [B,L,C], index [B,L], dim=1")
    if src.dtype != torch.float32:
        raise nat.SsfError("src must be float32")
    return src.contiguous()


@torch.no_grad()
def scatter_softmax(src, index, dim=1, dim_size=None):
    src = _prep(src, index, dim)
    B, L, C = src.shape
    n_seg = int(index.max()) + 1 if dim_size is None else int(dim_size)
    ws = _csr(index, n_seg)
    out = torch.zeros_like(src)
    nat.check(nat.lib().ssf_segment_softmax(nat.ptr(src), nat.ptr(ws), B, L, C, n_seg, nat.ptr(out), nat.stream()))
    return out


@torch.no_grad()
def scatter_sum(src, index, dim=1, dim_size=None):
    src = _prep(src, index, dim)
    B, L, C = src.shape
    n_seg = int(index.max()) + 1 if dim_size is None else int(dim_size)
    ws = _csr(index, n_seg)
    out = torch.empty(B, n_seg, C, dtype=torch.float32, device=src.device)
    nat.check(nat.lib().ssf_segment_sum(nat.ptr(src), nat.ptr(ws), B, L, C, n_seg, nat.ptr(out), nat.stream()))
    return out
