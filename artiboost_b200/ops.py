"""Thin Python wrappers over the tensor-core / elementwise entry points of the C-ABI (include/artiboost_b200.h)."""
from __future__ import annotations

from typing import Optional

import torch

from . import lib


def gemm_bf16(a: torch.Tensor, b: torch.Tensor, scale: Optional[torch.Tensor] = None, bias: Optional[torch.Tensor] = None,
              residual: Optional[torch.Tensor] = None, relu: bool = False, out_fp32: bool = False,
              out: Optional[torch.Tensor] = None, col_stats: Optional[tuple] = None) -> torch.Tensor:
    """out[M,N] = epilogue(a[M,K] @ b[N,K]^T) on tcgen05 tensor cores.  a, b: bf16, last dim contiguous."""
    lib.require_cuda(a, "a")
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16 and a.dim() == 2 and b.dim() == 2
    assert a.stride(1) == 1 and b.stride(1) == 1 and a.shape[1] == b.shape[1]
    M, K = a.shape
    N = b.shape[0]
    dev = a.device
    if out is None:
        out = torch.empty((M, N), dtype=torch.float32 if out_fp32 else torch.bfloat16, device=dev)
    assert out.stride(1) == 1 and out.shape == (M, N)
    if residual is not None:
        assert residual.dtype == torch.bfloat16 and residual.stride(1) == 1 and residual.shape == (M, N)
    cs, cq = col_stats if col_stats is not None else (None, None)
    p = lambda t: None if t is None else t.data_ptr()  # noqa: E731
    with torch.cuda.device(dev):
        rc = lib.load().ab_gemm_bf16(M, N, K, p(a), a.stride(0), p(b), b.stride(0), p(out), out.stride(0),
                                     int(out.dtype == torch.float32), p(scale), p(bias), p(residual),
                                     0 if residual is None else residual.stride(0), int(relu), p(cs), p(cq),
                                     lib.stream_ptr(dev))
    lib.check(rc, "ab_gemm_bf16")
    return out
