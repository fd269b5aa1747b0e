"""Feasibility probe for the halo-resident 3x3 convolution: a tcgen05 K-major SWIZZLE_128B A operand that starts `shift`
rows into a TMA-written tile.  Needs artiboost_b200/build/variants/ummatest.so (gemm_tc.cu with -DAB_UMMA_SHIFT_TEST)."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from artiboost_b200 import lib  # noqa: E402

L = lib.load()
L.ab_debug_umma_shift.restype = C.c_int
L.ab_debug_umma_shift.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
dev = torch.device("cuda", 0)
torch.manual_seed(0)
rows = 208
A = torch.randn((rows, 64), device=dev).to(torch.bfloat16).contiguous()
B = torch.randn((64, 64), device=dev).to(torch.bfloat16).contiguous()
for bo in (0, 1):
    line = []
    for shift in (0, 1, 2, 3, 5, 7, 8, 9, 13, 66, 67, 79):
        out = torch.zeros((128, 64), device=dev)
        rc = L.ab_debug_umma_shift(A.data_ptr(), rows, B.data_ptr(), out.data_ptr(), shift, bo, lib.stream_ptr(dev))
        torch.cuda.synchronize()
        assert rc == 0, (rc, L.ab_last_error())
        ref = A[shift:shift + 128].float() @ B.float().t()
        err = float((out - ref).abs().max())
        line.append("%d:%s" % (shift, "ok" if err < 1e-3 else "BAD(%.1f)" % err))
    print("base_offset field %s -> %s" % ("set" if bo else "zero", " ".join(line)))
