"""NHWC bf16 activations and the layer primitives of the clasbased network on top of the C-ABI.

Every convolution / transposed convolution / linear layer is `im2col rows x packed filter matrix` on the tcgen05
GEMM (ab_gemm_bf16) with the BatchNorm affine, bias, residual add and ReLU fused into its epilogue.  Parameters stay
ordinary fp32 `nn.Parameter`s in the reference's layouts (state_dict compatible); packed bf16 copies are cached and
rebuilt when a parameter's version counter changes.
"""
from __future__ import annotations

import weakref
from dataclasses import dataclass
from typing import Optional

import torch
import torch.nn as nn

from .. import lib, ops


@dataclass
class Act:
    """bf16 [B*H*W, C] view of an NHWC activation."""
    data: torch.Tensor
    B: int
    H: int
    W: int
    C: int

    def nchw(self, dtype=torch.float32) -> torch.Tensor:
        return self.data.view(self.B, self.H, self.W, self.C).permute(0, 3, 1, 2).to(dtype)


_cache = weakref.WeakKeyDictionary()  # module -> {tag: (versions, packed value)}; dies with the module (no id() reuse)
_param_epoch = [0]
IMPLICIT_CONV = True  # False: materialise im2col rows (ab_im2col_nhwc) and run the plain GEMM, for A/B comparison


def _cached(owner, tag, versions, build):
    """Packed copy of `owner`'s parameters, rebuilt when a parameter's storage / version / the parameter epoch changes."""
    slot = _cache.get(owner)
    if slot is None:
        slot = _cache[owner] = {}
    hit = slot.get(tag)
    if hit is not None and hit[0] == versions:
        return hit[1]
    val = build()
    slot[tag] = (versions, val)
    return val


def bump_params():
    """Parameters were rewritten behind autograd's back (the fused Adam kernel writes the flat buffer): invalidate
    every packed copy."""
    _param_epoch[0] += 1


def _ver(*tensors):
    return (_param_epoch[0],) + tuple((t.data_ptr(), t._version, tuple(t.shape)) for t in tensors if t is not None)


def _pad8(n: int) -> int:
    return (n + 7) // 8 * 8


def _pad16(n: int) -> int:
    """K of a filter matrix / materialised im2col matrix: rows of a multiple of 32 bytes start on DRAM-sector boundaries
    (the stem's K = 196 -> 208: a 400-byte pitch made every 128-byte TMA box row straddle five sectors)."""
    return (n + 15) // 16 * 16


def pack_conv_weight(w: torch.Tensor, cin_pad: int = 0) -> torch.Tensor:
    """[Cout, Cin, kh, kw] fp32 -> bf16 [Cout, Kp], K order (ky, kx, ci), zero padded to a multiple of 16.
    cin_pad > Cin pads the input-channel axis with zeros first (the 3-channel image travels as 4 channels)."""
    cout = w.shape[0]
    m = w.detach().permute(0, 2, 3, 1)
    if cin_pad > m.shape[3]:
        m = torch.nn.functional.pad(m, (0, cin_pad - m.shape[3]))
    m = m.reshape(cout, -1)
    kp = _pad16(m.shape[1])
    out = torch.zeros((cout, kp), dtype=torch.bfloat16, device=w.device)
    out[:, :m.shape[1]] = m.to(torch.bfloat16)
    return out


def packed_filters(conv: nn.Conv2d, cin_pad: int, with_dgrad: bool = False):
    """bf16 operand copies of conv.weight, cached per module and rebuilt by ONE kernel when the parameter changes:
    -> (wp [Cout, Kp] forward filter matrix, wd [Cin, kh*kw*Cout] data-gradient filter matrix or None)."""
    w = conv.weight
    cout, cin, kh, kw = w.shape
    cin_pad = max(cin_pad, cin)

    def build():
        lib.require_cuda(w, "conv.weight")
        kp = _pad16(kh * kw * cin_pad)
        wp = torch.empty((cout, kp), dtype=torch.bfloat16, device=w.device)
        wd = torch.empty((cin, kh * kw * cout), dtype=torch.bfloat16, device=w.device) if with_dgrad else None
        src = w.detach().float().contiguous()
        with torch.cuda.device(w.device):
            rc = lib.load().ab_pack_conv_filters(src.data_ptr(), cout, cin, kh, kw, cin_pad, kp, wp.data_ptr(),
                                                 None if wd is None else wd.data_ptr(), lib.stream_ptr(w.device))
        lib.check(rc, "ab_pack_conv_filters")
        return wp, wd

    return _cached(conv, ("packed", cin_pad, with_dgrad), _ver(w), build)


def bn_affine(bn, training: bool):
    """Folded (scale, bias) of an eval-mode / frozen BatchNorm (resnet.py:44-69, nn.BatchNorm2d eval)."""
    if bn is None:
        return None, None
    if training and isinstance(bn, nn.BatchNorm2d):
        raise NotImplementedError("training-mode BatchNorm needs batch statistics: run the model with gradients enabled "
                                  "(the training graph), or call .eval() / use FREEZE_BATCHNORM: true under torch.no_grad()")

    def build():
        eps = getattr(bn, "eps", 1e-5)
        scale = bn.weight.detach().float() * torch.rsqrt(bn.running_var.detach().float() + eps)
        bias = bn.bias.detach().float() - bn.running_mean.detach().float() * scale
        return scale.contiguous(), bias.contiguous()

    return _cached(bn, "bn", _ver(bn.weight, bn.bias, bn.running_mean, bn.running_var), build)


def image_to_act(image: torch.Tensor) -> Act:
    """f32 NCHW image -> bf16 NHWC with the channel count padded to a multiple of 4 (RGB -> 8 bytes per pixel)."""
    lib.require_cuda(image, "image")
    B, C, H, W = image.shape
    cp = (C + 3) // 4 * 4
    img = image.contiguous().float()
    out = torch.empty((B * H * W, cp), dtype=torch.bfloat16, device=image.device)
    with torch.cuda.device(image.device):
        rc = lib.load().ab_image_to_nhwc(img.data_ptr(), B, C, H, W, cp, out.data_ptr(), lib.stream_ptr(image.device))
    lib.check(rc, "ab_image_to_nhwc")
    return Act(out, B, H, W, cp)


def conv_bn_act(x: Act, conv: nn.Conv2d, bn=None, relu: bool = False, residual: Optional[Act] = None,
                training: bool = False, out_fp32: bool = False):
    """Conv2d (+ folded BN / conv bias) (+ residual) (+ ReLU).  -> Act (bf16) or fp32 [B*Ho*Wo, Cout] when out_fp32."""
    kh, kw = conv.kernel_size
    stride, pad = conv.stride[0], conv.padding[0]
    cout = conv.out_channels
    assert conv.stride[0] == conv.stride[1] and conv.padding[0] == conv.padding[1] and conv.groups == 1
    wp, _ = packed_filters(conv, x.C)
    scale, bias = bn_affine(bn, training)
    if conv.bias is not None:
        cb = conv.bias.detach().float()
        bias = cb if bias is None else bias + cb * scale
    Ho, Wo = (x.H + 2 * pad - kh) // stride + 1, (x.W + 2 * pad - kw) // stride + 1
    dev = x.data.device
    if IMPLICIT_CONV and x.C % 64 == 0 and not (kh == 1 and kw == 1 and stride == 1):
        # implicit GEMM: the activation is the A operand, gathered by TMA in im2col mode
        out = torch.empty((x.B * Ho * Wo, cout), dtype=torch.float32 if out_fp32 else torch.bfloat16, device=dev)
        p = lambda t: None if t is None else t.data_ptr()  # noqa: E731
        with torch.cuda.device(dev):
            rc = lib.load().ab_conv_bf16_nhwc(x.data.data_ptr(), x.B, x.H, x.W, x.C, wp.data_ptr(), cout, kh, kw, stride, pad,
                                              out.data_ptr(), cout, int(out_fp32), p(scale), p(bias),
                                              None if residual is None else residual.data.data_ptr(), cout, int(relu),
                                              None, None, lib.stream_ptr(dev))
        lib.check(rc, "ab_conv_bf16_nhwc")
        return out if out_fp32 else Act(out, x.B, Ho, Wo, cout)
    if kh == 1 and kw == 1 and stride == 1 and pad == 0 and x.C % 8 == 0:
        a = x.data
    else:
        kp = wp.shape[1]
        a = torch.empty((x.B * Ho * Wo, kp), dtype=torch.bfloat16, device=dev)
        with torch.cuda.device(dev):
            rc = lib.load().ab_im2col_nhwc(x.data.data_ptr(), x.B, x.H, x.W, x.C, kh, kw, stride, pad, kp, a.data_ptr(),
                                           lib.stream_ptr(dev))
        lib.check(rc, "ab_im2col_nhwc")
    out = ops.gemm_bf16(a, wp, scale=scale, bias=bias, residual=None if residual is None else residual.data, relu=relu,
                        out_fp32=out_fp32)
    return out if out_fp32 else Act(out, x.B, Ho, Wo, cout)


def maxpool3x3s2(x: Act, idx: Optional[torch.Tensor] = None) -> Act:
    """idx: optional uint8 [B*Ho*Wo, C] that receives the argmax tap of every output (for the backward kernel)."""
    Ho, Wo = (x.H + 2 - 3) // 2 + 1, (x.W + 2 - 3) // 2 + 1
    out = torch.empty((x.B * Ho * Wo, x.C), dtype=torch.bfloat16, device=x.data.device)
    with torch.cuda.device(x.data.device):
        rc = lib.load().ab_maxpool3x3s2_nhwc(x.data.data_ptr(), x.B, x.H, x.W, x.C, out.data_ptr(),
                                             None if idx is None else idx.data_ptr(), lib.stream_ptr(x.data.device))
    lib.check(rc, "ab_maxpool3x3s2_nhwc")
    return Act(out, x.B, Ho, Wo, x.C)


def avgpool(x: Act):
    """-> (fp32 [B, C], bf16 [B, C])"""
    dev = x.data.device
    f = torch.empty((x.B, x.C), dtype=torch.float32, device=dev)
    h = torch.empty((x.B, x.C), dtype=torch.bfloat16, device=dev)
    with torch.cuda.device(dev):
        rc = lib.load().ab_avgpool_nhwc(x.data.data_ptr(), x.B, x.H * x.W, x.C, f.data_ptr(), h.data_ptr(),
                                        lib.stream_ptr(dev))
    lib.check(rc, "ab_avgpool_nhwc")
    return f, h


def deconv4x4s2_bn_relu(x: Act, deconv: nn.ConvTranspose2d, bn, training: bool = False, relu: bool = True) -> Act:
    """ConvTranspose2d(k=4, s=2, p=1) + BN + ReLU (simplebaseline.py:152-175): X . W -> [B*H*W, 16*Cout], then the
    gather form of col2im with the BN affine and ReLU fused."""
    if deconv.kernel_size != (4, 4) or deconv.stride != (2, 2) or deconv.padding != (1, 1) or deconv.output_padding != (0, 0):
        raise NotImplementedError("only ConvTranspose2d(kernel 4, stride 2, padding 1) (NUM_DECONV_KERNELS: 4)")
    cout = deconv.out_channels

    def build():  # [Cin, Cout, ky, kx] -> [(ky, kx, co), ci]
        return deconv.weight.detach().permute(2, 3, 1, 0).reshape(16 * cout, -1).to(torch.bfloat16).contiguous()

    wp = _cached(deconv, "dw", _ver(deconv.weight), build)
    scale, bias = bn_affine(bn, training)
    if deconv.bias is not None:
        db = deconv.bias.detach().float()
        bias = db if bias is None else bias + db * scale
    ycol = ops.gemm_bf16(x.data, wp, out_fp32=True)
    dev = x.data.device
    out = torch.empty((x.B * 4 * x.H * x.W, cout), dtype=torch.bfloat16, device=dev)
    p = lambda t: None if t is None else t.data_ptr()  # noqa: E731
    with torch.cuda.device(dev):
        rc = lib.load().ab_deconv4x4s2_col2im(ycol.data_ptr(), x.B, x.H, x.W, cout, p(scale), p(bias), int(relu),
                                              out.data_ptr(), None, lib.stream_ptr(dev))
    lib.check(rc, "ab_deconv4x4s2_col2im")
    return Act(out, x.B, 2 * x.H, 2 * x.W, cout)


def linear(x_bf16: torch.Tensor, fc: nn.Linear, relu: bool = False, out_fp32: bool = False) -> torch.Tensor:
    """nn.Linear on the tensor-core GEMM; out features padded to a multiple of 8 internally."""
    n = fc.out_features
    npad = _pad8(n)

    def build():
        w = torch.zeros((npad, _pad8(fc.in_features)), dtype=torch.bfloat16, device=fc.weight.device)
        w[:n, :fc.in_features] = fc.weight.detach().to(torch.bfloat16)
        b = torch.zeros(npad, dtype=torch.float32, device=fc.weight.device)
        if fc.bias is not None:
            b[:n] = fc.bias.detach().float()
        return w, b

    w, b = _cached(fc, "fc", _ver(fc.weight, fc.bias), build)
    out = ops.gemm_bf16(x_bf16, w, bias=b, relu=relu, out_fp32=out_fp32)
    return out[:, :n]


def head_decode(logits_f32: torch.Tensor, B: int, ncls: int, D: int, H: int, W: int, with_lse: bool = False):
    """-> (kp3d, confd) or, with_lse, (kp3d, confd, lse): lse [B, ncls] (None when D % 4 != 0) lets the backward pass run in
    one sweep (ab_head_decode_bwd)."""
    dev = logits_f32.device
    kp3d = torch.empty((B, ncls, 3), dtype=torch.float32, device=dev)
    confd = torch.empty((B, ncls), dtype=torch.float32, device=dev)
    lse = torch.empty((B, ncls), dtype=torch.float32, device=dev) if with_lse and D % 4 == 0 else None
    with torch.cuda.device(dev):
        rc = lib.load().ab_head_decode(logits_f32.data_ptr(), B, ncls, D, H, W, kp3d.data_ptr(), confd.data_ptr(),
                                       None if lse is None else lse.data_ptr(), lib.stream_ptr(dev))
    lib.check(rc, "ab_head_decode")
    return (kp3d, confd, lse) if with_lse else (kp3d, confd)
