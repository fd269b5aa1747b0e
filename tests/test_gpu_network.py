"""GPU parity of the clasbased network path: tcgen05 GEMM and NHWC kernels vs plain PyTorch fp32 references of the same
ops, and the whole HybridBaseline vs outputs of the REFERENCE's own modules (tests/golden/network_*.npz, produced by
tests/golden/make_golden_network.py from /root/reference).

Tolerances: the tensor-core path multiplies bf16 operands (8-bit mantissa) and accumulates in fp32; activations are
stored as bf16 between layers.  Stated bounds: single GEMM / conv vs fp32 math on the SAME bf16-rounded operands
<= 1e-5 relative (fp32 output); end-to-end network vs the fp32 reference <= 3 mm on absolute 3-D keypoints / corners
(BASELINE.json: "network outputs within stated fp tolerance").
"""
import os
import sys

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import GOLDEN, golden

sys.path.insert(0, GOLDEN)
import netcfg  # noqa: E402

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module", autouse=True)
def _lib(lib_built):
    return lib_built


def bf(x):
    return x.to(torch.bfloat16)


# ----------------------------------------------------------------------------------------------------- GEMM
@pytest.mark.parametrize("M,N,K", [(128, 64, 64), (1, 8, 8), (333, 8, 152), (1000, 616, 256), (4096, 64, 576),
                                   (257, 136, 72), (64, 512, 4608), (8192, 256, 2304)])
def test_gemm_matches_fp32_matmul(M, N, K):
    from artiboost_b200 import ops
    g = torch.Generator(device=DEV).manual_seed(M + N + K)
    a = bf(torch.randn((M, K), device=DEV, generator=g) * 0.5)
    b = bf(torch.randn((N, K), device=DEV, generator=g) * 0.5)
    ref = a.float() @ b.float().T
    out = ops.gemm_bf16(a, b, out_fp32=True)
    torch.testing.assert_close(out, ref, rtol=1e-5, atol=1e-5 * ref.abs().max().item())
    out16 = ops.gemm_bf16(a, b)
    torch.testing.assert_close(out16.float(), ref, rtol=1e-2, atol=1e-2 * ref.abs().max().item())


def test_gemm_fused_epilogue_and_column_statistics():
    from artiboost_b200 import ops
    g = torch.Generator(device=DEV).manual_seed(7)
    M, N, K = 3000, 256, 1152
    a, b = bf(torch.randn((M, K), device=DEV, generator=g)), bf(torch.randn((N, K), device=DEV, generator=g) * 0.05)
    scale, bias = torch.rand(N, device=DEV, generator=g) + 0.5, torch.randn(N, device=DEV, generator=g)
    res = bf(torch.randn((M, N), device=DEV, generator=g))
    raw = a.float() @ b.float().T
    ref = torch.relu(raw * scale + bias + res.float())
    tiles = (M + 127) // 128   # per-row-tile partial statistics, garbage-initialised: every entry must be overwritten
    cs, cq = torch.full((tiles, N), float("nan"), device=DEV), torch.full((tiles, N), float("nan"), device=DEV)
    out = ops.gemm_bf16(a, b, scale=scale, bias=bias, residual=res, relu=True, out_fp32=True, col_stats=(cs, cq))
    torch.testing.assert_close(out, ref, rtol=1e-5, atol=1e-4)
    pad = torch.zeros((tiles * 128, N), device=DEV)
    pad[:M] = raw
    torch.testing.assert_close(cs, pad.view(tiles, 128, N).sum(1), rtol=1e-4, atol=1e-3)
    torch.testing.assert_close(cq, (pad * pad).view(tiles, 128, N).sum(1), rtol=1e-4, atol=1e-3)
    cs2, cq2 = torch.empty_like(cs), torch.empty_like(cq)
    ops.gemm_bf16(a, b, col_stats=(cs2, cq2))
    assert torch.equal(cs, cs2) and torch.equal(cq, cq2), "column statistics must be bit-reproducible (no atomics)"
    # strided operands / outputs (row pitch > row length)
    big = torch.zeros((M, N + 64), dtype=torch.bfloat16, device=DEV)
    ops.gemm_bf16(a[:, :576], b[:, :576], out=big[:, 8:8 + N])
    torch.testing.assert_close(big[:, 8:8 + N].float(), a[:, :576].float() @ b[:, :576].float().T, rtol=1e-2, atol=0.05)
    assert float(big[:, :8].abs().sum()) == 0 and float(big[:, 8 + N:].abs().sum()) == 0


def test_gemm_argument_errors():
    from artiboost_b200 import lib, ops
    a, b = torch.zeros((16, 12), dtype=torch.bfloat16, device=DEV), torch.zeros((8, 12), dtype=torch.bfloat16, device=DEV)
    with pytest.raises(lib.AbError):
        ops.gemm_bf16(a, b)  # K = 12 is not a multiple of 8
    out = ops.gemm_bf16(torch.zeros((0, 16), dtype=torch.bfloat16, device=DEV), torch.zeros((8, 16), dtype=torch.bfloat16, device=DEV))
    assert out.shape == (0, 8)


# ------------------------------------------------------------------------------------------------- primitives
def to_act(x_nchw):
    from artiboost_b200.models.nhwc import Act, image_to_act
    B, C, H, W = x_nchw.shape
    if C == 3:
        return image_to_act(x_nchw)  # the stem's path: padded to 4 channels
    return Act(bf(x_nchw).permute(0, 2, 3, 1).reshape(B * H * W, C).contiguous(), B, H, W, C)


@pytest.mark.parametrize("cin,cout,k,s,p,hw", [(3, 64, 7, 2, 3, 64), (3, 64, 7, 2, 3, 37), (40, 64, 3, 1, 1, 16), (64, 64, 3, 1, 1, 32), (64, 128, 3, 2, 1, 32),
                                               (64, 128, 1, 2, 0, 32), (256, 64, 1, 1, 0, 16), (128, 128, 3, 1, 1, 15),
                                               # the halo-resident 3x3 kernel (Cout <= 64): several tiles per image, ragged last
                                               # tile, windows of 5 ... 34 input rows, two channel blocks
                                               (64, 64, 3, 1, 1, 64), (64, 64, 3, 1, 1, 37), (128, 64, 3, 1, 1, 9), (256, 56, 3, 1, 1, 5),
                                               (64, 64, 3, 1, 1, 2)])
def test_conv_bn_relu_residual_matches_torch(cin, cout, k, s, p, hw):
    from artiboost_b200.models import nhwc
    torch.manual_seed(cin + cout + k)
    conv = torch.nn.Conv2d(cin, cout, k, s, p, bias=False).to(DEV)
    bn = torch.nn.BatchNorm2d(cout).to(DEV).eval()
    netcfg.randomise_bn(bn)
    x = torch.randn((3, cin, hw, hw), device=DEV)
    xb, wb = bf(x).float(), bf(conv.weight).float()
    raw = F.conv2d(xb, wb, None, s, p)
    res = torch.randn_like(raw)
    ref = torch.relu(F.batch_norm(raw, bn.running_mean, bn.running_var, bn.weight, bn.bias, False, 0.0, bn.eps) + bf(res).float())
    out = nhwc.conv_bn_act(to_act(x), conv, bn, relu=True, residual=to_act(res))
    assert (out.H, out.W, out.C) == (ref.shape[2], ref.shape[3], cout)
    torch.testing.assert_close(out.nchw(), ref, rtol=1e-2, atol=2e-2)  # bf16 output rounding
    # fp32 output, no BN: only accumulation-order differences remain
    out32 = nhwc.conv_bn_act(to_act(x), conv, None, out_fp32=True).view(3, out.H, out.W, cout).permute(0, 3, 1, 2)
    torch.testing.assert_close(out32, raw, rtol=1e-5, atol=1e-4)


def test_maxpool_avgpool_match_torch():
    from artiboost_b200.models import nhwc
    x = torch.randn((2, 64, 33, 32), device=DEV)
    ref = F.max_pool2d(bf(x).float(), 3, 2, 1)
    out = nhwc.maxpool3x3s2(to_act(x))
    torch.testing.assert_close(out.nchw(), ref, rtol=0, atol=0)
    f32, b16 = nhwc.avgpool(to_act(x))
    torch.testing.assert_close(f32, bf(x).float().mean(3).mean(2), rtol=1e-5, atol=1e-5)


def test_deconv_bn_relu_matches_torch():
    from artiboost_b200.models import nhwc
    torch.manual_seed(3)
    deconv = torch.nn.ConvTranspose2d(128, 64, 4, 2, 1, 0, bias=False).to(DEV)
    bn = torch.nn.BatchNorm2d(64).to(DEV).eval()
    netcfg.randomise_bn(bn)
    x = torch.randn((2, 128, 8, 8), device=DEV)
    raw = F.conv_transpose2d(bf(x).float(), bf(deconv.weight).float(), None, 2, 1)
    ref = torch.relu(F.batch_norm(raw, bn.running_mean, bn.running_var, bn.weight, bn.bias, False, 0.0, bn.eps))
    out = nhwc.deconv4x4s2_bn_relu(to_act(x), deconv, bn)
    assert (out.H, out.W, out.C) == (16, 16, 64)
    torch.testing.assert_close(out.nchw(), ref, rtol=1e-2, atol=2e-2)


def reference_head_decode(x, ncls, D, H, W):
    """anakin/models/simplebaseline.py:16-71,182-190 restated on a [B, ncls*D, H, W] logit tensor."""
    B = x.shape[0]
    x = x.reshape(B, ncls, -1)
    x = F.softmax(x, 2)
    confd = torch.max(x, dim=-1).values
    x = x / (x.sum(dim=-1, keepdim=True) + 1e-7)
    x = x.view(B, ncls, D, H, W)
    d_accu, v_accu, u_accu = x.sum(dim=[3, 4]), x.sum(dim=[2, 4]), x.sum(dim=[2, 3])
    wd, wv, wu = (torch.arange(n, dtype=x.dtype, device=x.device) / n for n in (D, H, W))
    uvd = torch.stack([(u_accu * wu).sum(-1), (v_accu * wv).sum(-1), (d_accu * wd).sum(-1)], dim=-1)
    return uvd, confd


def test_head_decode_matches_reference_formula_and_one_hot():
    from artiboost_b200.models import nhwc
    B, ncls, D, H, W = 3, 22, 28, 32, 32
    logits_nchw = 3.0 * torch.randn((B, ncls * D, H, W), device=DEV)
    ref_uvd, ref_confd = reference_head_decode(logits_nchw.double(), ncls, D, H, W)
    nhwc_logits = logits_nchw.permute(0, 2, 3, 1).reshape(B * H * W, ncls * D).contiguous()
    kp3d, confd = nhwc.head_decode(nhwc_logits, B, ncls, D, H, W)
    torch.testing.assert_close(kp3d.double(), ref_uvd, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(confd.double(), ref_confd, rtol=1e-4, atol=1e-7)
    # a one-hot heatmap integrates to the exact bin index / size
    hot = torch.full((1, ncls * D, H, W), -100.0, device=DEV)
    hot[0, 5 * D + 7, 11, 19] = 100.0
    kp, cf = nhwc.head_decode(hot.permute(0, 2, 3, 1).reshape(H * W, ncls * D).contiguous(), 1, ncls, D, H, W)
    torch.testing.assert_close(kp[0, 5], torch.tensor([19 / 32, 11 / 32, 7 / 28], device=DEV), rtol=1e-6, atol=1e-6)
    assert abs(float(cf[0, 5]) - 1.0) < 1e-6


@pytest.mark.parametrize("D", [28, 6])
def test_head_decode_backward_one_sweep_and_three_sweeps_match_autograd(D):
    """ab_head_decode_bwd with the forward's log-sum-exp (one sweep, D % 4 == 0) and without (three sweeps) vs torch autograd
    of the reference formula (fp64); D = 6 exercises the scalar kernels on both sides."""
    from artiboost_b200 import lib
    from artiboost_b200.models import nhwc
    B, ncls, H, W = 2, 22, 16, 16
    logits_nchw = (2.0 * torch.randn((B, ncls * D, H, W), device=DEV)).double().requires_grad_(True)
    ref_uvd, _ = reference_head_decode(logits_nchw, ncls, D, H, W)
    dk = torch.randn((B, ncls, 3), device=DEV)
    ref_uvd.backward(dk.double())
    ref_dl = logits_nchw.grad.permute(0, 2, 3, 1).reshape(B * H * W, ncls * D).float()
    lg = logits_nchw.detach().float().permute(0, 2, 3, 1).reshape(B * H * W, ncls * D).contiguous()
    kp3d, confd, lse = nhwc.head_decode(lg, B, ncls, D, H, W, with_lse=True)
    assert (lse is not None) == (D % 4 == 0)
    torch.testing.assert_close(kp3d.double(), ref_uvd.detach(), rtol=1e-4, atol=1e-5)
    L = lib.load()
    outs = []
    for use_lse in ([True, False] if lse is not None else [False]):
        dl = torch.empty(lg.shape, dtype=torch.bfloat16, device=DEV)
        lib.check(L.ab_head_decode_bwd(lg.data_ptr(), dk.data_ptr(), kp3d.data_ptr() if use_lse else None,
                                       lse.data_ptr() if use_lse else None, B, ncls, D, H, W, dl.data_ptr(), lib.stream_ptr(lg.device)),
                  "ab_head_decode_bwd")
        scale = float(ref_dl.abs().max())
        assert float((dl.float() - ref_dl).abs().max()) < 6e-3 * scale   # bf16 output
        outs.append(dl.float())
    if len(outs) == 2:
        assert float((outs[0] - outs[1]).abs().max()) < 6e-3 * float(ref_dl.abs().max())
    if lse is not None:  # lse is what it says
        ref_lse = torch.logsumexp(logits_nchw.detach().reshape(B, ncls, -1), dim=2)
        torch.testing.assert_close(lse.double(), ref_lse, rtol=1e-5, atol=1e-5)


# -------------------------------------------------------------------------------------------- whole network
def build(backbone):
    import artiboost_b200.models as M
    arch, preset = netcfg.arch_cfg(backbone)
    torch.manual_seed(netcfg.SEED)
    model = M.Arch({"ARCH": arch}, M.build_arch_model_list(arch, preset_cfg=preset)).eval()
    netcfg.randomise_bn(model)
    return model.to(DEV)


@pytest.mark.parametrize("backbone", ["ResNet34", "ResNet50"])
def test_hybridbaseline_matches_reference_modules(backbone):
    g = golden(f"network_{backbone.lower()}.npz")
    model = build(backbone)
    assert sum(p.numel() for p in model.parameters()) == int(g["n_params"])
    inp = {k: v.to(DEV) for k, v in netcfg.make_inputs(2).items()}
    with torch.no_grad():
        out = model(inp)["HybridBaseline"]
    for k in ("joints_3d_abs", "corners_3d_abs", "joints_3d", "corners_3d", "2d_uvd", "boxroot_3d_abs", "box_rot_rotmat"):
        assert tuple(out[k].shape) == g[k].shape, k
    graph_out = model(inp)["HybridBaseline"]  # same eval-mode arithmetic when autograd records the graph
    assert graph_out["joints_3d_abs"].requires_grad
    torch.testing.assert_close(graph_out["joints_3d_abs"].detach(), out["joints_3d_abs"], rtol=0, atol=1e-4)
    err_j = np.abs(out["joints_3d_abs"].cpu().numpy() - g["joints_3d_abs"]).max()
    err_c = np.abs(out["corners_3d_abs"].cpu().numpy() - g["corners_3d_abs"]).max()
    err_uvd = np.abs(out["2d_uvd"].cpu().numpy() - g["2d_uvd"]).max()
    err_R = np.abs(out["box_rot_rotmat"].cpu().numpy() - g["box_rot_rotmat"]).max()
    print(f"{backbone}: max |d joints| {err_j * 1e3:.3f} mm, |d corners| {err_c * 1e3:.3f} mm, |d uvd| {err_uvd:.2e}, |d R| {err_R:.2e}")
    assert err_j < 3e-3 and err_c < 3e-3, "3 mm bound on absolute keypoints / corners"
    assert err_uvd < 5e-3 and err_R < 3e-2
    # intermediate features follow the fp32 reference statistically and point-wise on the pooled feature
    with torch.no_grad():
        feats = model.model_list[0].backbone(image=inp["image"])
    l4m = feats["res_layer4_mean"].cpu().numpy()
    rel = np.abs(l4m - g["res_layer4_mean"]).max() / np.abs(g["res_layer4_mean"]).max()
    assert rel < 3e-2, rel
    assert feats["res_layer1"].shape[1:] == (64 if backbone == "ResNet34" else 256, 64, 64)
    np.testing.assert_allclose([feats["res_layer1"].mean().item(), feats["res_layer1"].std().item()], g["res_layer1_stat"], rtol=2e-2)


def test_registry_and_checkpoint_names_follow_the_reference():
    import artiboost_b200.models as M
    with pytest.raises(KeyError):
        M.build_backbone({"TYPE": "HRNet"})  # not in the reference either
    model = build("ResNet34")
    keys = list(model.state_dict().keys())
    for k in ("_model_list.0.backbone.conv1.weight", "_model_list.0.backbone.layer4.2.bn2.running_var",
              "_model_list.0.backbone.fc.weight", "_model_list.0.hybrid_head.deconv_layers.0.weight",
              "_model_list.0.hybrid_head.deconv_layers.4.running_mean", "_model_list.0.hybrid_head.final_layer.bias",
              "_model_list.0.box_head.layers.4.weight"):
        assert k in keys, k
    model.train()
    inp = {k: v.to(DEV) for k, v in netcfg.make_inputs(1).items()}
    with torch.no_grad(), pytest.raises(NotImplementedError):
        model(inp)  # train mode without a graph is refused loudly rather than silently using running statistics


@pytest.mark.parametrize("cout,cin,k,cin_pad", [(64, 64, 3, 64), (128, 64, 1, 64), (256, 128, 3, 128), (96, 32, 3, 32), (64, 3, 7, 4),
                                               (40, 24, 3, 24)])
def test_packed_filters_match_the_torch_packing(cout, cin, k, cin_pad):
    """ab_pack_conv_filters (tiled shared-memory transpose for channel counts that are multiples of 32 and 1 / 9 taps, the
    per-element kernel otherwise) vs the torch packing: forward matrix [Cout, Kp] in (ky, kx, ci) order and the
    data-gradient matrix [Cin, kh*kw*Cout] with reversed taps, bit for bit."""
    from artiboost_b200.models import nhwc
    torch.manual_seed(cout + cin + k)
    conv = torch.nn.Conv2d(cin, cout, k, 1, k // 2, bias=False).to(DEV)
    wp, wd = nhwc.packed_filters(conv, cin_pad, with_dgrad=True)
    ref = nhwc.pack_conv_weight(conv.weight.detach(), cin_pad=cin_pad)
    assert wp.shape == ref.shape and torch.equal(wp, ref)
    w = conv.weight.detach()
    ref_d = w.flip(2, 3).permute(1, 2, 3, 0).reshape(cin, k * k * cout).to(torch.bfloat16)   # [ci][(ky', kx', co)]
    assert wd.shape == ref_d.shape and torch.equal(wd, ref_d)


# ------------------------------------------------------------------------- persistent form (more than 2 tiles per SM)
def test_persistent_gemm_matches_fp32_matmul_ragged_and_fused():
    """Grids beyond four tiles per SM take gemm_bf16_tn_persistent_kernel (static tile scheduler, double-buffered TMEM
    accumulators, two epilogue groups, 32-byte row stores when the rows are 32-byte aligned).  Ragged M / N / K, every
    epilogue option, per-tile column statistics, aligned and unaligned output pitches."""
    from artiboost_b200 import ops
    g = torch.Generator(device=DEV).manual_seed(11)
    M, N, K = 80003, 136, 72          # 626 x 2 tiles of 128 x 128, 2 k-blocks (the second 8 columns wide)
    a, b = bf(torch.randn((M, K), device=DEV, generator=g) * 0.5), bf(torch.randn((N, K), device=DEV, generator=g) * 0.5)
    raw = a.float() @ b.float().T
    out = ops.gemm_bf16(a, b, out_fp32=True)
    torch.testing.assert_close(out, raw, rtol=1e-5, atol=1e-5 * raw.abs().max().item())
    scale, bias = torch.rand(N, device=DEV, generator=g) + 0.5, torch.randn(N, device=DEV, generator=g)
    res = bf(torch.randn((M, N), device=DEV, generator=g))
    ref = torch.relu(raw * scale + bias + res.float())
    tiles = (M + 127) // 128
    cs, cq = torch.full((tiles, N), float("nan"), device=DEV), torch.full((tiles, N), float("nan"), device=DEV)
    out16 = ops.gemm_bf16(a, b, scale=scale, bias=bias, residual=res, relu=True, col_stats=(cs, cq))
    torch.testing.assert_close(out16.float(), ref, rtol=1e-2, atol=1e-2 * ref.abs().max().item())
    pad = torch.zeros((tiles * 128, N), device=DEV)
    pad[:M] = raw
    torch.testing.assert_close(cs, pad.view(tiles, 128, N).sum(1), rtol=1e-4, atol=1e-3)
    torch.testing.assert_close(cq, (pad * pad).view(tiles, 128, N).sum(1), rtol=1e-4, atol=1e-3)
    cs2, cq2 = torch.empty_like(cs), torch.empty_like(cq)
    ops.gemm_bf16(a, b, col_stats=(cs2, cq2))
    assert torch.equal(cs, cs2) and torch.equal(cq, cq2)
    # N = 64 tiles (BN = 64, four stages), 16-column fragments all inside N: the 32-byte store path in bf16 and fp32 ...
    b64 = bf(torch.randn((64, K), device=DEV, generator=g) * 0.5)
    ref64 = a.float() @ b64.float().T
    torch.testing.assert_close(ops.gemm_bf16(a, b64, out_fp32=True), ref64, rtol=1e-5, atol=1e-5 * ref64.abs().max().item())
    torch.testing.assert_close(ops.gemm_bf16(a, b64).float(), ref64, rtol=1e-2, atol=1e-2 * ref64.abs().max().item())
    # ... and an output whose pitch is NOT a multiple of 32 bytes (16-byte stores), framed by guard columns
    big = torch.zeros((M, 64 + 24), dtype=torch.bfloat16, device=DEV)
    ops.gemm_bf16(a, b64, out=big[:, 8:72])
    torch.testing.assert_close(big[:, 8:72].float(), ref64, rtol=1e-2, atol=1e-2 * ref64.abs().max().item())
    assert float(big[:, :8].abs().sum()) == 0 and float(big[:, 72:].abs().sum()) == 0


@pytest.mark.parametrize("cin,cout,k,s,hw,B", [(128, 128, 3, 1, 32, 80), (64, 128, 3, 2, 64, 80), (64, 256, 1, 1, 32, 40)])
def test_persistent_implicit_gemm_convolution_matches_torch(cin, cout, k, s, hw, B):
    """The same kernel in im2col mode (and the 1x1 plain-GEMM route) at tile counts above 4 x 148, against F.conv2d on the
    bf16-rounded operands: fp32 output (accumulation order only) and the fused BatchNorm + residual + ReLU bf16 output."""
    from artiboost_b200.models import nhwc
    torch.manual_seed(cin + cout + k + s)
    conv = torch.nn.Conv2d(cin, cout, k, s, k // 2, bias=False).to(DEV)
    bn = torch.nn.BatchNorm2d(cout).to(DEV).eval()
    netcfg.randomise_bn(bn)
    x = torch.randn((B, cin, hw, hw), device=DEV)
    raw = F.conv2d(bf(x).float(), bf(conv.weight).float(), None, s, k // 2)
    assert (raw.shape[0] * raw.shape[2] * raw.shape[3] + 127) // 128 * ((cout + 127) // 128) > 4 * 148
    out32 = nhwc.conv_bn_act(to_act(x), conv, None, out_fp32=True).view(B, raw.shape[2], raw.shape[3], cout).permute(0, 3, 1, 2)
    torch.testing.assert_close(out32, raw, rtol=1e-5, atol=2e-4)
    res = torch.randn_like(raw)
    ref = torch.relu(F.batch_norm(raw, bn.running_mean, bn.running_var, bn.weight, bn.bias, False, 0.0, bn.eps) + bf(res).float())
    out = nhwc.conv_bn_act(to_act(x), conv, bn, relu=True, residual=to_act(res))
    torch.testing.assert_close(out.nchw(), ref, rtol=1e-2, atol=2e-2)
