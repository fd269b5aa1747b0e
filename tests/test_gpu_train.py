"""GPU parity of the TRAINING path (forward with batch statistics + backward kernels) vs torch autograd in fp32 on the same
bf16-rounded operands.

The tensor-core path stores activations, raw conv outputs and activation gradients in bf16 (relative step 2^-8), like
any bf16 mixed-precision training.  The torch reference is therefore evaluated in fp32 arithmetic but ROUNDED TO BF16 AT
THE SAME POINTS (tests/torch_ref.py: `rb`, whose backward rounds the gradient at the same place): with matched rounding
points ReLU masks agree and a layer's outputs and gradients agree to <= 3e-2 relative Frobenius error (measured 1e-3 ..
2e-2).  Against PURE fp32 autograd the whole-model loss agrees to 2 %, but gradient correlation decays with depth for
torch's own bf16-point evaluation exactly as for ours (BatchNorm backward subtracts the mean and x-hat components of dy,
which amplifies the bf16 rounding of dy); the whole-model test asserts that we are never further from the fp32 gradient
than torch's bf16-point evaluation is."""
import sys

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import GOLDEN

sys.path.insert(0, GOLDEN)
import netcfg  # noqa: E402
import torch_ref  # noqa: E402

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module", autouse=True)
def _lib(lib_built):
    return lib_built


def bf(x):
    return x.to(torch.bfloat16)


def close(a, b, tol=2e-2, name=""):
    """Relative Frobenius error.  (A max-norm bound is not meaningful for gradients: a ReLU mask can flip where the
    pre-activation is within bf16 rounding of zero, which moves a handful of elements by O(1) on both sides.)"""
    a, b = a.float(), b.float()
    err = (a - b).norm().item() / (b.norm().item() + 1e-30)
    assert err < tol, f"{name}: relative error {err:.3e} (|ref| {b.norm().item():.3e})"


def cosine(a, b):
    a, b = a.float().flatten(), b.float().flatten()
    return float(torch.dot(a, b) / (a.norm() * b.norm() + 1e-30))


def act_of(x_nchw, requires_grad=False):
    from artiboost_b200.models.nhwc import Act
    B, C, H, W = x_nchw.shape
    d = bf(x_nchw).permute(0, 2, 3, 1).reshape(B * H * W, C).contiguous()
    d.requires_grad_(requires_grad)
    return Act(d, B, H, W, C)


def nchw(mat, B, H, W):
    return mat.float().view(B, H, W, -1).permute(0, 3, 1, 2)


@pytest.mark.parametrize("cin,cout,k,s,p,hw,res", [(64, 64, 3, 1, 1, 16, True), (64, 128, 3, 2, 1, 16, False),
                                                   (64, 128, 1, 2, 0, 16, False), (128, 64, 1, 1, 0, 8, False),
                                                   (256, 256, 3, 1, 1, 8, True)])
def test_conv_bn_relu_train_forward_backward(cin, cout, k, s, p, hw, res):
    from artiboost_b200.models import train_ops
    torch.manual_seed(cin * 7 + cout + k + s)
    B = 4
    conv = torch.nn.Conv2d(cin, cout, k, s, p, bias=False).to(DEV)
    bn = torch.nn.BatchNorm2d(cout).to(DEV).train()
    with torch.no_grad():
        bn.weight.copy_(1 + 0.2 * torch.randn(cout, device=DEV))
        bn.bias.copy_(0.1 * torch.randn(cout, device=DEV))
        conv.weight.copy_(bf(conv.weight).float())
    x = bf(torch.randn((B, cin, hw, hw), device=DEV)).float()
    ho = (hw + 2 * p - k) // s + 1
    r = bf(torch.randn((B, cout, ho, ho), device=DEV)).float() if res else None
    dy = bf(torch.randn((B, cout, ho, ho), device=DEV)).float()
    # reference: torch autograd, fp32
    xr = x.clone().requires_grad_(True)
    rr = r.clone().requires_grad_(True) if res else None
    bn_ref = torch.nn.BatchNorm2d(cout).to(DEV).train()
    bn_ref.load_state_dict(bn.state_dict())
    wref = conv.weight.detach().clone().requires_grad_(True)
    raw = torch_ref.rb(F.conv2d(xr, wref, None, s, p))
    o = bn_ref(raw)
    yr = torch_ref.rb(torch.relu(o + rr if res else o))
    yr.backward(dy)
    # ours
    xa = act_of(x, True)
    ra = act_of(r, True) if res else None
    y = train_ops.conv_bn_act(xa, conv, bn, relu=True, residual=ra, training=True)
    y.data.backward(bf(dy).permute(0, 2, 3, 1).reshape(-1, cout))
    close(nchw(y.data.detach(), B, ho, ho), yr, 2e-2, "y")
    close(nchw(xa.data.grad, B, hw, hw), xr.grad, 3e-2, "dx")
    close(conv.weight.grad, wref.grad, 3e-2, "dw")
    close(bn.weight.grad, bn_ref.weight.grad, 3e-2, "dgamma")
    close(bn.bias.grad, bn_ref.bias.grad, 3e-2, "dbeta")
    assert cosine(conv.weight.grad, wref.grad) > 0.998 and cosine(nchw(xa.data.grad, B, hw, hw), xr.grad) > 0.998
    if res:
        close(nchw(ra.data.grad, B, ho, ho), rr.grad, 3e-2, "dres")
    close(bn.running_mean, bn_ref.running_mean, 1e-2, "running_mean")
    close(bn.running_var, bn_ref.running_var, 1e-2, "running_var")
    assert int(bn.num_batches_tracked) == 1


def test_stem_maxpool_train():
    from artiboost_b200.models import nhwc, train_ops
    torch.manual_seed(5)
    B, hw = 2, 32
    conv = torch.nn.Conv2d(3, 64, 7, 2, 3, bias=False).to(DEV)
    bn = torch.nn.BatchNorm2d(64).to(DEV).train()
    with torch.no_grad():
        conv.weight.copy_(bf(conv.weight).float())
    img = bf(torch.rand((B, 3, hw, hw), device=DEV) - 0.5).float()
    bn_ref = torch.nn.BatchNorm2d(64).to(DEV).train()
    wref = conv.weight.detach().clone().requires_grad_(True)
    yr = F.max_pool2d(torch_ref.rb(torch.relu(bn_ref(torch_ref.rb(F.conv2d(img, wref, None, 2, 3))))), 3, 2, 1)
    dy = bf(torch.randn_like(yr)).float()
    yr.backward(dy)
    x = nhwc.image_to_act(img)
    y = train_ops.maxpool3x3s2(train_ops.conv_bn_act(x, conv, bn, relu=True, training=True))
    y.data.backward(bf(dy).permute(0, 2, 3, 1).reshape(-1, 64))
    close(nchw(y.data.detach(), B, 8, 8), yr, 2e-2, "y")
    close(conv.weight.grad, wref.grad, 3e-2, "dw stem")
    assert cosine(conv.weight.grad, wref.grad) > 0.998


def test_deconv_head_decode_linear_train():
    from artiboost_b200.models import train_ops
    torch.manual_seed(9)
    B, cin, cout, hw = 2, 128, 64, 8
    deconv = torch.nn.ConvTranspose2d(cin, cout, 4, 2, 1, bias=False).to(DEV)
    bn = torch.nn.BatchNorm2d(cout).to(DEV).train()
    final = torch.nn.Conv2d(cout, 22 * 4, 1).to(DEV)
    with torch.no_grad():
        deconv.weight.copy_(bf(deconv.weight).float())
        final.weight.copy_(bf(final.weight * 4).float())
    x = bf(torch.randn((B, cin, hw, hw), device=DEV)).float()
    # reference
    xr = x.clone().requires_grad_(True)
    bn_ref = torch.nn.BatchNorm2d(cout).to(DEV).train()
    wd = deconv.weight.detach().clone().requires_grad_(True)
    wf, bfin = final.weight.detach().clone().requires_grad_(True), final.bias.detach().clone().requires_grad_(True)
    h = torch_ref.rb(torch.relu(bn_ref(torch_ref.rb(F.conv_transpose2d(xr, wd, None, 2, 1)))))
    logits = F.conv2d(h, wf, bfin)
    p = F.softmax(logits.reshape(B, 22, -1), 2)
    p = (p / (p.sum(-1, keepdim=True) + 1e-7)).view(B, 22, 4, 16, 16)
    u = (p.sum(dim=[2, 3]) * (torch.arange(16, device=DEV) / 16)).sum(-1)
    v = (p.sum(dim=[2, 4]) * (torch.arange(16, device=DEV) / 16)).sum(-1)
    d = (p.sum(dim=[3, 4]) * (torch.arange(4, device=DEV) / 4)).sum(-1)
    kr = torch.stack([u, v, d], -1)
    dk = torch.randn_like(kr)
    kr.backward(dk)
    # ours
    xa = act_of(x, True)
    ha = train_ops.deconv4x4s2_bn_relu(xa, deconv, bn, training=True)
    lg = train_ops.conv_bn_act(ha, final, None, relu=False, out_fp32=True)
    kp, confd = train_ops.head_decode(lg, B, 22, 4, 16, 16)
    kp.backward(dk)
    close(kp.detach(), kr, 2e-2, "kp3d")
    close(nchw(xa.data.grad, B, hw, hw), xr.grad, 3e-2, "dx")
    close(deconv.weight.grad, wd.grad, 3e-2, "dw deconv")
    close(final.weight.grad, wf.grad, 3e-2, "dw final")
    close(final.bias.grad, bfin.grad, 3e-2, "db final")
    assert cosine(deconv.weight.grad, wd.grad) > 0.998 and cosine(final.weight.grad, wf.grad) > 0.998
    # linear
    fc = torch.nn.Linear(512, 256).to(DEV)
    fc2 = torch.nn.Linear(256, 6).to(DEV)
    with torch.no_grad():
        fc.weight.copy_(bf(fc.weight).float())
        fc2.weight.copy_(bf(fc2.weight).float())
    z = bf(torch.randn((16, 512), device=DEV)).float()
    zr = z.clone().requires_grad_(True)
    w1, b1 = fc.weight.detach().clone().requires_grad_(True), fc.bias.detach().clone().requires_grad_(True)
    w2, b2 = fc2.weight.detach().clone().requires_grad_(True), fc2.bias.detach().clone().requires_grad_(True)
    outr = F.linear(torch.relu(F.linear(zr, w1, b1)), w2, b2)
    g = torch.randn_like(outr)
    outr.backward(g)
    za = z.clone().requires_grad_(True)
    out = train_ops.linear(train_ops.linear(za, fc, relu=True), fc2, relu=False, out_fp32=True)
    out.backward(g)
    close(out.detach(), outr, 2e-2, "linear out")
    close(fc.weight.grad, w1.grad, 3e-2, "dw1"); close(fc2.weight.grad, w2.grad, 3e-2, "dw2")
    close(fc.bias.grad, b1.grad, 3e-2, "db1"); close(fc2.bias.grad, b2.grad, 3e-2, "db2")
    close(za.grad, zr.grad, 3e-2, "dz")


@pytest.mark.parametrize("backbone", ["ResNet34", "ResNet50"])
def test_whole_model_gradients_follow_torch_autograd(backbone):
    import copy
    import artiboost_b200.models as M
    arch, preset = netcfg.arch_cfg(backbone, size=128)
    torch.manual_seed(netcfg.SEED)
    model = M.Arch({"ARCH": arch}, M.build_arch_model_list(arch, preset_cfg=preset)).to(DEV).train()
    ref = copy.deepcopy(model)
    inp = {k: v.to(DEV) for k, v in netcfg.make_inputs(8, size=128).items()}
    tgt_j = torch.randn((8, 21, 3), device=DEV) * 0.05 + torch.tensor([0, 0, 0.5], device=DEV)
    tgt_c = torch.randn((8, 8, 3), device=DEV) * 0.05 + torch.tensor([0, 0, 0.5], device=DEV)

    def loss_of(out):
        return F.mse_loss(out["joints_3d_abs"], tgt_j) + F.mse_loss(out["corners_3d_abs"], tgt_c)

    out = model(inp)["HybridBaseline"]
    loss = loss_of(out)
    loss.backward()
    # references: torch autograd with matched bf16 rounding points, and pure fp32
    ref32 = copy.deepcopy(ref)
    loss_r = loss_of(torch_ref.hybrid_forward(ref, inp, bf16_points=True))
    loss_r.backward()
    loss32 = loss_of(torch_ref.hybrid_forward(ref32, inp))
    loss32.backward()
    assert abs(loss.item() - loss_r.item()) / loss_r.item() < 5e-3     # forward agrees with the matched-point evaluation
    assert abs(loss.item() - loss32.item()) / loss32.item() < 3e-2     # and with fp32
    n, worst_gap, worst_head, torch_head = 0, (0.0, ""), 1.0, 1.0
    ours, theirs = [], []
    for (k, p), (_, q), (_, r) in zip(model.named_parameters(), ref32.named_parameters(), ref.named_parameters()):
        if q.grad is None:
            assert p.grad is None or float(p.grad.abs().sum()) == 0.0, k  # the unused ImageNet fc head
            continue
        assert p.grad is not None and torch.isfinite(p.grad).all(), k
        ours.append(p.grad.flatten()), theirs.append(q.grad.flatten())
        if p.dim() < 2:
            continue
        n += 1
        c_ours, c_torch = cosine(p.grad, q.grad), cosine(r.grad, q.grad)   # both measured against the fp32 gradient
        if c_torch - c_ours > worst_gap[0]:
            worst_gap = (c_torch - c_ours, k)
        if "final_layer" in k or "box_head.layers.4" in k:
            worst_head, torch_head = min(worst_head, c_ours), min(torch_head, c_torch)
    norm_ratio = float(torch.cat(ours).norm() / torch.cat(theirs).norm())
    print(f"{backbone}: loss ours {loss.item():.6f} / bf16-point torch {loss_r.item():.6f} / fp32 {loss32.item():.6f}; "
          f"{n} weight tensors, gradient-norm ratio to fp32 {norm_ratio:.4f}, output-layer cosine {worst_head:.4f}, "
          f"largest deficit vs torch's bf16-point gradient {worst_gap}")
    assert n > 40 and abs(norm_ratio - 1) < 0.05
    # never meaningfully further from the fp32 gradient than torch's own bf16-point evaluation (both are noisy estimates
    # of it at this tiny batch: the 0.06 margin is the observed run-to-run spread between two such estimates)
    assert worst_head > torch_head - 0.06 and worst_head > 0.85
    assert worst_gap[0] < 0.06, worst_gap
    # running statistics were updated like torch's
    bn, bn_r = model.model_list[0].backbone.bn1, ref32.model_list[0].backbone.bn1
    close(bn.running_mean, bn_r.running_mean, 2e-2, "stem running_mean")
    close(bn.running_var, bn_r.running_var, 2e-2, "stem running_var")


@pytest.mark.parametrize("M,C,f32", [(1, 8, False), (777, 64, False), (70001, 64, True), (4099, 256, False), (300, 2304, True)])
def test_column_statistics_are_exact_sums_and_reproducible(M, C, f32):
    """ab_col_stats (deterministic partial rows + second pass) vs fp64 sums; ragged M, C above the CTA width."""
    from artiboost_b200 import lib
    from artiboost_b200.models.train_ops import _stat_ws
    g = torch.Generator(device=DEV).manual_seed(M + C)
    x = torch.randn((M, C), device=DEV, generator=g)
    x = x if f32 else bf(x)
    L = lib.load()

    def run():
        s, q = torch.full((C,), float("nan"), device=DEV), torch.full((C,), float("nan"), device=DEV)
        lib.check(L.ab_col_stats(x.data_ptr(), int(f32), M, C, C, s.data_ptr(), q.data_ptr(), _stat_ws(C, DEV).data_ptr(),
                                 lib.stream_ptr(DEV)), "ab_col_stats")
        return s, q

    s, q = run()
    xd = x.double()
    torch.testing.assert_close(s.double(), xd.sum(0), rtol=1e-5, atol=1e-3)
    torch.testing.assert_close(q.double(), (xd * xd).sum(0), rtol=1e-5, atol=1e-3)
    s2, q2 = run()
    assert torch.equal(s, s2) and torch.equal(q, q2)


def test_maxpool_argmax_index_and_backward_match_torch():
    """First-maximum tie-break (bf16 activations tie often after ReLU): dx equals torch's max_pool2d backward."""
    from artiboost_b200.models import train_ops
    g = torch.Generator(device=DEV).manual_seed(3)
    x = torch.relu(torch.randn((3, 16, 37, 41), device=DEV, generator=g)).mul(4).round().div(4)  # many exact ties
    a = act_of(x, requires_grad=True)
    y = train_ops.maxpool3x3s2(a)
    xr = bf(x).float().requires_grad_(True)
    yr = F.max_pool2d(xr, 3, 2, 1)
    torch.testing.assert_close(nchw(y.data, 3, y.H, y.W), yr, rtol=0, atol=0)
    dy = bf(torch.randn(yr.shape, device=DEV, generator=g))
    y.data.backward(dy.permute(0, 2, 3, 1).reshape(-1, 16).contiguous())
    yr.backward(dy.float())
    close(nchw(a.data.grad, 3, 37, 41), xr.grad, tol=4e-3, name="maxpool dx")


def test_fused_stem_equals_conv_bn_relu_then_maxpool():
    """StemFn (bn1 + ReLU evaluated inside the pooling kernel, the normalised activation never stored) against ConvBNActFn +
    MaxPoolFn on the same input: pooled output, weight / gamma / beta gradients and the running statistics, bit for bit."""
    import copy
    from artiboost_b200.models import nhwc, train_ops
    torch.manual_seed(5)
    conv = torch.nn.Conv2d(3, 64, 7, 2, 3, bias=False).to(DEV)
    bn_a = torch.nn.BatchNorm2d(64).to(DEV).train()
    with torch.no_grad():
        bn_a.weight.uniform_(0.5, 1.5)
        bn_a.bias.normal_(0, 0.3)
    bn_b = copy.deepcopy(bn_a)
    image = torch.randn((6, 3, 70, 66), device=DEV)
    g = bf(torch.randn((6 * 18 * 17, 64), device=DEV))
    assert train_ops.stem_fusable(conv, bn_a)

    def run(bn, fused):
        conv.weight.grad = None
        x = nhwc.image_to_act(image)
        if fused:
            y = train_ops.stem_conv_bn_relu_maxpool(x, conv, bn)
        else:
            y = train_ops.maxpool3x3s2(train_ops.conv_bn_act(x, conv, bn, relu=True, training=True))
        assert (y.H, y.W, y.C) == (18, 17, 64)
        y.data.backward(g)
        return y.data.detach().clone(), conv.weight.grad.clone(), bn.weight.grad.clone(), bn.bias.grad.clone()

    ya, dwa, dga, dba = run(bn_a, False)
    yb, dwb, dgb, dbb = run(bn_b, True)
    assert torch.equal(ya, yb)
    assert torch.equal(dwa, dwb) and torch.equal(dga, dgb) and torch.equal(dba, dbb)
    assert torch.equal(bn_a.running_mean, bn_b.running_mean) and torch.equal(bn_a.running_var, bn_b.running_var)
    assert int(bn_a.num_batches_tracked) == int(bn_b.num_batches_tracked) == 1
