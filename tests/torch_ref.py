"""Plain torch.nn / autograd evaluation of the clasbased modules (the reference's own arithmetic: fp32 NCHW,
anakin/models/resnet.py:199-221, simplebaseline.py:177-190, mlp.py:24, hybridbaseline.py:37-96), run on the SAME parameter
containers as our model.  Test / baseline infrastructure only."""
import torch
import torch.nn.functional as F


def rb(x):
    """Round to bf16 and back.  autograd sends the gradient through the same two casts, so activation gradients are
    rounded to bf16 at the same points where our kernels store them as bf16."""
    return x.to(torch.bfloat16).float()


def backbone_head_bf16_points(model, image):
    """Same network, evaluated by torch in fp32 arithmetic but ROUNDED TO BF16 AT THE POINTS where the tensor-core path
    stores bf16: image, weights, raw conv outputs (before BatchNorm), layer outputs, MLP hidden activations.  With the
    rounding points matched, ReLU masks agree and gradients can be compared tightly."""
    hb = model.model_list[0]
    bb, head = hb.backbone, hb.hybrid_head

    def cbr(x, conv, bn, relu=True, res=None):
        raw = rb(F.conv2d(x, rb(conv.weight), None, conv.stride, conv.padding))
        o = bn(raw)
        if res is not None:
            o = o + res
        return rb(torch.relu(o) if relu else o)

    x = F.max_pool2d(cbr(rb(image), bb.conv1, bb.bn1), 3, 2, 1)
    for name in ("layer1", "layer2", "layer3", "layer4"):
        for blk in getattr(bb, name):
            r = x if blk.downsample is None else cbr(x, blk.downsample[0], blk.downsample[1], relu=False)
            if hasattr(blk, "conv3"):
                o = cbr(cbr(x, blk.conv1, blk.bn1), blk.conv2, blk.bn2)
                x = cbr(o, blk.conv3, blk.bn3, res=r)
            else:
                x = cbr(cbr(x, blk.conv1, blk.bn1), blk.conv2, blk.bn2, res=r)
    mean = x.mean(3).mean(2)
    h = x
    mods = list(head.deconv_layers)
    for i in range(0, len(mods), 3):
        raw = rb(F.conv_transpose2d(h, rb(mods[i].weight), None, 2, 1))
        h = rb(torch.relu(mods[i + 1](raw)))
    logits = F.conv2d(h, rb(head.final_layer.weight), head.final_layer.bias)
    p = F.softmax(logits.reshape(logits.shape[0], head.nclasses, -1), 2)
    confd = p.max(-1).values
    p = (p / (p.sum(-1, keepdim=True) + 1e-7)).view(p.shape[0], head.nclasses, head.depth_res, head.height_res, head.width_res)
    u = (p.sum(dim=[2, 3]) * (torch.arange(head.width_res, device=p.device) / head.width_res)).sum(-1)
    v = (p.sum(dim=[2, 4]) * (torch.arange(head.height_res, device=p.device) / head.height_res)).sum(-1)
    d = (p.sum(dim=[3, 4]) * (torch.arange(head.depth_res, device=p.device) / head.depth_res)).sum(-1)
    fcs = [m for m in hb.box_head.layers if isinstance(m, torch.nn.Linear)]
    z = rb(mean)
    for fc in fcs[:-1]:
        z = rb(torch.relu(F.linear(z, rb(fc.weight), fc.bias)))
    rot6d = F.linear(z, rb(fcs[-1].weight), fcs[-1].bias)
    return torch.stack([u, v, d], -1), confd, rot6d


def backbone_head(model, image):
    hb = model.model_list[0]
    bb, head = hb.backbone, hb.hybrid_head
    x = bb.maxpool(bb.relu(bb.bn1(bb.conv1(image))))
    for name in ("layer1", "layer2", "layer3", "layer4"):
        for blk in getattr(bb, name):
            r = x if blk.downsample is None else blk.downsample(x)
            if hasattr(blk, "conv3"):
                o = blk.relu(blk.bn1(blk.conv1(x)))
                o = blk.relu(blk.bn2(blk.conv2(o)))
                o = blk.bn3(blk.conv3(o))
            else:
                o = blk.relu(blk.bn1(blk.conv1(x)))
                o = blk.bn2(blk.conv2(o))
            x = blk.relu(o + r)
    mean = x.mean(3).mean(2)
    h = head.final_layer(head.deconv_layers(x))
    h = F.softmax(h.reshape(h.shape[0], head.nclasses, -1), 2)
    confd = h.max(-1).values
    h = (h / (h.sum(-1, keepdim=True) + 1e-7)).view(h.shape[0], head.nclasses, head.depth_res, head.height_res, head.width_res)
    u = (h.sum(dim=[2, 3]) * (torch.arange(head.width_res, device=h.device) / head.width_res)).sum(-1)
    v = (h.sum(dim=[2, 4]) * (torch.arange(head.height_res, device=h.device) / head.height_res)).sum(-1)
    d = (h.sum(dim=[3, 4]) * (torch.arange(head.depth_res, device=h.device) / head.depth_res)).sum(-1)
    rot6d = hb.box_head.layers(mean)
    return torch.stack([u, v, d], -1), confd, rot6d


def hybrid_forward(model, inputs, bf16_points=False):
    from artiboost_b200.models.transform import batch_uvd2xyz, compute_rotation_matrix_from_ortho6d
    hb = model.model_list[0]
    kp3d, confd, rot6d = (backbone_head_bf16_points if bf16_points else backbone_head)(model, inputs["image"])
    pose = batch_uvd2xyz(kp3d, inputs["root_joint"], inputs["cam_intr"], hb.inp_res)
    R = compute_rotation_matrix_from_ortho6d(rot6d)
    corners = torch.matmul(R, inputs["corners_can"].permute(0, 2, 1)).permute(0, 2, 1) + pose[:, 21:22]
    return {"joints_3d_abs": pose[:, :21], "corners_3d_abs": corners, "kp3d": kp3d, "box_rot_6d": rot6d}
