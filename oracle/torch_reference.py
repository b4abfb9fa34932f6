"""Plain PyTorch fp32 re-declaration of the reference's dense graph (TEST INFRASTRUCTURE ONLY).

The reference package itself cannot be imported (hard dependency on mmcv==0.6.2 / timm / pycocotools,
SURVEY.md 8c), so the layer graph is re-declared here from standard torch layers, following
  mmdet/models/backbones/resnet.py:203-239, 495-506   (Bottleneck, style='pytorch', BN eval)
  mmdet/models/necks/fpn.py:138-178                   (start_level=1, extra convs on inputs, GN, no act)
  mmdet/models/anchor_heads/orientedreppoints_head.py:148-171 (forward_single), :673-779 (get_bboxes)
  mmdet/ops/dcn/src/deform_conv_cuda_kernel.cu:84-115, 190-243 (deformable im2col) + the GEMM of
  deform_conv_cuda.cpp:231-236
  mmdet/core/post_processing/bbox_nms.py:93-182       (multiclass_rnms, with the class-offset trick)
PINNED (the reference has no tests of its own for these layers, so its code is run instead):
  * backbone / FPN / head graph: equal to 1e-15 (float64) to the reference's OWN ResNet / FPN / ConvModule /
    OrientedRepPointsHead modules imported from /root/reference with the mmcv plumbing stubbed
    (tests/golden/gen_golden_dense.py -> dense_ref.npz, tests/test_oracle_golden.py);
  * get_bboxes_single + multiclass_rnms: bit-identical detections / labels / order to the reference's OWN python functions
    run with the two CUDA ops replaced by their oracles (gen_golden_postprocess.py -> postprocess.npz);
  * deform_conv_ref: equal (1e-15, float64) to weight x the columns of the reference's OWN (modulated_)deformable_im2col
    kernels compiled as host C++ (oracle/build_ref.py, device_ops_ref.npz), and to torchvision.ops.deform_conv2d.
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import pyoracle as po


def deform_conv_ref(x, offset, weight, stride=1, padding=1, dilation=1, mask=None):
    """DCNv1/v2 forward, deformable_groups = groups = 1.  x [B,C,H,W], offset [B,2*KH*KW,Ho,Wo]
    (channel 2t = dy, 2t+1 = dx), weight [Cout,C,KH,KW] -> [B,Cout,Ho,Wo]."""
    B, C, H, W = x.shape
    Cout, _, KH, KW = weight.shape
    Ho = (H + 2 * padding - dilation * (KH - 1) - 1) // stride + 1
    Wo = (W + 2 * padding - dilation * (KW - 1) - 1) // stride + 1
    dev, dt = x.device, x.dtype
    hs = torch.arange(Ho, device=dev, dtype=dt).view(1, Ho, 1) * stride - padding
    ws = torch.arange(Wo, device=dev, dtype=dt).view(1, 1, Wo) * stride - padding
    cols = []
    xf = x.reshape(B, C, H * W)
    for i in range(KH):
        for j in range(KW):
            t = i * KW + j
            h_im = hs + i * dilation + offset[:, 2 * t]
            w_im = ws + j * dilation + offset[:, 2 * t + 1]
            valid = (h_im > -1) & (w_im > -1) & (h_im < H) & (w_im < W)
            h_low = torch.floor(h_im)
            w_low = torch.floor(w_im)
            lh, lw = h_im - h_low, w_im - w_low
            hh, hw = 1 - lh, 1 - lw
            h_low, w_low = h_low.long(), w_low.long()
            h_high, w_high = h_low + 1, w_low + 1

            def corner(hi, wi, ok):
                ok = ok & valid
                idx = (hi.clamp(0, H - 1) * W + wi.clamp(0, W - 1)).view(B, 1, Ho * Wo).expand(B, C, Ho * Wo)
                v = torch.gather(xf, 2, idx).view(B, C, Ho, Wo)
                return v * ok.view(B, 1, Ho, Wo).to(dt)

            v1 = corner(h_low, w_low, (h_low >= 0) & (w_low >= 0))
            v2 = corner(h_low, w_high, (h_low >= 0) & (w_high <= W - 1))
            v3 = corner(h_high, w_low, (h_high <= H - 1) & (w_low >= 0))
            v4 = corner(h_high, w_high, (h_high <= H - 1) & (w_high <= W - 1))
            w1, w2, w3, w4 = (hh * hw).unsqueeze(1), (hh * lw).unsqueeze(1), (lh * hw).unsqueeze(1), (lh * lw).unsqueeze(1)
            val = w1 * v1 + w2 * v2 + w3 * v3 + w4 * v4
            if mask is not None:
                val = val * mask[:, t].unsqueeze(1)
            cols.append(val)
    col = torch.stack(cols, 2)                                   # [B, C, KH*KW, Ho, Wo]
    col = col.reshape(B, C * KH * KW, Ho * Wo)
    out = torch.matmul(weight.reshape(Cout, C * KH * KW), col)   # row = c*9 + i*3 + j, as the reference's columns
    return out.view(B, Cout, Ho, Wo)


def _bn(x, sd, p):
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"],
                        False, 0.0, 1e-5)


def _gn(x, sd, p):
    return F.group_norm(x, 32, sd[p + ".weight"], sd[p + ".bias"], 1e-5)


def backbone(sd, img, blocks=(3, 4, 6, 3)):
    x = F.relu(_bn(F.conv2d(img, sd["backbone.conv1.weight"], None, 2, 3), sd, "backbone.bn1"))
    x = F.max_pool2d(x, 3, 2, 1)
    outs = []
    for li, nblk in enumerate(blocks):
        for b in range(nblk):
            p = "backbone.layer%d.%d" % (li + 1, b)
            s = 2 if (b == 0 and li > 0) else 1
            idt = x
            o = F.relu(_bn(F.conv2d(x, sd[p + ".conv1.weight"]), sd, p + ".bn1"))
            o = F.relu(_bn(F.conv2d(o, sd[p + ".conv2.weight"], None, s, 1), sd, p + ".bn2"))
            o = _bn(F.conv2d(o, sd[p + ".conv3.weight"]), sd, p + ".bn3")
            if b == 0:
                idt = _bn(F.conv2d(x, sd[p + ".downsample.0.weight"], None, s), sd, p + ".downsample.1")
            x = F.relu(o + idt)
        outs.append(x)
    return outs


def fpn(sd, feats):
    c3, c4, c5 = feats[1], feats[2], feats[3]
    lat = [_gn(F.conv2d(c, sd["neck.lateral_convs.%d.conv.weight" % i]), sd, "neck.lateral_convs.%d.gn" % i)
           for i, c in enumerate((c3, c4, c5))]
    for i in (2, 1):
        lat[i - 1] = lat[i - 1] + F.interpolate(lat[i], size=lat[i - 1].shape[2:], mode="nearest")
    outs = [_gn(F.conv2d(lat[i], sd["neck.fpn_convs.%d.conv.weight" % i], None, 1, 1), sd, "neck.fpn_convs.%d.gn" % i)
            for i in range(3)]
    outs.append(_gn(F.conv2d(c5, sd["neck.fpn_convs.3.conv.weight"], None, 2, 1), sd, "neck.fpn_convs.3.gn"))
    outs.append(_gn(F.conv2d(outs[-1], sd["neck.fpn_convs.4.conv.weight"], None, 2, 1), sd, "neck.fpn_convs.4.gn"))
    return outs


def head_single(sd, x, gradient_mul=0.3):
    base = np.arange(-1, 2).astype(np.float64)
    dcn_base = torch.tensor(np.stack([np.repeat(base, 3), np.tile(base, 3)], 1).reshape(-1)).view(1, -1, 1, 1).type_as(x)
    cls_feat, pts_feat = x, x
    for i in range(3):
        cls_feat = F.relu(_gn(F.conv2d(cls_feat, sd["bbox_head.cls_convs.%d.conv.weight" % i], None, 1, 1), sd,
                              "bbox_head.cls_convs.%d.gn" % i))
        pts_feat = F.relu(_gn(F.conv2d(pts_feat, sd["bbox_head.reg_convs.%d.conv.weight" % i], None, 1, 1), sd,
                              "bbox_head.reg_convs.%d.gn" % i))
    h = "bbox_head.reppoints_"
    init = F.conv2d(F.relu(F.conv2d(pts_feat, sd[h + "pts_init_conv.weight"], sd[h + "pts_init_conv.bias"], 1, 1)),
                    sd[h + "pts_init_out.weight"], sd[h + "pts_init_out.bias"])
    grad_mul = (1 - gradient_mul) * init + gradient_mul * init
    dcn_offset = grad_mul - dcn_base
    cls_out = F.conv2d(F.relu(deform_conv_ref(cls_feat, dcn_offset, sd[h + "cls_conv.weight"])),
                       sd[h + "cls_out.weight"], sd[h + "cls_out.bias"])
    refine = F.conv2d(F.relu(deform_conv_ref(pts_feat, dcn_offset, sd[h + "pts_refine_conv.weight"])),
                      sd[h + "pts_refine_out.weight"], sd[h + "pts_refine_out.bias"])
    refine = refine + init
    return cls_out, init, refine, dcn_offset


def forward_dense(sd, img, blocks=(3, 4, 6, 3)):
    """returns per level (cls_out, pts_init, pts_refine) in NCHW"""
    feats = fpn(sd, backbone(sd, img, blocks))
    return [head_single(sd, f)[:3] for f in feats], feats


def get_bboxes_single(cls_scores, pts_refine, strides=(8, 16, 32, 64, 128), nms_pre=2000, score_thr=0.05,
                      iou_thr=0.4, max_per_img=2000, scale_factor=1.0, nms="f64"):
    """orientedreppoints_head.py:707-779 + bbox_nms.py:93-182 for ONE image (inputs [C,H,W] per level).
    minaerarect = oracle; rnms = oracle NMS.  nms='f64': class-segmented fp64 NMS (what the offset trick means);
    nms='f32offset': the reference's literal fp32 rnms on class-offset coordinates."""
    mb, ms, mr = [], [], []
    for lvl, (cls, pts) in enumerate(zip(cls_scores, pts_refine)):
        s = strides[lvl]
        C, H, W = cls.shape
        scores = cls.permute(1, 2, 0).reshape(-1, C).sigmoid().cpu()   # sigmoid on the caller's device
        pp = pts.permute(1, 2, 0).reshape(-1, 18)
        xs = torch.arange(W, dtype=torch.float32) * s
        ys = torch.arange(H, dtype=torch.float32) * s
        points = torch.stack([xs.repeat(H), ys.view(-1, 1).repeat(1, W).view(-1)], 1)      # point_generator.py:14-22
        if nms_pre > 0 and scores.shape[0] > nms_pre:
            mx, _ = scores.max(1)
            _, idx = torch.sort(mx, descending=True, stable=True)                           # ties: lower index first
            idx = idx[:nms_pre]
            points, pp, scores = points[idx], pp[idx], scores[idx]
        p3 = pp.reshape(-1, 9, 2)
        xy = torch.cat([p3[:, :, 1:2], p3[:, :, 0:1]], 2).reshape(-1, 18)                   # (y,x) -> (x,y)
        rect, _, _ = po.minarearect(xy.numpy())
        boxes = torch.from_numpy(rect) * s + points.repeat(1, 4)
        rp = xy * s + points.repeat(1, 9)
        mb.append(boxes)
        ms.append(scores)
        mr.append(rp)
    mb, ms, mr = torch.cat(mb), torch.cat(ms), torch.cat(mr)
    mb = mb / mb.new_tensor(scale_factor)
    mr = mr / mr.new_tensor(scale_factor)
    valid = ms > score_thr
    nz = valid.nonzero()
    rows, labels = nz[:, 0], nz[:, 1]
    if rows.numel() == 0:
        return torch.zeros((0, 27)), torch.zeros((0,), dtype=torch.long)
    bb, sc, rr = mb[rows], ms[valid], mr[rows]
    dets = torch.cat([bb, sc[:, None]], 1).numpy()
    if nms == "f64":
        keep = []
        for c in range(ms.shape[1]):
            ids = np.nonzero(labels.numpy() == c)[0]
            if len(ids):
                keep.append(ids[po.nms_poly_f64(dets[ids], iou_thr, fast=True)])
        keep = np.sort(np.concatenate(keep))
    else:
        off = labels.to(bb) * (bb.max() + 1)
        d2 = torch.cat([bb + off[:, None], sc[:, None]], 1).numpy()
        keep = np.sort(po.nms_f32(d2, np.float32(iou_thr)))
    keep = torch.from_numpy(keep).long()
    out = torch.cat([rr[keep], bb[keep], sc[keep][:, None]], 1)
    lab = labels[keep]
    if keep.numel() > max_per_img:
        _, inds = torch.sort(out[:, -1], descending=True, stable=True)
        inds = inds[:max_per_img]
        out, lab = out[inds], lab[inds]
    return out, lab
