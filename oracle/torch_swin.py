"""PyTorch re-declaration of the reference's Swin-T backbone + its FPN variant (TEST INFRASTRUCTURE ONLY),
following mmdet/models/backbones/swin_transformer.py:21-631 (Mlp, window_partition/reverse, WindowAttention,
SwinTransformerBlock, PatchMerging, BasicLayer mask, PatchEmbed, SwinTransformer.forward with out_indices (1,2,3))
and mmdet/models/necks/fpn.py:138-178 with start_level=0, add_extra_convs=False.  PINNED: equal (float64, incl. the
window-padding case) to the reference's OWN SwinTransformer + FPN modules imported from /root/reference with timm / mmcv
plumbing stubbed (tests/golden/gen_golden_swin.py -> swin_ref.npz, tests/test_oracle_golden.py)."""
import numpy as np
import torch
import torch.nn.functional as F

DEPTHS, HEADS, EMBED, WS = (2, 2, 6, 2), (3, 6, 12, 24), 96, 7


def window_partition(x, ws):
    B, H, W, C = x.shape
    x = x.view(B, H // ws, ws, W // ws, ws, C)
    return x.permute(0, 1, 3, 2, 4, 5).contiguous().view(-1, ws, ws, C)


def window_reverse(windows, ws, H, W):
    B = int(windows.shape[0] / (H * W / ws / ws))
    x = windows.view(B, H // ws, W // ws, ws, ws, -1)
    return x.permute(0, 1, 3, 2, 4, 5).contiguous().view(B, H, W, -1)


def rel_index(ws=WS):
    ch, cw = torch.arange(ws), torch.arange(ws)
    coords = torch.stack(torch.meshgrid([ch, cw], indexing="ij"))
    cf = torch.flatten(coords, 1)
    rel = (cf[:, :, None] - cf[:, None, :]).permute(1, 2, 0).contiguous()
    rel[:, :, 0] += ws - 1
    rel[:, :, 1] += ws - 1
    rel[:, :, 0] *= 2 * ws - 1
    return rel.sum(-1)


def window_attention(xw, sd, p, heads, mask):
    B_, N, C = xw.shape
    qkv = F.linear(xw, sd[p + "attn.qkv.weight"], sd[p + "attn.qkv.bias"]).reshape(B_, N, 3, heads, C // heads).permute(2, 0, 3, 1, 4)
    return attention_core(qkv[0], qkv[1], qkv[2], sd[p + "attn.relative_position_bias_table"], heads, mask, sd, p)


def attention_core(q, k, v, table, heads, mask, sd=None, p=None):
    B_, _, N, hd = q.shape
    q = q * (hd ** -0.5)
    attn = q @ k.transpose(-2, -1)
    bias = table[rel_index().view(-1).to(table.device)].view(N, N, -1).permute(2, 0, 1).contiguous()
    attn = attn + bias.unsqueeze(0)
    if mask is not None:
        nW = mask.shape[0]
        attn = attn.view(B_ // nW, nW, heads, N, N) + mask.unsqueeze(1).unsqueeze(0)
        attn = attn.view(-1, heads, N, N)
    attn = attn.softmax(-1)
    x = (attn @ v).transpose(1, 2).reshape(B_, N, heads * hd)
    if sd is not None:
        x = F.linear(x, sd[p + "attn.proj.weight"], sd[p + "attn.proj.bias"])
    return x


def shift_mask(Hp, Wp, shift, device, ws=WS):
    img_mask = torch.zeros((1, Hp, Wp, 1), device=device)
    cnt = 0
    for h in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
        for w in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
            img_mask[:, h, w, :] = cnt
            cnt += 1
    mw = window_partition(img_mask, ws).view(-1, ws * ws)
    am = mw.unsqueeze(1) - mw.unsqueeze(2)
    return am.masked_fill(am != 0, float(-100.0)).masked_fill(am == 0, float(0.0))


def block(x, H, W, sd, p, heads, shift, mask_matrix):
    B, L, C = x.shape
    shortcut = x
    x = F.layer_norm(x, (C,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], 1e-5).view(B, H, W, C)
    pad_r, pad_b = (WS - W % WS) % WS, (WS - H % WS) % WS
    x = F.pad(x, (0, 0, 0, pad_r, 0, pad_b))
    _, Hp, Wp, _ = x.shape
    if shift > 0:
        sx = torch.roll(x, shifts=(-shift, -shift), dims=(1, 2))
        am = mask_matrix
    else:
        sx, am = x, None
    xw = window_partition(sx, WS).view(-1, WS * WS, C)
    aw = window_attention(xw, sd, p, heads, am).view(-1, WS, WS, C)
    sx = window_reverse(aw, WS, Hp, Wp)
    x = torch.roll(sx, shifts=(shift, shift), dims=(1, 2)) if shift > 0 else sx
    if pad_r > 0 or pad_b > 0:
        x = x[:, :H, :W, :].contiguous()
    x = shortcut + x.view(B, H * W, C)
    y = F.layer_norm(x, (C,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], 1e-5)
    y = F.linear(F.gelu(F.linear(y, sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"])), sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"])
    return x + y


def patch_merging(x, H, W, sd, p):
    B, L, C = x.shape
    x = x.view(B, H, W, C)
    if (H % 2 == 1) or (W % 2 == 1):
        x = F.pad(x, (0, 0, 0, W % 2, 0, H % 2))
    x = torch.cat([x[:, 0::2, 0::2, :], x[:, 1::2, 0::2, :], x[:, 0::2, 1::2, :], x[:, 1::2, 1::2, :]], -1)
    x = x.view(B, -1, 4 * C)
    x = F.layer_norm(x, (4 * C,), sd[p + "norm.weight"], sd[p + "norm.bias"], 1e-5)
    return F.linear(x, sd[p + "reduction.weight"])


def swin_forward(sd, img):
    """-> [C@1/8 (192), C@1/16 (384), C@1/32 (768)] in NCHW"""
    _, _, H, W = img.shape
    if W % 4:
        img = F.pad(img, (0, 4 - W % 4))
    if H % 4:
        img = F.pad(img, (0, 0, 0, 4 - H % 4))
    x = F.conv2d(img, sd["backbone.patch_embed.proj.weight"], sd["backbone.patch_embed.proj.bias"], stride=4)
    Wh, Ww = x.shape[2], x.shape[3]
    x = x.flatten(2).transpose(1, 2)
    x = F.layer_norm(x, (EMBED,), sd["backbone.patch_embed.norm.weight"], sd["backbone.patch_embed.norm.bias"], 1e-5)
    outs = []
    for i, (depth, heads) in enumerate(zip(DEPTHS, HEADS)):
        Hp = int(np.ceil(Wh / WS)) * WS
        Wp = int(np.ceil(Ww / WS)) * WS
        mask = shift_mask(Hp, Wp, WS // 2, x.device)
        for j in range(depth):
            x = block(x, Wh, Ww, sd, "backbone.layers.%d.blocks.%d." % (i, j), heads, 0 if j % 2 == 0 else WS // 2, mask)
        if i in (1, 2, 3):
            C = EMBED << i
            o = F.layer_norm(x, (C,), sd["backbone.norm%d.weight" % i], sd["backbone.norm%d.bias" % i], 1e-5)
            outs.append(o.view(-1, Wh, Ww, C).permute(0, 3, 1, 2).contiguous())
        if i < 3:
            x = patch_merging(x, Wh, Ww, sd, "backbone.layers.%d.downsample." % i)
            Wh, Ww = (Wh + 1) // 2, (Ww + 1) // 2
    return outs


def swin_fpn(sd, feats):
    gn = lambda t, p: F.group_norm(t, 32, sd[p + ".weight"], sd[p + ".bias"], 1e-5)   # noqa: E731
    lat = [gn(F.conv2d(c, sd["neck.lateral_convs.%d.conv.weight" % i]), "neck.lateral_convs.%d.gn" % i) for i, c in enumerate(feats)]
    for i in (2, 1):
        lat[i - 1] = lat[i - 1] + F.interpolate(lat[i], size=lat[i - 1].shape[2:], mode="nearest")
    outs = [gn(F.conv2d(lat[i], sd["neck.fpn_convs.%d.conv.weight" % i], None, 1, 1), "neck.fpn_convs.%d.gn" % i) for i in range(3)]
    outs.append(F.max_pool2d(outs[-1], 1, stride=2))
    outs.append(F.max_pool2d(outs[-1], 1, stride=2))
    return outs
