"""OrientedRepPointsDetector inference (simple_test path) over liborp_b200.so.

Mirrors mmdet/models/detectors/orientedreppoints_detector.py:37-46:
    x = extract_feat(img)            # single_stage.py:44-50: backbone -> neck
    outs = bbox_head(x)              # orientedreppoints_head.py:173: forward_single per level
    bbox_list = bbox_head.get_bboxes(*outs, img_meta, test_cfg, rescale)
    rbbox2result(...)                # core/bbox/transforms.py:356-375

Host code is Python; every layer is a kernel of this repository called through the C ABI
(include/orp_b200.h) on torch-owned device memory and the current torch stream.  Activations are NHWC.
Three arithmetic engines: 'f16x3' (tcgen05 tensor cores on fp16 hi/lo operand pairs, three MMAs per product,
fp32 accumulation in TMEM - fp32-faithful, the parity AND benchmark arithmetic), 'bf16' (tcgen05, single-pass bf16
operands - 3x the rate, ~1e-2 accuracy) and 'fp32' (CUDA-core FMAs).  There is no PyTorch/cuDNN fallback for any layer.
"""
import numpy as np
import torch

from . import _lib
from .weights import STAGE_BLOCKS, fold_bn

STRIDES = (8, 16, 32, 64, 128)


class ConvLayer:
    """weights of one convolution in kernel layout [Cout, KH, KW, Cin] (+ optional bias)"""

    def __init__(self, w_nchw, bias, stride, pad, device, pad_cin_to=None):
        w = w_nchw.permute(0, 2, 3, 1).contiguous()                   # [Cout, KH, KW, Cin]
        self.w_raw = w.float().cpu()                                  # unpadded, for the tensor-core operand prep
        if pad_cin_to is not None and w.shape[3] < pad_cin_to:
            w = torch.cat([w, w.new_zeros(*w.shape[:3], pad_cin_to - w.shape[3])], 3).contiguous()
        self.cout, self.kh, self.kw, self.cin = w.shape
        self.w = w.to(device=device, dtype=torch.float32).contiguous()
        self.bias = None if bias is None else bias.to(device=device, dtype=torch.float32).contiguous()
        self.stride, self.pad = stride, pad
        self.tc = None    # tensor-core operand cache (dense_tc), filled lazily by the bf16 engine


class Norm:
    def __init__(self, sd, prefix, device):
        self.gamma = sd[prefix + ".weight"].to(device=device, dtype=torch.float32).contiguous()
        self.beta = sd[prefix + ".bias"].to(device=device, dtype=torch.float32).contiguous()


class EngineF32:
    """fp32 CUDA-core kernels (csrc/dense_f32.cu)"""
    name = "fp32"
    act_dtype = torch.float32

    def __init__(self, device):
        self.device = device
        self.lib = _lib.lib()

    def prepare_input(self, img_nchw):
        n, c, h, w = img_nchw.shape
        x = torch.zeros((n, h, w, 4), dtype=torch.float32, device=self.device)
        x[..., :c] = img_nchw.to(self.device, torch.float32).permute(0, 2, 3, 1)
        return x

    def conv(self, x, L, relu=False, residual=None, want_stats=False):
        n, h, w, cin = x.shape
        assert cin == L.cin, (cin, L.cin)
        ho = (h + 2 * L.pad - L.kh) // L.stride + 1
        wo = (w + 2 * L.pad - L.kw) // L.stride + 1
        y = torch.empty((n, ho, wo, L.cout), dtype=torch.float32, device=self.device)
        stats = torch.zeros((n, 32, 2), dtype=torch.float64, device=self.device) if want_stats else None
        rc = self.lib.orp_conv2d_f32(_lib.ptr(x), n, h, w, cin, _lib.ptr(L.w), L.cout, L.kh, L.kw, L.stride, L.pad,
                                     _lib.ptr(L.bias), _lib.ptr(residual), int(relu), _lib.ptr(y), _lib.ptr(stats), 32,
                                     _lib.current_stream_ptr())
        _lib.check(rc, "orp_conv2d_f32")
        return (y, stats) if want_stats else y

    def gn(self, x, stats, norm, relu=False, up=None):
        n, h, w, c = x.shape
        y = torch.empty_like(x)
        rc = self.lib.orp_gn_apply_f32(_lib.ptr(x), n, h, w, c, _lib.ptr(stats), 32, _lib.ptr(norm.gamma),
                                       _lib.ptr(norm.beta), 1e-5, int(relu), _lib.ptr(up), _lib.ptr(y),
                                       _lib.current_stream_ptr())
        _lib.check(rc, "orp_gn_apply_f32")
        return y

    def conv_gn(self, x, L, norm, relu=False, up=None):
        y, st = self.conv(x, L, want_stats=True)
        return self.gn(y, st, norm, relu=relu, up=up)

    # multi-level forms (the head's weights are shared by the five FPN levels)
    def conv_multi(self, xs, L, relu=False, residual=None, out_f32=False, residual_f32=None):
        res = residual if residual is not None else residual_f32
        return [self.conv(x, L, relu=relu, residual=None if res is None else res[i]) for i, x in enumerate(xs)]

    def conv_gn_multi(self, xs, L, norm, relu=False):
        return [self.conv_gn(x, L, norm, relu=relu) for x in xs]

    def deform_conv_multi(self, xs, offsets, L, relu=False):
        return [self.deform_conv(x, o, L, relu=relu) for x, o in zip(xs, offsets)]

    def stem(self, x, L):
        return self.conv(x, L, relu=True)

    def maxpool(self, x):
        n, h, w, c = x.shape
        ho, wo = (h + 2 - 3) // 2 + 1, (w + 2 - 3) // 2 + 1
        y = torch.empty((n, ho, wo, c), dtype=torch.float32, device=self.device)
        rc = self.lib.orp_maxpool3x3s2_f32(_lib.ptr(x), n, h, w, c, _lib.ptr(y), _lib.current_stream_ptr())
        _lib.check(rc, "orp_maxpool3x3s2_f32")
        return y

    def deform_conv(self, x, offset, L, relu=False, mask=None):
        n, h, w, cin = x.shape
        y = torch.empty((n, h, w, L.cout), dtype=torch.float32, device=self.device)
        rc = self.lib.orp_deform_conv2d_f32(_lib.ptr(x), n, h, w, cin, _lib.ptr(offset), _lib.ptr(mask), _lib.ptr(L.w),
                                            L.cout, L.kh, L.kw, L.stride, L.pad, 1, _lib.ptr(L.bias), int(relu),
                                            _lib.ptr(y), _lib.current_stream_ptr())
        _lib.check(rc, "orp_deform_conv2d_f32")
        return y


class OrientedRepPointsDetector:
    """R-50 / R-101 + FPN(GN) + OrientedRepPointsHead, inference only."""

    def __init__(self, state_dict, depth=50, device="cuda", precision="fp32", test_cfg=None):
        self.device = torch.device(device)
        if self.device.type == "cuda" and self.device.index is None:      # 'cuda' -> the current device, with its index
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.depth = depth
        self.test_cfg = dict(nms_pre=2000, min_bbox_size=0, score_thr=0.05, nms=dict(type='rnms', iou_thr=0.4),
                             max_per_img=2000)                         # configs/dota/orientedrepoints_r50_demo.py:62-67
        if test_cfg:
            self.test_cfg.update(test_cfg)
        # configs/dota/orientedrepoints_r50_demo.py:72-73; used when simple_test() is given decoded uint8 HWC tiles
        self.img_norm_cfg = dict(mean=[123.675, 116.28, 103.53], std=[58.395, 57.12, 57.375], to_rgb=True)
        if precision == "fp32":
            self.eng = EngineF32(self.device)
        elif precision == "bf16":
            from .engine_tc import EngineTC
            self.eng = EngineTC(self.device)
        elif precision == "f16x3":
            from .engine_tc import EngineTCSplit
            self.eng = EngineTCSplit(self.device)
        else:
            raise ValueError("precision must be 'f16x3' (tensor cores, fp32-faithful), 'bf16' (tensor cores) or 'fp32' (CUDA cores)")
        self._load(state_dict)
        base = np.arange(-1, 2).astype(np.float64)
        off = np.stack([np.repeat(base, 3), np.tile(base, 3)], axis=1).reshape(-1)        # head :78-88 (dy,dx)
        self.dcn_base_offset = torch.tensor(off, dtype=torch.float32, device=self.device).view(1, 1, 1, 18)
        import ctypes
        self._dcn_base_host = (ctypes.c_float * 18)(*[float(v) for v in off])

    # ------------------------------------------------------------------ weights
    def _load(self, sd):
        d = self.device
        if self.depth == "swin_tiny":
            return self._load_swin(sd)

        def folded(conv, bn, stride, pad, pad_cin_to=None):
            w, b = fold_bn(sd[conv + ".weight"].float(), sd[bn + ".weight"].float(), sd[bn + ".bias"].float(),
                           sd[bn + ".running_mean"].float(), sd[bn + ".running_var"].float())
            return ConvLayer(w, b, stride, pad, d, pad_cin_to)

        self.stem = folded("backbone.conv1", "backbone.bn1", 2, 3, pad_cin_to=4)
        self.blocks = []
        for li, nblk in enumerate(STAGE_BLOCKS[self.depth]):
            stage = []
            for b in range(nblk):
                p = "backbone.layer%d.%d" % (li + 1, b)
                s = 2 if (b == 0 and li > 0) else 1
                blk = dict(c1=folded(p + ".conv1", p + ".bn1", 1, 0), c2=folded(p + ".conv2", p + ".bn2", s, 1),
                           c3=folded(p + ".conv3", p + ".bn3", 1, 0),
                           ds=folded(p + ".downsample.0", p + ".downsample.1", s, 0) if b == 0 else None)
                stage.append(blk)
            self.blocks.append(stage)
        self.lat = [(ConvLayer(sd["neck.lateral_convs.%d.conv.weight" % i].float(), None, 1, 0, d),
                     Norm(sd, "neck.lateral_convs.%d.gn" % i, d)) for i in range(3)]
        self.fpn = [(ConvLayer(sd["neck.fpn_convs.%d.conv.weight" % i].float(), None, 2 if i >= 3 else 1, 1, d),
                     Norm(sd, "neck.fpn_convs.%d.gn" % i, d)) for i in range(5)]
        self._load_head(sd)

    def _load_head(self, sd):
        d = self.device
        h = "bbox_head."
        self.cls_convs = [(ConvLayer(sd[h + "cls_convs.%d.conv.weight" % i].float(), None, 1, 1, d),
                           Norm(sd, h + "cls_convs.%d.gn" % i, d)) for i in range(3)]
        self.reg_convs = [(ConvLayer(sd[h + "reg_convs.%d.conv.weight" % i].float(), None, 1, 1, d),
                           Norm(sd, h + "reg_convs.%d.gn" % i, d)) for i in range(3)]
        r = h + "reppoints_"
        self.cls_dcn = ConvLayer(sd[r + "cls_conv.weight"].float(), None, 1, 1, d)
        self.cls_out = ConvLayer(sd[r + "cls_out.weight"].float(), sd[r + "cls_out.bias"].float(), 1, 0, d)
        self.init_conv = ConvLayer(sd[r + "pts_init_conv.weight"].float(), sd[r + "pts_init_conv.bias"].float(), 1, 1, d)
        self.init_out = ConvLayer(sd[r + "pts_init_out.weight"].float(), sd[r + "pts_init_out.bias"].float(), 1, 0, d)
        self.ref_dcn = ConvLayer(sd[r + "pts_refine_conv.weight"].float(), None, 1, 1, d)
        self.ref_out = ConvLayer(sd[r + "pts_refine_out.weight"].float(), sd[r + "pts_refine_out.bias"].float(), 1, 0, d)

    def _load_swin(self, sd):
        """Swin-T + FPN(in [192,384,768], start_level 0, no extra convs: P6/P7 = stride-2 subsampling, fpn.py:163-165)"""
        from .swin import SwinTiny
        d = self.device
        if self.eng.name not in ("bf16", "f16x3"):
            raise ValueError("the Swin-T backbone runs on the tensor-core engines ('f16x3' or 'bf16')")
        self.swin = SwinTiny(sd, d, self.eng)
        self.lat = [(ConvLayer(sd["neck.lateral_convs.%d.conv.weight" % i].float(), None, 1, 0, d),
                     Norm(sd, "neck.lateral_convs.%d.gn" % i, d)) for i in range(3)]
        self.fpn = [(ConvLayer(sd["neck.fpn_convs.%d.conv.weight" % i].float(), None, 1, 1, d),
                     Norm(sd, "neck.fpn_convs.%d.gn" % i, d)) for i in range(3)]
        self._load_head(sd)

    # ------------------------------------------------------------------ dense graph
    def normalize(self, img):
        """decoded uint8 HWC tiles [N,H,W,3] -> the pipeline's Normalize + ImageToTensor on the device
        (mmdet/datasets/pipelines/transforms.py:Normalize, formating.py:ImageToTensor); float NCHW input passes through"""
        if img.dtype != torch.uint8:
            return img
        c = self.img_norm_cfg
        x = img.to(self.device).float()
        if c["to_rgb"]:
            x = x.flip(-1)
        key = (tuple(c["mean"]), tuple(c["std"]))
        if getattr(self, "_norm_key", None) != key:                   # device constants, made once (not inside a graph capture)
            self._norm_mean = torch.tensor(c["mean"], dtype=torch.float32).to(self.device)
            self._norm_stdinv = torch.tensor([1.0 / v for v in c["std"]], dtype=torch.float64).float().to(self.device)
            self._norm_key = key
        return ((x - self._norm_mean) * self._norm_stdinv).permute(0, 3, 1, 2).contiguous()

    def extract_feat(self, img):
        e = self.eng
        if self.depth == "swin_tiny":
            # decoded uint8 tiles: Normalize + ImageToTensor are fused into the patch gather
            c3, c4, c5 = self.swin.forward(img, self.img_norm_cfg)
            l2 = e.conv_gn(c5, *self.lat[2])
            l1 = e.conv_gn(c4, *self.lat[1], up=l2)
            l0 = e.conv_gn(c3, *self.lat[0], up=l1)
            outs = [e.conv_gn(l0, *self.fpn[0]), e.conv_gn(l1, *self.fpn[1]), e.conv_gn(l2, *self.fpn[2])]
            outs.append(self.swin.subsample2(outs[-1]))
            outs.append(self.swin.subsample2(outs[-1]))
            return outs
        if img.dtype == torch.uint8 and hasattr(e, "stem_u8") and img.shape[1] % 2 == 0 and img.shape[2] % 2 == 0:
            x = e.maxpool(e.stem_u8(img, self.stem, self.img_norm_cfg))     # Normalize fused into the stem input transform
        else:
            x = e.prepare_input(self.normalize(img))
            x = e.maxpool(e.stem(x, self.stem))
        feats = []
        for stage in self.blocks:
            for blk in stage:
                idt = x if blk["ds"] is None else e.conv(x, blk["ds"])
                o = e.conv(x, blk["c1"], relu=True)
                o = e.conv(o, blk["c2"], relu=True)
                x = e.conv(o, blk["c3"], relu=True, residual=idt)
            feats.append(x)
        c3, c4, c5 = feats[1], feats[2], feats[3]
        l2 = e.conv_gn(c5, *self.lat[2])
        l1 = e.conv_gn(c4, *self.lat[1], up=l2)
        l0 = e.conv_gn(c3, *self.lat[0], up=l1)
        outs = [e.conv_gn(l0, *self.fpn[0]), e.conv_gn(l1, *self.fpn[1]), e.conv_gn(l2, *self.fpn[2])]
        outs.append(e.conv_gn(c5, *self.fpn[3]))
        outs.append(e.conv_gn(outs[-1], *self.fpn[4]))
        return outs

    def head(self, feats, gradient_mul=0.3):
        """forward_single (orientedreppoints_head.py:148-171) for all levels at once: the weights are shared,
        so every layer is ONE launch over the five levels.  Returns per level (cls, init, refine), fp32 NHWC."""
        e = self.eng
        cf, pf = list(feats), list(feats)
        for (lc, nc), (lr, nr) in zip(self.cls_convs, self.reg_convs):
            cf = e.conv_gn_multi(cf, lc, nc, relu=True)
            pf = e.conv_gn_multi(pf, lr, nr, relu=True)
        init = e.conv_multi(e.conv_multi(pf, self.init_conv, relu=True), self.init_out, out_f32=True)   # [N,H,W,18]
        # head :162-163, evaluated in fp32 exactly as written there - one launch for the five levels
        import ctypes
        offsets = [torch.empty_like(t) for t in init]
        n = len(init)
        pa = (ctypes.c_void_p * n)(*[t.data_ptr() for t in init])
        po = (ctypes.c_void_p * n)(*[t.data_ptr() for t in offsets])
        ne = (ctypes.c_longlong * n)(*[t.numel() for t in init])
        _lib.check(_lib.lib().orp_dcn_offsets_multi(n, pa, po, ne, float(gradient_mul), self._dcn_base_host, _lib.current_stream_ptr()),
                   "orp_dcn_offsets_multi")
        cls = e.conv_multi(e.deform_conv_multi(cf, offsets, self.cls_dcn, relu=True), self.cls_out, out_f32=True)
        ref = e.conv_multi(e.deform_conv_multi(pf, offsets, self.ref_dcn, relu=True), self.ref_out, out_f32=True,
                           residual_f32=init)
        return [(c.float(), i, r.float()) for c, i, r in zip(cls, init, ref)]

    def forward_dense(self, img):
        # every launch goes to the current stream of the current device: make that this detector's device
        with torch.cuda.device(self.device):
            if img.device != self.device:
                img = img.to(self.device)
            feats = self.extract_feat(img)
            return self.head(feats), feats

    # ------------------------------------------------------------------ CUDA graph of the dense graph
    def capture(self, img_shape, dtype=torch.float32):
        """Capture backbone + FPN + head for a fixed input shape into ONE CUDA graph (static buffers): the
        ~180 kernel launches of a step become a single graph launch, which removes the host-side launch
        cost that otherwise dominates a one-tile step.  simple_test() replays it when the shape matches."""
        shape = tuple(img_shape)
        torch.cuda.set_device(self.device)
        self._g_img = torch.zeros(shape, dtype=dtype, device=self.device)
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(2):                                      # warm-up: lazy weight prep, func attributes
                self.forward_dense(self._g_img)
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph):
            self._g_out = self.forward_dense(self._g_img)
        self._g_shape = (shape, dtype)
        return self

    def forward_dense_graph(self, img):
        with torch.cuda.device(self.device):
            self._g_img.copy_(img, non_blocking=True)
            self._graph.replay()
        return self._g_out

    # ------------------------------------------------------------------ simple_test
    def simple_test(self, img, img_metas=None, rescale=False, return_tensors=False):
        with torch.cuda.device(self.device):
            return self._simple_test(img, img_metas, rescale, return_tensors)

    def _simple_test(self, img, img_metas, rescale, return_tensors):
        from .core.get_bboxes import get_bboxes
        if getattr(self, "_g_shape", None) == (tuple(img.shape), img.dtype):
            outs, _ = self.forward_dense_graph(img)
        else:
            outs, _ = self.forward_dense(img)
        n = img.shape[0]
        if img_metas is None:
            img_metas = [dict(scale_factor=1.0) for _ in range(n)]
        nms_cfg = self.test_cfg['nms']
        if getattr(self, "fused_post", True) and nms_cfg.get('type', 'rnms') == 'rnms' and nms_cfg.get('mode', 'exact64') == 'exact64':
            from .core.get_bboxes import get_bboxes_fused
            dets, labels, counts = get_bboxes_fused([o[0] for o in outs], [o[2] for o in outs], STRIDES, img_metas,
                                                    self.test_cfg, rescale)
            if return_tensors == "padded":
                return dets, labels, counts
            cnt = counts.tolist()                                      # the one host sync of a step
            if any(c < 0 for c in cnt):
                raise _lib.OrpError("rotated NMS candidate list overflowed its buffer (orp_head_postprocess): results invalid")
            results = [(dets[i, :cnt[i]], labels[i, :cnt[i]]) for i in range(n)]
        else:
            results = get_bboxes([o[0] for o in outs], [o[2] for o in outs], STRIDES, img_metas, self.test_cfg, rescale)
        if return_tensors:
            return results
        from .core.transforms import rbbox2result
        return [rbbox2result(d, l, 16) for d, l in results]

    # ------------------------------------------------------------------ aug_test (multi-scale / flip merge)
    @staticmethod
    def rbbox_flip(rbboxes, img_shape, direction='horizontal'):
        """orientedreppoints_detector.py:48-73: x -> w - x - 1 (or y -> h - y - 1) on every vertex"""
        assert rbboxes.shape[-1] % 8 == 0
        flipped = rbboxes.clone()
        if direction == 'horizontal':
            flipped[..., 0::2] = img_shape[1] - rbboxes[..., 0::2] - 1
        elif direction == 'vertical':
            flipped[..., 1::2] = img_shape[0] - rbboxes[..., 1::2] - 1
        else:
            raise ValueError('Invalid flipping direction "{}"'.format(direction))
        return flipped

    def merge_aug_results(self, aug_bboxes, aug_scores, img_metas):
        """orientedreppoints_detector.py:81-110: map every view's boxes back (un-flip, / scale_factor), concatenate"""
        recovered = []
        for bboxes, info in zip(aug_bboxes, img_metas):
            m = info[0]
            b = self.rbbox_flip(bboxes, m['img_shape']) if m['flip'] else bboxes
            recovered.append(b / m['scale_factor'])
        bboxes = torch.cat(recovered, dim=0)
        if aug_scores is None:
            return bboxes
        return bboxes, torch.cat(aug_scores, dim=0)

    def aug_test(self, imgs, img_metas, rescale=False):
        """orientedreppoints_detector.py:112-144.  imgs: list of views, each ONE image (float NCHW [1,3,H,W] or uint8
        HWC [1,H,W,3]); img_metas: list of [dict(img_shape, scale_factor, flip)].  Raw candidates of all views
        (get_bboxes(nms=False), head :778-779) are concatenated and go through ONE multiclass_rnms."""
        from .core.bbox_nms import multiclass_rnms
        from .core.get_bboxes import get_bboxes
        from .core.transforms import rbbox2result
        aug_bboxes, aug_scores = [], []
        for img, meta in zip(imgs, img_metas):
            assert img.shape[0] == 1, "aug_test: one image per view"
            outs, _ = self.forward_dense(img)
            b, sc = get_bboxes([o[0] for o in outs], [o[2] for o in outs], STRIDES, meta, self.test_cfg, False, nms=False)[0]
            aug_bboxes.append(b)
            aug_scores.append(sc)
        merged_bboxes, merged_scores = self.merge_aug_results(aug_bboxes, aug_scores, img_metas)
        det_bboxes, det_labels = multiclass_rnms(merged_bboxes, merged_scores, self.test_cfg['score_thr'],
                                                 self.test_cfg['nms'], self.test_cfg['max_per_img'])
        if not rescale:
            det_bboxes = det_bboxes.clone()
            det_bboxes[:, :8] *= img_metas[0][0]['scale_factor']
        return rbbox2result(det_bboxes, det_labels, 16)
