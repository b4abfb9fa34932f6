"""bf16 tensor-core engine of the dense path: tcgen05 implicit-GEMM convolutions (csrc/dense_tc.cu) plus
their memory-bound companions (csrc/dense_bf16_misc.cu).  Same interface as detector.EngineF32."""
import ctypes

import torch

from . import _lib


class EngineTC:
    name = "bf16"
    act_dtype = torch.bfloat16

    def __init__(self, device):
        self.device = device
        self.lib = _lib.lib()

    # ------------------------------------------------------------------ weights
    def _tc(self, L):
        if L.tc is None:
            w = L.w_raw                                              # [Cout, KH, KW, Cin] fp32 (unpadded Cin)
            cout, kh, kw, cin = w.shape
            cout_p = ((cout + 31) // 32) * 32
            cin_p = ((cin + 63) // 64) * 64                          # per tap: whole 64-channel K blocks, zero padded
            wp4 = torch.zeros((cout_p, kh, kw, cin_p), dtype=torch.float32)
            wp4[:cout, :, :, :cin] = w
            wp = wp4.reshape(cout_p, kh * kw * cin_p)
            L.tc = dict(w=wp.to(self.device, torch.bfloat16).contiguous(), cout_p=cout_p)
        return L.tc

    def _stem_tc(self, L):
        if L.tc is None:
            w = L.w_raw                                              # [64, 7, 7, 3]
            wp = torch.zeros((64, 192), dtype=torch.float32)
            wp[:, :147] = w.reshape(64, 147)
            L.tc = dict(w=wp.to(self.device, torch.bfloat16).contiguous(), cout_p=64)
        return L.tc

    # ------------------------------------------------------------------ layers
    def prepare_input(self, img_nchw):
        img = img_nchw.to(self.device, torch.float32).contiguous()
        assert img.shape[1] == 3
        return img                                                   # the stem kernel reads the NCHW image directly

    def stem(self, img, L, materialise=True):
        tc = self._stem_tc(L)
        n, _, h, w = img.shape
        ho, wo = (h + 6 - 7) // 2 + 1, (w + 6 - 7) // 2 + 1
        y = torch.empty((n, ho, wo, 64), dtype=torch.bfloat16, device=self.device)
        # default: im2col rows in HBM + plain GEMM (557 us per 4 tiles); the direct producer variant
        # (orp_stem_conv_bf16, no HBM intermediate) is LSU-bound in its current scalar-gather form (1136 us)
        if materialise:
            cols = torch.empty((n, ho, wo, 192), dtype=torch.bfloat16, device=self.device)
            _lib.check(self.lib.orp_stem_im2col_bf16(_lib.ptr(img), n, h, w, _lib.ptr(cols), _lib.current_stream_ptr()),
                       "orp_stem_im2col_bf16")
            self._launch([cols], [y], tc, 64, 1, 1, 192, 1, 0, L.bias, True, False, False)
            return y
        _lib.check(self.lib.orp_stem_conv_bf16(_lib.ptr(img), n, h, w, _lib.ptr(tc["w"]), _lib.ptr(L.bias), 1, _lib.ptr(y),
                                               _lib.current_stream_ptr()), "orp_stem_conv_bf16")
        return y

    def _launch(self, xs, ys, tc, cout, kh, kw, cin, stride, pad, bias, relu, out_f32, deform, res=None, res32=None,
                offsets=None, stats=None):
        n = len(xs)
        arr = (_lib.TcProblem * n)()
        for i in range(n):
            arr[i].x = xs[i].data_ptr()
            arr[i].N, arr[i].H, arr[i].W = xs[i].shape[0], xs[i].shape[1], xs[i].shape[2]
            arr[i].out = ys[i].data_ptr()
            arr[i].residual_bf16 = res[i].data_ptr() if res is not None else None
            arr[i].residual_f32 = res32[i].data_ptr() if res32 is not None else None
            arr[i].offset = offsets[i].data_ptr() if offsets is not None else None
            arr[i].gn_stats = stats[i].data_ptr() if stats is not None else None
        rc = self.lib.orp_conv2d_bf16(n, arr, _lib.ptr(tc["w"]), cout, tc["cout_p"], kh, kw, cin, stride, pad,
                                      _lib.ptr(bias), int(relu), int(out_f32), int(deform), _lib.current_stream_ptr())
        _lib.check(rc, "orp_conv2d_bf16")

    def conv_multi(self, xs, L, relu=False, residual=None, out_f32=False, residual_f32=None, stats=None):
        """relu: False/True, or 2 for the exact-GELU epilogue (Swin MLP); stats: per-problem double [N,32,2] tensors
        (zeroed) that receive the GroupNorm statistics of the outputs"""
        tc = self._tc(L)
        ys = []
        for x in xs:
            n, h, w, cin = x.shape
            assert cin == L.w_raw.shape[3] and x.dtype == torch.bfloat16
            ho = (h + 2 * L.pad - L.kh) // L.stride + 1
            wo = (w + 2 * L.pad - L.kw) // L.stride + 1
            ys.append(torch.empty((n, ho, wo, L.cout), dtype=torch.float32 if out_f32 else torch.bfloat16,
                                  device=self.device))
        self._launch(xs, ys, tc, L.cout, L.kh, L.kw, L.w_raw.shape[3], L.stride, L.pad, L.bias, relu, out_f32, False,
                     res=residual, res32=residual_f32, stats=stats)
        return ys

    def conv(self, x, L, relu=False, residual=None, out_f32=False):
        return self.conv_multi([x], L, relu, None if residual is None else [residual], out_f32)[0]

    def gn_multi(self, xs, norm, relu=False, ups=None, stats=None):
        """GroupNorm(32, 256) (+ReLU, + nearest-upsampled top-down add) of several tensors sharing gamma / beta
        in one launch; stats: per-tensor double [N,32,2] sums (from the conv epilogue) or None to compute them"""
        st = _lib.current_stream_ptr()
        k = len(xs)
        if stats is None:
            stats = []
            for x in xs:
                n, h, w, c = x.shape
                s = torch.zeros((n, 32, 2), dtype=torch.float64, device=self.device)
                _lib.check(self.lib.orp_gn_stats_bf16(_lib.ptr(x), n, h * w, c, 32, _lib.ptr(s), st), "orp_gn_stats_bf16")
                stats.append(s)
        ys = [torch.empty_like(x) for x in xs]
        arr = (_lib.GnProblem * k)()
        for i, x in enumerate(xs):
            assert x.shape[3] == 256 and x.dtype == torch.bfloat16
            arr[i].x = x.data_ptr()
            arr[i].N, arr[i].H, arr[i].W = x.shape[0], x.shape[1], x.shape[2]
            arr[i].stats = stats[i].data_ptr()
            arr[i].up_src = ups[i].data_ptr() if ups is not None and ups[i] is not None else None
            arr[i].y = ys[i].data_ptr()
        _lib.check(self.lib.orp_gn_apply_bf16_multi(k, arr, 256, 32, _lib.ptr(norm.gamma), _lib.ptr(norm.beta), 1e-5,
                                                    int(relu), st), "orp_gn_apply_bf16_multi")
        return ys

    def gn(self, x, norm, relu=False, up=None, stats=None):
        return self.gn_multi([x], norm, relu=relu, ups=[up], stats=None if stats is None else [stats])[0]

    def conv_gn(self, x, L, norm, relu=False, up=None):
        st = torch.zeros((x.shape[0], 32, 2), dtype=torch.float64, device=self.device)
        y = self.conv_multi([x], L, stats=[st])[0]                     # statistics come out of the conv epilogue
        return self.gn(y, norm, relu=relu, up=up, stats=st)

    def conv_gn_multi(self, xs, L, norm, relu=False):
        sts = torch.zeros((len(xs), xs[0].shape[0], 32, 2), dtype=torch.float64, device=self.device)
        ys = self.conv_multi(xs, L, stats=[sts[i] for i in range(len(xs))])
        return self.gn_multi(ys, norm, relu=relu, stats=[sts[i] for i in range(len(xs))])

    def maxpool(self, x):
        n, h, w, c = x.shape
        ho, wo = (h + 2 - 3) // 2 + 1, (w + 2 - 3) // 2 + 1
        y = torch.empty((n, ho, wo, c), dtype=torch.bfloat16, device=self.device)
        _lib.check(self.lib.orp_maxpool3x3s2_bf16(_lib.ptr(x), n, h, w, c, _lib.ptr(y), _lib.current_stream_ptr()),
                   "orp_maxpool3x3s2_bf16")
        return y

    def deform_conv_multi(self, xs, offsets, L, relu=False):
        tc = self._tc(L)
        ys = [torch.empty((x.shape[0], x.shape[1], x.shape[2], L.cout), dtype=torch.bfloat16, device=self.device)
              for x in xs]
        self._launch(xs, ys, tc, L.cout, L.kh, L.kw, L.w_raw.shape[3], L.stride, L.pad, L.bias, relu, False, True,
                     offsets=offsets)
        return ys

    def deform_conv(self, x, offset, L, relu=False, mask=None):
        if mask is not None:
            raise NotImplementedError("DCNv2 modulation is served by the fp32 engine (orp_deform_conv2d_f32)")
        return self.deform_conv_multi([x], [offset], L, relu)[0]
