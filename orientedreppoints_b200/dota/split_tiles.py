"""Tile producer: mirror of DOTA_devkit/SplitOnlyImage_multi_process.py (splitbase.SplitSingle :51-87,
saveimagepatches :38-49; the demo splits with gap=200, subsize=1024, rates 1 / 0.5 / 1.5 :113-118).

The reference writes every tile to `<name>__<rate>__<left>___<up>.png` and the test data loader decodes it again;
here the decoded image goes to the device once and the uint8 HWC tile batch the detector consumes is cut out in HBM by
one kernel (`orp_split_tiles_u8`).  Tile names and origins are the reference's, so `dota/result_merge.py` (which parses
`__<rate>__<left>___<up>`) maps the detections back unchanged.
"""
import numpy as np
import torch

from .. import _lib


def tile_origins(width, height, subsize=1024, gap=200):
    """(left, up) of every tile in the reference's order (left outer loop, up inner loop; the last tile of a row /
    column is pulled back so that it ends at the image border; SplitOnlyImage_multi_process.py:67-87)"""
    if subsize <= gap:
        raise ValueError("gap must be smaller than subsize")
    slide = subsize - gap
    out = []
    left = 0
    while left < width:
        if left + subsize >= width:
            left = max(width - subsize, 0)
        up = 0
        while up < height:
            if up + subsize >= height:
                up = max(height - subsize, 0)
            out.append((left, up))
            if up + subsize >= height:
                break
            up += slide
        if left + subsize >= width:
            break
        left += slide
    return out


def tile_names(name, rate, origins):
    """`<name>__<rate>__<left>___<up>` (SplitOnlyImage_multi_process.py:60,77; str(rate) as python prints it)"""
    base = name + '__' + str(rate) + '__'
    return [base + str(l) + '___' + str(u) for l, u in origins]


def resize_image(img, rate):
    """rate != 1: cv2.resize(..., fx=rate, fy=rate, interpolation=cv2.INTER_CUBIC) on the host, exactly the reference's
    call (:54-58) - a one-off per image, not on the per-tile path"""
    if rate == 1:
        return img
    import cv2
    return cv2.resize(np.asarray(img), None, fx=rate, fy=rate, interpolation=cv2.INTER_CUBIC)


def split_image(img, name="img", rate=1, subsize=1024, gap=200, device=None):
    """img: decoded uint8 HWC image (numpy array or tensor, host or device) -> (tiles uint8 [T,subsize,subsize,C] on the
    device, names, origins).  Windows leaving the image are zero padded (padding=True, :44-47)."""
    if not torch.is_tensor(img):
        img = torch.from_numpy(np.ascontiguousarray(resize_image(img, rate)))
    elif rate != 1:
        img = torch.from_numpy(np.ascontiguousarray(resize_image(img.cpu().numpy(), rate)))
    if img.dtype != torch.uint8 or img.dim() != 3:
        raise TypeError("split_image: uint8 HWC image expected")
    dev = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
    img = img.to(dev).contiguous()
    h, w, c = img.shape
    origins = tile_origins(w, h, subsize, gap)
    org = torch.tensor(origins, dtype=torch.int32).reshape(-1, 2).to(dev)
    out = torch.empty((len(origins), subsize, subsize, c), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().orp_split_tiles_u8(_lib.ptr(img), h, w, c, _lib.ptr(org), len(origins), subsize, _lib.ptr(out),
                                                 _lib.current_stream_ptr()), "orp_split_tiles_u8")
    return out, tile_names(name, rate, origins), origins
