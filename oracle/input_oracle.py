"""TEST INFRASTRUCTURE ONLY -- CPU oracle of the test-time input pipeline.

Restates what `DefaultDatasetMapper.__call__` (tridet/data/dataset_mappers/dataset_mapper.py:100-153) does to an image
and its intrinsics at test time (augmentations = [ResizeShortestEdge], tridet/data/augmentations/build.py:35-44):

  * output shape        detectron2 `ResizeShortestEdge.get_transform` (un-vendored; restated)  resize_transform.py:91-94
  * image resampling    detectron2 `ResizeTransform.apply_image` -> `PIL.Image.resize((w, h), Image.BILINEAR)` for uint8
                        images.  The arithmetic is Pillow's `ImagingResample` (src/libImaging/Resample.c): separable
                        antialiased triangle filter, 8-bit fixed point with PRECISION_BITS = 22, horizontal pass rounded
                        to uint8 before the vertical pass.  Pillow IS installed in this image (12.2.0), so this
                        restatement is pinned bit-exactly against the real library in tests/test_input_pipeline.py.
  * intrinsics          `apply_imresize_intrinsics`  tridet/data/augmentations/resize_transform.py:13-21
  * HWC -> CHW tensor   dataset_mapper.py:126
"""
import math

import numpy as np
import torch

PRECISION_BITS = 32 - 8 - 2  # Resample.c


def resize_shortest_edge_shape(h, w, min_size, max_size):
    """detectron2 ResizeShortestEdge.get_transform with sample_style="choice" and one size (build.py:40-44)."""
    if isinstance(min_size, (list, tuple)):
        assert len(set(min_size)) == 1, "test-time resize uses one size"
        min_size = min_size[0]
    if min_size == 0:
        return h, w
    scale = min_size * 1.0 / min(h, w)
    if h < w:
        newh, neww = min_size, scale * w
    else:
        newh, neww = scale * h, min_size
    if max(newh, neww) > max_size:
        scale = max_size * 1.0 / max(newh, neww)
        newh = newh * scale
        neww = neww * scale
    return int(newh + 0.5), int(neww + 0.5)


def bilinear_coeffs(in_size, out_size):
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc for the bilinear (triangle, support 1) filter and the full
    box [0, in_size): returns (xmin [out], count [out], int32 coefficients [out][ksize])."""
    scale = float(in_size) / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    xmins = np.zeros(out_size, dtype=np.int32)
    counts = np.zeros(out_size, dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = np.zeros(ksize, dtype=np.float64)
        ww = 0.0
        for x in range(xmax):
            a = (x + xmin - center + 0.5) * ss
            if a < 0.0:
                a = -a
            w[x] = 1.0 - a if a < 1.0 else 0.0
            ww += w[x]
        for x in range(xmax):
            if ww != 0.0:
                w[x] /= ww
        for x in range(ksize):
            v = w[x] * (1 << PRECISION_BITS)
            kk[xx, x] = int(-0.5 + v) if w[x] < 0 else int(0.5 + v)
        xmins[xx] = xmin
        counts[xx] = xmax
    return xmins, counts, kk


def _clip8(acc):
    return np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)


def pil_resize_bilinear(img, new_h, new_w):
    """img: (H, W, 3) uint8 -> (new_h, new_w, 3) uint8, bit-identical to PIL.Image.resize((new_w, new_h), BILINEAR)."""
    H, W, _ = img.shape
    src = img.astype(np.int64)
    if new_w != W:  # horizontal pass (ImagingResampleHorizontal_8bpc)
        xmin, cnt, kx = bilinear_coeffs(W, new_w)
        acc = np.full((H, new_w, 3), 1 << (PRECISION_BITS - 1), dtype=np.int64)
        for k in range(kx.shape[1]):
            idx = np.minimum(xmin + k, W - 1)
            acc += src[:, idx, :] * (kx[:, k] * (k < cnt)).astype(np.int64)[None, :, None]
        src = _clip8(acc).astype(np.int64)
    if new_h != H:  # vertical pass (ImagingResampleVertical_8bpc)
        ymin, cnt, ky = bilinear_coeffs(H, new_h)
        acc = np.full((new_h, src.shape[1], 3), 1 << (PRECISION_BITS - 1), dtype=np.int64)
        for k in range(ky.shape[1]):
            idx = np.minimum(ymin + k, H - 1)
            acc += src[idx, :, :] * (ky[:, k] * (k < cnt)).astype(np.int64)[:, None, None]
        src = _clip8(acc).astype(np.int64)
    return src.astype(np.uint8)


def scale_intrinsics(K, h, w, new_h, new_w):
    """apply_imresize_intrinsics (resize_transform.py:13-21): rows 0 / 1 of the float32 matrix scaled by new_w / w and
    new_h / h (float32 factors)."""
    K = np.asarray(K, dtype=np.float32).reshape(3, 3)
    assert K[0, 1] == 0 and np.allclose(K, np.triu(K))
    f = np.float32([new_w / w, new_h / h, 1]).reshape(3, 1)
    return K * f


def map_input(raw, min_size, max_size):
    """DefaultDatasetMapper.__call__ for one test image: raw = {"image_hwc": (H, W, 3) uint8 BGR array,
    "intrinsics": 3x3}.  Returns the dict DD3D.forward consumes (image CHW uint8, scaled intrinsics, and the original
    height / width the detections are mapped back to by detector_postprocess)."""
    img = np.asarray(raw["image_hwc"])
    h, w = img.shape[:2]
    new_h, new_w = resize_shortest_edge_shape(h, w, min_size, max_size)
    out = pil_resize_bilinear(img, new_h, new_w)
    K = scale_intrinsics(np.asarray(raw["intrinsics"], dtype=np.float32), h, w, new_h, new_w)
    d = {k: v for k, v in raw.items() if k not in ("image_hwc", "intrinsics")}
    d.update(image=torch.as_tensor(np.ascontiguousarray(out.transpose(2, 0, 1))), intrinsics=torch.as_tensor(K),
             height=h, width=w)
    return d
