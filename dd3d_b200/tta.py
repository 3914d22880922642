"""DD3DB200WithTTA -- mirror of the reference's test-time-augmentation wrapper.

``tridet.modeling.dd3d.test_time_augmentation.DD3DWithTTA`` (test_time_augmentation.py:89-239) takes one mapped dataset
dict per image, builds ``len(TEST.AUG.MIN_SIZES) * (2 if FLIP else 1)`` augmented views on the CPU (PIL resize + numpy
flip, ``DatasetMapperTTA`` :24-87), runs the model on them in chunks of ``TEST.IMS_PER_BATCH // world_size``, maps every
detection back with numpy loops (:190-239) and reduces with one NMS (:160-171).  Here the views are produced on the GPU by
the fused resize(+flip)+normalise kernel straight into the engine (``dd3d_forward_resized``), the detections never leave
the device until the end, and the inverse transforms + merged NMS are one kernel pair (``dd3d_op_tta_merge``).  The host
code below only restates the transform bookkeeping (shapes, fp32 factors, intrinsics) in the reference's order.
"""
import ctypes as C

import numpy as np
import torch
from torch import nn

from . import lib as _lib
from .meta_arch import DD3DB200, NuscenesDD3DB200
from .structures import Boxes, Boxes3D, Instances


def build_views(image_hw, orig_hw, K_input, min_sizes, max_size, flip):
    """DatasetMapperTTA.__call__ (test_time_augmentation.py:50-87) + the inverse bookkeeping of
    _get_augmented_instances (:196-216) for one image.  Returns [(new_h, new_w, flip, lib.TtaView)]."""
    L = _lib.load()
    h, w = image_hw
    oh, ow = orig_hw
    K_input = np.asarray(K_input, dtype=np.float32).reshape(3, 3)
    pre = (oh, ow) != (h, w)  # pre_tfm = ResizeTransform(orig -> input) or NoOpTransform (:52-58)
    views = []
    nh_c, nw_c = C.c_int32(), C.c_int32()
    for min_size in min_sizes:
        _lib.check(L.dd3d_resize_shape(h, w, int(min_size), int(max_size), C.byref(nh_c), C.byref(nw_c)))
        nh, nw = nh_c.value, nw_c.value
        # apply_imresize_intrinsics (resize_transform.py:13-21)
        K_r = K_input * np.float32([nw / w, nh / h, 1]).reshape(3, 1)
        for f in ([0, 1] if flip else [0]):
            K_v = K_r.copy()
            if f:  # apply_hflip_intrinsics (flip_transform.py:8-10), width = width of the resized view
                K_v[0, 2] = nw - K_v[0, 2]
            # inv_tfm = (pre_tfm + tfms).inverse(): un-flip, un-resize, un-pre-resize (fvcore TransformList.inverse)
            K_o = K_v.copy()
            if f:
                K_o[0, 2] = nw - K_o[0, 2]
            K_o = K_o * np.float32([w / nw, h / nh, 1]).reshape(3, 1)
            if pre:
                K_o = K_o * np.float32([ow / w, oh / h, 1]).reshape(3, 1)
            v = _lib.TtaView()
            v.flip = f
            v.view_w = float(nw)
            # ResizeTransform.apply_coords of the inverses: coords * (new_w * 1.0 / w) with an fp32 coordinate array
            v.inv_sx[0], v.inv_sy[0] = np.float32(w * 1.0 / nw), np.float32(h * 1.0 / nh)
            v.inv_sx[1], v.inv_sy[1] = (np.float32(ow * 1.0 / w), np.float32(oh * 1.0 / h)) if pre else (1.0, 1.0)
            for i in range(9):
                v.K_view[i] = float(K_v.reshape(-1)[i])
                v.K_orig[i] = float(K_o.reshape(-1)[i])
            views.append((nh, nw, f, v))
    return views


class DD3DB200WithTTA(nn.Module):
    """Same constructor and call contract as DD3DWithTTA(cfg, model): ``__call__(batched_inputs)`` with mapped dataset
    dicts ("image" CHW uint8, "intrinsics", optional "height" / "width") -> ``[{"instances": Instances}]`` with
    pred_boxes, pred_boxes3d, pred_classes, scores, scores_3d on the original image, sorted by scores_3d."""
    def __init__(self, cfg, model, tta_mapper=None, world_size=1):
        super().__init__()
        assert isinstance(model, DD3DB200) and not isinstance(model, NuscenesDD3DB200), \
            "DD3DB200WithTTA only supports DD3DB200. Got a model of type {}".format(type(model))
        assert not model.postprocess_in_inference, \
            "To use test-time augmentation, `postprocess_in_inference` must be False."
        if tta_mapper is not None:
            raise NotImplementedError("custom tta_mapper: the views are generated on the device")
        if model.do_bev_nms:
            raise NotImplementedError("TTA with DO_BEV_NMS")
        self.cfg = cfg
        self.model = model
        self.nms_thresh = cfg.DD3D.FCOS2D.INFERENCE.NMS_THRESH
        self.min_sizes = list(cfg.TEST.AUG.MIN_SIZES)
        self.max_size = cfg.TEST.AUG.MAX_SIZE
        self.flip = bool(cfg.TEST.AUG.FLIP)
        self.batch_size = max(1, cfg.TEST.IMS_PER_BATCH // world_size)  # test_time_augmentation.py:116

    def __call__(self, batched_inputs):
        return [self._inference_one_image(x) for x in batched_inputs]

    @torch.no_grad()
    def _inference_one_image(self, x):
        model, L = self.model, _lib.load()
        device = model.device
        image = torch.as_tensor(x["image"])
        if image.dtype != torch.uint8:
            raise ValueError("TTA resamples uint8 images (PIL path of ResizeTransform.apply_image)")
        h, w = int(image.shape[1]), int(image.shape[2])
        orig = (int(x.get("height", h)), int(x.get("width", w)))
        views = build_views((h, w), orig, x["intrinsics"], self.min_sizes, self.max_size, self.flip)
        A, cap = len(views), model._desc.out_cap
        with torch.cuda.device(device):
            stream = torch.cuda.current_stream(device).cuda_stream
            hwc = image.to(device, non_blocking=True).permute(1, 2, 0).contiguous()
            dets = torch.empty((A, cap, _lib.DET_WORDS), dtype=torch.float32, device=device)
            counts = torch.empty((A, ), dtype=torch.int32, device=device)
            # chunks of `batch_size` views, each padded to the chunk's largest view like ImageList.from_tensors does for
            # one model(inputs) call (test_time_augmentation.py:118-133)
            for a0 in range(0, A, self.batch_size):
                chunk = views[a0:a0 + self.batch_size]
                B = len(chunk)
                model._plan(B, max(v[0] for v in chunk), max(v[1] for v in chunk))
                raw = hwc.unsqueeze(0).expand(B, h, w, 3).contiguous()
                raw_sizes = torch.tensor([[h, w]] * B, dtype=torch.int32)
                new_sizes = torch.tensor([[v[0], v[1]] for v in chunk], dtype=torch.int32)
                flips = torch.tensor([v[2] for v in chunk], dtype=torch.int32)
                K = torch.tensor([list(v[3].K_view) for v in chunk], dtype=torch.float32)
                sizes4 = torch.cat([new_sizes, new_sizes], 1).contiguous()
                _lib.check(L.dd3d_set_option(model._handle, b"do_postprocess", 0), model._handle)
                _lib.check(L.dd3d_set_option(model._handle, b"do_nms", int(model.do_nms)), model._handle)
                _lib.check(
                    L.dd3d_forward_resized(model._handle, C.c_void_p(raw.data_ptr()), h, w, C.c_void_p(raw_sizes.data_ptr()),
                                           C.c_void_p(new_sizes.data_ptr()), C.c_void_p(flips.data_ptr()),
                                           C.c_void_p(K.data_ptr()), C.c_void_p(sizes4.data_ptr()),
                                           C.c_void_p(dets[a0].data_ptr()), C.c_void_p(counts[a0:].data_ptr()),
                                           C.c_void_p(stream)), model._handle)
            mcap = int(L.dd3d_op_tta_merged_cap(A, cap))
            out = torch.empty((mcap, _lib.DET_WORDS), dtype=torch.float32, device=device)
            n_out = torch.zeros(1, dtype=torch.int32, device=device)
            self._flags = torch.zeros(1, dtype=torch.int32, device=device)
            scratch = torch.empty(int(L.dd3d_op_tta_merge_scratch_bytes(A, cap)), dtype=torch.uint8, device=device)
            varr = (_lib.TtaView * A)(*[v[3] for v in views])
            _lib.check(
                L.dd3d_op_tta_merge(C.c_void_p(dets.data_ptr()), C.c_void_p(counts.data_ptr()), varr, A, cap,
                                    float(self.nms_thresh), int(model.do_nms), C.c_void_p(scratch.data_ptr()),
                                    C.c_void_p(out.data_ptr()), C.c_void_p(n_out.data_ptr()),
                                    C.c_void_p(self._flags.data_ptr()), C.c_void_p(stream)), model._handle)
            n = int(n_out.item())  # the only synchronisation
        d, di = out[:n], out.view(torch.int32)[:n]
        # Boxes3D.from_vectors(vecs, orig_intrinsics): every detection keeps the inverse intrinsics recovered for its view
        inv_K = torch.stack([
            torch.linalg.inv(torch.tensor(list(v[3].K_orig), dtype=torch.float64).reshape(3, 3)).to(torch.float32)
            for v in views
        ]).to(device)
        inst = Instances(orig)
        inst.pred_boxes = Boxes(d[:, 0:4].clone())
        inst.pred_boxes3d = Boxes3D(d[:, 8:12].clone(), d[:, 12:14].clone(), d[:, 14:15].clone(), d[:, 15:18].clone(),
                                    inv_K[di[:, 7].to(torch.int64)])
        inst.pred_classes = di[:, 6].to(torch.int64)
        inst.scores = d[:, 4].clone()
        inst.scores_3d = d[:, 5].clone()
        return {"instances": inst}

    def overflow_flags(self):
        return self.model.overflow_flags() | int(self._flags.item())
