"""TEST INFRASTRUCTURE ONLY -- CPU oracle of test-time augmentation.

Restates tridet/modeling/dd3d/test_time_augmentation.py:
  * DatasetMapperTTA.__call__ (:38-87): per TEST.AUG.MIN_SIZES entry a ResizeShortestEdge view (PIL bilinear, restated
    bit-exactly in oracle/input_oracle.py) and, with FLIP, its horizontal mirror; intrinsics through
    apply_imresize_intrinsics (tridet/data/augmentations/resize_transform.py:13-21) and apply_hflip_intrinsics
    (flip_transform.py:8-10);
  * DD3DWithTTA._batch_inference (:118-133): the views run through the model in chunks of `batch_size`, each chunk zero
    padded to its largest view (ImageList.from_tensors);
  * _get_augmented_instances (:190-239): inverse transforms of 2-D boxes (fvcore apply_box, fp32), 3-D boxes
    (apply_hflip_box3d, flip_transform.py:28-55), intrinsics, Boxes3D.from_vectors (tridet/structures/boxes3d.py:176-216);
  * _inference_one_image (:151-184): one class-aware NMS on scores_3d over the union, output in that order.
The fvcore / detectron2 transform semantics are third-party (restated; pinned here through the reference's own TTA class
run under oracle/ref_standin.py -> tests/golden/tta_dla34.npz).
"""
import numpy as np
import torch

from oracle import input_oracle as IO
from oracle.dd3d_oracle import batched_nms_restated


def make_views(image_chw, orig_hw, K, min_sizes, max_size, flip):
    """-> list of dicts: image (CHW uint8 tensor), intrinsics (3x3 float32 tensor), new_hw, flip."""
    img = image_chw.permute(1, 2, 0).numpy()
    h, w = img.shape[:2]
    K = np.asarray(K, dtype=np.float32)
    views = []
    for s in min_sizes:
        nh, nw = IO.resize_shortest_edge_shape(h, w, s, max_size)
        r = IO.pil_resize_bilinear(img, nh, nw)
        K_r = IO.scale_intrinsics(K, h, w, nh, nw)
        for f in ([0, 1] if flip else [0]):
            im = np.ascontiguousarray(np.flip(r, axis=1)) if f else r
            K_v = K_r.copy()
            if f:
                K_v[0, 2] = nw - K_v[0, 2]
            views.append(dict(image=torch.from_numpy(np.ascontiguousarray(im.transpose(2, 0, 1))),
                              intrinsics=torch.from_numpy(K_v.copy()), new_hw=(nh, nw), flip=f))
    return views


def invert_view(det, view, image_hw, orig_hw):
    """One view's detections (oracle dict: box2d, quat, tvec, size, score, score3d, cls) -> on the original image:
    dict(box2d, quat, proj_ctr, depth, size, inv_K, score, score3d, cls)."""
    h, w = image_hw
    oh, ow = orig_hw
    nh, nw = view["new_hw"]
    box = det["box2d"].numpy().astype(np.float32).copy()
    x1, y1, x2, y2 = box[:, 0].copy(), box[:, 1].copy(), box[:, 2].copy(), box[:, 3].copy()
    if view["flip"]:  # HFlipTransform.apply_coords on the 4 corners, then min / max
        x1, x2 = np.float32(nw) - x2, np.float32(nw) - x1
    steps = [(np.float32(w * 1.0 / nw), np.float32(h * 1.0 / nh))]
    if (oh, ow) != (h, w):
        steps.append((np.float32(ow * 1.0 / w), np.float32(oh * 1.0 / h)))
    for fx, fy in steps:
        x1, x2, y1, y2 = x1 * fx, x2 * fx, y1 * fy, y2 * fy
    out_box = np.stack([np.minimum(x1, x2), np.minimum(y1, y2), np.maximum(x1, x2), np.maximum(y1, y2)], 1)
    quat = det["quat"].numpy().astype(np.float32)
    tvec = det["tvec"].numpy().astype(np.float32).copy()
    if view["flip"]:  # apply_hflip_box3d
        quat = np.stack([quat[:, 3], -quat[:, 2], -quat[:, 1], quat[:, 0]], 1)
        tvec[:, 0] = -tvec[:, 0]
    K_o = view["intrinsics"].numpy().astype(np.float32).copy()  # inv_tfm.apply_intrinsics
    if view["flip"]:
        K_o[0, 2] = nw - K_o[0, 2]
    K_o = K_o * np.float32([w / nw, h / nh, 1]).reshape(3, 1)
    if (oh, ow) != (h, w):
        K_o = K_o * np.float32([ow / w, oh / h, 1]).reshape(3, 1)
    proj = tvec @ K_o.T  # Boxes3D.from_vectors: intrinsics.dot(tvec)
    n = tvec.shape[0]
    extra = {k: det[k] for k in ("level", "pixel", "loc") if k in det}  # provenance, for matching in tests
    return dict(**extra, box2d=torch.from_numpy(out_box.astype(np.float32)), quat=torch.from_numpy(quat),
                proj_ctr=torch.from_numpy((proj[:, :2] / proj[:, 2:3]).astype(np.float32)),
                depth=torch.from_numpy(tvec[:, 2].copy()), size=det["size"], score=det["score"], score3d=det["score3d"],
                cls=det["cls"], inv_K=torch.from_numpy(np.linalg.inv(K_o).astype(np.float32))[None].expand(n, 3, 3),
                view=torch.full((n, ), view.get("index", 0), dtype=torch.long))


def merge(inverted, nms_thresh, do_nms=True):
    """Union in view order + batched_nms on scores_3d; survivors in descending score order (merged_instances[keep])."""
    cat = {k: torch.cat([d[k] for d in inverted], 0) for k in inverted[0]}
    if cat["box2d"].shape[0] and do_nms:
        keep = batched_nms_restated(cat["box2d"], cat["score3d"], cat["cls"], nms_thresh)
        cat = {k: v[keep] for k, v in cat.items()}
    return cat


def tta_forward(oracle, x, min_sizes, max_size, flip, batch_size, nms_thresh):
    """DD3DWithTTA._inference_one_image with `oracle` (a DD3DOracle) as the model."""
    image = x["image"]
    h, w = int(image.shape[1]), int(image.shape[2])
    orig = (int(x.get("height", h)), int(x.get("width", w)))
    views = make_views(image, orig, x["intrinsics"], min_sizes, max_size, flip)
    outs = []
    for a0 in range(0, len(views), batch_size):
        chunk = [{"image": v["image"], "intrinsics": v["intrinsics"]} for v in views[a0:a0 + batch_size]]
        outs.extend(oracle.forward(chunk, do_postprocess=False))
    inverted = []
    for a, (v, det) in enumerate(zip(views, outs)):
        v["index"] = a
        inverted.append(invert_view(det, v, (h, w), orig))
    return merge(inverted, nms_thresh)
