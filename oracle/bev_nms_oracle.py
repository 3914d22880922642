"""TEST INFRASTRUCTURE ONLY -- CPU oracle of the BEV rotated-NMS path.

Restates, with the reference file:line each step follows:
  * sensor->global transform of the decoded boxes   tridet/modeling/dd3d/postprocessing.py:22-55 (sample_bev_nms)
  * 3-D corners, top surface, BEV rectangle          tridet/structures/boxes3d.py:12-16,47-64; tridet/layers/bev_nms.py:51-96
  * class-aware greedy rotated NMS                   tridet/layers/bev_nms.py:99-133 -> detectron2 batched_nms_rotated
  * per-image application with dummy groups          tridet/modeling/dd3d/core.py:137-151, postprocessing.py:58-108

The rotated IoU itself lives in detectron2 (un-vendored, unpinned wheel): `box_iou_rotated_utils.h`
(get_rotated_vertices / get_intersection_points / convex_hull_graham / polygon_area, IoU = inter / (a1 + a2 - inter),
greedy suppression when IoU > threshold).  It is restated here from its published algorithm -- PARITY UNPINNED for
that third-party piece (no detectron2 build is available offline); tests pin it against an independent
Sutherland-Hodgman polygon clipper and exact closed-form cases instead.
"""
import math

import numpy as np
import torch

from oracle.dd3d_oracle import matrix_to_quaternion, quaternion_to_matrix

# boxes3d.py:12-16 (columns = the 8 corners; rows scale l, w, h)
BOX3D_CORNER_MAPPING = np.array([[1, 1, 1, 1, -1, -1, -1, -1], [1, -1, -1, 1, 1, -1, -1, 1], [1, 1, -1, -1, 1, 1, -1, -1]],
                                dtype=np.float32)


def to_global(quat, tvec, pose_quat, pose_tvec):
    """postprocessing.py:25-46: object->sensor then sensor->world; returns (quat_global, tvec_global)."""
    R_so = quaternion_to_matrix(quat)
    R_ws = quaternion_to_matrix(torch.as_tensor(pose_quat, dtype=torch.float32)[None])[0]
    R = torch.matmul(R_ws[None], R_so)
    t = torch.matmul(tvec, R_ws.T) + torch.as_tensor(pose_tvec, dtype=torch.float32)[None]
    return matrix_to_quaternion(R), t


def corners3d(quat, tvec, size):
    """GenericBoxes3D.corners (boxes3d.py:47-64): (R_q . (0.5 * template * [l, w, h]))^T + tvec; size is (W, L, H)."""
    lwh = size[:, [1, 0, 2]]
    tmpl = 0.5 * torch.tensor(BOX3D_CORNER_MAPPING).T  # (8, 3)
    c = lwh[:, None, :] * tmpl[None]
    R = quaternion_to_matrix(quat)
    return torch.matmul(c, R.transpose(1, 2)) + tvec[:, None, :]


def boxes3d_to_rotated_boxes(quat, tvec, size):
    """bev_nms.py:51-96 with pose_cam_global = identity (boxes already global) and VEHICLE_TO_BEV_ROTATION:
    BEV (x, y) = (-Y, -X) of the global frame; top surface = corners [0, 1, 5, 4]."""
    surf = corners3d(quat, tvec, size)[:, [0, 1, 5, 4], :]
    bev = torch.stack([-surf[..., 1], -surf[..., 0]], -1)  # [[0,-1,0],[-1,0,0],[0,0,-1]] @ p, first two rows
    length = (bev[:, 0] - bev[:, 3]).norm(dim=1)
    width = (bev[:, 0] - bev[:, 1]).norm(dim=1)
    center = bev[:, [0, 2]].mean(dim=1)
    fwd = bev[:, 0] - bev[:, 3]
    angle = torch.atan2(fwd[:, 0], fwd[:, 1]) * (180.0 / math.pi)
    return torch.stack([center[:, 0], center[:, 1], width, length, angle], 1)


# ------------------------------------------------------------------------------------------------ rotated IoU
def _vertices(box):
    x, y, w, h, a = (float(v) for v in box)
    th = a * math.pi / 180.0
    c, s = math.cos(th) * 0.5, math.sin(th) * 0.5
    p0 = (x + s * h + c * w, y + c * h - s * w)
    p1 = (x - s * h + c * w, y - c * h - s * w)
    return [p0, p1, (2 * x - p0[0], 2 * y - p0[1]), (2 * x - p1[0], 2 * y - p1[1])]


def _cross(a, b):
    return a[0] * b[1] - a[1] * b[0]


def _dot(a, b):
    return a[0] * b[0] + a[1] * b[1]


def _intersection_points(p1, p2):
    pts = []
    v1 = [(p1[(i + 1) % 4][0] - p1[i][0], p1[(i + 1) % 4][1] - p1[i][1]) for i in range(4)]
    v2 = [(p2[(i + 1) % 4][0] - p2[i][0], p2[(i + 1) % 4][1] - p2[i][1]) for i in range(4)]
    for i in range(4):  # edge x edge
        for j in range(4):
            det = _cross(v2[j], v1[i])
            if abs(det) <= 1e-14:
                continue
            v12 = (p2[j][0] - p1[i][0], p2[j][1] - p1[i][1])
            t1 = _cross(v2[j], v12) / det
            t2 = _cross(v1[i], v12) / det
            if 0.0 <= t1 <= 1.0 and 0.0 <= t2 <= 1.0:
                pts.append((p1[i][0] + v1[i][0] * t1, p1[i][1] + v1[i][1] * t1))

    def inside(pa, pb, vb):  # vertices of a inside rectangle b
        AB, DA = vb[0], vb[3]
        ABdotAB, ADdotAD = _dot(AB, AB), _dot(DA, DA)
        for q in pa:
            AP = (q[0] - pb[0][0], q[1] - pb[0][1])
            APdotAB, APdotAD = _dot(AP, AB), -_dot(AP, DA)
            if APdotAB >= 0 and APdotAD >= 0 and APdotAB <= ABdotAB and APdotAD <= ADdotAD:
                pts.append(q)

    inside(p1, p2, v2)
    inside(p2, p1, v1)
    return pts


def _convex_hull_graham(p):
    n = len(p)
    t = min(range(n), key=lambda i: (p[i][1], p[i][0]))
    start = p[t]
    q = [(x - start[0], y - start[1]) for x, y in p]
    q[0], q[t] = q[t], q[0]
    dist = [_dot(v, v) for v in q]

    def cmp_key(idx):
        return idx

    import functools

    def cmp(i, j):
        tmp = _cross(q[i], q[j])
        if abs(tmp) < 1e-6:
            return -1 if dist[i] < dist[j] else (1 if dist[i] > dist[j] else 0)
        return -1 if tmp > 0 else 1

    order = [0] + sorted(range(1, n), key=functools.cmp_to_key(cmp))
    q = [q[i] for i in order]
    dist = [dist[i] for i in order]
    k = 1
    while k < n and dist[k] <= 1e-8:
        k += 1
    if k == n:
        return [q[0]]
    hull = [q[0], q[k]]
    for i in range(k + 1, n):
        while len(hull) > 1 and _cross((q[i][0] - hull[-2][0], q[i][1] - hull[-2][1]),
                                       (hull[-1][0] - hull[-2][0], hull[-1][1] - hull[-2][1])) >= 0:
            hull.pop()
        hull.append(q[i])
    return hull


def _polygon_area(q):
    if len(q) <= 2:
        return 0.0
    a = 0.0
    for i in range(1, len(q) - 1):
        a += abs(_cross((q[i][0] - q[0][0], q[i][1] - q[0][1]), (q[i + 1][0] - q[0][0], q[i + 1][1] - q[0][1])))
    return a / 2.0


def rotated_iou(b1, b2):
    """detectron2 single_box_iou_rotated: boxes (cx, cy, w, h, angle_deg)."""
    a1, a2 = float(b1[2]) * float(b1[3]), float(b2[2]) * float(b2[3])
    if a1 < 1e-14 or a2 < 1e-14:
        return 0.0
    sx, sy = (float(b1[0]) + float(b2[0])) / 2.0, (float(b1[1]) + float(b2[1])) / 2.0  # centre shift for precision
    c1 = (float(b1[0]) - sx, float(b1[1]) - sy, float(b1[2]), float(b1[3]), float(b1[4]))
    c2 = (float(b2[0]) - sx, float(b2[1]) - sy, float(b2[2]), float(b2[3]), float(b2[4]))
    pts = _intersection_points(_vertices(c1), _vertices(c2))
    if len(pts) <= 2:
        return 0.0
    inter = _polygon_area(_convex_hull_graham(pts))
    return inter / (a1 + a2 - inter)


def nms_rotated(boxes, scores, classes, thr):
    """batched_nms_rotated: per-class greedy NMS in descending-score order (stable); suppress when IoU > thr.
    Returns kept indices in descending-score order."""
    order = torch.argsort(scores, descending=True, stable=True).tolist()
    removed = set()
    keep = []
    for a, i in enumerate(order):
        if i in removed:
            continue
        keep.append(i)
        for j in order[a + 1:]:
            if j in removed or int(classes[j]) != int(classes[i]):
                continue
            if rotated_iou(boxes[i], boxes[j]) > thr:
                removed.add(j)
    return torch.tensor(keep, dtype=torch.long)


def bev_nms_image(det, pose_quat, pose_tvec, thr):
    """core.py:137-151 for ONE image (dummy group = the image itself): det is the oracle's per-image dict after 2-D NMS
    (needs quat, tvec, size, score3d, cls).  Returns the kept indices as a sorted index tensor (mask semantics:
    survivors keep their original order, postprocessing.py:99-107)."""
    if det["quat"].shape[0] == 0:
        return torch.zeros(0, dtype=torch.long)
    q, t = to_global(det["quat"], det["tvec"], pose_quat, pose_tvec)
    rb = boxes3d_to_rotated_boxes(q, t, det["size"])
    keep = nms_rotated(rb, det["score3d"], det["cls"], thr)
    return torch.sort(keep).values


def sample_aggregate(dets, group_ids, poses, thr, max_dets=None):
    """NuscenesDD3D sample aggregation (nuscenes_dd3d.py:449-463 -> postprocessing.py:58-108): the detections of all
    images of the call are concatenated in list order, classes are offset per sample group so that only boxes of the
    same sample AND class compete (postprocessing.py:80-84), ONE rotated NMS in descending scores_3d order runs over the
    lot (:87-89), the first `max_dets` survivors OF THE WHOLE CALL are kept (:92-93 -- the reference truncates across
    all groups of the call, not per group; restated as is) and the survivors are split back per image in their original
    order (:99-107).

    dets: per-image dicts (quat, tvec, size, score3d, cls, ...); group_ids: per-image group index (images of one sample
    share it); poses: per-image (quat wxyz, tvec) global camera pose.  Returns the filtered per-image dicts with
    `quat_global` / `tvec_global` (pred_boxes3d_global, include_boxes3d_global=True) added."""
    num_classes_off = 1 + max([int(d["cls"].max()) for d in dets if d["cls"].numel()] + [0])
    qs, ts, rbs, cat_ids, scores, img_ids = [], [], [], [], [], []
    for i, (d, (pq, pt)) in enumerate(zip(dets, poses)):
        n = d["quat"].shape[0]
        if n:
            q, t = to_global(d["quat"], d["tvec"], pq, pt)
        else:
            q, t = torch.zeros(0, 4), torch.zeros(0, 3)
        qs.append(q)
        ts.append(t)
        rbs.append(boxes3d_to_rotated_boxes(q, t, d["size"]) if n else torch.zeros(0, 5))
        cat_ids.append(d["cls"].to(torch.long) + group_ids[i] * num_classes_off)
        scores.append(d["score3d"])
        img_ids.append(torch.full((n, ), i, dtype=torch.long))
    rb, cid, sc, img = torch.cat(rbs), torch.cat(cat_ids), torch.cat(scores), torch.cat(img_ids)
    keep = nms_rotated(rb, sc, cid, thr)
    if max_dets:
        keep = keep[:max_dets]
    mask = torch.zeros(rb.shape[0], dtype=torch.bool)
    mask[keep] = True
    out, off = [], 0
    for i, d in enumerate(dets):
        n = d["quat"].shape[0]
        m = mask[off:off + n]
        o = {k: v[m] for k, v in d.items()}
        o["quat_global"], o["tvec_global"] = qs[i][m], ts[i][m]
        out.append(o)
        off += n
    return out
