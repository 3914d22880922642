"""Test-time augmentation (test_time_augmentation.py): CPU pins of the oracle against the reference's
own DD3DWithTTA, GPU parity of dd3d_op_tta_merge and DD3DB200WithTTA."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR
from oracle import tta_oracle as T
from util import quat_dist, rel_err


def _tta_setup():
    from dd3d_b200.synthetic import make_state_dict
    from oracle.gen_golden import tta_case
    cfg, x = tta_case()
    return cfg, x, make_state_dict(cfg)


# ------------------------------------------------------------------------------------------------ CPU
def test_tta_oracle_matches_reference_golden():
    """fp32 oracle TTA == the reference's DD3DWithTTA(DD3D) on the fixture case: same detections in the same order."""
    from oracle.dd3d_oracle import DD3DOracle
    g = np.load(os.path.join(GOLDEN_DIR, "tta_dla34.npz"))
    cfg, x, sd = _tta_setup()
    out = T.tta_forward(DD3DOracle(cfg, sd), x, cfg.TEST.AUG.MIN_SIZES, cfg.TEST.AUG.MAX_SIZE, cfg.TEST.AUG.FLIP,
                        cfg.TEST.IMS_PER_BATCH, cfg.DD3D.FCOS2D.INFERENCE.NMS_THRESH)
    n = g["scores_3d"].shape[0]
    assert n > 100 and out["score3d"].shape[0] == n
    assert tuple(g["image_size"]) == (x["height"], x["width"])
    np.testing.assert_allclose(out["score3d"].numpy(), g["scores_3d"], rtol=1e-4, atol=1e-6)
    assert np.array_equal(out["cls"].numpy(), g["classes"])
    assert rel_err(out["box2d"], g["boxes"], floor=16.0) < 1e-4
    assert quat_dist(out["quat"], g["quat"]).max().item() < 1e-4
    assert rel_err(out["proj_ctr"], g["proj_ctr"], floor=16.0) < 1e-4
    assert rel_err(out["depth"], g["depth"].reshape(-1), floor=1.0) < 1e-4
    assert rel_err(out["inv_K"], g["inv_K"], floor=1e-3) < 1e-3
    # the fixture exercises every branch: flipped and unflipped survivors from every scale
    assert set(out["view"].tolist()) == set(range(6))


def test_build_views_matches_oracle_bookkeeping():
    """Host side of the mirror (dd3d_b200.tta.build_views) == the oracle's restatement of DatasetMapperTTA + the inverse
    intrinsics / factors, bit for bit (fp32)."""
    from dd3d_b200.tta import build_views
    cfg, x, _ = _tta_setup()
    h, w = x["image"].shape[1:]
    orig = (x["height"], x["width"])
    views = build_views((h, w), orig, x["intrinsics"], cfg.TEST.AUG.MIN_SIZES, cfg.TEST.AUG.MAX_SIZE, True)
    ref = T.make_views(x["image"], orig, x["intrinsics"], cfg.TEST.AUG.MIN_SIZES, cfg.TEST.AUG.MAX_SIZE, True)
    assert len(views) == len(ref) == 6
    for (nh, nw, f, v), r in zip(views, ref):
        assert (nh, nw) == r["new_hw"] and f == r["flip"]
        assert np.array_equal(np.float32(list(v.K_view)).reshape(3, 3), r["intrinsics"].numpy())
        one = dict(box2d=torch.tensor([[1.0, 2.0, 30.0, 40.0]]), quat=torch.tensor([[1.0, 0, 0, 0]]),
                   tvec=torch.tensor([[1.0, 2.0, 10.0]]), size=torch.ones(1, 3), score=torch.ones(1), score3d=torch.ones(1),
                   cls=torch.zeros(1, dtype=torch.long))
        inv = T.invert_view(one, r, (h, w), orig)
        K_o = np.linalg.inv(inv["inv_K"][0].numpy().astype(np.float64))
        np.testing.assert_allclose(np.float32(list(v.K_orig)).reshape(3, 3), K_o, rtol=1e-5, atol=1e-4)
        # 2-D inverse chain with the mirror's factors == the oracle's
        x1, x2 = (np.float32(nw) - np.float32(30.0), np.float32(nw) - np.float32(1.0)) if f else (np.float32(1.0), np.float32(30.0))
        for s in range(2):
            x1, x2 = x1 * np.float32(v.inv_sx[s]), x2 * np.float32(v.inv_sx[s])
        assert np.float32(inv["box2d"][0, 0]) == x1 and np.float32(inv["box2d"][0, 2]) == x2


# ------------------------------------------------------------------------------------------------ GPU
def _random_view_dets(seed, n, view, K_view):
    g = torch.Generator().manual_seed(seed)
    nh, nw = view["new_hw"]
    xy = torch.rand(n, 2, generator=g) * torch.tensor([nw * 0.8, nh * 0.8])
    box = torch.cat([xy, xy + torch.rand(n, 2, generator=g) * 40 + 2], 1)
    q = torch.randn(n, 4, generator=g)
    q = q / q.norm(dim=1, keepdim=True)
    pc = torch.rand(n, 2, generator=g) * torch.tensor([float(nw), float(nh)])
    depth = torch.rand(n, generator=g) * 40 + 3
    inv_K = torch.linalg.inv(K_view.double()).float()
    tvec = (torch.cat([pc, torch.ones(n, 1)], 1) @ inv_K.T) * depth[:, None]
    s = torch.rand(n, generator=g).sort(descending=True).values
    return dict(box2d=box, quat=q, proj_ctr=pc, depth=depth, tvec=tvec, size=torch.rand(n, 3, generator=g) + 1,
                score=torch.sqrt(s), score3d=s, cls=torch.randint(0, 3, (n, ), generator=g))


@pytest.mark.gpu
@pytest.mark.parametrize("seed,do_nms", [(0, 1), (1, 1), (2, 0)])
def test_tta_merge_kernel_vs_oracle(seed, do_nms):
    from dd3d_b200 import lib
    from dd3d_b200.tta import build_views
    L = lib.load()
    cfg, x, _ = _tta_setup()
    h, w = x["image"].shape[1:]
    orig = (x["height"], x["width"])
    sizes = [64, 96, 128, 160, 192]
    views = build_views((h, w), orig, x["intrinsics"], sizes, 100000, True)
    ref_views = T.make_views(x["image"], orig, x["intrinsics"], sizes, 100000, True)
    A, cap = len(views), 128
    rs = np.random.RandomState(seed)
    buf = torch.zeros(A, cap, 24)
    counts = torch.zeros(A, dtype=torch.int32)
    inverted = []
    for a, rv in enumerate(ref_views):
        n = int(rs.randint(0, 101)) if a != 3 else 0
        det = _random_view_dets(seed * 100 + a, n, rv, rv["intrinsics"])
        buf[a, :n, 0:4], buf[a, :n, 4], buf[a, :n, 5] = det["box2d"], det["score"], det["score3d"]
        buf[a, :n, 8:12], buf[a, :n, 12:14], buf[a, :n, 14], buf[a, :n, 15:18] = det["quat"], det["proj_ctr"], det["depth"], det["size"]
        buf.view(torch.int32)[a, :n, 6] = det["cls"].to(torch.int32)
        buf.view(torch.int32)[a, :n, 20] = torch.arange(n, dtype=torch.int32) + 1000 * a
        counts[a] = n
        rv["index"] = a
        inverted.append(T.invert_view(det, rv, (h, w), orig))
    ref = T.merge(inverted, 0.75, do_nms=bool(do_nms))
    d_d, d_c = buf.cuda(), counts.cuda()
    mcap = int(L.dd3d_op_tta_merged_cap(A, cap))
    out = torch.zeros(mcap, 24, device="cuda")
    n_out = torch.zeros(1, dtype=torch.int32, device="cuda")
    flags = torch.zeros(1, dtype=torch.int32, device="cuda")
    scratch = torch.empty(int(L.dd3d_op_tta_merge_scratch_bytes(A, cap)), dtype=torch.uint8, device="cuda")
    varr = (lib.TtaView * A)(*[v[3] for v in views])
    st = L.dd3d_op_tta_merge(C.c_void_p(d_d.data_ptr()), C.c_void_p(d_c.data_ptr()), varr, A, cap, 0.75, do_nms,
                             C.c_void_p(scratch.data_ptr()), C.c_void_p(out.data_ptr()), C.c_void_p(n_out.data_ptr()),
                             C.c_void_p(flags.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert st == 0
    torch.cuda.synchronize()
    assert int(flags.item()) == 0
    n = int(n_out.item())
    o = out[:n].cpu()
    assert n == ref["score3d"].shape[0]
    if do_nms:  # survivors in descending scores_3d order
        assert torch.equal(o[:, 5], ref["score3d"])
        order = slice(None)
    else:  # model.do_nms False: plain concatenation in view order (test_time_augmentation.py:163 skipped)
        order = slice(None)
        assert torch.equal(o[:, 5], ref["score3d"])
    assert torch.equal(o[:, 0:4], ref["box2d"][order])  # fp32 chain restated operation for operation
    assert torch.equal(o.view(torch.int32)[:, 7].long(), ref["view"][order])
    assert torch.equal(o[:, 8:12], ref["quat"][order])
    np.testing.assert_allclose(o[:, 12:14].numpy(), ref["proj_ctr"][order].numpy(), rtol=2e-5, atol=2e-3)
    np.testing.assert_allclose(o[:, 14].numpy(), ref["depth"][order].numpy(), rtol=1e-5, atol=1e-5)


@pytest.mark.gpu
def test_tta_forward_vs_emulating_oracle_and_golden():
    from dd3d_b200.meta_arch import DD3DB200
    from dd3d_b200.tta import DD3DB200WithTTA
    from oracle.dd3d_oracle import DD3DOracle
    cfg, x, sd = _tta_setup()
    model = DD3DB200(cfg).to("cuda")
    model.load_state_dict(sd)
    tta = DD3DB200WithTTA(cfg, model)
    assert tta.batch_size == 4
    inst = tta([x])[0]["instances"]
    assert tta.overflow_flags() == 0
    assert tuple(inst.image_size) == (x["height"], x["width"])
    s3 = inst.scores_3d.cpu()
    assert torch.all(s3[:-1] >= s3[1:])  # merged_instances[keep]: descending scores_3d
    emu = T.tta_forward(DD3DOracle(cfg, sd, emulate_bf16=True), x, cfg.TEST.AUG.MIN_SIZES, cfg.TEST.AUG.MAX_SIZE,
                        cfg.TEST.AUG.FLIP, cfg.TEST.IMS_PER_BATCH, cfg.DD3D.FCOS2D.INFERENCE.NMS_THRESH)
    n, m = len(inst), emu["score3d"].shape[0]
    assert abs(n - m) <= 0.1 * m
    # match on (class, rounded box) -- the merged Instances carry no per-pixel provenance
    def keys(box, cls):
        return [(int(c), ) + tuple(int(round(float(v) / 2.0)) for v in b) for b, c in zip(box, cls)]
    ka, kb = keys(inst.pred_boxes.tensor.cpu(), inst.pred_classes.cpu()), keys(emu["box2d"], emu["cls"])
    pos = {k: i for i, k in enumerate(kb)}
    pairs = [(i, pos[k]) for i, k in enumerate(ka) if k in pos]
    assert len(pairs) >= 0.7 * m
    ia, ib = (torch.tensor(p) for p in zip(*pairs))
    assert rel_err(inst.pred_boxes.tensor.cpu()[ia], emu["box2d"][ib], floor=16.0) < 0.1
    assert rel_err(inst.scores_3d.cpu()[ia], emu["score3d"][ib], floor=0.05) < 0.2
    assert rel_err(inst.pred_boxes3d.depth.cpu()[ia].reshape(-1), emu["depth"][ib], floor=1.0) < 0.05
    assert quat_dist(inst.pred_boxes3d.quat.cpu()[ia], emu["quat"][ib]).median().item() < 0.05
    g = np.load(os.path.join(GOLDEN_DIR, "tta_dla34.npz"))
    assert abs(n - g["scores_3d"].shape[0]) <= 0.15 * g["scores_3d"].shape[0]
