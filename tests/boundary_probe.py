"""Run as a subprocess by tests/test_boundary_reference_plumbing.py (build container only: needs /root/reference).

Drives the B200 mirror through the REFERENCE's own plumbing, everything imported unmodified from /root/reference under
the third-party stand-in (oracle/ref_standin.py):

 1. registry: with the stand-in installed BEFORE dd3d_b200 is imported, `dd3d_b200.meta_arch` registers DD3DB200 /
    NuscenesDD3DB200 in the same META_ARCH_REGISTRY the reference's DD3D registers in (core.py:18), and the model is built
    the way scripts/train.py:48 -> detectron2 build_model does: META_ARCH_REGISTRY.get(cfg.MODEL.META_ARCHITECTURE)(cfg),
    .to(cfg.MODEL.DEVICE) -- selected purely by the config string;
 2. types: the Instances it returns are the (stand-in) detectron2 Instances / Boxes and the reference's own Boxes3D;
 3. evaluator: the reference's KITTI3DEvaluator.process() (kitti_3d_evaluator.py:66-123: per-detection iteration of
    pred_boxes3d, .vectorize(), convert_3d_box_to_kitti) consumes those outputs; its JSON / KITTI rows are identical to the
    ones it produces from the reference model's own outputs for the same detections.
The detections fed through DD3DB200._wrap (the product code that turns the C ABI's packed [B][cap][24] buffer into
Instances) are the reference DD3D's own detections on the dla34 golden case, packed into the C-ABI layout -- there is no GPU
in the build container, and no reference on the GPU box.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.abspath(ROOT))
from oracle import ref_standin  # noqa: E402

ref_standin.install_evaluator_stubs()
from detectron2.modeling.meta_arch.build import META_ARCH_REGISTRY  # noqa: E402  (stand-in registry)
from tridet.modeling.dd3d.core import DD3D as RefDD3D  # noqa: E402  (registers the reference's DD3D)
import dd3d_b200.meta_arch as ma  # noqa: E402  registers DD3DB200 / NuscenesDD3DB200
from dd3d_b200 import lib  # noqa: E402
from dd3d_b200.config import get_cfg  # noqa: E402
from dd3d_b200.synthetic import make_state_dict  # noqa: E402
from oracle.gen_golden import case_inputs  # noqa: E402


def build_model(cfg):
    """detectron2.modeling.build_model, restated: registry lookup by config string, then .to(device)."""
    model = META_ARCH_REGISTRY.get(cfg.MODEL.META_ARCHITECTURE)(cfg)
    model.to(torch.device(cfg.MODEL.DEVICE))
    return model


def pack_words(outputs, cap):
    """reference Instances -> the C ABI's [B][cap][24] fp32 words + counts (include/dd3d_b200.h dd3d_det)."""
    B = len(outputs)
    out = torch.zeros(B, cap, lib.DET_WORDS, dtype=torch.float32)
    oi = out.view(torch.int32)
    counts = torch.zeros(B, dtype=torch.int32)
    for b, o in enumerate(outputs):
        inst = o["instances"]
        n = len(inst)
        counts[b] = n
        b3 = inst.pred_boxes3d
        out[b, :n, 0:4] = inst.pred_boxes.tensor
        out[b, :n, 4] = inst.scores
        out[b, :n, 5] = inst.scores_3d
        oi[b, :n, 6] = inst.pred_classes.to(torch.int32)
        oi[b, :n, 7] = inst.fpn_levels.to(torch.int32)
        out[b, :n, 8:12] = b3.quat
        out[b, :n, 12:14] = b3.proj_ctr
        out[b, :n, 14] = b3.depth.reshape(-1)
        out[b, :n, 15:18] = b3.size
        out[b, :n, 18:20] = inst.locations
    return out, counts


def main():
    # ---- 1. registry / build_model by config string
    assert META_ARCH_REGISTRY.get("DD3D") is RefDD3D
    assert META_ARCH_REGISTRY.get("DD3DB200") is ma.DD3DB200
    assert META_ARCH_REGISTRY.get("NuscenesDD3DB200") is ma.NuscenesDD3DB200
    cfg = get_cfg("dla34", "kitti_3d")
    cfg.MODEL.DEVICE = "cpu"
    sd = make_state_dict(cfg)
    ref_model = build_model(cfg).eval()  # MODEL.META_ARCHITECTURE == "DD3D": the reference's own class
    assert type(ref_model) is RefDD3D
    ref_model.load_state_dict(sd)
    cfg_b = get_cfg("dla34", "kitti_3d")
    cfg_b.MODEL.DEVICE = "cpu"
    cfg_b.MODEL.META_ARCHITECTURE = "DD3DB200"
    model = build_model(cfg_b)
    assert type(model) is ma.DD3DB200
    for attr in ("postprocess_in_inference", "do_nms", "do_bev_nms", "bev_nms_iou_thresh", "only_box2d", "num_classes",
                 "device"):  # what do_test / DD3DWithTTA poke (scripts/train.py:206-209, test_time_augmentation.py:107)
        assert hasattr(model, attr), attr
        assert getattr(model, attr) == getattr(ref_model, attr) or attr == "device", attr
    assert model.backbone.size_divisibility == ref_model.backbone.size_divisibility
    res = model.load_state_dict(ref_model.state_dict())  # Checkpointer(model).load(): the reference's own key set
    assert not res.missing_keys and not res.unexpected_keys
    try:
        model([{"image": torch.zeros(3, 8, 8, dtype=torch.uint8), "intrinsics": torch.eye(3) * 2}])
        raise SystemExit("forward on cpu must fail loudly")
    except RuntimeError as e:
        assert "no CPU path" in str(e)

    # ---- 2. + 3. reference evaluator on the mirror's output containers
    from detectron2.data.catalog import DatasetCatalog, MetadataCatalog
    from types import SimpleNamespace
    inputs = case_inputs("dla34")
    for i, x in enumerate(inputs):
        x["file_name"], x["image_id"] = f"img{i:06d}.png", i
    names = ["Car", "Pedestrian", "Cyclist", "Van", "Truck"]
    DatasetCatalog.register("kitti_probe", lambda: [{"file_name": x["file_name"]} for x in inputs])  # test split: no GT
    MetadataCatalog.register("kitti_probe", SimpleNamespace(thing_classes=names,
                                                            contiguous_id_to_name=dict(enumerate(names))))
    from tridet.evaluators.kitti_3d_evaluator import KITTI3DEvaluator
    with torch.no_grad():
        ref_out = ref_model(inputs)
    assert sum(len(o["instances"]) for o in ref_out) > 10
    words, counts = pack_words(ref_out, model._desc.out_cap)
    K = torch.stack([x["intrinsics"] for x in inputs]).reshape(len(inputs), 9)
    sizes = torch.tensor([[x["image"].shape[-2], x["image"].shape[-1], x.get("height", x["image"].shape[-2]),
                           x.get("width", x["image"].shape[-1])] for x in inputs], dtype=torch.int32)
    b200_out = model._wrap(words, counts, K, sizes, torch.device("cpu"))  # product code: packed buffer -> Instances
    from detectron2.structures import Boxes, Instances
    from tridet.structures.boxes3d import Boxes3D
    for o, r in zip(b200_out, ref_out):
        inst = o["instances"]
        assert type(inst) is Instances and type(inst.pred_boxes) is Boxes and type(inst.pred_boxes3d) is Boxes3D
        assert inst.image_size == r["instances"].image_size
        assert set(inst.get_fields()) == set(r["instances"].get_fields())
    rows = []
    for out in (ref_out, b200_out):
        ev = KITTI3DEvaluator("kitti_probe", iou_thresholds=[0.5, 0.7])
        ev.reset()
        ev.process(inputs, out)  # inference_on_dataset's per-batch call (scripts/train.py:228)
        rows.append((ev._predictions_as_json, ev._predictions_kitti_format))
    (ja, ka), (jb, kb) = rows
    assert len(ja) == len(jb) == sum(int(c) for c in counts)
    for a, b in zip(ja, jb):
        assert a["category"] == b["category"] and a["file_name"] == b["file_name"] and a["image_id"] == b["image_id"]
        np.testing.assert_allclose(a["bbox3d"], b["bbox3d"], rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(a["bbox"], b["bbox"], rtol=1e-6, atol=1e-4)
        assert abs(a["score"] - b["score"]) < 1e-6 and abs(a["score_3d"] - b["score_3d"]) < 1e-6
    for fa, fb in zip(ka, kb):
        assert fa.shape == fb.shape
        assert list(fa[0]) == list(fb[0])  # class names
        np.testing.assert_allclose(fa.iloc[:, 3:].to_numpy(dtype=np.float64), fb.iloc[:, 3:].to_numpy(dtype=np.float64),
                                   rtol=1e-5, atol=1e-5)
    print(f"BOUNDARY_OK detections={len(ja)} kitti_rows={sum(len(f) for f in ka)}")


if __name__ == "__main__":
    main()
