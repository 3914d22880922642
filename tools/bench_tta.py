"""Times DD3DB200WithTTA (test-time augmentation fully on the device) on one synthetic nuScenes-sized image:
10 views (5 scales x flip) -> merged detections.  Prints one JSON line.  python tools/bench_tta.py [--arch v2_99]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from dd3d_b200.config import get_cfg  # noqa: E402
from dd3d_b200.meta_arch import DD3DB200  # noqa: E402
from dd3d_b200.synthetic import make_inputs, make_state_dict  # noqa: E402
from dd3d_b200.tta import DD3DB200WithTTA  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--arch", default="v2_99", choices=["v2_99", "dla34"])
    ap.add_argument("--images", type=int, default=4)
    args = ap.parse_args()
    ds, H, W, focal = ("nuscenes", 896, 1593, 1266.4) if args.arch == "v2_99" else ("kitti_3d", 384, 1272, 721.5)
    cfg = get_cfg(args.arch, ds)
    cfg.DD3D.INFERENCE.DO_POSTPROCESS = False
    model = DD3DB200(cfg).to("cuda")
    model.load_state_dict(make_state_dict(cfg))
    tta = DD3DB200WithTTA(cfg, model, world_size=8)  # TEST.IMS_PER_BATCH // 8 views per model call, as shipped
    inputs = make_inputs(args.images + 1, H, W, focal)
    for x in inputs:
        x["height"], x["width"] = (900, 1600) if args.arch == "v2_99" else (375, 1242)
    tta([inputs[0]])  # warm-up: plans, resize tables
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n = 0
    for x in inputs[1:]:
        n += len(tta([x])[0]["instances"])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.images
    views = len(cfg.TEST.AUG.MIN_SIZES) * 2
    print(json.dumps({"metric": "tta_ms_per_image", "value": ms, "views_per_image": views,
                      "views_per_s": views / (ms * 1e-3), "arch": args.arch, "mapped_size": [H, W],
                      "min_sizes": list(cfg.TEST.AUG.MIN_SIZES), "views_per_model_call": tta.batch_size,
                      "merged_detections_per_image": n / args.images, "overflow_flags": tta.overflow_flags()}))


if __name__ == "__main__":
    main()
