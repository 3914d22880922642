"""TEST INFRASTRUCTURE -- generates tests/golden/*.npz by running the REAL reference DD3D.forward
(/root/reference/tridet/modeling/dd3d/core.py, imported unmodified under oracle/ref_standin.py) in the build
container.  /root/reference does not exist on the GPU box, so the vectors are committed.

    python -m oracle.gen_golden

Fixtures (all fp32 reference arithmetic, synthetic calibrated weights seed 0, inputs from
dd3d_b200.synthetic.make_inputs):
  golden_<arch>.npz : per image b -> boxes, scores, scores_3d, classes, levels, locations, quat, proj_ctr, depth, size,
                      tvec (post NMS / top-k / postprocess), plus the case description.
  kat_boxes3d.npz   : known-answer vector for predictions_to_boxes3d (fcos3d.py:16-52) from SURVEY.md 8c.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from dd3d_b200.config import get_cfg  # noqa: E402
from dd3d_b200.synthetic import make_inputs, make_state_dict  # noqa: E402
from oracle import ref_standin  # noqa: E402

CASES = {
    # arch: (dataset, B, H, W, focal, ragged crop of the last image (dh, dw), output size factor of the last image)
    "dla34": ("kitti_3d", 2, 192, 320, 721.5, (21, 34), 2.0),
    "v2_99": ("nuscenes", 2, 128, 192, 1266.4, (0, 0), 1.0),
}


def case_inputs(arch):
    ds, B, H, W, focal, (dh, dw), fac = CASES[arch]
    inputs = make_inputs(B, H, W, focal)
    if dh or dw:
        inputs[-1]["image"] = inputs[-1]["image"][:, :H - dh, :W - dw].contiguous()
    if fac != 1.0:
        inputs[-1]["height"] = int(round((H - dh) * fac))
        inputs[-1]["width"] = int(round((W - dw) * fac))
    return inputs


def main():
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    for arch, (ds, *_rest) in CASES.items():
        cfg = get_cfg(arch, ds)
        model = ref_standin.build_reference_model(cfg).eval()
        model.load_state_dict(make_state_dict(cfg))
        inputs = case_inputs(arch)
        with torch.no_grad():
            outs = model(inputs)
        blob = {}
        for b, o in enumerate(outs):
            inst = o["instances"]
            b3 = inst.pred_boxes3d
            blob.update({
                f"boxes{b}": inst.pred_boxes.tensor.numpy(), f"scores{b}": inst.scores.numpy(),
                f"scores_3d{b}": inst.scores_3d.numpy(), f"classes{b}": inst.pred_classes.numpy(),
                f"levels{b}": inst.fpn_levels.numpy(), f"locations{b}": inst.locations.numpy(),
                f"quat{b}": b3.quat.numpy(), f"proj_ctr{b}": b3.proj_ctr.numpy(), f"depth{b}": b3.depth.numpy(),
                f"size{b}": b3.size.numpy(), f"tvec{b}": b3.tvec.numpy(),
                f"image_size{b}": np.array(inst.image_size),
            })
            print(arch, "image", b, "detections", len(inst))
        np.savez_compressed(os.path.join(out_dir, f"golden_{arch}.npz"), **blob)

    # known-answer test for the 3-D decode, inputs from SURVEY.md 8c (reference function called verbatim)
    ref_standin.install()
    from tridet.modeling.dd3d.fcos3d import predictions_to_boxes3d
    K = torch.tensor([[721.5377, 0, 609.5593], [0, 721.5377, 172.854], [0, 0, 1]])
    quat = torch.tensor([[1, 0, 0, 0], [.5, -.5, .5, .5], [2, .2, -1, .3]])
    ctr = torch.tensor([[0, 0], [12.5, -3.25], [-40, 8]])
    depth = torch.tensor([10, 25.5, 200])
    size = torch.tensor([[0, 0, 0], [.3, -.2, .1], [-1.5, 2, .7]])
    loc = torch.tensor([[609.5593, 172.854], [800, 200], [64, 320]])
    canon = torch.tensor(get_cfg("dla34", "kitti_3d").DD3D.FCOS3D.CANONICAL_BOX3D_SIZES)[:3]
    inv_K = torch.inverse(K)[None].expand(3, 3, 3)
    b3 = predictions_to_boxes3d(quat, ctr, depth, size, loc, inv_K, canon, 0.1, 80.0, 500.0)
    np.savez(os.path.join(out_dir, "kat_boxes3d.npz"), K=K.numpy(), quat_in=quat.numpy(), ctr_in=ctr.numpy(),
             depth_in=depth.numpy(), size_in=size.numpy(), loc=loc.numpy(), canon=canon.numpy(), quat=b3.quat.numpy(),
             proj_ctr=b3.proj_ctr.numpy(), depth=b3.depth.numpy(), size=b3.size.numpy(), tvec=b3.tvec.numpy())
    print("KAT quat", b3.quat.numpy(), "tvec", b3.tvec.numpy())


if __name__ == "__main__":
    main()
