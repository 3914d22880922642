"""Bit-level determinism probe (smoke printed 13 detections in one process and 15 in another).

One process = one engine + one plan + two forwards of a fixed seeded case; prints ONE JSON line with a SHA-1 of the
preprocessed input, of every engine op's bf16 output, of every fp32 head map and of the packed detections (first and
second forward).  Run it in several fresh processes with different settings and diff the lines:

    python tools/determinism_probe.py --case dla34 --fill 0      # workspace zeroed at plan time
    python tools/determinism_probe.py --case dla34 --fill 255    # workspace poisoned with NaN patterns
    DD3D_NO_PDL=1 python tools/determinism_probe.py --case dla34 # no programmatic dependent launch

Cases: dla34 / v2_99 = the golden cases of oracle/gen_golden.py (small, ragged); dla34_full = 2 x 384x1280,
v2_99_full = 1 x 900x1600 (the BASELINE shapes).
"""
import argparse
import hashlib
import json
import os
import sys

import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.abspath(ROOT))
import ctypes as C  # noqa: E402

from dd3d_b200 import lib as _lib  # noqa: E402
from dd3d_b200.config import get_cfg  # noqa: E402
from dd3d_b200.meta_arch import DD3DB200  # noqa: E402
from dd3d_b200.synthetic import make_inputs, make_state_dict  # noqa: E402


def sha(t):
    t = t.contiguous()
    if t.dtype == torch.bfloat16:
        t = t.view(torch.int16)
    return hashlib.sha1(t.cpu().numpy().tobytes()).hexdigest()[:12]


def build_case(name):
    if name in ("dla34", "v2_99"):
        from oracle.gen_golden import CASES, case_inputs  # inputs only (no oracle arithmetic)
        cfg = get_cfg(name, CASES[name][0])
        return cfg, case_inputs(name)
    if name == "dla34_full":
        return get_cfg("dla34", "kitti_3d"), make_inputs(2, 384, 1280, 721.5, with_size=True)
    if name == "v2_99_full":
        return get_cfg("v2_99", "nuscenes"), make_inputs(1, 900, 1600, 1266.4, with_size=True)
    raise SystemExit(f"unknown case {name}")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", default="dla34")
    ap.add_argument("--fill", type=int, default=-1)
    ap.add_argument("--dirty", type=int, default=0, help="MiB of device memory to scribble on and free before planning")
    ap.add_argument("--label", default="")
    a = ap.parse_args()
    cfg, inputs = build_case(a.case)
    sd = make_state_dict(cfg)
    model = DD3DB200(cfg).to("cuda:0")
    model.load_state_dict(sd)
    if a.dirty:
        junk = torch.full((a.dirty << 18, ), float("nan"), device="cuda:0")
        torch.cuda.synchronize()
        del junk
        torch.cuda.empty_cache()
    L = _lib.load()
    h = model._engine()
    _lib.check(L.dd3d_set_option(h, b"workspace_fill", a.fill), h)
    _lib.check(L.dd3d_set_option(h, b"sparse_box3d", 0), h)  # the hashes below include the dense 3-D maps
    res = dict(label=a.label, case=a.case, fill=a.fill, no_pdl=os.environ.get("DD3D_NO_PDL", ""), runs=[])
    for it in range(2):
        out = model(inputs)
        torch.cuda.synchronize()
        r = {"counts": [len(o["instances"]) for o in out]}
        r["input"] = sha(model.get_tensor("input"))
        ops = []
        for i in range(L.dd3d_num_ops(h)):
            s = 0
            while True:
                try:
                    t = model.get_tensor(f"op{i}:{s}")
                except RuntimeError:
                    break
                nan = int(torch.isnan(t.float()).sum())
                ops.append(f"{i}:{s}:{sha(t)}" + (f":nan{nan}" if nan else ""))
                s += 1
        r["ops"] = ops
        for l in range(5):
            for n in ("cls", "box", "b3d"):
                t = model.get_tensor(f"{n}{l}")
                r[f"{n}{l}"] = sha(t) + (":nan" if torch.isnan(t).any() else "")
        dets = torch.cat([torch.cat([o["instances"].pred_boxes.tensor, o["instances"].scores_3d[:, None],
                                     o["instances"].pred_boxes3d.quat], 1) for o in out], 0)
        r["dets"] = sha(dets)
        r["flags"] = model.overflow_flags()
        res["runs"].append(r)
    print("PROBE " + json.dumps(res))


if __name__ == "__main__":
    main()
