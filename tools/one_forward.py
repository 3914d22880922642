"""Dev tool: one (or N) forwards of a bench workload, for ncu captures.  python tools/one_forward.py v2_99 32 [n]"""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from bench import WORKLOADS
from dd3d_b200.config import get_cfg
from dd3d_b200.meta_arch import DD3DB200
from dd3d_b200.synthetic import make_inputs, make_state_dict
wl, B = sys.argv[1], int(sys.argv[2])
n = int(sys.argv[3]) if len(sys.argv) > 3 else 1
arch, ds, _, H, W, focal, _ = WORKLOADS[wl]
cfg = get_cfg(arch, ds)
m = DD3DB200(cfg).to("cuda"); m.load_state_dict(make_state_dict(cfg))
inp = make_inputs(B, H, W, focal)
for _ in range(n):
    out = m(inp)
torch.cuda.synchronize()
print("detections", sum(len(o["instances"]) for o in out))
