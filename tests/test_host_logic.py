"""Host-side logic that needs no GPU: the lazy overall-motion entry of the MotionLearning output dict, the staging of
the reference's files for the CPU arm, and the reference arm of bench.py (its JSON contract)."""
import hashlib
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_overall_motion_is_formed_on_first_use():
    from simpledepthestimation_b200.modeling.meta_arch.MotionLearning import OverallMotion

    pose = torch.eye(4).repeat(2, 1, 1)
    pose[:, :3, 3] = torch.tensor([[0.1, 0.2, 0.3], [-0.1, 0.0, 0.5]])
    field = torch.rand(2, 3, 4, 5)
    om = OverallMotion(pose, field, (4, 5))
    assert om._t is None                                   # nothing materialised yet
    want = pose[:, :3, 3][:, :, None, None] + field         # MotionLearning.py:143-147
    assert torch.equal(om[0], want[0]) and om.shape == want.shape and torch.equal(om.detach(), want)
    rigid = OverallMotion(pose, None, (4, 5))
    assert rigid.shape == (2, 3, 4, 5) and torch.equal(rigid[1][:, 2, 3], pose[1, :3, 3])


@pytest.mark.skipif(not os.path.isdir("/root/reference/detectron2"), reason="reference tree not present")
def test_make_ref_stages_the_reference_files_unmodified():
    subprocess.run(["bash", os.path.join(ROOT, "oracle", "make_ref.sh")], check=True, capture_output=True)
    staged = os.path.join(ROOT, "oracle", "_ref", "detectron2")
    n = 0
    for dirpath, _, files in os.walk(staged):
        for f in files:
            a = os.path.join(dirpath, f)
            b = os.path.join("/root/reference/detectron2", os.path.relpath(a, staged))
            assert hashlib.sha256(open(a, "rb").read()).hexdigest() == hashlib.sha256(open(b, "rb").read()).hexdigest(), f
            n += 1
    assert n == 12
    # and it is ignored by git (never enters history)
    out = subprocess.run(["git", "check-ignore", "oracle/_ref/detectron2/geometry/camera.py"], cwd=ROOT, capture_output=True, text=True)
    assert out.stdout.strip() != ""


def test_bench_reference_arm_prints_one_contract_line():
    """python bench.py --impl reference: one JSON line with the contract's keys, timed on the host CPU (the reference's own
    files when staged, else the port)."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-500:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "Mpix/s" and d["value"] > 0 and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and "workload" in d["config"]
