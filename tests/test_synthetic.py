import torch

from simpledepthestimation_b200.synthetic import CONFIGS, euler_pose, mono_inputs, motion_inputs


def test_generator_is_deterministic_and_in_range():
    a, b = mono_inputs(2, 32, 64, seed=7), mono_inputs(2, 32, 64, seed=7)
    assert torch.equal(a["img"], b["img"]) and torch.equal(a["depth"][2], b["depth"][2])
    assert 0.0 <= float(a["img"].min()) and float(a["img"].max()) <= 1.0
    assert float(a["depth"][0].min()) > 0.09 and float(a["depth"][0].max()) <= 80.0
    assert [tuple(d.shape) for d in a["depth"]] == [(2, 1, 32, 64), (2, 1, 16, 32), (2, 1, 8, 16), (2, 1, 4, 8)]
    m = motion_inputs(1, 16, 32)
    assert m["motion"].shape == (2, 3, 16, 32) and m["pose_vec"].shape == (2, 6)
    assert set(CONFIGS) == {"cfg1", "cfg2", "cfg3", "cfg4", "cfg5"}


def test_euler_pose_is_a_rigid_transform():
    v = torch.tensor([[0.1, -0.2, 0.3, 0.02, -0.01, 0.03]], dtype=torch.float64)
    T = euler_pose(v)
    R = T[0, :3, :3]
    assert float((R @ R.T - torch.eye(3, dtype=torch.float64)).abs().max()) < 1e-14
    assert abs(float(torch.det(R)) - 1.0) < 1e-14
    assert torch.equal(T[0, :3, 3], v[0, :3])


def test_per_sample_oracle_assembly_equals_the_batch_oracle():
    """tests/test_fullsize_gpu.py evaluates the oracle sample by sample at the full GPU sizes; the assembly (loss =
    mean of the per-sample losses, gradients = per-sample gradients / B) must reproduce the batch oracle."""
    import torch
    from helpers import port_mono_from_vec, rel_err
    from test_fullsize_gpu import _oracle_per_sample

    inp = mono_inputs(3, 32, 64, seed=4)
    full = port_mono_from_vec(inp, torch.float64)
    parts = _oracle_per_sample(inp, torch.float64)
    assert abs(parts["rec_loss"] - float(full["rec_loss"])) < 1e-13
    assert abs(parts["smooth_loss"] - float(full["smooth_loss"])) < 1e-15
    for a, b in zip(parts["grad_depth"] + parts["grad_pose_vec"], full["grad_depth"] + full["grad_pose_vec"]):
        assert rel_err(a, b) < 1e-12
    for a, b in zip(parts["argmin"], full["argmin"]):
        assert torch.equal(a.reshape(b.shape), b)
