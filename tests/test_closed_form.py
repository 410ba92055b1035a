"""Unit checks of the identities the kernels rely on (CPU, fp64)."""
import torch

from oracle import closed_form as cf
from oracle import port


def test_reflect_box_adjoint_is_the_transpose():
    torch.manual_seed(0)
    for (h, w) in [(2, 2), (3, 5), (7, 4), (16, 9)]:
        x = torch.randn(1, 1, h, w, dtype=torch.float64)
        c = torch.randn(1, 1, h, w, dtype=torch.float64)
        lhs = (cf.box_reflect(x) * c).sum()
        rhs = (x * cf.box_reflect_adjoint(c)).sum()
        assert abs(float(lhs - rhs)) < 1e-10 * max(1.0, abs(float(lhs)))


def test_smoothness_homogeneity_shortcut_matches_autograd():
    """L(inv / mean(inv)) = L(inv) / mean(inv) and its gradient."""
    torch.manual_seed(1)
    d = (torch.rand(3, 1, 12, 20, dtype=torch.float64) * 30 + 0.05).requires_grad_()
    img = torch.rand(3, 3, 12, 20, dtype=torch.float64)
    ref = port.smoothness(d, img)
    ref.backward()
    loss, gd = cf.smoothness_fwd_bwd(d.detach()[:, 0], img)
    assert abs(float(loss - ref)) < 1e-14
    assert float((gd - d.grad[:, 0]).abs().max()) < 1e-14


def test_photometric_grad_matches_autograd_with_clamped_ssim():
    torch.manual_seed(2)
    S = torch.rand(2, 3, 9, 11, dtype=torch.float64).requires_grad_()
    A = torch.rand(2, 3, 9, 11, dtype=torch.float64)
    g_pe = torch.rand(2, 9, 11, dtype=torch.float64)
    pe = port.photometric_error(S, A)
    (pe[:, 0] * g_pe).sum().backward()
    gS = cf.photometric_grad_S(S.detach(), A, g_pe, 0.85, 1e-4, 9e-4)
    assert float((gS - S.grad).abs().max()) < 1e-12


def test_identical_source_gives_zero_identity_candidate():
    """Edge case (i) of SURVEY.md App. C: source == target -> identity candidate exactly 0."""
    A = torch.rand(1, 3, 8, 8, dtype=torch.float64)
    assert float(port.photometric_error(A, A).abs().max()) == 0.0
