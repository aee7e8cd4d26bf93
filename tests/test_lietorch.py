"""CPU: the SE3 type layer against the identities of the reference's own test script
(ramp/lietorch/run_tests.py:16-52), float64, atol 1e-8."""
import torch

from rampvo_b200.lietorch import SE3, cat, stack


def test_exp_log():
    torch.manual_seed(0)
    a = 0.5 * torch.randn(64, 6, dtype=torch.float64)
    assert (SE3.exp(a).log() - a).abs().max() < 1e-8          # run_tests.py:16
    tiny = 1e-9 * torch.randn(8, 6, dtype=torch.float64)       # Taylor branches
    assert (SE3.exp(tiny).log() - tiny).abs().max() < 1e-12


def test_inverse_and_mul():
    torch.manual_seed(1)
    X = SE3.Random(32, dtype=torch.float64)
    I = SE3.IdentityLike(X)
    assert ((X * X.inv()).log() - I.log()).abs().max() < 1e-8  # run_tests.py:23
    Y = SE3.Random(32, dtype=torch.float64)
    p = torch.randn(32, 3, dtype=torch.float64)
    assert ((X * Y).act(p) - X.act(Y.act(p))).abs().max() < 1e-8


def test_adjoint_identity():
    torch.manual_seed(2)
    X = SE3.Random(16, dtype=torch.float64)
    a = torch.randn(16, 6, dtype=torch.float64)
    b = 1e-5 * torch.randn(16, 6, dtype=torch.float64)
    # X exp(b) X^-1 = exp(Adj(X) b)  =>  a . Adj(X) b == adjT(a) . b   (run_tests.py:30)
    lhs = (a * (X * SE3.exp(b) * X.inv()).log()).sum(-1)
    rhs = (X.adjT(a) * b).sum(-1)
    assert (lhs - rhs).abs().max() < 1e-8


def test_act_matches_matrix():
    torch.manual_seed(3)
    X = SE3.Random(1, 8, dtype=torch.float64)
    p = torch.randn(1, 8, 4, dtype=torch.float64)
    q1 = X.act(p)
    q2 = (X.matrix() @ p[..., None])[..., 0]                   # run_tests.py:44
    assert (q1 - q2).abs().max() < 1e-8


def test_plumbing():
    X = SE3.Identity(1, 4)
    assert X.shape == (1, 4) and X.data.shape == (1, 4, 7)
    X[:, 1] = SE3.exp(torch.ones(1, 6))
    assert cat([X, X], 1).shape == (1, 8) and stack([X[0], X[0]], 0).shape == (2, 4)
    assert X.retr(torch.zeros(1, 4, 6)).data.allclose(X.data)
    assert X[:, [0, 1]].inv().view((2,)).shape == (2,)
