"""CPU: pin the oracles (oracle/savi_numpy.py, oracle/savi_torch.py) against the
reference-generated fixtures in tests/golden/ (made by tests/golden/make_golden.py)."""
import numpy as np
import pytest
import torch

from oracle import savi_numpy as O
from oracle import savi_torch as OT
from tests._util import FIXTURES, err, grad_scale, load_fixture


@pytest.mark.parametrize("name", FIXTURES)
@pytest.mark.parametrize("folded", [False, True])
def test_numpy_forward_matches_reference(name, folded):
    fx = load_fixture(name)
    x = fx["x"].astype(np.float64)
    s, a = O.forward(fx["params"], x, fx["noise"].astype(np.float64), fx["I"], fx["heads"], folded=folded)
    tol = 1e-12 if fx["sub"] == 1 else 2e-7     # big fixture is stored rounded to fp32
    assert err(s, fx["slots_f64"]) < tol
    assert err(a[:, :, ::fx["sub"]], fx["attn_f64"]) < tol
    assert np.allclose(a.sum(-1), 1.0, atol=1e-12)


@pytest.mark.parametrize("name", FIXTURES)
def test_numpy_backward_matches_reference_autograd(name):
    fx = load_fixture(name)
    x = fx["x"].astype(np.float64)
    _, _, sv = O.forward(fx["params"], x, fx["noise"].astype(np.float64), fx["I"], fx["heads"], keep=True)
    ga = None if fx["g_attn"] is None else fx["g_attn"].astype(np.float64)
    dx, G, _ = O.backward(fx["params"], sv, fx["g_slots"].astype(np.float64), ga)
    tol = 1e-11 if fx["sub"] == 1 else 5e-7
    assert err(dx[:, :, ::fx["sub"]], fx["dx_f64"]) < tol
    gs = grad_scale(fx["grads"])
    assert set(G) == set(fx["grads"])
    for k, g in fx["grads"].items():
        assert np.abs(G[k] - g).max() / gs < tol, k


@pytest.mark.parametrize("name", FIXTURES)
def test_torch_port_matches_reference(name):
    fx = load_fixture(name)
    P = {k: torch.from_numpy(v).double() for k, v in fx["params"].items()}
    t = lambda a: None if a is None else torch.from_numpy(a).double()
    s, a, dx, G = OT.forward_backward(P, t(fx["x"]), t(fx["noise"]), fx["I"], fx["heads"], t(fx["g_slots"]), t(fx["g_attn"]))
    tol = 1e-11 if fx["sub"] == 1 else 5e-7
    assert err(s.numpy(), fx["slots_f64"]) < tol
    assert err(a.numpy()[:, :, ::fx["sub"]], fx["attn_f64"]) < tol
    assert err(dx.numpy()[:, :, ::fx["sub"]], fx["dx_f64"]) < tol
    gs = grad_scale(fx["grads"])
    for k, g in fx["grads"].items():
        assert np.abs(G[k].numpy() - g).max() / gs < tol, k


def test_c1_reference_fp32_noise_floor():
    """The reference's own fp32 CPU forward vs its fp64 forward: fixes the achievable tolerance."""
    fx = load_fixture("c1")
    assert err(fx["slots_f32"], fx["slots_f64"]) < 5e-6
    assert err(fx["attn_f32"], fx["attn_f64"]) < 5e-6


def test_norm_slots_bias_grad_is_structurally_zero():
    fx = load_fixture("tiny_a")
    assert np.abs(fx["grads"]["norm_slots.bias"]).max() < 1e-12 * grad_scale(fx["grads"]) + 1e-13


def test_weight_fold_algebra_and_chain_rule():
    """The tcgen05 kernels multiply by the folded weights wqk = Ds^-1/2 Wk^T Wq and wg = Wih Wv and map the
    gradients of the folded weights back to the four parameters (savi_wgrad.cu: launch_fold_grads).  fp64 check of both
    identities against the two-step products and their autograd-free closed form."""
    rng = np.random.default_rng(11)
    R, D, Ds = 40, 12, 8
    s = Ds ** -0.5
    Wq, Wk, Wv = rng.standard_normal((Ds, Ds)), rng.standard_normal((Ds, D)), rng.standard_normal((Ds, D))
    Wih = rng.standard_normal((3 * Ds, Ds))
    st, ux = rng.standard_normal((R, Ds)), rng.standard_normal((R, D))          # s~ rows, Ux rows (all steps and clips stacked)
    dqk, dgi = rng.standard_normal((R, D)), rng.standard_normal((R, 3 * Ds))    # upstream gradients of qk and gi
    # forward: two-step (reference order) vs folded
    wqk = s * Wk.T @ Wq                     # [D, Ds]
    wg = Wih @ Wv                           # [3Ds, D]
    np.testing.assert_allclose(s * (st @ Wq.T) @ Wk, st @ wqk.T, rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose((ux @ Wv.T) @ Wih.T, ux @ wg.T, rtol=1e-12, atol=1e-12)
    # backward of the two-step form
    q = st @ Wq.T
    dq = s * dqk @ Wk.T
    dWq_ref, dWk_ref = dq.T @ st, s * q.T @ dqk
    U = ux @ Wv.T
    dU = dgi @ Wih
    dWih_ref, dWv_ref = dgi.T @ U, dU.T @ ux
    # backward through the folded weights + chain rule (what the CUDA path computes)
    dwqk, dwg = dqk.T @ st, dgi.T @ ux
    np.testing.assert_allclose(s * Wk @ dwqk, dWq_ref, rtol=1e-11, atol=1e-11)
    np.testing.assert_allclose(s * Wq @ dwqk.T, dWk_ref, rtol=1e-11, atol=1e-11)
    np.testing.assert_allclose(dwg @ Wv.T, dWih_ref, rtol=1e-11, atol=1e-11)
    np.testing.assert_allclose(Wih.T @ dwg, dWv_ref, rtol=1e-11, atol=1e-11)
    # and the activations' gradients
    np.testing.assert_allclose(dqk @ wqk, dq @ Wq, rtol=1e-11, atol=1e-11)      # d s~
    np.testing.assert_allclose(dgi @ wg, dU @ Wv, rtol=1e-11, atol=1e-11)       # d Ux
