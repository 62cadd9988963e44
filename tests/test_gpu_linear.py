"""GPU numerics: fp32 dense layer kernels (fwd / dgrad / wgrad) vs torch CPU fp32 autograd."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

ACTS = {"none": lambda x: x, "tanh": torch.tanh, "leaky_relu": F.leaky_relu, "elu": F.elu}


def _tol(ref):
    return dict(rtol=1e-5, atol=1e-5 * max(1.0, float(ref.abs().max())))


@pytest.mark.parametrize("M,K,N,act", [(1, 1, 1, "none"), (5, 3, 2, "tanh"), (128, 8, 128, "tanh"), (257, 64, 64, "tanh"),
                                       (1000, 66, 37, "leaky_relu"), (513, 530, 512, "elu"), (64, 28224, 256, "leaky_relu"),
                                       (4096, 64, 1, "none"), (300, 128, 4, "none"), (77, 100, 3, "tanh")])
def test_linear_fwd_bwd(M, K, N, act):
    from ppo_exploration_b200 import models as PM
    torch.manual_seed(M + K + N)
    x = torch.randn(M, K)
    w = (torch.randn(K, N) / np.sqrt(K)).requires_grad_(True)           # in-major
    b = torch.randn(N).requires_grad_(True)
    xr = x.clone().requires_grad_(True)
    y_ref = ACTS[act](xr @ w + b)
    dy = torch.randn(M, N)
    y_ref.backward(dy)

    dev = "cuda"
    xd, wd, bd, dyd = x.to(dev), w.detach().to(dev), b.detach().to(dev), dy.to(dev)
    y = torch.empty(M, N, device=dev)
    PM.linear_fwd(xd.data_ptr(), K, wd.data_ptr(), bd.data_ptr(), M, K, N, PM.ACT[act], y.data_ptr(), N)
    torch.testing.assert_close(y.cpu(), y_ref.detach(), **_tol(y_ref.detach()))
    # pre-activation gradient (what the previous layer's dgrad epilogue would have produced)
    dpre = torch.autograd.grad(ACTS[act](xr @ w + b), [xr], dy, retain_graph=False)[0] if False else None
    pre = (x @ w.detach() + b.detach())
    pre.requires_grad_(True)
    ACTS[act](pre).backward(dy)
    dpre = pre.grad
    dpre_d = dpre.to(dev).contiguous()
    dw = torch.empty(K, N, device=dev); db = torch.empty(N, device=dev)
    sc = PM._Scratch(torch.device(dev))
    PM.linear_bwd_weight(sc, xd.data_ptr(), K, dpre_d.data_ptr(), N, M, K, N, dw.data_ptr(), db.data_ptr())
    torch.testing.assert_close(dw.cpu(), w.grad, **_tol(w.grad))
    torch.testing.assert_close(db.cpu(), b.grad, **_tol(b.grad))
    dx = torch.empty(M, K, device=dev)
    PM.linear_bwd_data(dpre_d.data_ptr(), N, wd.data_ptr(), M, K, N, None, K, 0, dx.data_ptr(), K)
    torch.testing.assert_close(dx.cpu(), xr.grad, **_tol(xr.grad))


@pytest.mark.parametrize("act", ["tanh", "leaky_relu", "elu"])
def test_dgrad_activation_epilogue(act):
    from ppo_exploration_b200 import models as PM
    torch.manual_seed(0)
    M, K, N = 200, 48, 24
    pre = torch.randn(M, K, requires_grad=True)
    h = ACTS[act](pre)
    w = torch.randn(K, N)
    dy = torch.randn(M, N)
    (h @ w).backward(dy)
    dev = "cuda"
    dx = torch.empty(M, K, device=dev)
    hd, wd, dyd = h.detach().to(dev), w.to(dev), dy.to(dev)
    PM.linear_bwd_data(dyd.data_ptr(), N, wd.data_ptr(), M, K, N, hd.data_ptr(), K, PM.ACT[act], dx.data_ptr(), K)
    torch.testing.assert_close(dx.cpu(), pre.grad, **_tol(pre.grad))


def test_strided_batched_matches_loop():
    """actor|critic|int_critic second layers as one strided-batched launch."""
    from ppo_exploration_b200 import models as PM
    torch.manual_seed(1)
    M, h, G = 333, 64, 3
    H1 = torch.randn(M, G * h)
    W2, b2 = torch.randn(G, h, h) / 8, torch.randn(G, h)
    dev = "cuda"
    H1d, W2d, b2d = H1.to(dev), W2.to(dev), b2.to(dev)
    H2 = torch.empty(M, G * h, device=dev)
    PM.linear_fwd(H1d.data_ptr(), G * h, W2d.data_ptr(), b2d.data_ptr(), M, h, h, PM.ACT["tanh"], H2.data_ptr(), G * h,
                  batch=G, sx=h, sw=h * h, sb=h, sy=h)
    ref = torch.cat([torch.tanh(H1[:, g * h:(g + 1) * h] @ W2[g] + b2[g]) for g in range(G)], dim=1)
    torch.testing.assert_close(H2.cpu(), ref, **_tol(ref))
    dP2 = torch.randn(M, G * h)
    dW2 = torch.empty(G, h, h, device=dev); db2 = torch.empty(G, h, device=dev)
    sc = PM._Scratch(torch.device(dev))
    dP2d = dP2.to(dev)
    PM.linear_bwd_weight(sc, H1d.data_ptr(), G * h, dP2d.data_ptr(), G * h, M, h, h, dW2.data_ptr(), db2.data_ptr(),
                         batch=G, sx=h, sdy=h, sdw=h * h, sdb=h)
    refW = torch.stack([H1[:, g * h:(g + 1) * h].t() @ dP2[:, g * h:(g + 1) * h] for g in range(G)])
    torch.testing.assert_close(dW2.cpu(), refW, **_tol(refW))
    torch.testing.assert_close(db2.cpu(), dP2.view(M, G, h).sum(0), **_tol(refW))


def test_wgrad_is_deterministic():
    from ppo_exploration_b200 import models as PM
    torch.manual_seed(2)
    M, K, N = 131072, 64, 64
    x, dy = torch.randn(M, K, device="cuda"), torch.randn(M, N, device="cuda")
    sc = PM._Scratch(torch.device("cuda"))
    outs = []
    for _ in range(2):
        dw = torch.empty(K, N, device="cuda"); db = torch.empty(N, device="cuda")
        PM.linear_bwd_weight(sc, x.data_ptr(), K, dy.data_ptr(), N, M, K, N, dw.data_ptr(), db.data_ptr())
        outs.append((dw.clone(), db.clone()))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    ref = (x.double().t() @ dy.double()).float()
    torch.testing.assert_close(outs[0][0], ref, rtol=1e-4, atol=1e-2)
