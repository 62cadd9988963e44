"""GPU numerics: fused policy-MLP forward/backward (mlp_fused.cu) vs torch CPU fp32 autograd and vs the
layer-by-layer kernels, over the shapes the learners use (models.py:137-213)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _policy(D, h, space, intrinsic):
    import ppo_exploration_b200 as ppx
    env = ppx.SyntheticVecEnv(4, D, space, seed=0)
    torch.manual_seed(D + h)
    return ppx.models.Policy(env, h, intrinsic_model=intrinsic, device="cuda")


def _torch_nets(pol):
    sd = pol.state_dict()
    nets = {}
    for g, o in zip(pol.names, pol.outs):
        seq = torch.nn.Sequential(torch.nn.Linear(pol.state_dim, pol.hidden_size), torch.nn.Tanh(),
                                  torch.nn.Linear(pol.hidden_size, pol.hidden_size), torch.nn.Tanh(),
                                  torch.nn.Linear(pol.hidden_size, o))
        seq.load_state_dict({k[len(g) + 1:]: v for k, v in sd.items() if k.startswith(g + ".")})
        nets[g] = seq
    return nets


def _tol(ref):
    return dict(rtol=1e-5, atol=2e-5 * max(1.0, float(ref.abs().max())))


@pytest.mark.parametrize("M,D,h,box,intrinsic", [(1, 4, 64, False, False), (127, 8, 64, True, False), (128, 8, 64, True, True),
                                                 (1000, 3, 128, True, False), (4099, 32, 64, False, True),
                                                 (131072, 8, 64, True, False), (300, 6, 128, True, True), (515, 17, 64, True, True)])
def test_fused_mlp_matches_autograd(M, D, h, box, intrinsic):
    import ppo_exploration_b200 as ppx
    from ppo_exploration_b200 import models as PM
    space = ppx.Box((2,)) if box else ppx.Discrete(18)
    pol = _policy(D, h, space, intrinsic)
    assert pol.mlp._fused_args()["ok"], "fused path must be the one under test"
    pol.mlp._fused_args()["tc"] = False                      # the SIMT pair (mlp_fused.cu); tensor-core pair: test_gpu_mlp_tc.py
    nets = _torch_nets(pol)
    g = torch.Generator().manual_seed(M)
    x = torch.randn(M, D, generator=g)
    d_outs = [torch.randn(M, o, generator=g) / M for o in pol.outs]
    outs = pol.forward_raw(x.cuda())
    pol.bank.grad.fill_(float("nan"))
    pol.mlp.backward([d.cuda().contiguous() for d in d_outs])
    torch.cuda.synchronize()
    got = pol.state_dict(grad=True)
    for gi, (name, o) in enumerate(zip(pol.names, pol.outs)):
        y = nets[name](x)
        torch.testing.assert_close(outs[gi].cpu(), y.detach(), **_tol(y.detach()))
        y.backward(d_outs[gi])
        for k, prm in nets[name].named_parameters():
            torch.testing.assert_close(got[f"{name}.{k}"], prm.grad, **_tol(prm.grad))
    for name in pol.bank.offsets:
        if name != "action_log_std":
            assert torch.isfinite(pol.bank.view(name, grad=True)).all(), name


def test_fused_matches_layerwise_bitwise_shapes():
    """Same activations layout as the layer-by-layer path (H1/H2 saved [M, G*h]) and close results."""
    import ppo_exploration_b200 as ppx
    from ppo_exploration_b200 import models as PM
    pol = _policy(8, 64, ppx.Box((2,)), True)
    pol.mlp._fused_args()["tc"] = False
    x = torch.randn(777, 8, device="cuda")
    outs_f = [o.clone() for o in pol.forward_raw(x)]
    H1f, H2f = pol.mlp._saved[1].clone(), pol.mlp._saved[2].clone()
    d = [torch.randn_like(o) for o in outs_f]
    pol.mlp.backward(d)
    gf = pol.bank.grad.clone()
    pol.mlp._fa["ok"] = False
    try:
        outs_l = pol.forward_raw(x)
        torch.testing.assert_close(pol.mlp._saved[1], H1f, rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(pol.mlp._saved[2], H2f, rtol=1e-5, atol=1e-6)
        for a, b in zip(outs_f, outs_l):
            torch.testing.assert_close(a, b, rtol=1e-5, atol=1e-5)
        pol.mlp.backward(d)
        n = pol.bank.offsets["action_log_std"]
        torch.testing.assert_close(gf[:n], pol.bank.grad[:n], rtol=1e-4, atol=1e-4 * float(gf[:n].abs().max()))
    finally:
        pol.mlp._fa["ok"] = True
