"""GPU numerics: tensor-core fused policy-MLP forward/backward (mlp_tc.cu, tcgen05 3xTF32) vs torch CPU autograd
(fp32 and fp64) and vs the SIMT fused pair, over the h = 64 shapes the learners use (models.py:137-213)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _policy(D, h, space, intrinsic):
    import ppo_exploration_b200 as ppx
    env = ppx.SyntheticVecEnv(4, D, space, seed=0)
    torch.manual_seed(D + h)
    return ppx.models.Policy(env, h, intrinsic_model=intrinsic, device="cuda")


def _torch_nets(pol, dtype=torch.float32):
    sd = pol.state_dict()
    nets = {}
    for g, o in zip(pol.names, pol.outs):
        seq = torch.nn.Sequential(torch.nn.Linear(pol.state_dim, pol.hidden_size), torch.nn.Tanh(),
                                  torch.nn.Linear(pol.hidden_size, pol.hidden_size), torch.nn.Tanh(),
                                  torch.nn.Linear(pol.hidden_size, o))
        seq.load_state_dict({k[len(g) + 1:]: v for k, v in sd.items() if k.startswith(g + ".")})
        nets[g] = seq.to(dtype)
    return nets


def _tol(ref):
    return dict(rtol=1e-5, atol=2e-5 * max(1.0, float(ref.abs().max())))


SHAPES = [(1, 4, True, False), (127, 8, True, False), (128, 8, True, True), (129, 8, True, True), (4099, 32, True, True),
          (131072, 8, True, False), (515, 17, True, True), (777, 11, False, False), (20000, 3, False, True)]


@pytest.mark.parametrize("M,D,box,intrinsic", SHAPES)
def test_tc_mlp_matches_autograd(M, D, box, intrinsic):
    import ppo_exploration_b200 as ppx
    space = ppx.Box((2,)) if box else ppx.Discrete(3)
    pol = _policy(D, 64, space, intrinsic)
    assert pol.mlp._fused_args()["tc"], "tensor-core path must be the one under test"
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


def test_tc_mlp_error_vs_fp64_not_worse_than_simt():
    """3xTF32 with per-tile TMEM drains stays fp32-equivalent: its error against an fp64 reference is within 2x of the
    exact-fp32 SIMT pair's (plus a floor), for outputs and for every gradient."""
    import ppo_exploration_b200 as ppx
    M, D = 65536, 8
    pol = _policy(D, 64, ppx.Box((2,)), True)
    nets = _torch_nets(pol, torch.float64)
    g = torch.Generator().manual_seed(3)
    x = torch.randn(M, D, generator=g)
    d_outs = [torch.randn(M, o, generator=g) / M for o in pol.outs]
    ref_out, ref_grad = [], {}
    for gi, name in enumerate(pol.names):
        y = nets[name](x.double())
        y.backward(d_outs[gi].double())
        ref_out.append(y.detach())
        for k, prm in nets[name].named_parameters():
            ref_grad[f"{name}.{k}"] = prm.grad

    def run(tc):
        pol.mlp._fused_args()["tc"] = tc
        outs = [o.clone() for o in pol.forward_raw(x.cuda())]
        pol.mlp.backward([d.cuda().contiguous() for d in d_outs])
        torch.cuda.synchronize()
        grads = pol.state_dict(grad=True)
        e_out = max(float((o.cpu().double() - r).abs().max() / r.abs().max()) for o, r in zip(outs, ref_out))
        e_grad = max(float((grads[k].double() - r).abs().max() / r.abs().max()) for k, r in ref_grad.items())
        return e_out, e_grad

    eo_tc, eg_tc = run(True)
    eo_simt, eg_simt = run(False)
    print(f"rel err vs fp64: tc out {eo_tc:.2e} grad {eg_tc:.2e} | simt out {eo_simt:.2e} grad {eg_simt:.2e}")
    assert eo_tc < max(2 * eo_simt, 2e-6) and eg_tc < max(2 * eg_simt, 4e-6)


def test_tc_mlp_deterministic_and_value_head():
    """Two launches give bit-identical gradients; the in-kernel value-head gradient matches the explicit one."""
    import ppo_exploration_b200 as ppx
    M = 5000
    pol = _policy(8, 64, ppx.Box((2,)), False)
    assert pol.mlp._fused_args()["tc"]
    x = torch.randn(M, 8, device="cuda")
    outs = pol.forward_raw(x)
    v = outs[1].reshape(-1).contiguous()
    ov = v + 0.3 * torch.randn_like(v)
    R = torch.randn_like(v)
    branch = torch.tensor([0.25, 0.75], dtype=torch.float64, device="cuda")
    clip, scale = 0.2, 0.5
    d_act = torch.randn(M, 2, device="cuda") / M
    dd = v - ov
    vc = ov + dd.clamp(-clip, clip)
    passed = ((dd >= -clip) & (dd <= clip)).float()
    d_v = (scale * (0.25 * (-2 * (R - v)) + 0.75 * (-2 * (R - vc)) * passed) / M).reshape(M, 1).contiguous()
    pol.mlp.backward([d_act, d_v])
    g_explicit = pol.bank.grad.clone()
    pol.forward_raw(x)
    pol.mlp.backward([d_act, None], value_heads={1: (v, ov, R, branch.data_ptr(), scale)}, clip_range=clip, B_total=M)
    g_inkernel = pol.bank.grad.clone()
    pol.forward_raw(x)
    pol.mlp.backward([d_act, None], value_heads={1: (v, ov, R, branch.data_ptr(), scale)}, clip_range=clip, B_total=M)
    torch.cuda.synchronize()
    n = pol.bank.offsets["action_log_std"]
    n = n[0] if isinstance(n, tuple) else n
    assert torch.equal(g_inkernel[:n], pol.bank.grad[:n])
    torch.testing.assert_close(g_inkernel[:n], g_explicit[:n], rtol=1e-5, atol=1e-6 * float(g_explicit[:n].abs().max()))


@pytest.mark.parametrize("M,D", [(1, 4), (129, 8), (1000, 17), (4099, 32)])
def test_tc_mlp_stays_inside_its_buffers(M, D):
    """Every scratch buffer of the pair (tile-transposed activations, outputs, backward partials, sum-of-squares
    partials) is followed by a guard band; no kernel of the forward / backward / reduce sequence may touch it
    (ragged last tiles, D not a multiple of 4)."""
    import ppo_exploration_b200 as ppx
    pol = _policy(D, 64, ppx.Box((2,)), True)
    assert pol.mlp._fused_args()["tc"]
    sc, guard, SENT = pol.mlp.scratch, 4096, -12345.0
    taken = {}
    orig = sc.get

    def get(name, numel, dtype=torch.float32):
        b = sc.bufs.get(name)
        if name not in taken or taken[name][1] < numel or b.dtype != dtype:
            b = torch.full((int(numel) + guard,), SENT, dtype=dtype, device=sc.device)
            sc.bufs[name] = b
            taken[name] = (b, int(numel))
        return b

    sc.get = get
    try:
        x = torch.randn(M, D, device="cuda")
        outs = pol.forward_raw(x)
        pol.mlp.backward([torch.randn_like(o) for o in outs], with_sumsq=True)
        torch.cuda.synchronize()
    finally:
        sc.get = orig
    assert {"pmlp.H1t", "pmlp.H2t", "pmlp.fused_ws_tc"} <= set(taken)
    for name, (b, n) in taken.items():
        assert bool((b[n:] == SENT).all()), f"{name}: wrote past its {n} elements"
