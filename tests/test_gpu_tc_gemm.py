"""GPU numerics: tcgen05 3xTF32 dense-layer kernel vs an fp64 reference (fp32-equivalent accuracy)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

ACTS = {0: lambda x: x, 1: torch.tanh, 2: F.leaky_relu, 3: F.elu}


def _run(M, R, N, act, dgrad, seed=0):
    from ppo_exploration_b200 import _lib as L
    g = torch.Generator(device="cuda").manual_seed(seed)
    A = torch.randn(M, R, device="cuda", generator=g)
    W = torch.randn(N, R, device="cuda", generator=g) / np.sqrt(R)           # B operand [N, R], R contiguous
    bias = torch.randn(N, device="cuda", generator=g)
    H = torch.tanh(torch.randn(M, N, device="cuda", generator=g))
    hi, lo = torch.empty_like(W), torch.empty_like(W)
    L.call("ppx_tc_split", W.data_ptr(), N, R, hi.data_ptr(), lo.data_ptr(), None, None, L.stream())
    assert torch.equal(hi + lo, W) and torch.equal(hi.view(torch.int32) & 0x1FFF, torch.zeros_like(hi, dtype=torch.int32))
    C = torch.full((M, N), float("nan"), device="cuda")
    assert L.call("ppx_tc_supported", M, R, N, R, R, A.data_ptr(), hi.data_ptr()) == 1
    L.call("ppx_tc_linear", A.data_ptr(), R, hi.data_ptr(), lo.data_ptr(), R, M, R, N, bias.data_ptr(), H.data_ptr(), N, act,
           dgrad, None, None, 0.0, C.data_ptr(), N, L.stream())
    torch.cuda.synchronize()
    acc = A.double() @ W.double().t()
    if dgrad:
        d = {0: torch.ones_like(H), 1: 1 - H * H, 2: torch.where(H > 0, 1.0, 0.01), 3: torch.where(H > 0, 1.0, H + 1)}[act]
        ref = acc * d.double()
    else:
        ref = ACTS[act](acc + bias.double())
    return C.double(), ref, acc


@pytest.mark.parametrize("M,R,N,act,dgrad", [(128, 32, 64, 0, 0), (128, 64, 64, 1, 0), (1000, 8, 128, 1, 0), (300, 64, 100, 2, 0),
                                             (131072, 64, 64, 1, 0), (131072, 64, 64, 1, 1), (257, 100, 16, 3, 1),
                                             (64, 36, 40, 0, 0), (4096, 3136, 512, 2, 0), (777, 28224, 256, 2, 0), (300, 1000, 128, 1, 1)])
def test_tc_linear_matches_fp64(M, R, N, act, dgrad):
    C, ref, acc = _run(M, R, N, act, dgrad)
    assert torch.isfinite(C).all()
    scale = float(acc.abs().max())
    err = float((C - ref).abs().max())
    # 3xTF32: dropped lo*lo term ~2^-22 per product + fp32 accumulation -> well inside 1e-5 of the output scale
    assert err <= 4e-6 * max(scale, 1.0), (err, scale)


@pytest.mark.parametrize("M,R,N,act,dgrad", [(4096, 28224, 128, 2, 0), (4096, 28224, 384, 0, 0), (1024, 3136, 256, 1, 0),
                                             (1023, 1024, 100, 3, 0), (512, 2048, 64, 1, 1)])
def test_tc_linear_split_k_matches_fp64_and_is_deterministic(M, R, N, act, dgrad):
    """Few output tiles + a long reduction (the wide first layers at minibatch size): ppx_tc_linear_ws cuts the reduction
    across CTAs.  Same accuracy bound as the single-pass kernel, bit-identical from run to run (fixed-order finish)."""
    from ppo_exploration_b200 import _lib as L
    g = torch.Generator(device="cuda").manual_seed(M + N)
    A = torch.randn(M, R, device="cuda", generator=g)
    W = torch.randn(N, R, device="cuda", generator=g) / np.sqrt(R)
    bias = torch.randn(N, device="cuda", generator=g)
    H = torch.tanh(torch.randn(M, N, device="cuda", generator=g))
    hi, lo = torch.empty_like(W), torch.empty_like(W)
    L.call("ppx_tc_split", W.data_ptr(), N, R, hi.data_ptr(), lo.data_ptr(), None, None, L.stream())
    need = int(L.call("ppx_tc_linear_workspace", M, R, N))
    assert need > 0, "these shapes are meant to take the split-K path"
    ws = torch.full((need,), float("nan"), device="cuda")
    outs = []
    for _ in range(2):
        C = torch.full((M, N), float("nan"), device="cuda")
        L.call("ppx_tc_linear_ws", A.data_ptr(), R, hi.data_ptr(), lo.data_ptr(), R, M, R, N, bias.data_ptr(), H.data_ptr(), N, act,
               dgrad, None, None, 0.0, C.data_ptr(), N, ws.data_ptr(), need, L.stream())
        outs.append(C)
    torch.cuda.synchronize()
    assert torch.equal(outs[0], outs[1])
    acc = A.double() @ W.double().t()
    if dgrad:
        d = {0: torch.ones_like(H), 1: 1 - H * H, 2: torch.where(H > 0, 1.0, 0.01), 3: torch.where(H > 0, 1.0, H + 1)}[act]
        ref = acc * d.double()
    else:
        ref = ACTS[act](acc + bias.double())
    assert torch.isfinite(outs[0]).all()
    err, scale = float((outs[0].double() - ref).abs().max()), float(acc.abs().max())
    assert err <= 4e-6 * max(scale, 1.0), (err, scale)


def test_tc_split_transposed():
    from ppo_exploration_b200 import _lib as L
    W = torch.randn(70, 45, device="cuda")
    hiT, loT = torch.empty(45, 70, device="cuda"), torch.empty(45, 70, device="cuda")
    hi, lo = torch.empty_like(W), torch.empty_like(W)
    L.call("ppx_tc_split", W.data_ptr(), 70, 45, hi.data_ptr(), lo.data_ptr(), hiT.data_ptr(), loT.data_ptr(), L.stream())
    assert torch.equal(hi.t().contiguous(), hiT) and torch.equal(lo.t().contiguous(), loT) and torch.equal(hi + lo, W)


def test_wide_bonus_net_forward_uses_tensor_cores_and_matches_fp64():
    """RND bonus (models.py:261-267) on a wide observation: the first layers take the tcgen05 path (K >= 256);
    result vs an fp64 torch evaluation of the same weights, and vs the SIMT fp32 kernels."""
    import ppo_exploration_b200 as ppx
    from ppo_exploration_b200 import models as PM
    torch.manual_seed(0)
    D, h, M = 1024, 128, 300
    rnd = ppx.RndNetwork(D, hidden_size=h, device="cuda")
    sd = {}
    for name, layers in (("predictor", rnd.p_layers), ("target", rnd.t_layers)):
        for i, (K, N, _) in enumerate(layers):
            sd[f"{name}.{2 * i}.weight"] = torch.randn(N, K) / np.sqrt(K)
            sd[f"{name}.{2 * i}.bias"] = 0.1 * torch.randn(N)
    rnd.load_state_dict(sd)
    assert 0 in rnd.predictor.tc and 0 in rnd.target.tc, "wide first layers must have tensor-core shadows"
    obs = torch.randn(M, D)
    got = rnd.int_reward(obs.cuda()).cpu().double()

    def mlp(x, name, acts):
        for i, a in enumerate(acts):
            x = x @ sd[f"{name}.{2 * i}.weight"].double().t() + sd[f"{name}.{2 * i}.bias"].double()
            x = {"l": F.leaky_relu, "e": F.elu, "n": lambda z: z}[a](x)
        return x
    x = obs.double()
    want = (mlp(x, "predictor", "llen") - mlp(x, "target", "lln")).pow(2).squeeze(-1)
    torch.testing.assert_close(got, want, rtol=1e-5, atol=1e-6 * float(want.abs().max()))
    rnd.predictor.tc, rnd.target.tc = {}, {}
    simt = rnd.int_reward(obs.cuda()).cpu().double()
    torch.testing.assert_close(got, simt, rtol=2e-5, atol=1e-6 * float(want.abs().max()))


def test_fused_observation_normalisation_matches_separate_pass():
    """RND bonus on RAW observations with a RunningMeanStd: the normalisation fused into the tcgen05 A-split stage gives
    the same bonus as normalize_obs + forward (same f64 formula), and both match the fp64 evaluation."""
    import ppo_exploration_b200 as ppx
    torch.manual_seed(1)
    D, h, M = 1000, 128, 513                                   # K not a multiple of the 32-wide k-block
    rnd = ppx.RndNetwork(D, hidden_size=h, device="cuda")
    sd = {}
    for name, layers in (("predictor", rnd.p_layers), ("target", rnd.t_layers)):
        for i, (K, N, _) in enumerate(layers):
            sd[f"{name}.{2 * i}.weight"] = torch.randn(N, K) / np.sqrt(K)
            sd[f"{name}.{2 * i}.bias"] = 0.1 * torch.randn(N)
    rnd.load_state_dict(sd)
    rms = ppx.RunningMeanStd(shape=(D,), device="cuda")
    rms.set_state(np.random.RandomState(0).rand(D) * 0.5, np.random.RandomState(1).rand(D) * 0.2 + 1e-3, 100.0)
    obs = torch.rand(M, D).cuda() * 3 - 1                      # some entries clip at +-5 sigma
    from ppo_exploration_b200 import models as PM
    assert rnd.predictor.can_fuse_norm(obs)
    old, PM.TC_FUSE_NORM = PM.TC_FUSE_NORM, True
    try:
        fused = rnd.int_reward(obs, rms=rms).cpu().double()
    finally:
        PM.TC_FUSE_NORM = old
    sep = rnd.int_reward(ppx.normalize_obs(obs, rms)).cpu().double()
    torch.testing.assert_close(fused, sep, rtol=1e-6, atol=1e-9)
    mean, var = torch.tensor(rms.mean), torch.tensor(rms.var)
    x = ((obs.cpu().double() - mean) / torch.sqrt(var + 1e-10)).clamp(-5, 5).float().double()

    def mlp(z, name, acts):
        for i, a in enumerate(acts):
            z = z @ sd[f"{name}.{2 * i}.weight"].double().t() + sd[f"{name}.{2 * i}.bias"].double()
            z = {"l": F.leaky_relu, "e": F.elu, "n": lambda q: q}[a](z)
        return z
    want = (mlp(x, "predictor", "llen") - mlp(x, "target", "lln")).pow(2).squeeze(-1)
    torch.testing.assert_close(fused, want, rtol=1e-5, atol=1e-6 * float(want.abs().max()))


def test_wide_policy_first_layer_on_tensor_cores():
    """Atari-shaped flat observations (D >= 256): the policy MLPs' shared first layer runs on tcgen05; outputs and
    gradients still match torch autograd."""
    import ppo_exploration_b200 as ppx
    torch.manual_seed(2)
    D, h, M = 1024, 64, 384
    env = ppx.SyntheticVecEnv(4, D, ppx.Discrete(6), seed=0)
    pol = ppx.models.Policy(env, h, intrinsic_model=True, device="cuda")
    assert pol.mlp.tc1 is not None and not pol.mlp.fused()
    x = torch.randn(M, D)
    outs = pol.forward_raw(x.cuda())
    sd = pol.state_dict()
    d_outs = [torch.randn(M, o) / M for o in pol.outs]
    pol.mlp.backward([d.cuda().contiguous() for d in d_outs])
    got = pol.state_dict(grad=True)
    for gi, (name, o) in enumerate(zip(pol.names, pol.outs)):
        seq = torch.nn.Sequential(torch.nn.Linear(D, h), torch.nn.Tanh(), torch.nn.Linear(h, h), torch.nn.Tanh(), torch.nn.Linear(h, o))
        seq.load_state_dict({k[len(name) + 1:]: v for k, v in sd.items() if k.startswith(name + ".")})
        y = seq(x)
        torch.testing.assert_close(outs[gi].cpu(), y.detach(), rtol=1e-5, atol=2e-5 * max(1.0, float(y.abs().max())))
        y.backward(d_outs[gi])
        for k, prm in seq.named_parameters():
            torch.testing.assert_close(got[f"{name}.{k}"], prm.grad, rtol=1e-5, atol=2e-5 * max(1.0, float(prm.grad.abs().max())))


@pytest.mark.parametrize("M,K,N,ldx_pad", [(4096, 3136, 512, 0), (1024, 1000, 100, 8), (256, 28224, 128, 0)])
def test_tc_wgrad_matches_fp64(M, K, N, ldx_pad):
    """dW = X^T dY and dbias = colsum(dY) through the tensor-core path (samples = reduction dimension) vs fp64."""
    from ppo_exploration_b200 import _lib as L
    g = torch.Generator(device="cuda").manual_seed(M + K)
    Xfull = torch.randn(M, K + ldx_pad, device="cuda", generator=g)
    X = Xfull[:, :K]
    dY = torch.randn(M, N, device="cuda", generator=g) / M
    assert L.call("ppx_tc_wgrad_supported", M, K, N, X.data_ptr(), dY.data_ptr()) == 1
    ws = torch.empty(L.call("ppx_tc_wgrad_workspace", M, K, N), device="cuda")
    dW = torch.full((K, N), float("nan"), device="cuda")
    db = torch.full((N,), float("nan"), device="cuda")
    L.call("ppx_tc_wgrad", X.data_ptr(), K + ldx_pad, dY.data_ptr(), N, M, K, N, dW.data_ptr(), db.data_ptr(), ws.data_ptr(), L.stream())
    torch.cuda.synchronize()
    want = X.double().t() @ dY.double()
    scale = float(want.abs().max())
    assert float((dW.double() - want).abs().max()) <= 4e-6 * max(scale, 1e-3)
    wb = dY.double().sum(0)
    assert float((db.double() - wb).abs().max()) <= 1e-5 * max(float(wb.abs().max()), 1e-3)
