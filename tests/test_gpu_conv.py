"""GPU numerics of the conv front-end.  The reference has no live conv arithmetic (its only conv spec
is the dead draft .ipynb_checkpoints/models-checkpoint.py:48-66, :93-121), so parity is DEFINED against torch.nn.Conv2d /
nn.Linear evaluated in fp64 on the CPU: forward features, every weight / bias gradient and the input gradient, at 1e-5 of
each tensor's scale."""
import numpy as np
import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu


def _close(got, ref, name, tol=1e-5):
    ref = ref.detach().double().cpu()
    got = got.detach().double().cpu().reshape(ref.shape)
    scale = max(float(ref.abs().max()), 1e-30)
    err = float((got - ref).abs().max())
    assert err <= tol * scale, f"{name}: max err {err:.3e} vs scale {scale:.3e} (rel {err / scale:.2e})"


@pytest.mark.parametrize("nchw,C,H,W,k,s", [(1, 4, 84, 84, 8, 4), (0, 32, 20, 20, 4, 2), (0, 64, 9, 9, 3, 1), (1, 3, 11, 7, 3, 2),
                                          (0, 5, 6, 9, 2, 1)])
def test_im2col_and_col2im_against_unfold_fold(nchw, C, H, W, k, s):
    """ppx_im2col == F.unfold (patch order (c,kh,kw) for NCHW input, (kh,kw,c) for NHWC); ppx_col2im == F.fold, i.e. the
    adjoint of im2col (every input element sums the patch entries that cover it)."""
    import torch.nn.functional as F
    from ppo_exploration_b200 import _lib as L
    N = 3
    g = torch.Generator().manual_seed(C * 100 + k)
    x_nchw = torch.randn(N, C, H, W, generator=g)
    OH, OW = (H - k) // s + 1, (W - k) // s + 1
    K = C * k * k
    x_dev = (x_nchw if nchw else x_nchw.permute(0, 2, 3, 1)).contiguous().cuda()
    cols = torch.full((N * OH * OW, K), float("nan"), device="cuda")
    L.call("ppx_im2col", x_dev.data_ptr(), nchw, N, C, H, W, k, k, s, cols.data_ptr(), L.stream())
    ref = F.unfold(x_nchw, k, stride=s).transpose(1, 2).reshape(N * OH * OW, C, k, k)            # rows of (c, kh, kw)
    ref = ref.reshape(N * OH * OW, K) if nchw else ref.permute(0, 2, 3, 1).reshape(N * OH * OW, K)
    assert torch.equal(cols.cpu(), ref)
    d = torch.randn(N * OH * OW, K, generator=g)
    dx = torch.full((x_dev.numel(),), float("nan"), device="cuda")
    L.call("ppx_col2im", d.cuda().data_ptr(), nchw, N, C, H, W, k, k, s, dx.data_ptr(), L.stream())
    d_ckk = d.reshape(N, OH * OW, C, k, k) if nchw else d.reshape(N, OH * OW, k, k, C).permute(0, 1, 4, 2, 3)
    want = F.fold(d_ckk.reshape(N, OH * OW, K).transpose(1, 2).double(), (H, W), k, stride=s)   # [N, C, H, W]
    want = want if nchw else want.permute(0, 2, 3, 1)
    _close(dx, want.contiguous(), "col2im", tol=1e-6)


def _torch_trunk(kind, hidden):
    act = nn.ReLU if kind == "actor_critic" else nn.LeakyReLU
    last_k = 3 if kind == "actor_critic" else 2          # the RND draft's third conv is 2/1 (models-checkpoint.py:100)
    return nn.Sequential(nn.Conv2d(4, 32, 8, 4), act(), nn.Conv2d(32, 64, 4, 2), act(), nn.Conv2d(64, 64, last_k, 1), act(),
                         nn.Flatten(), nn.Linear((7 if last_k == 3 else 8) ** 2 * 64, hidden), nn.ReLU())


@pytest.mark.parametrize("kind,N,hidden", [("actor_critic", 48, 512), ("rnd", 20, 256)])
def test_conv_trunk_forward_backward_vs_torch_fp64(kind, N, hidden):
    """The Nature-CNN trunk of CnnActorCritic (draft :51-62) and the RND conv stack (:97-108), forward + backward."""
    from ppo_exploration_b200 import models as PM
    torch.manual_seed(3)
    ref = _torch_trunk(kind, hidden)
    for m in ref.modules():                                   # the draft's init (:76-79)
        if isinstance(m, (nn.Conv2d, nn.Linear)):
            nn.init.orthogonal_(m.weight, np.sqrt(2))
            m.bias.data.uniform_(-0.1, 0.1)
    a = "relu" if kind == "actor_critic" else "leaky_relu"
    convs = [(32, 8, 4, a), (64, 4, 2, a), (64, 3 if kind == "actor_critic" else 2, 1, a)]
    fc = (hidden, "relu")
    dev = torch.device("cuda")
    bank = PM.ParamBank(PM.ConvTrunk.specs("trunk", (4, 84, 84), convs, fc), dev)
    trunk = PM.ConvTrunk(bank, "trunk", (4, 84, 84), convs, fc, PM._Scratch(dev))
    trunk.load_torch(ref.state_dict())
    x = torch.rand(N, 4, 84, 84)                              # frames in [0, 1]
    feat, saved = trunk.forward(x.reshape(N, -1).cuda())
    ref64 = ref.double()
    x64 = x.double().requires_grad_(True)
    want = ref64(x64)
    _close(feat, want, "features")
    d_feat = torch.randn(N, hidden, generator=torch.Generator().manual_seed(5))
    want.backward(d_feat.double())
    dx = trunk.backward(saved, d_feat.cuda(), need_dx=True)
    got = trunk.grads_as_torch()
    for name, p in ref64.named_parameters():
        _close(got[name], p.grad, name)
    _close(dx, x64.grad.reshape(N, -1), "d_input")
