"""Adam oracle against torch.optim.Adam itself (the reference's optimiser, scene/gaussian_model.py:230), and the
host-side behaviour of FusedAdam that does not need a GPU."""
import numpy as np
import pytest
import torch

from oracle import adam as oadam

# (name, per-point shape, lr) as GaussianModel.training_setup builds them (scene/gaussian_model.py:215-223,
# arguments/__init__.py defaults: position_lr_init 0.00016 * spatial_lr_scale, feature_lr 0.0025, ...)
GROUPS = [("xyz", (3,), 0.00016 * 5.0), ("f_dc", (1, 3), 0.0025), ("f_rest", (15, 3), 0.0025 / 20.0),
          ("opacity", (1,), 0.05), ("scaling", (3,), 0.005), ("rotation", (4,), 0.001), ("ins_feat", (6,), 0.001)]


def make_params(P, seed, device="cpu"):
    g = torch.Generator().manual_seed(seed)
    return [torch.nn.Parameter(torch.randn(P, *shape, generator=g).to(device)) for _, shape, _ in GROUPS]


def make_grads(params, seed, it):
    g = torch.Generator().manual_seed(seed * 1000 + it)
    grads = []
    for p in params:
        x = torch.randn(p.shape, generator=g) * 10 ** float(torch.randint(-6, 1, (1,), generator=g))
        x[torch.rand(p.shape[0], generator=g) < 0.4] = 0            # Gaussians outside the view: zero gradient
        grads.append(x)
    return grads


def test_adam_oracle_vs_torch():
    P, steps = 301, 12
    params = make_params(P, 1)
    opt = torch.optim.Adam([{"params": [p], "lr": lr, "name": n} for p, (n, _, lr) in zip(params, GROUPS)],
                           lr=0.0, eps=1e-15, foreach=False)
    mine = [p.detach().numpy().copy() for p in params]
    m = [np.zeros_like(a) for a in mine]
    v = [np.zeros_like(a) for a in mine]
    for it in range(steps):
        grads = make_grads(params, 1, it)
        frozen = it % 5 == 4                                         # xyz frozen now and then: .grad None -> skipped
        for k, (p, g) in enumerate(zip(params, grads)):
            p.grad = None if (k == 0 and frozen) else g.clone()
        opt.step()
        for k, g in enumerate(grads):
            if k == 0 and frozen:
                continue
            step = int(opt.state[params[k]]["step"])
            oadam.adam_step(mine[k], g.numpy(), m[k], v[k], step, GROUPS[k][2])
    for k, p in enumerate(params):
        st = opt.state[p]
        # float32 rounding order differs (lerp / addcmul / addcdiv contractions): 1e-5 relative, 1e-6 of the range
        for mine_k, ref in ((mine[k], p.detach().numpy()), (m[k], st["exp_avg"].numpy()), (v[k], st["exp_avg_sq"].numpy())):
            np.testing.assert_allclose(mine_k, ref, rtol=1e-5, atol=1e-6 * np.abs(ref).max(), err_msg=GROUPS[k][0])


def test_adam_tensor_layout_matches_header():
    import ctypes as C
    from opengaussian_b200 import _lib
    assert C.sizeof(_lib.AdamTensor) == 72 and _lib.AdamTensor.n.offset == 32 and _lib.AdamTensor.eps.offset == 64


def test_fused_adam_is_a_torch_adam_and_refuses_cpu():
    from opengaussian_b200 import _lib
    from opengaussian_b200.optim import FusedAdam
    params = make_params(8, 2)
    opt = FusedAdam([{"params": [p], "lr": lr, "name": n} for p, (n, _, lr) in zip(params, GROUPS)], lr=0.0, eps=1e-15)
    assert isinstance(opt, torch.optim.Adam)
    assert [g["name"] for g in opt.param_groups] == [n for n, _, _ in GROUPS]
    assert opt.param_groups[0]["eps"] == 1e-15 and opt.param_groups[3]["lr"] == 0.05
    opt.step()                                                       # no gradients anywhere: nothing to do
    assert len(opt.state) == 0
    sd = opt.state_dict()
    torch.optim.Adam([{"params": [p], "lr": 0.0} for p in params]).load_state_dict(sd)   # same state layout
    params[0].grad = torch.zeros_like(params[0])
    with pytest.raises(_lib.OgsError):
        opt.step()                                                   # no CPU path
    with pytest.raises(NotImplementedError):
        FusedAdam(params, weight_decay=0.1)
