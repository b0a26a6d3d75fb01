"""FusedAdam (csrc/adam.cu through the C ABI) against torch.optim.Adam -- the optimiser the reference builds at
scene/gaussian_model.py:230 -- on the same parameters and gradients, and against the CPU oracle.
Tolerance: float32 rounding order only (1e-5 relative, 1e-6 of the tensor's range), written below."""
import copy

import numpy as np
import pytest
import torch

from test_optim_cpu import GROUPS, make_grads, make_params

pytestmark = pytest.mark.gpu


def _close(a, b, what):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    tol = 1e-5 * b.abs() + 1e-6 * b.abs().max()
    assert bool(((a - b).abs() <= tol).all()), (what, float((a - b).abs().max()), float(b.abs().max()))


def _groups(params):
    return [{"params": [p], "lr": lr, "name": n} for p, (n, _, lr) in zip(params, GROUPS)]


@pytest.mark.parametrize("P", [1, 1000, 50_001])
def test_fused_adam_vs_torch_adam(P):
    from opengaussian_b200.optim import FusedAdam
    ref_p = make_params(P, 3, "cuda")
    my_p = [torch.nn.Parameter(p.detach().clone()) for p in ref_p]
    ref = torch.optim.Adam(_groups(ref_p), lr=0.0, eps=1e-15)
    mine = FusedAdam(_groups(my_p), lr=0.0, eps=1e-15)
    for it in range(8):
        grads = make_grads(ref_p, 3, it)
        frozen = it % 4 == 3                                  # xyz without a gradient: skipped, step count not bumped
        for k, g in enumerate(grads):
            g = g.cuda()
            ref_p[k].grad = None if (k == 0 and frozen) else g.clone()
            my_p[k].grad = None if (k == 0 and frozen) else g.clone()
        if it == 5:                                           # update_learning_rate (scene/gaussian_model.py:236-247)
            for opt in (ref, mine):
                opt.param_groups[6]["lr"] = opt.param_groups[6]["lr"] * 0 + 0.0001
        ref.step()
        mine.step()
        ref.zero_grad(set_to_none=True)
        mine.zero_grad(set_to_none=True)
    for k, (a, b) in enumerate(zip(my_p, ref_p)):
        name = GROUPS[k][0]
        _close(a, b, name)
        _close(mine.state[a]["exp_avg"], ref.state[b]["exp_avg"], name + ".exp_avg")
        _close(mine.state[a]["exp_avg_sq"], ref.state[b]["exp_avg_sq"], name + ".exp_avg_sq")
        assert float(mine.state[a]["step"]) == float(ref.state[b]["step"])


def test_fused_adam_vs_oracle_and_state_surgery():
    """Oracle parity, then the reference's optimiser edits: prune (scene/gaussian_model.py:372-386) and
    state_dict round trip (:98, :120); unaligned views take the scalar path."""
    from opengaussian_b200.optim import FusedAdam
    from oracle import adam as oadam
    P = 777
    params = make_params(P, 4, "cuda")
    opt = FusedAdam(_groups(params), lr=0.0, eps=1e-15)
    mine = [p.detach().cpu().numpy().copy() for p in params]
    m = [np.zeros_like(a) for a in mine]
    v = [np.zeros_like(a) for a in mine]
    for it in range(3):
        grads = make_grads(params, 4, it)
        for p, g in zip(params, grads):
            p.grad = g.cuda()
        opt.step()
        for k, g in enumerate(grads):
            oadam.adam_step(mine[k], g.numpy(), m[k], v[k], it + 1, GROUPS[k][2])
    for k, p in enumerate(params):
        _close(p, torch.from_numpy(mine[k]), GROUPS[k][0])
    # prune: keep every other Gaussian, exactly as _prune_optimizer rewrites the state
    keep = torch.arange(P, device="cuda") % 2 == 0
    for group in opt.param_groups:
        old = group["params"][0]
        st = opt.state.pop(old)
        st["exp_avg"] = st["exp_avg"][keep]
        st["exp_avg_sq"] = st["exp_avg_sq"][keep]
        group["params"][0] = torch.nn.Parameter(old[keep].detach().requires_grad_(True))
        opt.state[group["params"][0]] = st
    sd = copy.deepcopy(opt.state_dict())          # state_dict() hands out the live tensors; the twin needs its own
    twin_p = [torch.nn.Parameter(g["params"][0].detach().clone()) for g in opt.param_groups]
    twin = torch.optim.Adam(_groups(twin_p), lr=0.0, eps=1e-15)
    twin.load_state_dict(sd)
    for g_mine, p_twin in zip(opt.param_groups, twin_p):
        grad = torch.randn_like(p_twin)
        g_mine["params"][0].grad = grad.clone()
        p_twin.grad = grad.clone()
    opt.step()
    twin.step()
    for g_mine, p_twin, (name, _, _) in zip(opt.param_groups, twin_p, GROUPS):
        _close(g_mine["params"][0], p_twin, name + " after prune")
        assert float(opt.state[g_mine["params"][0]]["step"]) == 4.0


def test_fused_adam_grad_scale_and_single_launch():
    from opengaussian_b200 import _lib
    from opengaussian_b200.optim import FusedAdam
    a = make_params(4099, 5, "cuda")
    b = [torch.nn.Parameter(p.detach().clone()) for p in a]
    oa, ob = FusedAdam(_groups(a), lr=0.0, eps=1e-15), FusedAdam(_groups(b), lr=0.0, eps=1e-15)
    for k, (p, q, g) in enumerate(zip(a, b, make_grads(a, 5, 0))):
        p.grad = g.cuda()
        if k == 2:                                            # a gradient that is only 4-byte aligned: scalar path
            buf = torch.empty(g.numel() + 1, device="cuda")
            buf[1:] = g.cuda().view(-1)
            p.grad = buf[1:].view(g.shape)
            assert p.grad.data_ptr() % 16 == 4 and p.grad.is_contiguous()
        q.grad = g.cuda() * 0.25
    _lib.lib().ogs_profile_enable(1)
    _lib.profile_read()
    oa.step(grad_scale=0.25)
    ms, launches = _lib.profile_read()["adam"]
    _lib.lib().ogs_profile_enable(0)
    assert launches == 1                                      # seven groups, one kernel
    ob.step()
    for p, q in zip(a, b):
        assert torch.equal(p, q)                              # 0.25 is a power of two: scaling commutes exactly
