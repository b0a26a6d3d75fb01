"""CPU restatement (numpy float32) of the optimiser step the reference takes at train.py:609 --
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The algorithm lives in a third-party dependency: torch.optim.Adam (torch 2.11, torch/optim/adam.py::
_single_tensor_adam), built by the reference as `torch.optim.Adam(l, lr=0.0, eps=1e-15)`
(scene/gaussian_model.py:230).  torch IS installed in the build container and on the GPU box, so this restatement
is pinned directly: tests/test_optim_cpu.py runs torch.optim.Adam (CPU, foreach=False) beside it on the reference's
seven parameter groups.
"""
import numpy as np


def adam_step(param, grad, exp_avg, exp_avg_sq, step, lr, beta1=0.9, beta2=0.999, eps=1e-15, grad_scale=1.0):
    """One update of one tensor, in place on float32 numpy arrays; `step` is the count AFTER the increment."""
    f = np.float32
    g = grad.astype(f) * f(grad_scale)
    exp_avg += (g - exp_avg) * f(1 - beta1)                       # exp_avg.lerp_(grad, 1 - beta1)
    exp_avg_sq *= f(beta2)                                        # exp_avg_sq.mul_(beta2)
    exp_avg_sq += f(1 - beta2) * g * g                            #   .addcmul_(grad, grad, value=1 - beta2)
    bias_correction1 = 1 - beta1 ** step
    bias_correction2 = 1 - beta2 ** step
    step_size = lr / bias_correction1
    denom = np.sqrt(exp_avg_sq) / f(bias_correction2 ** 0.5) + f(eps)
    param -= f(step_size) * (exp_avg / denom)                     # param.addcdiv_(exp_avg, denom, value=-step_size)
    return param
