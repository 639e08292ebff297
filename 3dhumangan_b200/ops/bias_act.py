"""bias_act: y = clamp(act(x + b) * gain)  (reference: lib/components/ops/bias_act.py:52-86, bias_act.cu:24-165).

Same Python signature as the reference op.  The reference defaults to `impl='ref'` (plain torch
ops) and its CUDA plugin cannot even be built; here there is exactly one
implementation, the sm_100a kernel behind `hg_bias_act`.  Forward only in this round.
"""
import math

import torch

from .. import abi

# name -> (id, default alpha, default gain)   (bias_act.py:22-32)
ACTIVATIONS = {
    "linear": (1, 0.0, 1.0), "relu": (2, 0.0, math.sqrt(2)), "lrelu": (3, 0.2, math.sqrt(2)), "tanh": (4, 0.0, 1.0),
    "sigmoid": (5, 0.0, 1.0), "elu": (6, 0.0, 1.0), "selu": (7, 0.0, 1.0), "softplus": (8, 0.0, 1.0),
    "swish": (9, 0.0, math.sqrt(2)),
}


def bias_act(x, b=None, dim=1, act="linear", alpha=None, gain=None, clamp=None, impl="cuda"):
    if act not in ACTIVATIONS:
        raise RuntimeError(f"bias_act: unknown activation {act!r}")
    if torch.is_grad_enabled() and (x.requires_grad or (b is not None and b.requires_grad)):
        raise RuntimeError("hg3d: bias_act backward is not built yet; call under torch.no_grad()")
    aid, dalpha, dgain = ACTIVATIONS[act]
    alpha = float(dalpha if alpha is None else alpha)
    gain = float(dgain if gain is None else gain)
    clamp = float(-1 if clamp is None else clamp)
    if b is not None:
        if b.ndim != 1 or not (0 <= dim < x.ndim) or b.shape[0] != x.shape[dim]:
            raise RuntimeError("bias_act: bias must be 1-D and match x.shape[dim]")
        b = b.detach().float().contiguous()
    xin = x.detach().float().contiguous()
    y = torch.empty_like(xin)
    step = 1
    for s in xin.shape[dim + 1:]:
        step *= s
    with torch.cuda.device_of(xin):
        abi.call("hg_bias_act", abi.ptr(xin), abi.ptr(b), abi.ptr(y), xin.numel(), step,
                                        xin.shape[dim] if b is not None else 1, aid, alpha, gain, clamp, abi.stream())
    return y.to(x.dtype)
