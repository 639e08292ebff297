"""Fixture for the two native ops (SURVEY.md rows a'1, a'2): outputs of the UNMODIFIED reference's own pure-torch implementations
`bias_act(..., impl='ref')` (lib/components/ops/bias_act.py:90-121) and `upfirdn2d / upsample2d / downsample2d / filter2d
(..., impl='ref')` (lib/components/ops/upfirdn2d.py:165-211, 281-390) on seeded inputs, rebuilt from the recipes in `cases()`.

    python tests/golden/make_golden_ops.py      # writes tests/golden/native_ops.npz   (build container only)
"""
import importlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))

ACTS = ["linear", "relu", "lrelu", "tanh", "sigmoid", "elu", "selu", "softplus", "swish"]


def bias_act_cases():
    """(x, b, dim, act, alpha, gain, clamp)"""
    g = torch.Generator().manual_seed(101)
    out = []
    for i, act in enumerate(ACTS):
        x = torch.randn(2, 5, 6, 7, generator=g) * 2
        b = torch.randn(5, generator=g)
        out.append((x, b, 1, act, None, None, None))
        out.append((x, None, 1, act, 0.3 if act == "lrelu" else None, 1.7, 0.9))
    x = torch.randn(4, 9, generator=g)
    out.append((x, torch.randn(9, generator=g), 1, "lrelu", 0.1, None, 0.5))
    out.append((x, torch.randn(4, generator=g), 0, "swish", None, 0.5, None))
    return out


def upfirdn_cases():
    """(kind, x, taps, kwargs): kind in upfirdn2d / upsample2d / downsample2d / filter2d"""
    g = torch.Generator().manual_seed(202)
    x = torch.randn(2, 3, 11, 14, generator=g)
    xs = torch.randn(1, 2, 16, 16, generator=g)
    sym6 = [0.015404109327027373, 0.0034907120842174702, -0.11799011114819057, -0.048311742585633, 0.4910559419267466,
            0.787641141030194]
    sym6 = sym6 + sym6[::-1]
    out = [
        ("upfirdn2d", x, [1.0, 3.0, 3.0, 1.0], dict(up=2, down=1, padding=[2, 1, 2, 1], gain=4.0)),
        ("upfirdn2d", x, [1.0, 3.0, 3.0, 1.0], dict(up=1, down=2, padding=[1, 1, 1, 1])),
        ("upfirdn2d", x, [1.0, 2.0, 1.0], dict(up=[2, 1], down=[1, 2], padding=[1, 0, 2, -1], flip_filter=True, gain=0.5)),
        ("upfirdn2d", x, sym6, dict(up=2, down=1, padding=[6, 5, 6, 5], gain=4.0)),
        ("upfirdn2d", x, sym6, dict(up=1, down=2, padding=[5, 5, 5, 5])),
        ("upfirdn2d", x, None, dict(up=2, down=1, padding=0)),
        ("upsample2d", xs, sym6, dict(up=2)),
        ("downsample2d", xs, sym6, dict(down=2)),
        ("upsample2d", xs, [1.0, 3.0, 3.0, 1.0], dict(up=2, padding=1, flip_filter=True, gain=2.0)),
        ("filter2d", xs, [1.0, 4.0, 6.0, 4.0, 1.0], dict(padding=2)),
    ]
    # a non-separable (2-D) filter
    out.append(("upfirdn2d", x, torch.tensor([[1.0, 2.0, 0.5], [0.0, 1.0, 3.0]]).tolist(), dict(up=1, down=1, padding=[1, 1, 0, 1])))
    return out


def main():
    sys.path.insert(0, os.path.join(ROOT, "oracle", "shims"))
    sys.path.insert(0, os.environ.get("HG_REFERENCE", "/root/reference"))
    ba = importlib.import_module("lib.components.ops.bias_act")
    uf = importlib.import_module("lib.components.ops.upfirdn2d")
    # the reference module imports `nv_misc` but calls `misc.assert_shape / suppress_tracer_warnings` (a NameError as shipped:
    # its only callers are in augment.py); bind the name it means, nothing else is touched
    if not hasattr(uf, "misc"):
        uf.misc = uf.nv_misc
    if not hasattr(ba, "misc"):
        ba.misc = ba.nv_misc
    arrays = {}
    for i, (x, b, dim, act, alpha, gain, clamp) in enumerate(bias_act_cases()):
        y = ba.bias_act(x, b, dim=dim, act=act, alpha=alpha, gain=gain, clamp=clamp, impl="ref")
        arrays[f"bias_act_{i}"] = y.numpy()
    for i, (kind, x, taps, kw) in enumerate(upfirdn_cases()):
        f = None if taps is None else uf.setup_filter(taps, normalize=True, flip_filter=False, gain=1, separable=None)
        y = getattr(uf, kind)(x, f, impl="ref", **kw)
        arrays[f"upfirdn_{i}"] = y.numpy()
        if f is not None:
            arrays[f"filter_{i}"] = f.numpy()
    np.savez(os.path.join(HERE, "native_ops.npz"), **arrays)
    print(len(arrays), "arrays", {k: v.shape for k, v in list(arrays.items())[:4]})


if __name__ == "__main__":
    main()
