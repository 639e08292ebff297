"""upfirdn2d: pad / zero-insert up-sample / FIR / decimate  (reference: lib/components/ops/upfirdn2d.py:117-161,
upfirdn2d.cu:29-375).  Same Python signatures (`upfirdn2d`, `setup_filter`, `filter2d`, `upsample2d`,
`downsample2d`); one implementation, the sm_100a kernel behind `hg_upfirdn2d`.  Forward only in this round.
"""
import numpy as np
import torch

from .. import abi


def _pair(v):
    if isinstance(v, int):
        return v, v
    v = list(v)
    assert len(v) == 2
    return int(v[0]), int(v[1])


def _padding(p):
    if isinstance(p, int):
        p = [p, p]
    p = [int(v) for v in p]
    if len(p) == 2:
        p = [p[0], p[0], p[1], p[1]]
    assert len(p) == 4
    return p


def setup_filter(f, device=torch.device("cpu"), normalize=True, flip_filter=False, gain=1, separable=None):
    """upfirdn2d.py:66-113."""
    if f is None:
        f = 1
    f = torch.as_tensor(f, dtype=torch.float32)
    assert f.ndim in (0, 1, 2) and f.numel() > 0
    if f.ndim == 0:
        f = f[np.newaxis]
    if separable is None:
        separable = f.ndim == 1 and f.numel() >= 8
    if f.ndim == 1 and not separable:
        f = f.ger(f)
    assert f.ndim == (1 if separable else 2)
    if normalize:
        f = f / f.sum()
    if flip_filter:
        f = f.flip(list(range(f.ndim)))
    f = f * (gain ** (f.ndim / 2))
    return f.to(device=device)


def _run(x, f2d, upx, upy, downx, downy, px0, px1, py0, py1, flip, gain):
    B, C, H, W = x.shape
    fH, fW = f2d.shape
    outH = (H * upy + py0 + py1 - fH) // downy + 1
    outW = (W * upx + px0 + px1 - fW) // downx + 1
    y = torch.empty(B, C, max(outH, 0), max(outW, 0), dtype=torch.float32, device=x.device)
    if y.numel() == 0:
        return y
    with torch.cuda.device_of(x):
        abi.call("hg_upfirdn2d", abi.ptr(x), abi.ptr(f2d), abi.ptr(y), B * C, H, W, outH, outW, fH, fW, upx, upy,
                                         downx, downy, px0, py0, int(bool(flip)), float(gain), abi.stream())
    return y


def upfirdn2d(x, f, up=1, down=1, padding=0, flip_filter=False, gain=1, impl="cuda"):
    if torch.is_grad_enabled() and x.requires_grad:
        raise RuntimeError("hg3d: upfirdn2d backward is not built yet; call under torch.no_grad()")
    assert x.ndim == 4
    upx, upy = _pair(up)
    downx, downy = _pair(down)
    px0, px1, py0, py1 = _padding(padding)
    xin = x.detach().float().contiguous()
    if f is None:
        f = torch.ones(1, 1, dtype=torch.float32, device=x.device)
    f = f.to(device=x.device, dtype=torch.float32).contiguous()
    if f.ndim == 2:
        y = _run(xin, f, upx, upy, downx, downy, px0, px1, py0, py1, flip_filter, gain)
    else:  # separable: a [1,fw] pass then a [fh,1] pass, gain split as in upfirdn2d.py:243-244
        g = float(gain) ** 0.5
        y = _run(xin, f[None, :].contiguous(), upx, 1, downx, 1, px0, px1, 0, 0, flip_filter, g)
        y = _run(y, f[:, None].contiguous(), 1, upy, 1, downy, 0, 0, py0, py1, flip_filter, g)
    return y.to(x.dtype)


def _filter_size(f):
    if f is None:
        return 1, 1
    return (int(f.shape[-1]), int(f.shape[0]))


def filter2d(x, f, padding=0, flip_filter=False, gain=1, impl="cuda"):
    px0, px1, py0, py1 = _padding(padding)
    fw, fh = _filter_size(f)
    p = [px0 + fw // 2, px1 + (fw - 1) // 2, py0 + fh // 2, py1 + (fh - 1) // 2]
    return upfirdn2d(x, f, padding=p, flip_filter=flip_filter, gain=gain)


def upsample2d(x, f, up=2, padding=0, flip_filter=False, gain=1, impl="cuda"):
    upx, upy = _pair(up)
    px0, px1, py0, py1 = _padding(padding)
    fw, fh = _filter_size(f)
    p = [px0 + (fw + upx - 1) // 2, px1 + (fw - upx) // 2, py0 + (fh + upy - 1) // 2, py1 + (fh - upy) // 2]
    return upfirdn2d(x, f, up=up, padding=p, flip_filter=flip_filter, gain=gain * upx * upy)


def downsample2d(x, f, down=2, padding=0, flip_filter=False, gain=1, impl="cuda"):
    downx, downy = _pair(down)
    px0, px1, py0, py1 = _padding(padding)
    fw, fh = _filter_size(f)
    p = [px0 + (fw - downx + 1) // 2, px1 + (fw - downx) // 2, py0 + (fh - downy + 1) // 2, py1 + (fh - downy) // 2]
    return upfirdn2d(x, f, down=down, padding=p, flip_filter=flip_filter, gain=gain)
