"""Fixture for the class-balanced segmentation loss: values of the UNMODIFIED reference method
PhaseTrainer._calculate_segmentation_loss (lib/trainers/phase_trainer.py:203-256) on seeded inputs.

    python tests/golden/make_golden_loss.py      # writes tests/golden/seg_loss.npz   (build container only)
"""
import importlib
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def cases():
    g = torch.Generator().manual_seed(77)
    out = []
    for B, L, H, W in [(2, 26, 16, 16), (1, 26, 8, 12), (3, 7, 8, 8)]:
        seg = torch.randn(B, L, H, W, generator=g) * 2
        gt = torch.randint(0, L, (B, H, W), generator=g)
        out.append((seg, gt, L))
    seg, gt, L = out[0]
    out.append((seg, torch.zeros_like(gt), L))        # all background: plain cross entropy branch
    return out


def main():
    sys.path.insert(0, os.path.join(ROOT, "oracle", "shims"))
    sys.path.insert(0, os.environ.get("HG_REFERENCE", "/root/reference"))
    pt = importlib.import_module("lib.trainers.phase_trainer")
    me = types.SimpleNamespace(device="cpu")
    vals, grads = [], []
    for seg, gt, L in cases():
        s = seg.clone().requires_grad_(True)
        loss, _, _ = pt.PhaseTrainer._calculate_segmentation_loss(me, s, gt, {"label_dim": L})
        loss.backward()
        vals.append(float(loss))
        grads.append(float(s.grad.double().norm()))
    np.savez(os.path.join(HERE, "seg_loss.npz"), loss=np.array(vals), grad_norm=np.array(grads))
    print(vals, grads)


if __name__ == "__main__":
    main()
