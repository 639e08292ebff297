"""Gradient fixtures from the UNMODIFIED reference (build container only; see make_golden.py for the set-up).

    python tests/golden/make_golden_grads.py      # writes tests/golden/*_grads.npz

For the generator case `g_tiny_dense` and the discriminator case `d_tiny`: run the reference module under autograd
on the CPU, loss = sum(output * w) with seeded weights w, `loss.backward()`, and store for EVERY parameter that
received a gradient its L2 norm and its projection on a seeded random direction (a checksum of checksums: the full
gradients are ~10 MB), plus a few small gradients in full (and the image gradient of the discriminator).
`tests/test_oracle_pin.py` checks autograd through the oracle against these, which pins the oracle's BACKWARD to the
reference's; the GPU tests then compare the kernels with autograd through the oracle.
"""
import copy
import importlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
import make_golden  # noqa: E402

FULL_G = ["neural_field.sigma_layer.weight", "neural_field.color_layer_linear.weight", "synthesis_network.to_rgbs.m3d_8.linear.weight",
          "synthesis_network.network.m3d_0.spade_0.first_norm.weight", "synthesis_input.network.0.weight",
          "neural_field_mapping_network.network.6.bias"]
FULL_D = ["layer_up_last.weight", "output_layer.bias", "body_down.0.conv_s.bias"]


def loss_weights(shape, seed):
    return torch.randn(shape, generator=torch.Generator().manual_seed(seed))


def direction(name, shape):
    return torch.randn(shape, generator=torch.Generator().manual_seed(sum(name.encode()) * 7919 % (2 ** 31)))


def summarise(named_grads, full_names):
    names = sorted(named_grads)
    norms = np.array([float(named_grads[n].double().norm()) for n in names])
    dots = np.array([float((named_grads[n].double() * direction(n, named_grads[n].shape).double()).sum()) for n in names])
    out = {"names": np.array(names), "norms": norms, "dots": dots}
    for n in full_names:
        if n in named_grads:
            out["full:" + n] = named_grads[n].numpy()
    return out


def main():
    pkg = importlib.import_module("3dhumangan_b200")
    from oracle import port
    gens, discs, impl = make_golden.reference_modules()

    name = "g_tiny_dense"
    cfg, params, cond, z, B = make_golden.build_case(pkg, port, name)
    meta = dict(cfg)
    meta["neural_field_cls"] = getattr(impl, meta["neural_field_cls"])
    G = gens.Map3DGenerator(**meta)
    G.load_state_dict(copy.deepcopy(params), strict=True)
    G.set_device("cpu")
    G.train()
    torch.manual_seed(1234)
    out = G(z, cond, **meta)
    loss = (out["rgbs"] * loss_weights(out["rgbs"].shape, 1)).sum() + (out["rgbs_render"] * loss_weights(out["rgbs_render"].shape, 2)).sum()
    loss.backward()
    grads = {n: p.grad.detach().clone() for n, p in G.named_parameters() if p.grad is not None}
    np.savez_compressed(os.path.join(HERE, name + "_grads.npz"), loss=np.array(float(loss)), **summarise(grads, FULL_G))
    print(name, "loss", float(loss), len(grads), "gradients")

    over, pseed, Bd = make_golden.D_CASES["d_tiny"]
    cfg = pkg.configs.baseline_config("C2")
    cfg.update(over)
    params = port.init_discriminator_params(cfg, seed=pseed)
    D = discs.UNetDiscriminator(**cfg)
    D.load_state_dict(params, strict=True)
    D.train()
    img = torch.randn(Bd, 3, cfg["gen_height"], cfg["gen_width"], generator=torch.Generator().manual_seed(pseed)).clamp(-1, 1)
    img.requires_grad_(True)
    o = D(img, None, alpha=1.0)
    loss = sum((o[k] * loss_weights(o[k].shape, 3 + i)).sum() for i, k in enumerate(("prediction", "segments", "latents")))
    loss.backward()
    grads = {n: p.grad.detach().clone() for n, p in D.named_parameters() if p.grad is not None}
    np.savez_compressed(os.path.join(HERE, "d_tiny_grads.npz"), loss=np.array(float(loss)), image_grad=img.grad.numpy(),
                        **summarise(grads, FULL_D))
    print("d_tiny loss", float(loss), len(grads), "gradients")


if __name__ == "__main__":
    main()
