"""Rebuild the exact inputs of a golden fixture from its recipe (tests/golden/manifest.json)."""
import importlib
import json
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")


def manifest():
    with open(os.path.join(GOLD, "manifest.json")) as f:
        return json.load(f)


def load(name):
    return {k: torch.from_numpy(v) for k, v in np.load(os.path.join(GOLD, name + ".npz")).items()}


def generator_case(name):
    """-> cfg, params, cond, z, (u, noise), golden outputs."""
    pkg = importlib.import_module("3dhumangan_b200")
    from oracle import port
    m = manifest()[name]
    base, over, pseed, sg, sb, B, noise_std = m["recipe"]
    cfg = pkg.configs.baseline_config(base)
    cfg.update(over)
    cfg["nerf_noise"] = noise_std
    params = port.init_generator_params(cfg, seed=pseed, sigma_gain=sg, sigma_bias=sb)
    cond = pkg.synthetic.make_conditions(B, seed=11 + pseed)
    z = torch.randn(B, cfg["latent_dim"], generator=torch.Generator().manual_seed(100 + pseed))
    torch.manual_seed(m["rng_seed"])
    u, noise = pkg.rng.draw_render_noise(B, cfg["render_width"] * cfg["render_height"], cfg["num_steps"], "cpu",
                                         cfg["sample_dist"])
    return cfg, params, cond, z, (u, noise), load(name)


def discriminator_case(name):
    pkg = importlib.import_module("3dhumangan_b200")
    from oracle import port
    over, pseed, B = manifest()[name]["recipe"]
    cfg = pkg.configs.baseline_config("C2")
    cfg.update(over)
    params = port.init_discriminator_params(cfg, seed=pseed)
    img = torch.randn(B, 3, cfg["gen_height"], cfg["gen_width"], generator=torch.Generator().manual_seed(pseed)).clamp(-1, 1)
    return cfg, params, img, load(name)


def rel_l2(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))
