"""Curricula of the hot path's inputs, with the reference's dict semantics.

Mirrors `configs/__init__.py:37-72` and `configs/map3d.py` of the reference: a curriculum is a
dict whose int keys are "from this step on" overrides and whose str keys are constants;
`extract_metadata(curriculum, step)` merges them; the merged dict is splatted as **kwargs into
every constructor / forward call (SURVEY.md §5 "Config / flags" — the kwargs-splat convention is
part of the drop-in boundary).  The three shipped curricula differ only in a handful of keys, so
they are generated from one base table here.
"""
from __future__ import annotations

import copy
import math

_PHASE = lambda rotate, r1: {"name": "uncond", "uncond": True, "rotate": rotate, "gen_modal": "rgbs", "do_r1": r1}

_BASE = {
    "trainer": "PhaseTrainer",
    # 8-phase schedule: R1 on phases 3 and 7 (configs/map3d.py:104-113)
    "phases": [_PHASE(False, False), _PHASE(True, False), _PHASE(True, False), _PHASE(False, True),
               _PHASE(False, False), _PHASE(True, False), _PHASE(False, False), _PHASE(True, True)],
    "2d_coords_input": True, "2d_semantic_input": False, "2d_latent_input": False,
    "neural_field_latent_input": False, "use_mixed_precision": True, "lock_view_dependence": True,
    "num_steps": 32,
    "ray_start": -0.5, "ray_end": 0.55, "side_length": 2.85, "depth_length": 1.05,
    "vis_rotate": math.pi / 6, "fade_steps": 1, "sample_dist": "gaussian",
    "h_stddev": 0.4, "v_stddev": 0.1, "h_mean": 0, "v_mean": 0, "coordinate_mode": "fix_body",
    "betas": (0, 0.9), "unique_lr": True, "appearance_codes_lr_mul": 1.0, "mapping_net_lr_mul": 0.05,
    "neural_field_lr_mul": 0.05, "weight_decay": 0,
    "gan_lambda": 0, "r1_lambda": 0, "photometric_lambda": 0, "perceptual_lambda": [0, 0, 0, 0],
    "latent_lambda": 0, "z_lambda": 0, "pos_lambda": 0, "semantic_lambda": 0, "segmentation_lambda": 1,
    "input_dim": 3, "output_dim": 3, "semantic_dim": 0, "geo_feature_dim": 31, "label_dim": 26,
    "grad_clip": 1.0,
    "neural_field_cls": "COORDCONCATSIREN", "generator": "Map3DGenerator", "map3d_mode": "mixed",
    "neural_field_blocks": 4, "synthesis_blocks": 9, "mod_blocks": [0, 1, 2],
    "spatial_normalization": "batch_norm", "discriminator": "UNetDiscriminator",
    "condition_modal_disc_real": "body_segments", "condition_modal_disc_gen": "rasterized_segments",
    "condition_modal_gen": "rasterized_segments",
    "ada_aug": dict(xflip=1, rotate90=0, rotate_max=0.05, xint=0, scale=1, rotate=1, aniso=1, xfrac=0,
                    brightness=1, contrast=1, saturation=1),
    "ada_target": 0.6, "ada_interval": 0, "ada_kimg": 20, "ada_alpha_thresh": 0.5,
    "dataset": "SHHQDataset", "dataset_length": 10, "dataroot": "./datasets/shhq_example_dataset",
    "joints": list(range(24)), "white_back": True, "clamp_mode": "relu", "z_dist": "gaussian",
    "hierarchical_sample": False, "learnable_dist": False, "last_back": False, "eval_last_back": True,
}


def _curriculum(name, dims, gen_hw, render_hw, steps, **extra):
    c = copy.deepcopy(_BASE)
    c.update(name=name, latent_dim=dims, hidden_dim=dims, feature_dim=dims,
             gen_height=gen_hw[0], gen_width=gen_hw[1], render_height=render_hw[0], render_width=render_hw[1])
    c.update(extra)
    c.update(steps)
    return c


_SLOW = {"batch_size": 32, "batch_split": 1, "gen_lr": 5e-5, "disc_lr": 2e-4}
_FAST = {"batch_size": 32, "batch_split": 1, "gen_lr": 1e-4, "disc_lr": 4e-4}

# configs/map3d.py:3-95, :98-191, :194-290
MAP3DBN = _curriculum("map3dbn", 384, (256, 128), (64, 32),
                      {0: dict(_FAST), int(140e3 + 1): dict(_SLOW), int(300e3 + 1): {}}, r1_lambda=0.25)
MAP3DBN512 = _curriculum("map3dbn512", 256, (512, 256), (96, 48), {0: dict(_SLOW), int(300e3 + 1): {}})
MAP3DBN512L = _curriculum("map3dbn512l", 420, (512, 256), (96, 48), {0: dict(_SLOW), int(300e3 + 1): {}},
                          map3d_mode="isolated", legacy_mode=True, dataset_length=219047,
                          dataroot="./datasets/shhq_train_40000")


def extract_metadata(curriculum, current_step):
    """configs/__init__.py:37-46: newest int-keyed override <= step, then every str-keyed constant."""
    out = {}
    for s in sorted((k for k in curriculum if isinstance(k, int)), reverse=True):
        if s <= current_step:
            out.update(curriculum[s])
            break
    out.update({k: v for k, v in curriculum.items() if not isinstance(k, int)})
    return out


def get_config(opt):
    """configs/__init__.py:49-72.  Resolves `neural_field_cls` to the class (idempotent here)."""
    from .modules import implicit
    config = globals()[opt.config]
    if isinstance(config["neural_field_cls"], str):
        config["neural_field_cls"] = getattr(implicit, config["neural_field_cls"])
    tune = getattr(opt, "tune", "")
    if not tune:
        return config
    if tune == "lr":
        gen_lr, disc_lr = [(1e-4, 4e-4), (2e-4, 2e-4), (1e-4, 2e-4), (1e-4, 1e-4)][opt.variant]
        for k in config:
            if isinstance(k, int):
                config[k]["gen_lr"], config[k]["disc_lr"] = gen_lr, disc_lr
        config["name"] = f"{config['name']}_G_lr={gen_lr}_D_lr={disc_lr}"
    elif tune == "map3d_mode":
        mode = ["isolated", "mixed", "all"][opt.variant]
        config["map3d_mode"] = mode
        config["name"] = f"{config['name']}_map3d_mode={mode}"
    else:
        raise NotImplementedError
    return config


# --------------------------------------------------------------------------------------------
# BASELINE.json configs pinned by SURVEY.md §8(d): square variants keep the render:gen ratio.
# --------------------------------------------------------------------------------------------
def baseline_config(which: str):
    """Merged metadata dict (str keys only) for BASELINE config `which` in {C1,C2,C2native,C5,tiny}."""
    table = {
        "C1": (MAP3DBN, dict(gen_height=256, gen_width=256, render_height=64, render_width=64)),
        "C2": (MAP3DBN512, dict(gen_height=512, gen_width=512, render_height=96, render_width=96)),
        "C2native": (MAP3DBN512, {}),
        "C5": (MAP3DBN512, dict(gen_height=1024, gen_width=1024, render_height=192, render_width=192, num_steps=128)),
        "tiny": (MAP3DBN512, dict(gen_height=64, gen_width=64, render_height=16, render_width=16)),
    }
    cur, over = table[which]
    meta = extract_metadata(copy.deepcopy({k: v for k, v in cur.items()}), 0)
    meta.update(over)
    meta["nerf_noise"] = 0.0
    return meta
