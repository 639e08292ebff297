"""App-level plumbing on the GPU: the call sequence of the reference's `apps/sample_from_generator.py:70-149` + `generate_frames`
(:24-58) executed against the drop-in import paths (`3dhumangan_b200/dropin` ahead of everything else on sys.path), i.e. BASELINE
config C1: MAP3DBN (hidden 384), 1 x 256x128 (and the square 256x256 variant), random z, truncation 0.7, eval-mode
BatchNorm, `last_back = eval_last_back`, checkpoint = a plain `state_dict` loaded strictly.

The reference app itself cannot travel to the GPU box (it needs the dataset, smplx and the pytorch3d rasteriser for the
conditions); the dataset + preprocessor are replaced by `synthetic.make_conditions`, everything
from `configs.get_config` to the uint8 frames follows the app line by line."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r'''
import importlib, math, os, sys, tempfile
import numpy as np, torch
sys.path.insert(0, %(root)r); sys.path.insert(0, %(dropin)r)
import configs, lib.generators                                     # the drop-in packages
device = torch.device("cuda")
class opt: config = "MAP3DBN"; tune = ""; variant = 0
config = configs.get_config(opt)
config = {k: v for k, v in config.items() if type(k) is str}
config.update(%(over)r)
config["truncation_psi"] = 0.7; config["v_stddev"] = 0; config["h_stddev"] = 0
config["last_back"] = config.get("eval_last_back", False)
config["nerf_noise"] = 0
assert config["hidden_dim"] == 384 and config["last_back"] is True

# a "released checkpoint": the state_dict of a generator whose BatchNorm running statistics / spectral-norm vectors have seen data
pkg = importlib.import_module("3dhumangan_b200")
torch.manual_seed(0)
G0 = getattr(lib.generators, config["generator"])(**config).to(device)
G0.set_device(device); G0.train()
cond = {k: v.to(device) for k, v in pkg.synthetic.make_conditions(2, seed=3).items()}
with torch.no_grad():
    for _ in range(4):
        G0(torch.randn(2, config["latent_dim"], device=device), cond, **dict(config, nerf_noise=0.5, last_back=False))
path = os.path.join(tempfile.mkdtemp(), "generator_state.pth")
torch.save(G0.state_dict(), path)

checkpoint = torch.load(path)
assert isinstance(checkpoint, dict)
generator = getattr(lib.generators, config["generator"])(**config).to(device)
generator.load_state_dict(checkpoint)                              # strict
generator.set_device(device)
generator.eval()

n_angles = 3
torch.manual_seed(1); torch.cuda.manual_seed(1)
z = torch.randn((1, config["latent_dim"]), device=device).repeat_interleave(n_angles, dim=0).to(generator.device)
conditions = {k: v[:1].repeat_interleave(n_angles, dim=0) for k, v in cond.items()}
frames = torch.zeros(n_angles, 3, config["gen_height"], config["gen_width"]).float().to(generator.device)
with torch.no_grad():
    for i in range(n_angles):
        sub = {k: v[i:i + 1] for k, v in conditions.items()}
        out = generator.staged_forward(z[i:i + 1], sub, **config)
        frames[i:i + 1] = out["rgbs"]
        assert out["depths"].device.type == "cpu" and out["depths"].shape == (1, 1, config["render_height"], config["render_width"])
frames = frames * 0.5 + 0.5
frames = torch.clamp(frames * 255, 0, 255).to(torch.uint8).permute(0, 2, 3, 1).cpu().numpy()
assert frames.shape == (n_angles, config["gen_height"], config["gen_width"], 3) and frames.dtype == np.uint8
assert frames.std() > 0
print("frames", frames.shape, int(frames.min()), int(frames.max()))
print("ok")
'''


@pytest.mark.parametrize("over", [{}, dict(gen_height=256, gen_width=256, render_height=64, render_width=64)])
def test_sample_app_flow_through_dropin(over):
    code = SCRIPT % dict(root=ROOT, dropin=os.path.join(ROOT, "3dhumangan_b200", "dropin"), over=over)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0 and out.stdout.strip().endswith("ok"), (out.stdout[-1000:], out.stderr[-3000:])
