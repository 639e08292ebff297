"""Stand-alone GPU probe (not collected): per-block error of the synthesis network on a multi-tile-per-CTA case."""
import importlib, os, sys
_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(_ROOT, 'tests')); sys.path.insert(0, _ROOT)
import torch, torch.nn.functional as F
from golden_util import rel_l2
from oracle import port
pkg = importlib.import_module("3dhumangan_b200")
syn = importlib.import_module("3dhumangan_b200.modules.synthesis_ops")
cfg = pkg.configs.baseline_config("C2"); cfg.update(gen_height=int(sys.argv[1]) if len(sys.argv) > 1 else 128, gen_width=int(sys.argv[2]) if len(sys.argv) > 2 else 96, render_height=24, render_width=18, map3d_mode="mixed")
B = 2
params = port.init_generator_params(cfg, seed=21)
g = torch.Generator().manual_seed(21)
fmap = torch.rand(B, 256, 24, 18, generator=g) * 2 - 0.5
fstyle = torch.randn(B, 1, 256, generator=g)
Hg, Wg = cfg["gen_height"], cfg["gen_width"]
with torch.no_grad():
    style = F.interpolate(fmap, (Hg, Wg), mode="bilinear"); x0 = port.synthesis_input(params, B, Hg, Wg)
    ref_rgb, ref_int = port.synthesis_network(params, x0, style, fstyle, cfg, training=True, return_internal=True)
gp = {k: v.cuda() for k, v in params.items()}
feat_lr = fmap.permute(0, 2, 3, 1).reshape(B, -1, 256).contiguous().cuda()
rgb, internal = syn.synthesis_forward(gp, feat_lr, fstyle.cuda(), cfg, training=True, passes=3, return_internal=True)
torch.cuda.synchronize()
for k in range(9):
    a, b = internal[f"m3d_{k}"].cpu(), ref_int[f"m3d_{k}"]
    e = rel_l2(a, b)
    bad = ((a - b).abs() > 1e-2 * b.abs().max()).reshape(B, 256, -1)
    print(k, f"{e:.3e}", "bad px per image:", bad.any(1).sum(-1).tolist(), "first bad tiles:", sorted(set((bad.any(1)[0].nonzero().flatten() // 128).tolist()))[:8], "finite", bool(torch.isfinite(a).all()))
print("rgb", rel_l2(rgb.cpu(), ref_rgb))
