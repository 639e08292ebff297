"""Stand-alone GPU probe (not collected by pytest): graph vs eager divergence, call by call."""
import importlib, sys, os
_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(_ROOT, 'tests')); sys.path.insert(0, _ROOT)
import torch
from golden_util import generator_case, manifest, rel_l2
pkg = importlib.import_module("3dhumangan_b200")
gen = importlib.import_module("3dhumangan_b200.modules.generator")
rng = importlib.import_module("3dhumangan_b200.rng")
cfg, params, cond, z, _, gold = generator_case("g_tiny_dense")
cg = {k: v.cuda() for k, v in cond.items()}
torch.manual_seed(1)
u, noise = rng.draw_render_noise(z.shape[0], 64, cfg["num_steps"], "cpu", cfg["sample_dist"])
ud, nd = u.cuda(), noise.cuda()
rng.draw_render_noise = lambda *a, **k: (ud, nd)
def mk():
    G = gen.Map3DGenerator(**cfg).cuda(); G.load_state_dict(params); G.set_device("cuda"); G.train(); return G
Ge, Gg = mk(), mk()
ku = "synthesis_network.network.m3d_0.conv_0.weight_u"; kr = "synthesis_network.network.m3d_0.spade_0.first_norm.running_mean"
with torch.no_grad():
    for i in range(3):
        oe = Ge(z.cuda(), cg, **cfg); og = Gg(z.cuda(), cg, **dict(cfg, hg_cuda_graph=True))
        torch.cuda.synchronize()
        print(i, "rgbs", rel_l2(og["rgbs"].cpu(), oe["rgbs"].cpu()), "render", rel_l2(og["rgbs_render"].cpu(), oe["rgbs_render"].cpu()),
              "u", rel_l2(Gg.state_dict()[ku].cpu(), Ge.state_dict()[ku].cpu()), "rm", rel_l2(Gg.state_dict()[kr].cpu(), Ge.state_dict()[kr].cpu()))
