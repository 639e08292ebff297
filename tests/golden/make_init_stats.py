"""Per-tensor initialisation statistics (mean, std, abs-max) of the UNMODIFIED reference modules, averaged over seeds, for
every floating-point parameter of Map3DGenerator / UNetDiscriminator (MAP3DBN512).  Build container only
(needs /root/reference + oracle/shims).  tests/test_cpu_boundary.py compares this repo's modules with them."""
import copy, json, os, sys
HERE = os.path.dirname(os.path.abspath(__file__)); ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle", "shims")); sys.path.insert(0, "/root/reference")
import torch, configs, lib.generators, lib.discriminators, lib.implicit_funcitions
cur = copy.deepcopy(configs.MAP3DBN512)
meta = configs.extract_metadata(cur, 0)
meta["neural_field_cls"] = getattr(lib.implicit_funcitions, meta["neural_field_cls"])
out = {"G": {}, "D": {}}
SEEDS = 3
for seed in range(SEEDS):
    torch.manual_seed(seed)
    for key, m in (("G", lib.generators.Map3DGenerator(**meta)), ("D", lib.discriminators.UNetDiscriminator(**meta))):
        for n, p in m.named_parameters():
            if p.is_floating_point():
                r = out[key].setdefault(n, [0.0, 0.0, 0.0])
                r[0] += float(p.mean()) / SEEDS
                r[1] += float(p.std()) / SEEDS if p.numel() > 1 else 0.0
                r[2] += float(p.abs().max()) / SEEDS
json.dump(out, open(os.path.join(HERE, "init_stats.json"), "w"))
print(len(out["G"]), len(out["D"]))
