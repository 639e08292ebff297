"""Dump the reference's state_dict schema (names, shapes, dtypes, order) for every shipped curriculum.
Build container only (needs /root/reference + oracle/shims)."""
import copy, json, os, sys
HERE = os.path.dirname(os.path.abspath(__file__)); ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle", "shims")); sys.path.insert(0, "/root/reference")
import torch, configs, lib.generators, lib.discriminators, lib.implicit_funcitions
out = {}
for name in ("MAP3DBN", "MAP3DBN512", "MAP3DBN512L"):
    cur = copy.deepcopy(getattr(configs, name))
    meta = configs.extract_metadata(cur, 0)
    if name == "MAP3DBN512L":
        meta["dataset_length"] = 16        # the 219047-row latent pool is 368 MB; its shape is [dataset_length, L]
    meta["neural_field_cls"] = getattr(lib.implicit_funcitions, meta["neural_field_cls"])
    G = lib.generators.Map3DGenerator(**meta)
    D = lib.discriminators.UNetDiscriminator(**meta)
    out[name] = {"G": [[k, list(v.shape), str(v.dtype)] for k, v in G.state_dict().items()],
                 "D": [[k, list(v.shape), str(v.dtype)] for k, v in D.state_dict().items()],
                 "G_params": [k for k, _ in G.named_parameters()], "D_params": [k for k, _ in D.named_parameters()]}
json.dump(out, open(os.path.join(HERE, "state_dict_schema.json"), "w"))
print({k: (len(v["G"]), len(v["D"])) for k, v in out.items()})
