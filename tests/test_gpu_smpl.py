"""SMPL skinning + pose conditions on the device (3dhumangan_b200/smpl.py, csrc/smpl.cu) against the oracle, which is pinned to the
reference's own functions by tests/test_oracle_pin.py."""
import importlib

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("pose2rot", [True, False])
def test_lbs_and_conditions_match_oracle(pose2rot):
    from oracle import smpl_port as sp
    smpl = importlib.import_module("3dhumangan_b200.smpl")
    model = smpl.SMPLModel.synthetic("cuda", V=6890, J=24, NB=10, seed=3)
    m = {k: getattr(model, k).cpu() for k in ("v_template", "shapedirs", "posedirs", "J_regressor", "lbs_weights")}
    parents = model.parents.cpu().long()
    g = torch.Generator().manual_seed(4)
    B = 5
    betas = torch.randn(B, 10, generator=g)
    pose = torch.randn(B, 24, 3, generator=g) * 0.5
    pose[0, 3] = 0                                              # a zero rotation (the 1e-8 guard of batch_rodrigues)
    rot = sp.batch_rodrigues(pose.reshape(-1, 3)).reshape(B, 24, 3, 3)
    ref = sp.lbs(betas, pose.reshape(B, -1) if pose2rot else rot, m["v_template"], m["shapedirs"], m["posedirs"], m["J_regressor"], parents,
                 m["lbs_weights"], pose2rot=pose2rot)
    out = smpl.lbs(betas, pose if pose2rot else rot, model, pose2rot=pose2rot)
    torch.cuda.synchronize()
    for name, r in zip(("fk_matrices", "tpose_vertices", "vertices", "joints_shaped", "joints"), ref):
        got = out[name].cpu()
        assert got.shape == r.shape, name
        assert float((got - r).abs().max()) < 2e-5 * max(1.0, float(r.abs().max())), (name, float((got - r).abs().max()))
    orig_cam = torch.stack([1.2 + 0.2 * torch.rand(B, generator=g), torch.ones(B), 0.1 * torch.randn(B, generator=g), 0.1 * torch.randn(B, generator=g)], 1)
    cref = sp.conditions_fix_body(orig_cam, ref[4], rot, ref[1], ref[0], m["lbs_weights"], m["v_template"])
    cond = smpl.conditions_fix_body(orig_cam, out, model)
    for k, r in cref.items():
        got = cond[k].cpu()
        assert got.shape == r.shape, k
        assert float((got - r).abs().max()) < 5e-5 * max(1.0, float(r.abs().max())), (k, float((got - r).abs().max()))
    ang = torch.randn(3, B, generator=g) * 0.3
    c2w_ref, _ = sp.cam2world_fix_body(rot, cref["R"], cref["T"], ang[0], ang[1], ang[2])
    c2w = smpl.cam2world_fix_body(cond, ang[0], ang[1], ang[2]).cpu()
    assert float((c2w - c2w_ref).abs().max()) < 5e-5 * float(c2w_ref.abs().max())


def test_conditions_feed_the_generator(pkg):
    """The dict produced on the device is what Map3DGenerator.forward reads (keys, shapes, dtypes): one forward through it."""
    smpl = importlib.import_module("3dhumangan_b200.smpl")
    gen = importlib.import_module("3dhumangan_b200.modules.generator")
    model = smpl.SMPLModel.synthetic("cuda", seed=5)
    B = 2
    g = torch.Generator().manual_seed(6)
    out = smpl.lbs(torch.randn(B, 10, generator=g) * 0.5, torch.randn(B, 24, 3, generator=g) * 0.2, model)
    cond = smpl.conditions_fix_body(torch.tensor([[1.4, 1.4, 0.0, 0.0]] * B), out, model)
    cond["cam2world_matrices"] = smpl.cam2world_fix_body(cond, torch.zeros(B), torch.zeros(B), torch.zeros(B))
    cfg = pkg.configs.baseline_config("tiny")
    G = gen.Map3DGenerator(**cfg).cuda().train()
    G.set_device(torch.device("cuda:0"))
    with torch.no_grad():
        img = G(torch.randn(B, cfg["latent_dim"], device="cuda"), cond, **cfg)["rgbs"]
    assert img.shape == (B, 3, cfg["gen_height"], cfg["gen_width"]) and torch.isfinite(img).all()
