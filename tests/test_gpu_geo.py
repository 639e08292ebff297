"""Ray sampling + nearest-vertex search + geometry features (geo.cu) vs the oracle."""
from importlib import import_module

import pytest
import torch

pytestmark = pytest.mark.gpu


def _setup(pkg, B, seed=5):
    cond = pkg.synthetic.make_conditions(B, seed=seed)
    return cond, {k: v.cuda() for k, v in cond.items()}


@pytest.mark.parametrize("brute", [False, True])
@pytest.mark.parametrize("legacy", [False, True])
def test_knn_bit_exact_and_features(pkg, port, legacy, brute):
    """Same input points on both sides: nearest index and squared distance must be bit-exact."""
    abi = import_module("3dhumangan_b200.abi")
    abi.require_device()
    B, N = 2, 6000
    cond, cg = _setup(pkg, B)
    g = torch.Generator().manual_seed(0)
    pts = (torch.rand(B, N, 3, generator=g) - 0.5) * torch.tensor([1.6, 3.0, 1.2])
    # a few points exactly on vertices and exactly between two vertices (tie handling)
    pts[0, :50] = cond["vertices"][0, 100:150]
    pts[1, :50] = 0.5 * (cond["vertices"][1, 0:50] + cond["vertices"][1, 50:100])
    geo_ref, idx_ref = port.geo_features(pts, cond["skeletons_xyz"], cond["vertices"], cond["tpose_vertices"],
                                         cond["fk_matrices"], cond["lbs_weights"], legacy)
    d2_ref, _ = port.knn1(pts, cond["vertices"])
    vik = abi.vertex_ik(cg["fk_matrices"], cg["lbs_weights"])
    out = abi.geo_features(cg["vertices"], cg["tpose_vertices"], cg["skeletons_xyz"], vik, input_scaler=2 / 2.85,
                           legacy_mode=legacy, points_in=pts.cuda(), want_nearest=True, brute_force=brute)
    torch.cuda.synchronize()
    assert torch.equal(out["nearest"].cpu().long(), idx_ref), "nearest-vertex indices differ"
    assert torch.equal(out["nearest_d2"].cpu(), d2_ref), "squared distances not bit-exact"
    rec = out["rec"].cpu()
    assert torch.allclose(rec[..., :3], pts * (2 / 2.85), rtol=1e-6, atol=0)
    err = (rec[..., 3:34] - geo_ref).abs().max() / geo_ref.abs().max()
    assert err < 1e-5, f"geo feature max error {err:.2e}"
    assert (rec[..., 34:] == 0).all()


def test_vertex_ik_matches_oracle(pkg):
    abi = import_module("3dhumangan_b200.abi")
    cond, cg = _setup(pkg, 3)
    ref = torch.einsum("bij,bjkl->bikl", cond["lbs_weights"], torch.inverse(cond["fk_matrices"])).reshape(3, -1, 16)
    got = abi.vertex_ik(cg["fk_matrices"], cg["lbs_weights"]).cpu()
    assert (got - ref).abs().max() < 2e-6 * ref.abs().max()


def test_ray_sampling_matches_oracle(pkg, port):
    abi = import_module("3dhumangan_b200.abi")
    B, Rw, Rh, S = 2, 6, 12, 16
    cond, cg = _setup(pkg, B)
    cfg = pkg.configs.baseline_config("C2")
    focals, scales = cond["intrinsics"][:, 0, 0], cond["scales"]
    pts, z, d = port.initial_rays(focals, scales, S, Rw, Rh, cfg["ray_start"], cfg["ray_end"])
    u = torch.rand(B, Rw * Rh, S, 1, generator=torch.Generator().manual_seed(3))
    pw, zj = port.jitter_and_transform(pts, z, d, cond["cam2world_matrices"], u)
    xs = torch.linspace(-Rw / Rh, Rw / Rh, Rw)
    ys = torch.linspace(-1, 1, Rh)
    zs = torch.linspace(cfg["ray_start"], cfg["ray_end"], S)
    vik = abi.vertex_ik(cg["fk_matrices"], cg["lbs_weights"])
    out = abi.geo_features(cg["vertices"], cg["tpose_vertices"], cg["skeletons_xyz"], vik, input_scaler=1.0,
                           xs=xs.cuda(), ys=ys.cuda(), zs=zs.cuda(), focals=focals.cuda(), scales=scales.cuda(),
                           cam2world=cg["cam2world_matrices"], jitter=u.reshape(B, -1).cuda(), want_points=True,
                           want_nearest=True)
    torch.cuda.synchronize()
    p = out["points"].cpu().reshape(B, Rw * Rh, S, 3)
    # a few ulp of fp32 at coordinate magnitude ~13 (camera distance) before the transform
    assert (p - pw).abs().max() < 2e-5
    assert (out["z_vals"].cpu().reshape(B, Rw * Rh, S, 1) - zj).abs().max() < 4e-6
    # nearest index agrees wherever the oracle's top-2 gap is not at rounding level
    flat = pw.reshape(B, -1, 3)
    d2, idx = port.knn1(flat, cond["vertices"])
    mism = out["nearest"].cpu().long() != idx
    assert mism.float().mean() < 1e-3


def test_pruned_knn_equals_brute_force_on_ray_samples(pkg):
    """Exact pruning (Morton clusters + box lower bounds) must return the brute-force answer bit for bit,
    including points far outside the body (most ray samples) and exact duplicates of vertices."""
    abi = import_module("3dhumangan_b200.abi")
    B = 2
    cond, cg = _setup(pkg, B, seed=8)
    g = torch.Generator().manual_seed(1)
    pts = (torch.rand(B, 40000, 3, generator=g) - 0.5) * torch.tensor([3.0, 3.0, 1.5])
    pts[0, :3000] = cond["vertices"][0, :3000]
    pts[1, :3000] = cond["vertices"][1, 3000:6000] + 1e-4
    vik = abi.vertex_ik(cg["fk_matrices"], cg["lbs_weights"])
    kw = dict(input_scaler=1.0, points_in=pts.cuda(), want_nearest=True)
    a = abi.geo_features(cg["vertices"], cg["tpose_vertices"], cg["skeletons_xyz"], vik, brute_force=False, **kw)
    b = abi.geo_features(cg["vertices"], cg["tpose_vertices"], cg["skeletons_xyz"], vik, brute_force=True, **kw)
    assert torch.equal(a["nearest"], b["nearest"])
    assert torch.equal(a["nearest_d2"], b["nearest_d2"])
    assert torch.equal(a["rec"], b["rec"])
