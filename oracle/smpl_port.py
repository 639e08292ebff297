"""CPU oracle for the SMPL skinning / pose-condition step in front of the hot path.

TEST INFRASTRUCTURE ONLY (imported by tests/ only).

`lbs` restates lib/components/smpl.py:11-107 of the reference.  That function composes four helpers of the third-party
package `smplx` (`smplx[all]`, unpinned in doc/INSTALL.md:14; not vendored by the reference and absent from this image):
blend_shapes, vertices2joints, batch_rodrigues, batch_rigid_transform -- restated below from smplx's published `lbs.py`.
PINNING: `tests/test_oracle_pin.py::test_smpl_lbs_*` runs the reference's OWN `lbs` body (imported from /root/reference, with
these helper restatements injected for the missing package) against this file; the helpers themselves are "parity unpinned"
(no smplx, no SMPL_NEUTRAL.pkl, no reference test at that boundary).  `conditions_fix_body` restates
`SHHQDataset._preprocess_smpl_fix_body` (lib/data/datasets.py:117-181) batch-wise and is pinned to the reference's method by
the fixture tests/golden/smpl_conditions.npz; `cam2world_fix_body` restates `SHHQPreprocessor._forward_fix_body`
(lib/data/preprocessor.py:72-98) with pytorch3d's documented `euler_angles_to_matrix(convention="XYZ")` (R = Rx Ry Rz)."""
import math

import torch
import torch.nn.functional as F


# ---- smplx.lbs helpers (published algorithm) -----------------------------------------------------------------------
def blend_shapes(betas, shape_disps):
    return torch.einsum("bl,mkl->bmk", betas, shape_disps)


def vertices2joints(J_regressor, vertices):
    return torch.einsum("bik,ji->bjk", vertices, J_regressor)


def batch_rodrigues(rot_vecs, epsilon=1e-8):
    n = rot_vecs.shape[0]
    angle = torch.norm(rot_vecs + 1e-8, dim=1, keepdim=True)
    rot_dir = rot_vecs / angle
    cos, sin = torch.cos(angle)[:, None], torch.sin(angle)[:, None]
    rx, ry, rz = torch.split(rot_dir, 1, dim=1)
    zeros = torch.zeros((n, 1), dtype=rot_vecs.dtype, device=rot_vecs.device)
    K = torch.cat([zeros, -rz, ry, rz, zeros, -rx, -ry, rx, zeros], dim=1).view(n, 3, 3)
    ident = torch.eye(3, dtype=rot_vecs.dtype, device=rot_vecs.device)[None]
    return ident + sin * K + (1 - cos) * torch.bmm(K, K)


def batch_rigid_transform(rot_mats, joints, parents, dtype=torch.float32):
    joints = joints.unsqueeze(-1)
    rel = joints.clone()
    rel[:, 1:] -= joints[:, parents[1:]]
    tm = torch.cat([F.pad(rot_mats.reshape(-1, 3, 3), [0, 0, 0, 1]), F.pad(rel.reshape(-1, 3, 1), [0, 0, 0, 1], value=1.0)], dim=2)
    tm = tm.reshape(-1, joints.shape[1], 4, 4)
    chain = [tm[:, 0]]
    for i in range(1, parents.shape[0]):
        chain.append(torch.matmul(chain[int(parents[i])], tm[:, i]))
    transforms = torch.stack(chain, dim=1)
    posed = transforms[:, :, :3, 3]
    jh = F.pad(joints, [0, 0, 0, 1])
    rel_t = transforms - F.pad(torch.matmul(transforms, jh), [3, 0, 0, 0, 0, 0, 0, 0])
    return posed, rel_t


# ---- lib/components/smpl.py:11-107 -------------------------------------------------------------------------------------
def lbs(betas, pose, v_template, shapedirs, posedirs, J_regressor, parents, lbs_weights, pose2rot=True):
    B = max(betas.shape[0], pose.shape[0])
    v_shaped = v_template + blend_shapes(betas, shapedirs)
    J = vertices2joints(J_regressor, v_shaped)
    ident = torch.eye(3, dtype=betas.dtype)
    if pose2rot:
        rot = batch_rodrigues(pose.view(-1, 3)).view(B, -1, 3, 3)
        feat = (rot[:, 1:] - ident).view(B, -1)
    else:
        feat = (pose[:, 1:].view(B, -1, 3, 3) - ident).view(B, -1)
        rot = pose.view(B, -1, 3, 3)
    v_posed = torch.matmul(feat, posedirs).view(B, -1, 3) + v_shaped
    J_t, A = batch_rigid_transform(rot, J, parents)
    W = lbs_weights[None].expand(B, -1, -1)
    T = torch.matmul(W, A.view(B, J_regressor.shape[0], 16)).view(B, -1, 4, 4)
    vh = torch.cat([v_posed, torch.ones(B, v_posed.shape[1], 1, dtype=betas.dtype)], dim=2)
    verts = torch.matmul(T, vh[..., None])[:, :, :3, 0]
    return A, v_shaped, verts, J, J_t


# ---- lib/data/datasets.py:117-181 (batched; `pred` = the SMPL prediction a sample carries) -------------------------------
def conditions_fix_body(orig_cam, joints, full_pose_rot, tpose_vertices_shaped, fk_matrices, lbs_weights, smpl_tpose_vertices,
                        joint_ids=tuple(range(24))):
    B = orig_cam.shape[0]
    fov = math.pi * 12 / 180
    focal = 1.0 / math.tan(fov / 2)
    sx, tx, ty = orig_cam[:, 0] / 2.0, orig_cam[:, 2], orig_cam[:, 3]
    K = torch.diag(torch.tensor([focal, focal, 1.0, 1.0])).float()[None].expand(B, 4, 4).clone()
    R = torch.eye(4)[None].expand(B, 4, 4).clone()
    T = torch.eye(4)[None].expand(B, 4, 4).clone()
    T[:, 0, 3], T[:, 1, 3], T[:, 2, 3] = tx, ty, focal / sx
    cano_rot = torch.tensor([[1.0, 0.0, 0.0], [0.0, math.cos(math.pi), -math.sin(math.pi)], [0.0, math.sin(math.pi), math.cos(math.pi)]],
                            dtype=torch.float64)
    cano = torch.eye(4, dtype=torch.float64)[None].expand(B, 4, 4).clone()
    cano[:, :3, :3] = cano_rot @ torch.linalg.inv(full_pose_rot[:, 0].double())
    fk = torch.einsum("bij,bnjk->bnik", cano, fk_matrices.double())
    vfk = torch.einsum("vn,bnjk->bvjk", lbs_weights.double(), fk)
    vh = F.pad(tpose_vertices_shaped.double(), (0, 1), value=1.0)
    vertices = torch.einsum("bvij,bvj->bvi", vfk, vh)[..., :3]
    sk = joints[:, list(joint_ids)].double()
    sk = torch.einsum("bij,bnj->bni", cano, F.pad(sk, (0, 1), value=1.0))[..., :3]
    tp = smpl_tpose_vertices.float().clone()
    tp[..., 1] += 0.35
    return {"scales": sx.float(), "skeletons_xyz": sk.float(), "intrinsics": K, "vertices": vertices.float(),
            "tpose_vertices": tp[None].expand(B, -1, -1).contiguous(), "full_pose": full_pose_rot.float(), "fk_matrices": fk.float(),
            "lbs_weights": lbs_weights.float()[None].expand(B, -1, -1).contiguous(), "cano_matrices": cano.float(), "R": R, "T": T}


def euler_xyz_to_matrix(e):
    """pytorch3d.transforms.euler_angles_to_matrix(e, "XYZ") = Rx(e0) @ Ry(e1) @ Rz(e2) (pytorch3d 0.6.2 documented contract)."""
    def rot(axis, a):
        c, s, o, z = torch.cos(a), torch.sin(a), torch.ones_like(a), torch.zeros_like(a)
        m = {"X": (o, z, z, z, c, -s, z, s, c), "Y": (c, z, s, z, o, z, -s, z, c), "Z": (c, -s, z, s, c, z, z, z, o)}[axis]
        return torch.stack(m, -1).reshape(a.shape + (3, 3))
    return rot("X", e[..., 0]) @ rot("Y", e[..., 1]) @ rot("Z", e[..., 2])


def cam2world_fix_body(full_pose_rot, R, T, h_rotation, v_rotation, r_rotation):
    """lib/data/preprocessor.py:72-98 -> (cam2world [B,4,4], R_raster [B,3,3])."""
    B = R.shape[0]
    euler = torch.zeros(B, 3)
    euler[:, 1] = -h_rotation
    euler[:, 0] = math.pi - v_rotation
    euler[:, 2] = -r_rotation
    Rb = full_pose_rot[:, 0] @ euler_xyz_to_matrix(euler)
    body = F.pad(Rb, (0, 1, 0, 1))
    body[:, -1, -1] = 1.0
    w2c = torch.bmm(torch.bmm(R, T), body)
    return torch.inverse(w2c.float()), torch.inverse(Rb)
