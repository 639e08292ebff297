"""Restatement of the documented contract of pytorch3d.ops.knn_points / knn_gather (0.6.2).

knn_points(p1 [N,P1,D], p2 [N,P2,D], K=1) -> (dists [N,P1,K] squared L2, idx [N,P1,K] int64, None)
Direct-difference form in fp32, summed x,y,z in that order with one rounding per operation
(no FMA contraction); ties go to the lowest index (torch.min over dim returns the first
minimal element on CPU).  knn_gather(x [N,P2,U], idx [N,P1,K]) -> [N,P1,K,U].
"""
from collections import namedtuple

import torch

_KNN = namedtuple("KNN", "dists idx knn")


def knn_points(p1, p2, lengths1=None, lengths2=None, K=1, version=-1, return_nn=False, return_sorted=True):
    assert K == 1, "only K=1 is used by the reference hot path (smpl.py:220)"
    p1 = p1.float()
    p2 = p2.float()
    N, P1, _ = p1.shape
    dists = torch.empty(N, P1, 1, dtype=torch.float32, device=p1.device)
    idx = torch.empty(N, P1, 1, dtype=torch.int64, device=p1.device)
    chunk = 4096
    for n in range(N):
        vx, vy, vz = p2[n, :, 0][None], p2[n, :, 1][None], p2[n, :, 2][None]
        for s in range(0, P1, chunk):
            q = p1[n, s:s + chunk]
            dx = q[:, 0:1] - vx
            dy = q[:, 1:2] - vy
            dz = q[:, 2:3] - vz
            d2 = dx * dx + dy * dy + dz * dz
            m, i = torch.min(d2, dim=1)
            # torch.min does not document first-index tie-breaking on every backend: enforce it.
            first = (d2 == m[:, None]).to(torch.int8).argmax(dim=1)
            dists[n, s:s + chunk, 0] = m
            idx[n, s:s + chunk, 0] = first
    return _KNN(dists, idx, None)


def knn_gather(x, idx, lengths=None):
    N, P2, U = x.shape
    _, P1, K = idx.shape
    g = torch.gather(x[:, :, None, :].expand(N, P2, K, U), 1, idx[:, :, :, None].expand(N, P1, K, U))
    return g
