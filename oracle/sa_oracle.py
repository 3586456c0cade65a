"""TEST INFRASTRUCTURE - CPU oracle for the Ev2Hands set-abstraction encoder.

This is a restatement, in plain PyTorch on the CPU, of the algorithm in the
reference's ``src/Ev2Hands/model/pointnet2_utils.py``.  It exists only as the
checker for the CUDA path: ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it; the
product package ``ev2hands_b200`` never does.

Parity status: PINNED.  ``tests/golden/make_golden.py`` imports the real
reference from ``/root/reference`` in the build container, runs it on seeded
inputs, and commits the outputs under ``tests/golden/``;
``tests/test_oracle_cpu.py`` checks every function here against those vectors
(indices bit-exact, features to 1e-6).

It deliberately keeps the reference's *cost structure* (Python-level FPS loop,
materialised ``[B,S,N]`` distance matrix, full sort in the ball query, channel
first ``[B,C,K,S]`` convolutions) so that timing it is a fair stand-in for
timing the reference on the same host cores.

Function <-> reference line map
    pairwise_sqdist       pointnet2_utils.py:19-40   (square_distance)
    take_rows             pointnet2_utils.py:43-60   (index_points)
    fps                   pointnet2_utils.py:63-84   (farthest_point_sample)
    ball_query            pointnet2_utils.py:87-107  (query_ball_point)
    group_all             pointnet2_utils.py:141-158 (sample_and_group_all)
    shared_mlp_max        pointnet2_utils.py:193-199, :253-257
    sa_msg_forward        pointnet2_utils.py:224-262 (PointNetSetAbstractionMsg.forward)
    sa_all_forward        pointnet2_utils.py:176-202 (PointNetSetAbstraction.forward, group_all)
    three_nn_weights      pointnet2_utils.py:294-301 (3 nearest sources + inverse-distance weights)
    fp_forward            pointnet2_utils.py:276-315 (PointNetFeaturePropagation.forward)
    decoder_forward       TEHNet.py:184-186          (fp3 -> fp2 -> fp1)
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

BN_EPS = 1e-5   # nn.BatchNorm2d default, used by pointnet2_utils.py:172,219


def pairwise_sqdist(src: torch.Tensor, dst: torch.Tensor) -> torch.Tensor:
    """[B,S,3],[B,N,3] -> [B,S,N] in the expanded form the reference uses.

    The rounding of this exact expression (sgemm with K=3, then two broadcast
    adds) decides ball-query membership, so it is kept term for term."""
    cross = torch.matmul(src, dst.transpose(1, 2))
    d = -2 * cross
    d += (src * src).sum(-1).unsqueeze(2)
    d += (dst * dst).sum(-1).unsqueeze(1)
    return d


def take_rows(table: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """table [B,N,C], idx [B,...] int64 -> [B,...,C]."""
    b = torch.arange(table.shape[0], device=table.device).view(-1, *([1] * (idx.dim() - 1)))
    return table[b.expand_as(idx), idx]


def fps(xyz: torch.Tensor, n_sample: int, start: torch.Tensor | None = None) -> torch.Tensor:
    """Iterative farthest point sampling, [B,N,3] -> int64 [B,n_sample].

    ``start`` replays the reference's ``torch.randint`` draw; when None the
    draw is made here from the same (CPU default) generator."""
    B, N, _ = xyz.shape
    if start is None:
        start = torch.randint(0, N, (B,), dtype=torch.long)
    dev = xyz.device                                    # (runs on any device; the reference timing uses the CPU)
    picked = torch.zeros(B, n_sample, dtype=torch.long, device=dev)
    best = torch.full((B, N), 1e10, device=dev)
    cur = start.clone().to(dev)
    rows = torch.arange(B, device=dev)
    for i in range(n_sample):
        picked[:, i] = cur
        c = xyz[rows, cur].unsqueeze(1)
        d = ((xyz - c) ** 2).sum(-1)
        closer = d < best
        best[closer] = d[closer]
        cur = best.max(-1).indices
    return picked


def ball_query(radius: float, n_neighbor: int, xyz: torch.Tensor, centres: torch.Tensor) -> torch.Tensor:
    """First ``n_neighbor`` in-radius point indices per centre, ascending, padded
    with the first hit.  [B,N,3],[B,S,3] -> int64 [B,S,n_neighbor]."""
    B, N, _ = xyz.shape
    S = centres.shape[1]
    ids = torch.arange(N, dtype=torch.long, device=xyz.device).expand(B, S, N).clone()
    ids[pairwise_sqdist(centres, xyz) > radius ** 2] = N
    ids = ids.sort(dim=-1).values[:, :, :n_neighbor]
    first = ids[:, :, :1].expand(-1, -1, n_neighbor)
    empty = ids == N
    ids[empty] = first[empty]
    return ids


def group_all(xyz: torch.Tensor, feats: torch.Tensor | None):
    """One group holding every point; channels [xyz(uncentred), feats]."""
    B, N, C = xyz.shape
    centre = torch.zeros(B, 1, C, device=xyz.device)
    g = xyz.view(B, 1, N, C)
    if feats is not None:
        g = torch.cat([g, feats.view(B, 1, N, -1)], dim=-1)
    return centre, g


def shared_mlp_max(x: torch.Tensor, layers) -> torch.Tensor:
    """x [B,C,K,S]; layers = [(W[Co,Ci,1,1], b, gamma, beta, mean, var)...];
    1x1 conv + eval-mode BatchNorm + ReLU per layer, then max over K."""
    for layer in layers:
        w, b, gamma, beta, mean, var = (t.to(x.device) for t in layer)
        x = F.relu(F.batch_norm(F.conv2d(x, w, b), mean, var, gamma, beta, False, 0.1, BN_EPS))
    return x.max(dim=2).values


def _layers(state: dict, conv_fmt: str, bn_fmt: str, n: int):
    out = []
    for j in range(n):
        c, b = conv_fmt.format(j=j), bn_fmt.format(j=j)
        out.append(tuple(torch.as_tensor(state[k]) for k in (
            c + ".weight", c + ".bias", b + ".weight", b + ".bias",
            b + ".running_mean", b + ".running_var")))
    return out


def sa_msg_forward(state: dict, spec: dict, xyz_cf: torch.Tensor, feats_cf: torch.Tensor | None,
                   start: torch.Tensor | None = None, return_aux: bool = False):
    """Multi-scale set abstraction.  xyz_cf [B,3,N], feats_cf [B,D,N] (channel
    first, as the model passes them) -> (centres [B,3,S], feats [B,sum(D'),S])."""
    xyz = xyz_cf.permute(0, 2, 1).contiguous()
    feats = feats_cf.permute(0, 2, 1).contiguous() if feats_cf is not None else None
    B, N, _ = xyz.shape
    S = spec["npoint"]
    fps_idx = fps(xyz, S, start)
    centres = take_rows(xyz, fps_idx)
    pooled, ball = [], []
    for i, (r, K) in enumerate(zip(spec["radius_list"], spec["nsample_list"])):
        gi = ball_query(r, K, xyz, centres)
        ball.append(gi)
        rel = take_rows(xyz, gi) - centres.view(B, S, 1, 3)
        g = torch.cat([take_rows(feats, gi), rel], dim=-1) if feats is not None else rel
        g = g.permute(0, 3, 2, 1).contiguous()
        layers = _layers(state, "conv_blocks.%d.{j}" % i, "bn_blocks.%d.{j}" % i, len(spec["mlp_list"][i]))
        pooled.append(shared_mlp_max(g, layers))
    out = (centres.permute(0, 2, 1).contiguous(), torch.cat(pooled, dim=1))
    if return_aux:
        return out + ({"fps_idx": fps_idx, "ball_idx": ball},)
    return out


def sa_all_forward(state: dict, spec: dict, xyz_cf: torch.Tensor, feats_cf: torch.Tensor | None):
    """group_all set abstraction: [B,3,N],[B,D,N] -> ([B,3,1] zeros, [B,D',1])."""
    xyz = xyz_cf.permute(0, 2, 1).contiguous()
    feats = feats_cf.permute(0, 2, 1).contiguous() if feats_cf is not None else None
    centre, g = group_all(xyz, feats)
    g = g.permute(0, 3, 2, 1).contiguous()
    layers = _layers(state, "mlp_convs.{j}", "mlp_bns.{j}", len(spec["mlp"]))
    return centre.permute(0, 2, 1).contiguous(), shared_mlp_max(g, layers)


def encoder_forward(states: dict, specs: dict, events: torch.Tensor, starts: dict, return_aux: bool = False):
    """sa1 -> sa2 -> sa3 exactly as TEHNet.forward wires them (TEHNet.py:172-181):
    xyz is the first three channels, sa1's features are all five channels."""
    l0_xyz = events[:, :3, :]
    aux = {}
    r1 = sa_msg_forward(states["sa1"], specs["sa1"], l0_xyz, events, starts["sa1"], return_aux)
    r2 = sa_msg_forward(states["sa2"], specs["sa2"], r1[0], r1[1], starts["sa2"], return_aux)
    l3_xyz, l3 = sa_all_forward(states["sa3"], specs["sa3"], r2[0], r2[1])
    if return_aux:
        aux = {"sa1": r1[2], "sa2": r2[2], "l1_xyz": r1[0], "l1_points": r1[1],
               "l2_xyz": r2[0], "l2_points": r2[1]}
        return l3, aux
    return l3


def regressor_sa_forward(states: dict, specs: dict, xyz_cf: torch.Tensor, hand_feats: torch.Tensor,
                         start: torch.Tensor, return_aux: bool = False):
    """MANORegressor's sa1 -> sa2 (TEHNet.py:75-79)."""
    r1 = sa_msg_forward(states["sa1"], specs["sa1"], xyz_cf, hand_feats, start, return_aux)
    _, out = sa_all_forward(states["sa2"], specs["sa2"], r1[0], r1[1])
    if return_aux:
        return out, {"sa1": r1[2], "l1_xyz": r1[0], "l1_points": r1[1]}
    return out


def three_nn_weights(xyz1: torch.Tensor, xyz2: torch.Tensor):
    """xyz1 [B,N,3] queries, xyz2 [B,S,3] sources (S >= 3) -> (idx int64 [B,N,3], weight [B,N,3]):
    the three nearest sources by the expanded squared distance (full sort, as the reference
    does) and the normalised inverse-distance weights 1 / (d + 1e-8)."""
    d = pairwise_sqdist(xyz1, xyz2)
    d, idx = d.sort(dim=-1)
    d, idx = d[:, :, :3], idx[:, :, :3]
    recip = 1.0 / (d + 1e-8)
    norm = torch.sum(recip, dim=2, keepdim=True)
    return idx, recip / norm


def fp_forward(state: dict, spec: dict, xyz1_cf: torch.Tensor, xyz2_cf: torch.Tensor,
               points1_cf: torch.Tensor | None, points2_cf: torch.Tensor, return_aux: bool = False):
    """Feature propagation: xyz1 [B,3,N], xyz2 [B,3,S], points1 [B,D1,N] or None, points2 [B,D2,S]
    -> [B,D',N] (inverse-distance interpolation of points2 onto xyz1, concat, Conv1d+BN1d+ReLU stack)."""
    xyz1 = xyz1_cf.permute(0, 2, 1).contiguous()
    xyz2 = xyz2_cf.permute(0, 2, 1).contiguous()
    p2 = points2_cf.permute(0, 2, 1).contiguous()
    B, N, _ = xyz1.shape
    S = xyz2.shape[1]
    aux = {}
    if S == 1:
        up = p2.repeat(1, N, 1)
    else:
        idx, w = three_nn_weights(xyz1, xyz2)
        aux = {"idx": idx, "weight": w}
        up = torch.sum(take_rows(p2, idx) * w.view(B, N, 3, 1), dim=2)
    h = up if points1_cf is None else torch.cat([points1_cf.permute(0, 2, 1).contiguous(), up], dim=-1)
    h = h.permute(0, 2, 1).contiguous()
    for layer in _layers(state, "mlp_convs.{j}", "mlp_bns.{j}", len(spec["mlp"])):
        w_, b_, gamma, beta, mean, var = (t.to(h.device) for t in layer)
        h = F.relu(F.batch_norm(F.conv1d(h, w_, b_), mean, var, gamma, beta, False, 0.1, BN_EPS))
    return (h, aux) if return_aux else h


def decoder_forward(states: dict, specs: dict, l0_xyz, l1_xyz, l2_xyz, l3_xyz, l1_points, l2_points, l3_points):
    """fp3 -> fp2 -> fp1 as TEHNet.forward wires them (TEHNet.py:184-186) -> per-point features [B,256,N]."""
    l2p = fp_forward(states["fp3"], specs["fp3"], l2_xyz, l3_xyz, l2_points, l3_points)
    l1p = fp_forward(states["fp2"], specs["fp2"], l1_xyz, l2_xyz, l1_points, l2p)
    return fp_forward(states["fp1"], specs["fp1"], l0_xyz, l1_xyz, None, l1p), l1p, l2p
