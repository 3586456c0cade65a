"""The network around the set-abstraction path, wired like the reference's ``TEHNet`` so that the training
configuration (BASELINE.json configs[3]) and the rows next to the path (N2 classifier + attention,
N4 regressor head) can be exercised end to end on the CUDA kernels of this package.

Reference: ``src/Ev2Hands/model/TEHNet.py`` - ``AttentionBlock`` :9-27, ``MANORegressor`` :30-112, ``TEHNet``
:115-197.  Submodule names, constructor arguments and tensor shapes follow it exactly: they are the checkpoint
contract (``demo.py:84`` loads with ``strict=True``; ``tests/test_host_cpu.py`` compares the ``state_dict`` layouts
against the real reference).  A reference user does not need this file - pointing ``TEHNet.py:6`` at
``ev2hands_b200`` is the whole integration - it exists because ``/root/reference`` is not on the
GPU box and ``manopth`` / ``mesh_intersection`` are not importable anywhere here.

What is NOT the reference's: ``StandInManoLayer`` (MANO's assets are licence gated: a random but fixed linear-blend
-skinning layer of MANO's shape - 778 vertices, 16 joints + 5 fingertips, 10 shape and ``n_cmps`` pose components -
with the call signature of ``model/utils.py:25``) and ``training_losses`` (``losses.py:145-206`` without
``loss_interpen``, which needs ``mesh_intersection``).  Parity of those two parts is therefore UNPINNED; everything
else is compared with reference outputs in the tests.
"""
from __future__ import annotations

import os
from types import SimpleNamespace

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _capi
from . import pointnet2_utils as _pu
from .pointnet2_utils import PointNetFeaturePropagation, PointNetSetAbstraction, PointNetSetAbstractionMsg


# ---------------------------------------------------------------------------------------------------------------
# segmentation classifier + per-hand query convolutions + attention on the CUDA kernels
# (eval mode, no autograd; training keeps the PyTorch modules).  Reference: TEHNet.py:135-166 (modules), :188-192 (use).
# ---------------------------------------------------------------------------------------------------------------
def _bn_affine(bn):
    """eval-mode BatchNorm1d as y = a x + b (fp64 on the device)"""
    a = bn.weight.double() / torch.sqrt(bn.running_var.double() + bn.eps)
    return a, bn.bias.double() - bn.running_mean.double() * a


def _pack_conv1d(w, bias, mode):
    """Conv1d weight [Cout, Cin, T] (+ bias [Cout]), fp64 -> (packed tensor-core image, padded fp32 bias, T, Cin, Cout)"""
    cout, cin, taps = w.shape
    rows, cols = (taps * cin + 15) // 16 * 16, (cout + 127) // 128 * 128
    wt = torch.zeros((rows, cols), dtype=torch.float32, device=w.device)
    wt[:taps * cin, :cout] = w.permute(2, 1, 0).reshape(taps * cin, cout).float()      # row t * Cin + ci = W[:, ci, t]
    b = torch.zeros((cols,), dtype=torch.float32, device=w.device)
    b[:cout] = bias.float()
    return {"packed": _capi.tc_pack(wt, taps * cin, cout, mode), "bias": b, "taps": taps, "cin": cin, "cout": cout}


class _HeadWeights:
    """folded + packed weights of the classifier and the two query convolutions, per (device, arithmetic mode),
    rebuilt when any source tensor changed (data_ptr / _version), like pointnet2_utils._FoldedMLP"""

    def __init__(self):
        self.cache = {}

    @staticmethod
    def _key(net):
        k = []
        for seq in (net.classifier, net.left_query_conv, net.right_query_conv):
            for t in list(seq.parameters()) + list(seq.buffers()):
                k.append((t.data_ptr(), t._version))
        return tuple(k)

    def get(self, net, mode):
        dev = net.classifier[0].weight.device
        key = self._key(net)
        with _pu._CACHE_LOCK:
            hit = self.cache.get((dev, mode))
            if hit is not None and hit[0] == key:
                return hit[1]
        with torch.no_grad():
            out = {}
            c1, bn, c2 = net.classifier[0], net.classifier[2], net.classifier[4]
            a, b = _bn_affine(bn)
            out["cls1"] = _pack_conv1d(c1.weight.double(), c1.bias.double(), mode)
            # Conv -> ReLU -> BN -> (Dropout) -> Conv with kernel size 1: the BatchNorm folds into the second convolution
            w2 = c2.weight.double()[:, :, 0]
            out["cls2"] = _pack_conv1d((w2 * a[None, :]).unsqueeze(-1), c2.bias.double() + w2 @ b, mode)
            for side, seq in (("left", net.left_query_conv), ("right", net.right_query_conv)):
                q1, bn1, q2, bn2 = seq[0], seq[2], seq[4], seq[5]
                a1, b1 = _bn_affine(bn1)
                a2, b2 = _bn_affine(bn2)
                d = {"q1": _pack_conv1d(q1.weight.double(), q1.bias.double(), mode),
                     # the BatchNorm after the ReLU cannot fold into the next 3-tap convolution (its zero padding is applied
                     # AFTER the BatchNorm): it is applied in the first convolution's epilogue instead
                     "post_a": a1.float().contiguous(), "post_b": b1.float().contiguous(),
                     # Conv -> BN folds as usual
                     "q2": _pack_conv1d(q2.weight.double() * a2[:, None, None], q2.bias.double() * a2 + b2, mode)}
                out[side] = d
        with _pu._CACHE_LOCK:
            self.cache[(dev, mode)] = (key, out)
        return out


def _conv_rows(x, M, ld_x, L, rows_per_seq, relu, post=None):
    y = torch.empty((M, _pu._pad4(L["cout"])), dtype=torch.float32, device=x.device)
    _capi.conv1d_tc(x, M, ld_x, L["cin"], L["taps"], rows_per_seq, L["packed"], L["bias"], L["cout"], relu,
                    None if post is None else post[0], None if post is None else post[1], y, y.shape[1], 0, L["mode"])
    return y


class AttentionBlock(nn.Module):
    """class-wise attention pooling (TEHNet.py:9-27): key = segmentation logits [B,C,N], value = point features
    [B,D,N], query [B,D,N] -> context [B,C,N] = softmax_over_classes(D^-1/2 key query^T) value.  (The scale is the
    VALUE's channel count: the reference reassigns ``KC`` from ``value.shape`` before using it, :17-22.)"""

    def forward(self, key, value, query):
        B, C = key.shape[:2]
        D = value.shape[1]
        sim = torch.bmm(key.reshape(B, C, -1), query.permute(0, 2, 1)) * (D ** -0.5)       # [B,C,D]
        return torch.bmm(F.softmax(sim, dim=1), value.reshape(B, D, -1))


class MANORegressor(nn.Module):
    """one hand's head (TEHNet.py:30-112): set abstraction over the hand's attention features, then a two-layer
    regressor of (global_orient 3, hand_pose n_pose, betas n_shape, transl 3) fed to the MANO layer."""

    def __init__(self, n_inp_features=4, n_pose_params=6, n_shape_params=10):
        super().__init__()
        self.sa1 = PointNetSetAbstractionMsg(128, [0.4, 0.8], [64, 128], n_inp_features, [[128, 128, 256], [128, 196, 256]])
        self.sa2 = PointNetSetAbstraction(npoint=None, radius=None, nsample=None, in_channel=512 + 3, mlp=[256, 512], group_all=True)
        self.n_pose_params = n_pose_params
        self.n_mano_params = n_pose_params + n_shape_params
        self.mano_regressor = nn.Sequential(
            nn.Linear(512, 1024), nn.ReLU(), nn.BatchNorm1d(1024), nn.Dropout(0.3),
            nn.Linear(1024, 3 + self.n_mano_params + 3))

    def forward(self, xyz, features, mano_hand, previous_mano_params=None):
        l1_xyz, l1_points = self.sa1(xyz, features)
        _, l2_points = self.sa2(l1_xyz, l1_points)
        params = self.mano_regressor(l2_points.squeeze(-1))
        dev = mano_hand.shapedirs.device
        args = {"global_orient": params[:, :3].to(dev), "hand_pose": params[:, 3:3 + self.n_pose_params].to(dev),
                "betas": params[:, 3 + self.n_pose_params:-3].to(dev), "transl": params[:, -3:].to(dev)}
        out = mano_hand(**args)
        res = {"vertices": out.vertices, "j3d": out.joints}
        res.update(args)
        if not self.training:
            import numpy as np
            res["faces"] = np.tile(mano_hand.faces, (xyz.shape[0], 1, 1))
        return res


class TEHNet(nn.Module):
    """events [B, 3+extra, N] -> {'class_logits' [B,4,N], 'left': {...}, 'right': {...}} (TEHNet.py:115-197)."""

    def __init__(self, n_pose_params, num_classes=4):
        super().__init__()
        extra = 1 + int(os.getenv("ERPC", 0))
        self.sa1 = PointNetSetAbstractionMsg(512, [0.1, 0.2, 0.4], [32, 64, 128], 3 + extra,
                                             [[32, 32, 64], [64, 64, 128], [64, 96, 128]])
        self.sa2 = PointNetSetAbstractionMsg(128, [0.4, 0.8], [64, 128], 128 + 128 + 64, [[128, 128, 256], [128, 196, 256]])
        self.sa3 = PointNetSetAbstraction(npoint=None, radius=None, nsample=None, in_channel=512 + 3, mlp=[256, 512, 1024], group_all=True)
        self.fp3 = PointNetFeaturePropagation(in_channel=1536, mlp=[256, 256])
        self.fp2 = PointNetFeaturePropagation(in_channel=576, mlp=[256, 128])
        self.fp1 = PointNetFeaturePropagation(128, [128, 128, 256])
        self.classifier = nn.Sequential(nn.Conv1d(256, 256, 1), nn.ReLU(), nn.BatchNorm1d(256), nn.Dropout(0.3),
                                        nn.Conv1d(256, num_classes, 1))
        self.attention_block = AttentionBlock()
        self.left_mano_regressor = MANORegressor(n_pose_params=n_pose_params)
        self.right_mano_regressor = MANORegressor(n_pose_params=n_pose_params)
        self.mhlnes = int(os.getenv("MHLNES", 0))

        def query_conv():
            return nn.Sequential(nn.Conv1d(256, 256, 3, 1, 1), nn.ReLU(), nn.BatchNorm1d(256), nn.Dropout(0.1),
                                 nn.Conv1d(256, 256, 3, 1, 1), nn.BatchNorm1d(256))
        self.left_query_conv = query_conv()
        self.right_query_conv = query_conv()
        self._head_weights = _HeadWeights()

    def _heads_cuda(self, l0_points):
        """classifier, query convolutions and attention of both hands on the tensor-core / CUDA kernels:
        l0_points [B,256,N] -> (class_logits [B,C,N], left features [B,C,N], right features [B,C,N])"""
        B, D, N = l0_points.shape
        M = B * N
        prec = _pu.get_mlp_precision()
        mode = _pu._layer_mode({"tf32x3": _capi.TC_TF32X3, "bf16": _capi.TC_BF16}[prec])
        W = self._head_weights.get(self, mode)
        rec = getattr(l0_points, "_ev2h_fp_rows", None)
        if rec is not None and rec[3] == l0_points._version and rec[2] == D:
            x, ld = rec[0].view(M, rec[1]), rec[1]                      # the decoder's own rows
        else:
            x, ld = _pu._to_rows(l0_points).view(M, D), D
        for L in (W["cls1"], W["cls2"], W["left"]["q1"], W["left"]["q2"], W["right"]["q1"], W["right"]["q2"]):
            L["mode"] = mode
        h = _conv_rows(x, M, ld, W["cls1"], N, relu=True)
        logits = _conv_rows(h, M, h.shape[1], W["cls2"], N, relu=False)                 # [M, pad4(C)]
        C = W["cls2"]["cout"]
        seg = torch.empty((B, C, N), dtype=torch.float32, device=x.device)
        _capi.transpose(logits, (N * logits.shape[1], logits.shape[1], 1), B, N, C, seg, C * N, N, 0)
        feats = []
        for side in ("left", "right"):
            q = _conv_rows(x, M, ld, W[side]["q1"], N, relu=True, post=(W[side]["post_a"], W[side]["post_b"]))
            q = _conv_rows(q, M, q.shape[1], W[side]["q2"], N, relu=False)
            feats.append(_capi.class_attention(logits, logits.shape[1], q, q.shape[1], x, ld, B, N, C, D, D ** -0.5))
        return seg, feats[0], feats[1]

    def trunk(self, xyz):
        """everything up to the per-hand attention features: -> (l0_xyz, class_logits, left_feat, right_feat)"""
        l0_points = xyz
        l0_xyz = xyz[:, :3, :]
        if self.mhlnes:
            l0_xyz[:, -1, :] = xyz[:, 3:, :].mean(1)
        l1_xyz, l1_points = self.sa1(l0_xyz, l0_points)
        l2_xyz, l2_points = self.sa2(l1_xyz, l1_points)
        l3_xyz, l3_points = self.sa3(l2_xyz, l2_points)
        l2_points = self.fp3(l2_xyz, l3_xyz, l2_points, l3_points)
        l1_points = self.fp2(l1_xyz, l2_xyz, l1_points, l2_points)
        l0_points = self.fp1(l0_xyz, l1_xyz, None, l1_points)
        if (l0_points.is_cuda and not _pu._wants_autograd(self, l0_points) and _pu.get_mlp_precision() in ("tf32x3", "bf16")
                and self.classifier[0].kernel_size == (1,) and l0_points.shape[1] % 32 == 0):
            with torch.no_grad():
                seg_out, left, right = self._heads_cuda(l0_points)
            return l0_xyz, seg_out, left, right
        seg_out = self.classifier(l0_points)
        left = self.attention_block(seg_out, l0_points, self.left_query_conv(l0_points))
        right = self.attention_block(seg_out, l0_points, self.right_query_conv(l0_points))
        return l0_xyz, seg_out, left, right

    def forward(self, xyz, mano_hands):
        l0_xyz, seg_out, left_feat, right_feat = self.trunk(xyz)
        left = self.left_mano_regressor(l0_xyz, left_feat, mano_hands["left"])
        right = self.right_mano_regressor(l0_xyz, right_feat, mano_hands["right"])
        return {"class_logits": seg_out, "left": left, "right": right}


# ---------------------------------------------------------------------------------------------------------------
# stand-ins for the parts of the training step whose reference implementation cannot run here (parity unpinned)
# ---------------------------------------------------------------------------------------------------------------
def _rodrigues(aa):
    """axis-angle [..., 3] -> rotation matrices [..., 3, 3]"""
    angle = torch.sqrt((aa * aa).sum(-1, keepdim=True) + 1e-16)
    axis = aa / angle
    c, s = torch.cos(angle)[..., None], torch.sin(angle)[..., None]
    x, y, z = axis.unbind(-1)
    zero = torch.zeros_like(x)
    K = torch.stack([zero, -z, y, z, zero, -x, -y, x, zero], -1).reshape(aa.shape[:-1] + (3, 3))
    eye = torch.eye(3, device=aa.device, dtype=aa.dtype).expand_as(K)
    return eye + s * K + (1 - c) * (K @ K)


class StandInManoLayer(nn.Module):
    """A MANO-shaped linear-blend-skinning layer with random, fixed, seeded assets.  Same call signature and output
    object as the reference's ``SmplxAdapter.__call__`` (model/utils.py:25-31): (global_orient [B,3], hand_pose
    [B,n_cmps] PCA coefficients, betas [B,10], transl [B,3]) -> .vertices [B,778,3], .joints [B,21,3] in metres."""

    N_VERTS, N_JOINTS, N_TIPS = 778, 16, 5
    PARENTS = (-1, 0, 1, 2, 0, 4, 5, 0, 7, 8, 0, 10, 11, 0, 13, 14)

    def __init__(self, side: str = "right", n_cmps: int = 6, seed: int = 0):
        super().__init__()
        g = torch.Generator().manual_seed(seed + (1 if side == "left" else 0))
        V, J = self.N_VERTS, self.N_JOINTS
        self.register_buffer("v_template", 0.08 * torch.rand(V, 3, generator=g) - 0.04)
        self.register_buffer("shapedirs", 0.004 * torch.randn(V, 3, 10, generator=g))
        self.register_buffer("posedirs", 0.0005 * torch.randn((J - 1) * 9, V * 3, generator=g))
        jr = torch.rand(J, V, generator=g) ** 8
        self.register_buffer("j_regressor", jr / jr.sum(1, keepdim=True))
        w = torch.rand(V, J, generator=g) ** 6
        self.register_buffer("lbs_weights", w / w.sum(1, keepdim=True))
        self.register_buffer("hand_components", 0.5 * torch.randn(n_cmps, (J - 1) * 3, generator=g))
        self.register_buffer("hand_mean", 0.1 * torch.randn((J - 1) * 3, generator=g))
        self.register_buffer("tip_ids", torch.randint(0, V, (self.N_TIPS,), generator=g))
        f = torch.randint(0, V, (1538, 3), generator=g)
        self.faces = f.numpy()

    def forward(self, global_orient, hand_pose, betas, transl):
        B, V, J = global_orient.shape[0], self.N_VERTS, self.N_JOINTS
        full_pose = torch.cat([global_orient, self.hand_mean + hand_pose @ self.hand_components], 1).view(B, J, 3)
        R = _rodrigues(full_pose)                                                       # [B,16,3,3]
        v_shaped = self.v_template + torch.einsum("vck,bk->bvc", self.shapedirs, betas)
        joints = torch.einsum("jv,bvc->bjc", self.j_regressor, v_shaped)
        pose_feat = (R[:, 1:] - torch.eye(3, device=R.device)).reshape(B, -1)
        v_posed = v_shaped + (pose_feat @ self.posedirs).view(B, V, 3)
        # kinematic chain: world transform of every joint, relative to the rest pose
        rel = joints.clone()
        par = torch.tensor(self.PARENTS[1:], device=joints.device)
        rel[:, 1:] = joints[:, 1:] - joints[:, par]
        T = torch.cat([torch.cat([R, rel.unsqueeze(-1)], -1),
                       torch.tensor([0.0, 0.0, 0.0, 1.0], device=R.device).expand(B, J, 1, 4)], -2)       # [B,16,4,4]
        chain = [T[:, 0]]
        for j in range(1, J):
            chain.append(chain[self.PARENTS[j]] @ T[:, j])
        G = torch.stack(chain, 1)
        j_world = G[:, :, :3, 3]
        rest = torch.cat([joints, torch.zeros(B, J, 1, device=joints.device)], -1).unsqueeze(-1)
        G = G - F.pad(G @ rest, (3, 0))
        Tv = torch.einsum("vj,bjrc->bvrc", self.lbs_weights, G)
        verts = (Tv[:, :, :3, :3] @ v_posed.unsqueeze(-1)).squeeze(-1) + Tv[:, :, :3, 3]
        joints21 = torch.cat([j_world, verts[:, self.tip_ids]], 1)
        t = transl.unsqueeze(1)
        return SimpleNamespace(vertices=verts + t, joints=joints21 + t)


def create_standin_mano_layers(device, n_cmps: int = 6, seed: int = 0):
    """{'left': layer, 'right': layer} like ``create_mano_layers`` (model/utils.py:13-45), from random assets."""
    return {side: StandInManoLayer(side, n_cmps, seed).to(device) for side in ("left", "right")}


def _masked_mean(loss_fn, out, target, mask):
    """mean of an element-wise loss over the samples selected by ``mask`` (losses.py:122-136); 0 if none is."""
    mask = mask.to(out.dtype)
    if float(mask.sum()) == 0:
        return out.new_zeros(())
    per = loss_fn(out, target, reduction="none").reshape(out.shape[0], -1)
    return (per * mask[:, None]).sum() / (mask.sum() * per.shape[1])


def training_losses(outs: dict, targets: dict, hands: dict, n_cmps: int = 6) -> dict:
    """The supervised (``mano_gt``) branch of the reference's criterion, ``losses.py:145-206``, term for term and with
    its weights, EXCEPT ``loss_interpen`` (needs ``mesh_intersection``'s BVH kernels; absent here).  ``targets`` is the
    batch dict of ``Ev2HandSDataset`` (``erpc.py:251-296``); the ground-truth joints come from the same MANO layers."""
    losses = {}
    for side in ("left", "right"):
        t = targets[side]
        gt = hands[side](global_orient=t["global_orient"], hand_pose=t["hand_pose"][:, :n_cmps], betas=t["shape"], transl=t["trans"])
        t["j3d"], t["vertices"] = gt.joints, gt.vertices
    both = targets["handedness"].sum(1) == 2
    L, R = outs["left"], outs["right"]
    tl, tr = targets["left"], targets["right"]
    losses["loss_inter_shape"] = _masked_mean(F.mse_loss, L["betas"], R["betas"], both)
    losses["loss_inter_transl"] = 100 * _masked_mean(F.mse_loss, L["transl"] - R["transl"], tl["trans"] - tr["trans"], both)
    losses["loss_inter_j3d"] = 100 * _masked_mean(F.mse_loss, L["j3d"] - R["j3d"], tl["j3d"] - tr["j3d"], both)
    acc = {k: 0.0 for k in ("loss_global_orient", "loss_hand_pose", "loss_rj3d", "loss_j3d", "loss_shape", "loss_transl")}
    for side in ("left", "right"):
        o, t = outs[side], targets[side]
        m = t["valid"]
        acc["loss_global_orient"] = acc["loss_global_orient"] + 10 * _masked_mean(F.mse_loss, o["global_orient"], t["global_orient"], m)
        acc["loss_hand_pose"] = acc["loss_hand_pose"] + 10 * _masked_mean(F.mse_loss, o["hand_pose"], t["hand_pose"][:, :n_cmps], m)
        acc["loss_rj3d"] = acc["loss_rj3d"] + 0.01 * _masked_mean(F.l1_loss, (o["j3d"][:, 1:] - o["j3d"][:, :1]) * 1000,
                                                                   (t["j3d"][:, 1:] - t["j3d"][:, :1]) * 1000, m)
        acc["loss_j3d"] = acc["loss_j3d"] + 0.01 * _masked_mean(F.l1_loss, o["j3d"] * 1000, t["j3d"] * 1000, m)
        acc["loss_shape"] = acc["loss_shape"] + 10 * _masked_mean(F.mse_loss, o["betas"], t["shape"], m)
        acc["loss_transl"] = acc["loss_transl"] + 10 * _masked_mean(F.l1_loss, o["transl"], t["trans"], m)
    losses.update(acc)
    w = torch.tensor([1.0, 30.0, 30.0, 10.0], device=outs["class_logits"].device)
    losses["loss_class_logits"] = F.cross_entropy(outs["class_logits"], targets["class_logits"], weight=w, ignore_index=0)
    return losses


def make_training_batch(batch: int, n_points: int = 2048, seed: int = 0, device="cpu") -> dict:
    """Synthetic batch with the fields and shapes of ``Ev2HandSDataset.__getitem__`` after collation (erpc.py:251-296):
    events [B,5,N], per-point labels [B,N] in {0..3}, per-hand MANO parameters, handedness, mano_gt = 1."""
    import numpy as np
    from . import synth
    rs = np.random.RandomState(seed)
    ev = torch.from_numpy(synth.make_windows(batch, n_points, seed=1234 + 4 + seed))
    hand = lambda: {"global_orient": torch.from_numpy(rs.randn(batch, 3).astype("float32")) * 0.5,      # noqa: E731
                    "hand_pose": torch.from_numpy(rs.randn(batch, 45).astype("float32")) * 0.3,
                    "shape": torch.from_numpy(rs.randn(batch, 10).astype("float32")) * 0.5,
                    "trans": torch.from_numpy(rs.randn(batch, 3).astype("float32")) * 0.1,
                    "valid": torch.from_numpy(rs.rand(batch) < 0.9)}
    b = {"mano_gt": torch.ones(batch), "events": ev, "class_logits": torch.from_numpy(rs.randint(0, 4, size=(batch, n_points)).astype("int64")),
         "left": hand(), "right": hand()}
    b["handedness"] = torch.stack([b["left"]["valid"], b["right"]["valid"]], 1).int()

    def to(x):
        return {k: to(v) for k, v in x.items()} if isinstance(x, dict) else x.to(device)
    return to(b)
