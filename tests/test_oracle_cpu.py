"""Pins the two oracles (oracle/sa_oracle.py in torch, oracle/sa_oracle_c.c in C)
to the golden vectors the real reference produced (tests/golden/make_golden.py).
Indices must be bit-exact; features agree to 1e-6 of the tensor's max."""
import numpy as np
import pytest
import torch

from ev2hands_b200 import synth
from oracle import c_oracle, sa_oracle


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def test_fps_matches_reference(golden):
    g = golden("fps_ball")
    xyz, start = g["xyz"], g["start"]
    want = g["fps_idx"].astype(np.int64)
    assert np.array_equal(sa_oracle.fps(t(xyz), 512, t(start)).numpy(), want)
    assert np.array_equal(c_oracle.fps(xyz, 512, start), want)


def test_sqdist_bitwise(golden):
    g = golden("fps_ball")
    xyz = g["xyz"][:1]
    centres = xyz[:, g["fps_idx"][0, :8].astype(np.int64)]
    want = g["sqdist_w0_first8"]
    got_c = c_oracle.sqdist(centres, xyz)[0]
    assert np.array_equal(got_c.view(np.uint32), want.view(np.uint32))
    got_t = sa_oracle.pairwise_sqdist(t(centres), t(xyz))[0].numpy()
    assert np.array_equal(got_t.view(np.uint32), want.view(np.uint32))


@pytest.mark.parametrize("r,k", [(0.1, 32), (0.2, 64), (0.4, 128)])
def test_ball_query_matches_reference(golden, r, k):
    g = golden("fps_ball")
    xyz = g["xyz"]
    idx = g["fps_idx"].astype(np.int64)
    centres = np.stack([xyz[b, idx[b]] for b in range(xyz.shape[0])])
    want = g["ball_r%g" % r].astype(np.int64)
    assert np.array_equal(c_oracle.ball_query(r, k, xyz, centres), want)
    assert np.array_equal(sa_oracle.ball_query(r, k, t(xyz), t(centres)).numpy(), want)


def test_edge_cases(golden):
    g = golden("edge")
    # identical points: FPS returns start then index 0 forever; ball = first K indices
    assert np.array_equal(c_oracle.fps(g["same_xyz"], 16, g["same_start"]), g["same_fps"])
    assert np.array_equal(sa_oracle.fps(t(g["same_xyz"]), 16, t(g["same_start"])).numpy(), g["same_fps"])
    assert np.array_equal(c_oracle.ball_query(0.2, 8, g["same_xyz"], g["same_xyz"][:, :4]), g["same_ball"])
    # S == N on a coarse grid (exact ties everywhere)
    assert np.array_equal(c_oracle.fps(g["grid_xyz"], 96, g["grid_start"]), g["grid_fps"])
    gx = g["grid_xyz"]
    gc = np.stack([gx[b, g["grid_fps"][b, :24]] for b in range(2)])
    assert np.array_equal(c_oracle.ball_query(0.5, 16, gx, gc), g["grid_ball_r0.5_k16"])
    assert np.array_equal(sa_oracle.ball_query(0.5, 16, t(gx), t(gc)).numpy(), g["grid_ball_r0.5_k16"])
    # ragged N, radii whose fp32 square rounds up, singleton and all-inclusive balls
    ox = g["odd_xyz"]
    assert np.array_equal(c_oracle.fps(ox, 40, g["odd_start"]), g["odd_fps"])
    oc = np.stack([ox[b, g["odd_fps"][b]] for b in range(2)])
    for r, k in [(0.3, 16), (0.7, 48), (1e-3, 4), (4.0, 301)]:
        want = g["odd_ball_r%g_k%d" % (r, k)]
        assert np.array_equal(c_oracle.ball_query(r, k, ox, oc), want), (r, k)
        assert np.array_equal(sa_oracle.ball_query(r, k, t(ox), t(oc)).numpy(), want), (r, k)
    # centre with no neighbour: every slot is N (the reference's sentinel survives)
    assert np.array_equal(c_oracle.ball_query(0.2, 4, ox, g["far_centres"]), g["far_ball_r0.2_k4"])
    assert (g["far_ball_r0.2_k4"] == 301).all()


def _close(got, want, tol=1e-6):
    scale = np.abs(want).max()
    return np.abs(got - want).max() <= tol * scale


def test_encoder_matches_reference(golden):
    g = golden("encoder")
    states = {n: synth.random_state_for(synth.ENCODER_SPECS[n], seed=int(s))
              for n, s in zip(("sa1", "sa2", "sa3"), g["weight_seeds"])}
    starts = {"sa1": t(g["start_sa1"]), "sa2": t(g["start_sa2"])}
    with torch.no_grad():
        l3, aux = sa_oracle.encoder_forward(states, synth.ENCODER_SPECS, t(g["events"]), starts, return_aux=True)
    assert np.array_equal(aux["sa1"]["fps_idx"].numpy(), g["fps_sa1"].astype(np.int64))
    assert np.array_equal(aux["sa2"]["fps_idx"].numpy(), g["fps_sa2"].astype(np.int64))
    assert np.array_equal(aux["sa2"]["ball_idx"][0].numpy(), g["ball_sa2_r0.4"].astype(np.int64))
    assert np.array_equal(aux["sa2"]["ball_idx"][1].numpy(), g["ball_sa2_r0.8"].astype(np.int64))
    assert np.array_equal(aux["l1_xyz"].numpy(), g["l1_xyz"])
    assert np.array_equal(aux["l2_xyz"].numpy(), g["l2_xyz"])
    assert _close(aux["l1_points"][0].numpy(), g["l1_points_w0"])
    assert _close(aux["l2_points"][0].numpy(), g["l2_points_w0"])
    assert _close(l3.numpy(), g["l3_points"])
    # C index oracle on the second layer's shapes (N=512 -> S=128)
    xyz1 = np.ascontiguousarray(g["l1_xyz"].transpose(0, 2, 1))
    f2 = c_oracle.fps(xyz1, 128, g["start_sa2"])
    assert np.array_equal(f2, g["fps_sa2"].astype(np.int64))
    c2 = np.stack([xyz1[b, f2[b]] for b in range(2)])
    assert np.array_equal(c_oracle.ball_query(0.8, 128, xyz1, c2), g["ball_sa2_r0.8"].astype(np.int64))


def test_regressor_matches_reference(golden):
    g = golden("regressor")
    states = {n: synth.random_state_for(synth.REGRESSOR_SPECS[n], seed=int(s))
              for n, s in zip(("sa1", "sa2"), g["weight_seeds"])}
    with torch.no_grad():
        out, aux = sa_oracle.regressor_sa_forward(states, synth.REGRESSOR_SPECS, t(g["events"][:, :3]),
                                                  t(g["hand_feats"]), t(g["start_sa1"]), return_aux=True)
    assert np.array_equal(aux["l1_xyz"].numpy(), g["r1_xyz"])
    assert _close(aux["l1_points"][0].numpy(), g["r1_points_w0"])
    assert _close(out.numpy(), g["r2_points"])


def test_f64_mlp_yardstick_agrees_with_torch():
    rs = np.random.RandomState(5)
    spec = dict(kind="all", in_channel=19, mlp=[24, 40])
    st = synth.random_state_for(spec, seed=9)
    x = rs.randn(3, 7, 19).astype(np.float32)            # [G,K,C]
    layers_np = [(st["mlp_convs.%d.weight" % j], st["mlp_convs.%d.bias" % j], st["mlp_bns.%d.weight" % j],
                  st["mlp_bns.%d.bias" % j], st["mlp_bns.%d.running_mean" % j], st["mlp_bns.%d.running_var" % j])
                 for j in range(2)]
    got = c_oracle.mlp_max_f64(x, layers_np)
    xt = t(x).permute(2, 1, 0).unsqueeze(0).contiguous()   # [1,C,K,G]
    want = sa_oracle.shared_mlp_max(xt, [tuple(t(a) for a in l) for l in layers_np])[0].numpy().T
    assert np.abs(got - want).max() <= 1e-5 * np.abs(want).max()


def test_decoder_matches_reference(golden):
    """fp3 -> fp2 -> fp1: 3-NN indices bit-exact, weights and features to 1e-6."""
    e, g = golden("encoder"), golden("decoder")
    states = {n: synth.random_state_for(synth.DECODER_SPECS[n], seed=int(s))
              for n, s in zip(("fp3", "fp2", "fp1"), g["weight_seeds"])}
    f1, f2, f3 = (t(a) for a in synth.decoder_test_features(2, seed=int(g["feature_seed"])))
    l0_xyz = t(e["events"])[:, :3, :]
    l1_xyz, l2_xyz, l3_xyz = t(e["l1_xyz"]), t(e["l2_xyz"]), t(e["l3_xyz"])
    with torch.no_grad():
        d0, d1, d2 = sa_oracle.decoder_forward(states, synth.DECODER_SPECS, l0_xyz, l1_xyz, l2_xyz, l3_xyz, f1, f2, f3)
        for tag, q, src in (("fp2", l1_xyz, l2_xyz), ("fp1", l0_xyz, l1_xyz)):
            idx, w = sa_oracle.three_nn_weights(q.permute(0, 2, 1).contiguous(), src.permute(0, 2, 1).contiguous())
            assert np.array_equal(idx.numpy(), g[tag + "_idx"].astype(np.int64))
            assert _close(w.numpy(), g[tag + "_weight"])
    assert _close(d2.numpy(), g["d2"])
    assert _close(d1.numpy()[:, :, ::2], g["d1_every2"])
    assert _close(d0.numpy()[:, :, ::16], g["d0_every16"])


# ------------------------------------------------------------------ event windows ----
def _untied_columns(rec, idx):
    """columns of a window whose source pixel's mean time is unique (the reference's argsort is unstable)"""
    tt = rec[:, 2]
    u, c = np.unique(tt, return_counts=True)
    return ~np.isin(tt, u[c > 1])[idx]


@pytest.mark.parametrize("case,mode", [("stream", "stream"), ("erpc", "erpc"), ("erpct", "erpc"), ("erpcpad", "erpc")])
def test_window_oracle_matches_reference(golden, case, mode):
    """oracle/window_oracle.py against outputs of the reference's own __getitem__ methods
    (tests/golden/make_window_golden.py): bit for bit, except where equal mean times leave the reference's
    unstable sort a choice ("erpct"), where time and every untied point must still agree exactly."""
    from oracle import window_oracle as wo
    g = golden("windows")
    ev = g[case + "_events"]
    for b, (s, c) in enumerate(zip(g[case + "_starts"], g[case + "_counts"])):
        rec = wo.aggregate(ev[s:s + c], mode)
        assert rec.shape[0] == g[case + "_M"][b]
        got = wo.sample_normalize(rec, g[case + "_idx"][b])
        want = g[case + "_windows"][b]
        if case == "erpct":
            keep = _untied_columns(rec, g[case + "_idx"][b])
            assert 0 < keep.sum() < keep.size
            assert np.array_equal(got[2], want[2]) and np.array_equal(got[:, keep], want[:, keep])
        else:
            assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_window_oracle_sums_like_add_at():
    """the per-pixel time sum is float32(double(cell) + t) event by event (np.add.at on a float32 grid)"""
    from oracle import window_oracle as wo
    ev = np.array([[3, 2, 16777217.0, 1], [3, 2, 3.0, 0], [3, 2, 1e-3, 1], [5, 0, 7.0, 0]], dtype=np.float64)
    rec = wo.aggregate(ev, "erpc")
    acc = np.float32(0)
    for v in ev[:3, 2]:
        acc = np.float32(np.float64(acc) + v)
    want_t = np.float32(acc / np.float32(3)) * np.float32(1e-6)
    assert rec.shape == (2, 5)
    # sorted by mean time: pixel (5,0) first, rebased to it
    assert rec[0].tolist()[:2] == [5.0, 0.0] and rec[1].tolist()[:2] == [3.0, 2.0]
    assert rec[1, 2] == np.float32(want_t - np.float32(np.float32(7.0) * np.float32(1e-6)))
    assert rec[1, 3] == 2 and rec[1, 4] == 1
