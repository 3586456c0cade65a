"""Deterministic synthetic event windows and random encoder weights.

The reference builds its network input in ``src/Ev2Hands/dataset/erpc.py:170-255``:
2048 raw events are accumulated per pixel on the 346x260 sensor grid (mean
timestamp, positive count, negative count), the surviving pixels are re-sampled
to exactly N points *with replacement* (``erpc.py:213``) and x / y / t are
normalised to [-1, 1] (``erpc.py:23-37``).  There is no dataset in the tree
(``.MISSING_LARGE_BLOBS``), so every test and benchmark of this repo runs on
windows drawn here from the same recipe: two Gaussian "hand" blobs plus sparse
uniform noise.  Everything uses ``numpy.random.RandomState`` (a frozen
algorithm) so a seed means the same bytes on every machine.
"""
from __future__ import annotations

import numpy as np

SENSOR_W = 346   # src/settings.py:21
SENSOR_H = 260   # src/settings.py:22


def make_windows(n_windows: int, n_points: int = 2048, seed: int = 1234,
                 mode: str = "events", raw_events: int | None = None) -> np.ndarray:
    """Return float32 ``[n_windows, 5, n_points]`` = (x, y, t, n_pos, n_neg).

    mode="events": clustered, duplicate-heavy windows imitating erpc.py.
    mode="uniform": i.i.d. uniform points in [-1, 1]^3 (sparse worst case).
    """
    rs = np.random.RandomState(seed)
    out = np.empty((n_windows, 5, n_points), dtype=np.float32)
    if mode == "uniform":
        out[:, :3] = (rs.rand(n_windows, 3, n_points) * 2.0 - 1.0).astype(np.float32)
        out[:, 3:] = rs.randint(0, 4, size=(n_windows, 2, n_points)).astype(np.float32)
        return out
    if mode != "events":
        raise ValueError("mode must be 'events' or 'uniform'")
    n_raw = int(raw_events if raw_events is not None else n_points)
    for w in range(n_windows):
        n_noise = max(1, n_raw // 32)
        n_sig = n_raw - n_noise
        n_a = n_sig // 2
        centres = np.stack([rs.uniform(60, SENSOR_W - 60, size=2),
                            rs.uniform(50, SENSOR_H - 50, size=2)], axis=1)
        sigma = rs.uniform(18.0, 32.0, size=2)
        xs = np.concatenate([
            rs.normal(centres[0, 0], sigma[0], n_a),
            rs.normal(centres[1, 0], sigma[1], n_sig - n_a),
            rs.uniform(0, SENSOR_W, n_noise)])
        ys = np.concatenate([
            rs.normal(centres[0, 1], sigma[0], n_a),
            rs.normal(centres[1, 1], sigma[1], n_sig - n_a),
            rs.uniform(0, SENSOR_H, n_noise)])
        xi = np.clip(xs, 0, SENSOR_W - 1).astype(np.int64)
        yi = np.clip(ys, 0, SENSOR_H - 1).astype(np.int64)
        ts = rs.uniform(0.0, 5.0, n_raw)                      # ms inside the window
        pol = rs.rand(n_raw) < 0.5

        flat = yi * SENSOR_W + xi
        uniq, inv = np.unique(flat, return_inverse=True)
        cnt = np.bincount(inv, minlength=uniq.size).astype(np.float64)
        t_avg = np.bincount(inv, weights=ts, minlength=uniq.size) / cnt
        n_pos = np.bincount(inv, weights=pol.astype(np.float64), minlength=uniq.size)
        n_neg = cnt - n_pos

        order = np.argsort(t_avg, kind="stable")
        px = (uniq % SENSOR_W)[order].astype(np.float32)
        py = (uniq // SENSOR_W)[order].astype(np.float32)
        pt = t_avg[order].astype(np.float32)
        pp = n_pos[order].astype(np.float32)
        pn = n_neg[order].astype(np.float32)
        pt = pt - pt[0]

        pick = rs.randint(0, px.size, size=n_points)          # with replacement
        px, py, pt, pp, pn = px[pick], py[pick], pt[pick], pp[pick], pn[pick]

        px = np.float32(2.0) * (px / np.float32(SENSOR_W)) - np.float32(1.0)
        py = np.float32(2.0) * (py / np.float32(SENSOR_H)) - np.float32(1.0)
        t_min, t_max = pt.min(), pt.max()
        span = t_max - t_min if t_max > t_min else np.float32(1.0)
        pt = np.float32(2.0) * ((pt - t_min) / span) - np.float32(1.0)
        out[w, 0], out[w, 1], out[w, 2], out[w, 3], out[w, 4] = px, py, pt, pp, pn
    return out


def make_start_indices(n_windows: int, n_points: int, seed: int = 0) -> np.ndarray:
    """FPS start indices, int64 ``[n_windows]``; stands in for the reference's
    ``torch.randint(0, N, (B,))`` (pointnet2_utils.py:75) with a portable RNG."""
    return np.random.RandomState(seed).randint(0, n_points, size=n_windows).astype(np.int64)


def random_sa_state(prefix_convs: str, prefix_bns: str, channel_lists, in_channels,
                    seed: int) -> dict:
    """Random conv + BatchNorm tensors for one set-abstraction module.

    ``channel_lists`` is a list (one entry per radius scale) of MLP widths and
    ``in_channels`` the matching first-layer input widths.  Names follow the
    reference ``state_dict`` (pointnet2_utils.py:167-173, :210-222).  BN running
    statistics and affine terms are randomised because at their init values the
    fold is an identity and folding bugs would hide (SURVEY.md 8c caveat 3).
    For a single-scale module pass prefix patterns without ``{i}``.
    """
    rs = np.random.RandomState(seed)
    state = {}
    for i, (widths, cin) in enumerate(zip(channel_lists, in_channels)):
        last = cin
        for j, cout in enumerate(widths):
            bound = 1.0 / np.sqrt(last)
            cp = prefix_convs.format(i=i, j=j)
            bp = prefix_bns.format(i=i, j=j)
            state[cp + ".weight"] = rs.uniform(-bound, bound, (cout, last, 1, 1)).astype(np.float32)
            state[cp + ".bias"] = rs.uniform(-bound, bound, (cout,)).astype(np.float32)
            state[bp + ".weight"] = rs.uniform(0.5, 1.5, (cout,)).astype(np.float32)
            state[bp + ".bias"] = (0.1 * rs.randn(cout)).astype(np.float32)
            state[bp + ".running_mean"] = (0.1 * rs.randn(cout)).astype(np.float32)
            state[bp + ".running_var"] = rs.uniform(0.5, 1.5, (cout,)).astype(np.float32)
            state[bp + ".num_batches_tracked"] = np.array(0, dtype=np.int64)
            last = cout
    return state


# The five set-abstraction instances of the model and their constructor
# arguments (TEHNet.py:127-129 for the encoder, TEHNet.py:43-44 for each hand's
# regressor).  ``in_channel`` for "msg" excludes the +3 the module adds itself.
ENCODER_SPECS = {
    "sa1": dict(kind="msg", npoint=512, radius_list=[0.1, 0.2, 0.4], nsample_list=[32, 64, 128],
                in_channel=5, mlp_list=[[32, 32, 64], [64, 64, 128], [64, 96, 128]]),
    "sa2": dict(kind="msg", npoint=128, radius_list=[0.4, 0.8], nsample_list=[64, 128],
                in_channel=320, mlp_list=[[128, 128, 256], [128, 196, 256]]),
    "sa3": dict(kind="all", in_channel=515, mlp=[256, 512, 1024]),
}
REGRESSOR_SPECS = {
    "sa1": dict(kind="msg", npoint=128, radius_list=[0.4, 0.8], nsample_list=[64, 128],
                in_channel=4, mlp_list=[[128, 128, 256], [128, 196, 256]]),
    "sa2": dict(kind="all", in_channel=515, mlp=[256, 512]),
}


# The decoder's feature-propagation blocks (TEHNet.py:130-133), SURVEY.md section 8f row N1.
DECODER_SPECS = {
    "fp3": dict(kind="fp", in_channel=1536, mlp=[256, 256]),
    "fp2": dict(kind="fp", in_channel=576, mlp=[256, 128]),
    "fp1": dict(kind="fp", in_channel=128, mlp=[128, 128, 256]),
}


def decoder_test_features(batch: int, seed: int):
    """Seeded stand-ins for (l1_points [B,320,512], l2_points [B,512,128], l3_points [B,1024,1]), the feature
    inputs of fp2 / fp3 (TEHNet.py:184-185), for the decoder parity fixtures."""
    rs = np.random.RandomState(seed)
    return tuple(rs.randn(batch, c, n).astype(np.float32) for c, n in ((320, 512), (512, 128), (1024, 1)))


def random_state_for(spec: dict, seed: int) -> dict:
    if spec["kind"] == "fp":      # Conv1d / BatchNorm1d stack: weights are [Cout, Cin, 1]
        st = random_sa_state("mlp_convs.{j}", "mlp_bns.{j}", [spec["mlp"]], [spec["in_channel"]], seed)
        return {k: (v.reshape(v.shape[:3]) if k.endswith(".weight") and v.ndim == 4 else v) for k, v in st.items()}
    if spec["kind"] == "msg":
        cins = [spec["in_channel"] + 3] * len(spec["mlp_list"])
        return random_sa_state("conv_blocks.{i}.{j}", "bn_blocks.{i}.{j}",
                               spec["mlp_list"], cins, seed)
    return random_sa_state("mlp_convs.{j}", "mlp_bns.{j}", [spec["mlp"]],
                           [spec["in_channel"]], seed)


def random_head_state(seed: int, num_classes: int = 4, width: int = 256) -> dict:
    """Random weights for the heads that follow the decoder, named like the reference's
    ``TEHNet.state_dict()`` (TEHNet.py:135-166): ``classifier`` (Conv1d k=1, ReLU, BatchNorm1d, Dropout, Conv1d k=1)
    and ``left_query_conv`` / ``right_query_conv`` (Conv1d k=3, ReLU, BatchNorm1d, Dropout, Conv1d k=3, BatchNorm1d).
    BatchNorm statistics and affine terms are randomised like in ``random_sa_state``."""
    rs = np.random.RandomState(seed)
    st = {}

    def conv(name, cout, cin, k):
        bound = 1.0 / np.sqrt(cin * k)
        st[name + ".weight"] = rs.uniform(-bound, bound, (cout, cin, k)).astype(np.float32)
        st[name + ".bias"] = rs.uniform(-bound, bound, (cout,)).astype(np.float32)

    def bn(name, c):
        st[name + ".weight"] = rs.uniform(0.5, 1.5, (c,)).astype(np.float32)
        st[name + ".bias"] = (0.1 * rs.randn(c)).astype(np.float32)
        st[name + ".running_mean"] = (0.1 * rs.randn(c)).astype(np.float32)
        st[name + ".running_var"] = rs.uniform(0.5, 1.5, (c,)).astype(np.float32)
        st[name + ".num_batches_tracked"] = np.array(0, dtype=np.int64)

    conv("classifier.0", width, width, 1)
    bn("classifier.2", width)
    conv("classifier.4", num_classes, width, 1)
    for side in ("left", "right"):
        p = side + "_query_conv"
        conv(p + ".0", width, width, 3)
        bn(p + ".2", width)
        conv(p + ".4", width, width, 3)
        bn(p + ".5", width)
    return st


def make_raw_events(n_events: int, seed: int = 0, t0: float = 0.0, duration: float = 5.0e6,
                    extra_columns: int = 0, unique_times: bool = False) -> np.ndarray:
    """A raw event stream, float64 ``[n_events, 4 + extra_columns]`` = (x, y, t, polarity, ...), time ordered:
    the rows the reference's window builders consume (``dataset/erpc.py:170-176`` reads x, y, t [ns], p and two
    bookkeeping columns from the HDF5 table; ``dataset/evaluation_stream.py:108-127`` rows of x, y, t [ms], p).
    Same spatial recipe as ``make_windows``: two Gaussian blobs plus 1/32 uniform noise on the 346x260 sensor."""
    rs = np.random.RandomState(seed)
    n_noise = max(1, n_events // 32)
    n_sig = n_events - n_noise
    n_a = n_sig // 2
    centres = np.stack([rs.uniform(60, SENSOR_W - 60, size=2), rs.uniform(50, SENSOR_H - 50, size=2)], axis=1)
    sigma = rs.uniform(18.0, 32.0, size=2)
    xs = np.concatenate([rs.normal(centres[0, 0], sigma[0], n_a), rs.normal(centres[1, 0], sigma[1], n_sig - n_a),
                         rs.uniform(0, SENSOR_W, n_noise)])
    ys = np.concatenate([rs.normal(centres[0, 1], sigma[0], n_a), rs.normal(centres[1, 1], sigma[1], n_sig - n_a),
                         rs.uniform(0, SENSOR_H, n_noise)])
    order = rs.permutation(n_events)
    out = np.zeros((n_events, 4 + extra_columns), dtype=np.float64)
    out[:, 0] = np.clip(xs, 0, SENSOR_W - 1).astype(np.int64)[order]
    out[:, 1] = np.clip(ys, 0, SENSOR_H - 1).astype(np.int64)[order]
    if unique_times:      # no two events share a timestamp (integer ticks drawn without replacement)
        out[:, 2] = t0 + np.sort(rs.choice(int(duration), size=n_events, replace=False)).astype(np.float64)
    else:
        out[:, 2] = t0 + np.sort(np.floor(rs.uniform(0.0, duration, n_events)))
    out[:, 3] = (rs.rand(n_events) < 0.5).astype(np.float64)
    return out
