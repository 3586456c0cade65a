"""The encoder stack of the model: sa1 -> sa2 -> sa3, wired exactly as
``TEHNet.forward`` does (reference ``src/Ev2Hands/model/TEHNet.py:127-129`` for the
constructor arguments, ``:172-181`` for the data flow), plus one hand regressor's
``sa1 -> sa2`` (``TEHNet.py:43-44``, ``:75-79``).  This is the unit BASELINE.json's
metric ("encoder event-windows/s") is measured on.
"""
from __future__ import annotations

import os
import threading

import torch
import torch.nn as nn

from .pointnet2_utils import (PointNetFeaturePropagation, PointNetSetAbstraction, PointNetSetAbstractionMsg,
                              _check_inputs, _wants_autograd)

# Inference: sa2's FPS and ball query need only sa1's centres, so they run on a second stream beside sa1's
# grouping + MLP kernels (EV2H_GEOM_STREAM=0 keeps everything on one stream).
_GEOM_STREAM = os.environ.get("EV2H_GEOM_STREAM", "1") != "0"
_side_streams = {}


_side_lock = threading.Lock()


def _side_stream(device):
    """one side stream per (device, host thread): nn.DataParallel runs a replica per thread"""
    key = (device, threading.get_ident())
    with _side_lock:
        st = _side_streams.get(key)
        if st is None:
            st = _side_streams[key] = torch.cuda.Stream(device=device)
    return st


class SetAbstractionEncoder(nn.Module):
    """events [B, 3+extra, N] -> per-window features [B, 1024].

    ``extra`` follows TEHNet.__init__ (``1 + int(os.getenv('ERPC', 0))``, TEHNet.py:122);
    the dataset modules set ERPC=1, which gives the 5-channel windows used everywhere."""

    def __init__(self, extra_channels: int | None = None):
        super().__init__()
        if extra_channels is None:
            extra_channels = 1 + int(os.getenv("ERPC", 1))
        self.sa1 = PointNetSetAbstractionMsg(512, [0.1, 0.2, 0.4], [32, 64, 128], 3 + extra_channels,
                                             [[32, 32, 64], [64, 64, 128], [64, 96, 128]])
        self.sa2 = PointNetSetAbstractionMsg(128, [0.4, 0.8], [64, 128], 128 + 128 + 64,
                                             [[128, 128, 256], [128, 196, 256]])
        self.sa3 = PointNetSetAbstraction(npoint=None, radius=None, nsample=None, in_channel=512 + 3,
                                          mlp=[256, 512, 1024], group_all=True)

    def forward(self, events, fps_starts=None, return_levels=False):
        """``fps_starts`` = optional (start_sa1, start_sa2) int64 [B] tensors replacing the two
        ``torch.randint`` draws the reference makes (in this order) per forward."""
        s1, s2 = fps_starts if fps_starts is not None else (None, None)
        l0_xyz = events[:, :3, :]
        if events.is_cuda and not _wants_autograd(self, events):
            return self._forward_inference(l0_xyz, events, s1, s2, return_levels)
        l1_xyz, l1_points = self.sa1(l0_xyz, events, fps_start=s1)
        l2_xyz, l2_points = self.sa2(l1_xyz, l1_points, fps_start=s2)
        _, l3_points = self.sa3(l2_xyz, l2_points)
        out = l3_points.squeeze(-1)
        if return_levels:
            return out, {"l1_xyz": l1_xyz, "l1_points": l1_points, "l2_xyz": l2_xyz, "l2_points": l2_points}
        return out


    def _forward_inference(self, l0_xyz, events, s1, s2, return_levels):
        """The same data flow on the kernels' point-major rows (``forward_rows``): each layer's pooled rows
        [features | centre xyz] are the next layer's input rows as they are, and the channel-first tensors of the
        reference's API are only materialised for ``return_levels``.  sa2's FPS + ball query run on a side stream
        as soon as sa1's centres exist."""
        _check_inputs(l0_xyz, events)
        B, _, N = events.shape
        # the reference draws sa1's start indices, then sa2's, from the CPU generator (pointnet2_utils.py:75)
        if s1 is None:
            s1 = torch.randint(0, N, (B,), dtype=torch.long)
        if s2 is None:
            s2 = torch.randint(0, self.sa1.npoint, (B,), dtype=torch.long)
        with torch.no_grad():
            # sa1's front end: sampling in ranges with the ball query of each finished range on a side stream, when it applies
            piped = self.sa1._pipelined_ok(l0_xyz, events.shape[1])
            g1 = self.sa1._fps_ball_pipelined(l0_xyz, events, events.shape[1], s1) if piped else self.sa1._fps(l0_xyz, s1)
            l1_xyz = g1["new_xyz"]
            d1 = sum(convs[-1].out_channels for convs in self.sa1.conv_blocks)
            if _GEOM_STREAM:
                main, side = torch.cuda.current_stream(events.device), _side_stream(events.device)
                # tensors allocated on the side stream stay referenced (g2) until main has waited for it below;
                # the side stream starts every forward behind main, which orders any reuse of their memory
                side.wait_stream(main)
                with torch.cuda.stream(side):
                    g2 = self.sa2._ball(self.sa2._fps(l1_xyz, s2), l1_xyz, None, d1, fused=True)
                if not piped:
                    self.sa1._ball(g1, l0_xyz, events, events.shape[1], fused=True)
                _, r1 = self.sa1.forward_rows(l0_xyz, events, geom=g1)
                main.wait_stream(side)
            else:
                if not piped:
                    self.sa1._ball(g1, l0_xyz, events, events.shape[1], fused=True)
                _, r1 = self.sa1.forward_rows(l0_xyz, events, geom=g1)
                g2 = self.sa2._ball(self.sa2._fps(l1_xyz, s2), l1_xyz, None, d1, fused=True)
            l2_xyz, r2 = self.sa2.forward_rows(l1_xyz, r1, geom=g2)
            out = self.sa3.forward_rows(l2_xyz, r2).squeeze(-1)
        if return_levels:
            return out, {"l1_xyz": l1_xyz, "l1_points": r1.cf(), "l2_xyz": l2_xyz, "l2_points": r2.cf()}
        return out


class FeaturePropagationDecoder(nn.Module):
    """fp3 -> fp2 -> fp1 as TEHNet wires them (constructor arguments TEHNet.py:130-133, data flow :184-186):
    the encoder's three levels -> per-point features [B, 256, N] (the input of the segmentation classifier)."""

    def __init__(self):
        super().__init__()
        self.fp3 = PointNetFeaturePropagation(in_channel=1536, mlp=[256, 256])
        self.fp2 = PointNetFeaturePropagation(in_channel=576, mlp=[256, 128])
        self.fp1 = PointNetFeaturePropagation(128, [128, 128, 256])

    def forward(self, l0_xyz, l1_xyz, l2_xyz, l3_xyz, l1_points, l2_points, l3_points, return_levels=False):
        l2_up = self.fp3(l2_xyz, l3_xyz, l2_points, l3_points)
        l1_up = self.fp2(l1_xyz, l2_xyz, l1_points, l2_up)
        l0_up = self.fp1(l0_xyz, l1_xyz, None, l1_up)
        return (l0_up, l1_up, l2_up) if return_levels else l0_up


class RegressorSetAbstraction(nn.Module):
    """The set-abstraction half of MANORegressor (TEHNet.py:43-44): (xyz [B,3,N],
    hand features [B,4,N]) -> [B, 512]."""

    def __init__(self, n_inp_features: int = 4):
        super().__init__()
        self.sa1 = PointNetSetAbstractionMsg(128, [0.4, 0.8], [64, 128], n_inp_features,
                                             [[128, 128, 256], [128, 196, 256]])
        self.sa2 = PointNetSetAbstraction(npoint=None, radius=None, nsample=None, in_channel=512 + 3,
                                          mlp=[256, 512], group_all=True)

    def forward(self, xyz, features, fps_start=None):
        l1_xyz, l1_points = self.sa1(xyz, features, fps_start=fps_start)
        _, l2_points = self.sa2(l1_xyz, l1_points)
        return l2_points.squeeze(-1)


class GraphedForward:
    """One forward of ``fn(*inputs)`` captured into a CUDA graph and replayed: a step of the encoder is ~35 short
    launches, and replaying them as a graph removes the host launch cost and most of the gaps between dependent
    kernels.  Inputs are copied into static buffers, the outputs are the static result tensors (valid until the
    next call).  Shapes are fixed at capture; everything ``fn`` does must be stream-ordered (the C ABI is)."""

    def __init__(self, fn, *example_inputs, warmup: int = 3):
        self.fn = fn
        for i, x in enumerate(example_inputs):
            if not (torch.is_tensor(x) and x.is_cuda):
                # a host tensor (e.g. FPS start indices) would be uploaded inside the capture from a temporary pinned
                # block and every replay would re-read that recycled memory: all inputs are static device buffers
                raise RuntimeError("GraphedForward: input %d is not a CUDA tensor; every input (the windows AND the FPS "
                                   "start indices) must be device resident - they are refilled by copy before each replay" % i)
        self.static_in = [x.clone() for x in example_inputs]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):               # folds / packs the weights and warms the allocator outside the capture
                fn(*self.static_in)
        torch.cuda.current_stream().wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.static_out = fn(*self.static_in)

    def __call__(self, *inputs):
        for dst, src in zip(self.static_in, inputs):
            if dst.data_ptr() != src.data_ptr():
                dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.static_out


def load_numpy_state(module: nn.Module, state: dict):
    """strict load of a {name: numpy array} state dict (ev2hands_b200.synth.random_state_for)."""
    module.load_state_dict({k: torch.from_numpy(v.copy()) if hasattr(v, "dtype") else v for k, v in state.items()},
                           strict=True)
    return module
