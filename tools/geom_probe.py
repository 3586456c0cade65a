"""Small launches of the geometry / head kernels added in round 2, for compute-sanitizer (racecheck, memcheck, synccheck):
pruned and cluster FPS on a long window, sampling in ranges + ball query over centre ranges, conv1d over rows, class
attention.  Each result is compared with the exhaustive kernel / a torch expression so a wrong answer also fails."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ev2hands_b200 import _capi, synth           # noqa: E402

dev = torch.device("cuda:0")
torch.backends.cudnn.allow_tf32 = False          # the torch expressions below are the fp32 reference
torch.backends.cuda.matmul.allow_tf32 = False
ev = torch.from_numpy(synth.make_windows(2, 4500, seed=3)).to(dev)
x = ev[:, :3, :]
start = torch.from_numpy(synth.make_start_indices(2, 4500, seed=1)).to(dev)
ref = _capi.fps(x, _capi.cf_strides(x), start, 2, 4500, 48, variant=1)[0]
for v in (2, 3):
    got = _capi.fps(x, _capi.cf_strides(x), start, 2, 4500, 48, variant=v)[0]
    print("fps variant", v, "identical to exhaustive:", bool(torch.equal(ref, got)))
    assert torch.equal(ref, got)

# conv1d over rows (3 taps) and class attention against torch
torch.manual_seed(0)
B, N, D, C = 2, 96, 64, 4
feat = torch.randn(B, D, N, device=dev)
conv = torch.nn.Conv1d(D, D, 3, 1, 1).to(dev)
from ev2hands_b200 import tehnet as th, pointnet2_utils as pu      # noqa: E402
mode = pu._layer_mode(_capi.TC_TF32X3)
L = th._pack_conv1d(conv.weight.detach(), conv.bias.detach(), mode)
L["mode"] = mode
rows = pu._to_rows(feat).view(B * N, D)
y = th._conv_rows(rows, B * N, D, L, N, relu=False)
want = conv(feat).permute(0, 2, 1).reshape(B * N, D)
err = float((y[:, :D] - want).abs().max() / want.abs().max())
print("conv1d rows rel err", err)
assert err < 1e-4
key = torch.randn(B * N, C, device=dev)
ctx = _capi.class_attention(key, C, rows, D, rows, D, B, N, C, D, D ** -0.5)
sim = torch.bmm(key.view(B, N, C).permute(0, 2, 1), rows.view(B, N, D)) * D ** -0.5
want = torch.bmm(torch.softmax(sim, 1), rows.view(B, N, D).permute(0, 2, 1))
err = float((ctx - want).abs().max() / want.abs().max())
print("class attention rel err", err)
assert err < 1e-4
torch.cuda.synchronize()
print("ok")
