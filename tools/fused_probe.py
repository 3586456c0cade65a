"""GPU probe for the fused set-abstraction kernel: small cases, progress printed before each launch."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import ev2hands_b200 as e2h
from ev2hands_b200 import synth, _capi
from ev2hands_b200.encoder import load_numpy_state

dev = "cuda:0"
which = sys.argv[1:] or ["a32", "a64", "a128", "b"]
def run(mod, xyz, pts, start, tag):
    outs = {}
    for prec, fused in (("fp32", False), ("tf32x3", True), ("bf16", True)):
        e2h.set_mlp_precision(prec)
        print("launch", tag, prec, flush=True)
        with torch.no_grad():
            _, o = mod(xyz, pts, fps_start=start)
        torch.cuda.synchronize()
        outs[prec] = o
        if prec != "fp32":
            err = ((o - outs["fp32"]).abs().max() / outs["fp32"].abs().max()).item()
            print("  ", tag, prec, "rel err vs fp32 path:", err, flush=True)
B, N = 2, 2048
ev = torch.from_numpy(synth.make_windows(B, N, seed=11)).to(dev)
start = torch.from_numpy(synth.make_start_indices(B, N, 3))
for w in which:
    if w.startswith("a"):
        K = int(w[1:])
        widths = {32: [32, 32, 64], 64: [64, 64, 128], 128: [64, 96, 128]}[K]
        m = e2h.PointNetSetAbstractionMsg(512, [0.3], [K], 5, [widths])
        st = synth.random_sa_state("conv_blocks.{i}.{j}", "bn_blocks.{i}.{j}", [widths], [8], seed=K)
        load_numpy_state(m, st); m = m.to(dev).eval()
        run(m, ev[:, :3], ev, start, w)
    else:
        m = e2h.PointNetSetAbstractionMsg(128, [0.4, 0.8], [64, 128], 320, [[128, 128, 256], [128, 196, 256]])
        st = synth.random_state_for(synth.ENCODER_SPECS["sa2"], seed=5)
        load_numpy_state(m, st); m = m.to(dev).eval()
        xyz = ev[:, :3, :512].contiguous()
        pts = torch.randn(B, 320, 512, device=dev).abs()
        run(m, xyz, pts, torch.from_numpy(synth.make_start_indices(B, 512, 4)), w)
