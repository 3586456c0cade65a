"""Small end-to-end run for compute-sanitizer: one encoder + decoder forward on 2 windows in every precision."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ev2hands_b200 as e2h
from ev2hands_b200 import synth
from ev2hands_b200.encoder import load_numpy_state
dev = "cuda:0"
enc = e2h.SetAbstractionEncoder()
for i, n in enumerate(("sa1", "sa2", "sa3")):
    load_numpy_state(getattr(enc, n), synth.random_state_for(synth.ENCODER_SPECS[n], seed=100 + i))
dec = e2h.FeaturePropagationDecoder()
for i, n in enumerate(("fp3", "fp2", "fp1")):
    load_numpy_state(getattr(dec, n), synth.random_state_for(synth.DECODER_SPECS[n], seed=300 + i))
enc, dec = enc.to(dev).eval(), dec.to(dev).eval()
ev = torch.from_numpy(synth.make_windows(2, 2048, seed=3)).to(dev)
for prec in ("tf32x3", "bf16", "fp32"):
    e2h.set_mlp_precision(prec)
    with torch.no_grad():
        l3, lv = enc(ev, return_levels=True)
        d0 = dec(ev[:, :3, :], lv["l1_xyz"], lv["l2_xyz"], torch.zeros(2, 3, 1, device=dev), lv["l1_points"], lv["l2_points"], l3.unsqueeze(-1))
    torch.cuda.synchronize()
    print(prec, float(l3.abs().sum()), float(d0.abs().sum()), flush=True)

# event-window construction (both modes, ragged counts, a hot pixel)
import numpy as np
for mode, cols in (("stream", 0), ("erpc", 2)):
    raw = synth.make_raw_events(5000, seed=5, t0=1.0e6, duration=5.0e6, extra_columns=cols)
    raw[100:400, :2] = (40, 50)
    wb = e2h.EventWindowBuilder(mode, n_events=2048)
    np.random.seed(1)
    w = wb(torch.from_numpy(raw).to(dev), [0, 1000, 2900, 4999], [2048, 1500, 2100 if mode == "stream" else 2048, 1], check=True)
    torch.cuda.synchronize()
    print(mode, float(w[:3].abs().sum()), int(wb.last_n_pixels.sum()), flush=True)
