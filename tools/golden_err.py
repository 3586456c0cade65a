"""Print the feature errors of every precision against the reference's golden vectors."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import ev2hands_b200 as e2h
from ev2hands_b200 import synth
from ev2hands_b200.encoder import load_numpy_state
g = dict(np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "encoder.npz")))
enc = e2h.SetAbstractionEncoder()
for n, s in zip(("sa1", "sa2", "sa3"), g["weight_seeds"]):
    load_numpy_state(getattr(enc, n), synth.random_state_for(synth.ENCODER_SPECS[n], seed=int(s)))
enc = enc.cuda().eval()
ev = torch.from_numpy(g["events"]).cuda()
def rel(a, b):
    a = a.detach().cpu().double().numpy(); b = np.asarray(b, dtype=np.float64)
    return np.abs(a - b).max() / np.abs(b).max()
for prec in ("fp32", "tf32x3", "bf16"):
    e2h.set_mlp_precision(prec)
    with torch.no_grad():
        out, lv = enc(ev, fps_starts=(torch.from_numpy(g["start_sa1"]), torch.from_numpy(g["start_sa2"])), return_levels=True)
    print("%-7s l1 %.2e  l2 %.2e  l3 %.2e" % (prec, rel(lv["l1_points"][0], g["l1_points_w0"]), rel(lv["l2_points"][0], g["l2_points_w0"]), rel(out, g["l3_points"][:, :, 0])))
