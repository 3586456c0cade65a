"""Timeline of one CTA of sa_fused_tc_kernel (EXP_TRACE build: EV2H_LIB=exp/libev2h_TRACE.so).
Prints clock-ordered events of CTA 0 for three steady-state tiles of every fused launch of one encoder step."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import ev2hands_b200 as e2h
from ev2hands_b200 import _capi, synth
import bench
prec = sys.argv[1] if len(sys.argv) > 1 else "tf32x3"
e2h.set_mlp_precision(prec)
dev = torch.device("cuda:0")
enc = bench.build_encoder(dev)
B = 64
ev = torch.from_numpy(synth.make_windows(B, 2048, seed=1236)).to(dev)
s1 = torch.from_numpy(synth.make_start_indices(B, 2048, 0)).to(dev); s2 = torch.from_numpy(synth.make_start_indices(B, 512, 1)).to(dev)
with torch.no_grad():
    for _ in range(2): enc(ev, fps_starts=(s1, s2))
torch.cuda.synchronize()
orig = _capi.sa_msg_fused
buf = torch.zeros(512, 16, dtype=torch.int64, device=dev)
ROLE = {1: "load", 2: "strm", 3: "issu", 4: "epil"}
EV = {(1, 1): "chunk computed", (1, 2): "slot granted", (1, 3): "arrived a_full", (1, 5): "acc_full1 seen (pool)", (1, 6): "pool done",
      (2, 1): "b_empty ok -> copy", (3, 1): "acc_empty ok", (3, 2): "a_full ok", (3, 3): "b_full ok", (3, 6): "mmas issued", (3, 7): "commits issued",
      (4, 1): "acc_full0 seen", (4, 2): "chunk converted", (4, 3): "slot granted", (4, 4): "arrived a_full",
      (4, 5): "acc_full1 seen", (4, 6): "pool done"}
def wrapped(*a, **k):
    buf.zero_(); _capi.lib().ev2h_fused_set_debug_buffer(buf.data_ptr())
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); orig(*a, **k); e1.record(); torch.cuda.synchronize()
    _capi.lib().ev2h_fused_set_debug_buffer(None)
    d = buf.cpu().numpy().reshape(-1)
    tr = np.concatenate([d[4800 + r * 800 + 1: 4800 + r * 800 + 1 + int(d[4800 + r * 800])] for r in range(4)]); n = len(tr)
    evs = sorted(((int(x) & 0xFFFFFFFFFF, int(x) >> 40) for x in tr))
    print("==== %s K=%d widths=%s: %.3f ms, %d events" % (prec, a[6], a[18], e0.elapsed_time(e1), n))
    if not evs: return
    t0 = evs[0][0]
    for t, tag in evs:
        role, e, it, c = tag // 1000000, (tag // 10000) % 100, (tag // 100) % 100, tag % 100
        g = ""
        if role in (2, 3): g = "L%d c%d" % (2 + c // 50, c % 50)
        else: g = "c%d" % c
        print("%8d  %s  it%d %-6s %s" % (t - t0, ROLE.get(role, "?"), it, g, EV.get((role, e), "?")))
_capi.sa_msg_fused = wrapped
with torch.no_grad():
    enc(ev, fps_starts=(s1, s2))
