"""Where does the UMMA issuer wait?  Runs the bench encoder once per fused launch with the debug counters on."""
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
# patch sa_msg_fused to capture per-launch counters
orig = _capi.sa_msg_fused
buf = torch.zeros(512, 16, dtype=torch.int64, device=dev)
def wrapped(*a, **k):
    buf.zero_(); _capi.lib().ev2h_fused_set_debug_buffer(buf.data_ptr())
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record(); orig(*a, **k); t1.record(); torch.cuda.synchronize()
    _capi.lib().ev2h_fused_set_debug_buffer(None)
    d = buf.cpu().numpy().astype(np.float64)
    act = d[d[:, 5] > 0]
    m = act.mean(0)
    print("%s K=%d widths=%s: %.3f ms | issuer cycles/CTA total=%.0f tiles=%.1f | wait a_full g0=%.1f%% g1=%.1f%% g2=%.1f%% b_full=%.1f%% acc_empty=%.1f%% | commit=%.1f%% issue+other=%.1f%% | cycles/tile=%.0f" % (
        prec, a[6], a[18], t0.elapsed_time(t1), m[5], m[6], 100*m[0]/m[5], 100*m[1]/m[5], 100*m[2]/m[5], 100*m[3]/m[5], 100*m[4]/m[5],
        100*m[7]/m[5], 100*(m[5]-m[:5].sum()-m[7])/m[5], m[5]/max(m[6],1)), flush=True)
    t = max(m[6], 1)
    print("      epilogue warp 0, cycles/tile: wait acc_full[0]=%.0f  convert(incl slot wait)=%.0f  slot wait=%.0f  wait acc_full[1]=%.0f  pool=%.0f  total=%.0f" % (
        m[8]/t, m[9]/t, m[10]/t, m[11]/t, m[12]/t, m[13]/t), flush=True)
_capi.sa_msg_fused = wrapped
import ev2hands_b200.pointnet2_utils as pu
with torch.no_grad():
    enc(ev, fps_starts=(s1, s2))
