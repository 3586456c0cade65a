"""Where the whole-network eval step (bench.py `full_network_eval`) spends its time: event brackets around the
sections of TEHNet.forward (eager, one stream), device time against the wall time of the loop.  GPU only; run through gpurun."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("ERPC", "1")
from ev2hands_b200 import synth, tehnet as th           # noqa: E402
from ev2hands_b200 import pointnet2_utils as pu         # noqa: E402

B = int(os.environ.get("B", "64"))
dev = torch.device("cuda:0")
torch.manual_seed(0)
net = th.TEHNet(n_pose_params=6).to(dev).eval()
hands = th.create_standin_mano_layers(dev)
ev = torch.from_numpy(synth.make_windows(B, 2048, seed=5)).to(dev)


def ms(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n, 1e3 * (time.perf_counter() - t0) / n


with torch.no_grad():
    print("whole forward, eager: %.3f ms device, %.3f ms wall" % ms(lambda: net(ev, hands)))
    print("trunk (encoder + decoder + heads), eager: %.3f ms device, %.3f ms wall" % ms(lambda: net.trunk(ev)))
    l0_xyz, seg, left, right = net.trunk(ev)
    print("one hand regressor: %.3f ms device, %.3f ms wall" % ms(lambda: net.left_mano_regressor(l0_xyz, left, hands["left"])))
    reg = net.left_mano_regressor
    print("  its sa1: %.3f ms device, %.3f ms wall" % ms(lambda: reg.sa1(l0_xyz, left)))
    a1 = reg.sa1(l0_xyz, left)
    print("  its sa2 (group_all): %.3f ms device, %.3f ms wall" % ms(lambda: reg.sa2(*a1)))
    a2 = reg.sa2(*a1)[1]
    print("  its FC head: %.3f ms device, %.3f ms wall" % ms(lambda: reg.mano_regressor(a2.squeeze(-1))))
    p = reg.mano_regressor(a2.squeeze(-1))
    print("  stand-in MANO layer: %.3f ms device, %.3f ms wall" % ms(
        lambda: hands["left"](global_orient=p[:, :3], hand_pose=p[:, 3:9], betas=p[:, 9:-3], transl=p[:, -3:])))
    # trunk sections
    xyz = ev[:, :3, :]
    print("encoder sa1: %.3f / %.3f" % ms(lambda: net.sa1(xyz, ev)))
    l1 = net.sa1(xyz, ev)
    print("encoder sa2: %.3f / %.3f" % ms(lambda: net.sa2(*l1)))
    l2 = net.sa2(*l1)
    print("encoder sa3: %.3f / %.3f" % ms(lambda: net.sa3(*l2)))
    l3 = net.sa3(*l2)
    print("fp3: %.3f / %.3f" % ms(lambda: net.fp3(l2[0], l3[0], l2[1], l3[1])))
    p2 = net.fp3(l2[0], l3[0], l2[1], l3[1])
    print("fp2: %.3f / %.3f" % ms(lambda: net.fp2(l1[0], l2[0], l1[1], p2)))
    p1 = net.fp2(l1[0], l2[0], l1[1], p2)
    print("fp1: %.3f / %.3f" % ms(lambda: net.fp1(xyz, l1[0], None, p1)))
    p0 = net.fp1(xyz, l1[0], None, p1)
    print("heads (classifier, query convolutions, attention): %.3f / %.3f" % ms(lambda: net._heads_cuda(p0)))
