"""Where the training step (bench.py cfg4: whole network, batch 32 on one GPU) spends its device time: torch.profiler
kernel table of two steps, fp32 convolutions (TF32 off) and - for comparison - PyTorch's default (cuDNN TF32 on).  GPU only."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("ERPC", "1")
from ev2hands_b200 import tehnet, trainer           # noqa: E402

dev = torch.device("cuda:0")
B = int(os.environ.get("B", "32"))


def run(tf32):
    torch.backends.cudnn.allow_tf32 = tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0)
    net = tehnet.TEHNet(n_pose_params=6).to(dev).train()
    hands = tehnet.create_standin_mano_layers(dev, n_cmps=6)
    opt = torch.optim.Adam(net.parameters(), lr=1e-3)
    red = trainer.BucketedGradReducer(net.parameters(), n_buckets=4)
    full = tehnet.make_training_batch(B, 2048, seed=0)

    def cut(x):
        return {k: cut(v) for k, v in x.items()} if isinstance(x, dict) else x.to(dev)
    batch = cut(full)

    def step():
        b = {k: (dict(v) if isinstance(v, dict) else v) for k, v in batch.items()}
        return trainer.train_step(net, hands, b, opt, red, tehnet.training_losses)
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3):
        step()
    z.record()
    torch.cuda.synchronize()
    print("cudnn TF32 %s: %.1f ms per step" % (tf32, a.elapsed_time(z) / 3))
    if not tf32:
        from torch.profiler import profile, ProfilerActivity
        with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
            step()
            torch.cuda.synchronize()
        print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=35, max_name_column_width=70))
    red.remove()


run(False)
run(True)
