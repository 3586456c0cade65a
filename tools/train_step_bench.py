#!/usr/bin/env python
"""Training-step benchmark of the set-abstraction encoder (BASELINE config 4 shape).

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/train_step_bench.py [--batch 32]

Encoder forward + backward in train() mode (BatchNorm batch statistics, per replica like the reference's
nn.DataParallel, train.py:68) on event windows sharded over the ranks, Adam step, and ONE flattened NCCL
all-reduce of the gradients per step.  MANO LBS and the mesh-intersection loss are not importable in this
environment (manopth / mesh_intersection absent, MANO files licence gated), so the loss is a stand-in
(squared error of the 1024 encoder features against random targets); that is stated in the output.
"""
import argparse
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ev2hands_b200 as e2h  # noqa: E402
from ev2hands_b200 import sharding, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32, help="global batch (config 4: 32)")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)                                   # same initial weights on every rank
    enc = e2h.SetAbstractionEncoder().to(dev).train()
    opt = torch.optim.Adam(enc.parameters(), lr=1e-3)      # train.py:23,56
    ev_all = torch.from_numpy(synth.make_windows(args.batch, 2048, seed=1234 + 4))
    tgt_all = torch.randn(args.batch, 1024)
    (ev, tgt) = sharding.shard((ev_all, tgt_all), rank, world)
    ev, tgt = ev.to(dev), tgt.to(dev)
    n_params = sum(p.numel() for p in enc.parameters())

    def step():
        opt.zero_grad(set_to_none=True)
        out = enc(ev)                                      # FPS starts drawn from the CPU generator, like the reference
        loss = ((out - tgt) ** 2).mean()
        loss.backward()
        sharding.allreduce_mean_grads(list(enc.parameters()))
        opt.step()
        return loss

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(args.steps):
        loss = step()
    t1.record()
    torch.cuda.synchronize()
    ms = sharding.max_over_ranks([t0.elapsed_time(t1) / args.steps], device=dev)[0]
    if rank == 0:
        print(json.dumps({"metric": "encoder training steps/s", "value": 1e3 / ms, "ms_per_step": ms, "n_gpus": world,
                          "global_batch": args.batch, "windows_per_s": args.batch * 1e3 / ms, "params": n_params,
                          "grad_allreduce_bytes": 4 * n_params, "loss": float(loss),
                          "note": "encoder sa1-sa3 only; stand-in loss (MANO / mesh-intersection not available); "
                                  "conv/BN/ReLU in PyTorch (batch statistics), CUDA kernels for FPS, ball query, grouping "
                                  "fwd/bwd and max-pool fwd/bwd; one flattened NCCL all-reduce per step"}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
