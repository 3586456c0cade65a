"""Host logic of the multi-GPU path on CPU: world_size-2 gloo processes shard a batch and its FPS
start indices, run the per-window oracle on their shard, and the union must equal the unsharded run;
timings agree through a max all-reduce; gradients average through one flattened all-reduce."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ev2hands_b200 import sharding, synth


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 64, 1024, 1025):
        for world in (1, 2, 3, 8):
            spans = [sharding.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_bounds(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import c_oracle
    torch.set_num_threads(1)
    B, N, S = 6, 256, 32
    ev = synth.make_windows(B, N, seed=9)
    start = synth.make_start_indices(B, N, seed=4)
    xyz = np.ascontiguousarray(ev[:, :3].transpose(0, 2, 1))
    (xs, ss) = sharding.shard((torch.from_numpy(xyz), torch.from_numpy(start)), rank, world)
    idx = c_oracle.fps(xs.numpy(), S, ss.numpy())
    np.save(os.path.join(out_dir, "fps_%d.npy" % rank), idx)
    # timing agreement: slowest rank wins
    got = sharding.max_over_ranks([1.0 + rank, 5.0 - rank])
    assert got == [float(world), 5.0]
    # gradient averaging through one flattened all-reduce
    lin = torch.nn.Linear(3, 2)
    with torch.no_grad():
        lin.weight.fill_(0.5)
        lin.bias.zero_()
    lin(torch.full((4, 3), float(rank + 1))).sum().backward()
    n = sharding.allreduce_mean_grads(lin.parameters())
    assert n == 8
    want_w = np.mean([4.0 * (r + 1) for r in range(world)])
    assert torch.allclose(lin.weight.grad, torch.full((2, 3), want_w))
    assert torch.allclose(lin.bias.grad, torch.full((2,), 4.0))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_matches_unsharded(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    from oracle import c_oracle
    B, N, S = 6, 256, 32
    ev = synth.make_windows(B, N, seed=9)
    start = synth.make_start_indices(B, N, seed=4)
    xyz = np.ascontiguousarray(ev[:, :3].transpose(0, 2, 1))
    whole = c_oracle.fps(xyz, S, start)
    parts = np.concatenate([np.load(os.path.join(str(tmp_path), "fps_%d.npy" % r)) for r in range(world)])
    assert np.array_equal(whole, parts)


def test_window_shards_reproduce_the_unsharded_batch():
    """Event windows shard like the encoder batch: every rank holds only its slice of the raw event
    table, the draws are made for the global batch and sliced with the same bounds; checked with the CPU oracle."""
    from oracle import window_oracle as wo
    ev = synth.make_raw_events(4000, seed=21)
    starts = np.array([0, 300, 900, 1500, 1501, 2500, 1900])        # overlapping, not sorted by start
    counts = np.array([800, 700, 600, 900, 10, 1500, 64])
    rs = np.random.RandomState(3)
    recs = [wo.aggregate(ev[s:s + c], "stream") for s, c in zip(starts, counts)]
    idx = np.stack([rs.randint(0, r.shape[0], size=128) for r in recs])
    whole = wo.build_windows(ev, starts, counts, idx, "stream")
    for world in (1, 2, 3, 8):
        parts = []
        for rank in range(world):
            row_lo, row_hi, ls, lc, (lo, hi) = sharding.shard_windows(starts, counts, rank, world)
            if hi == lo:
                assert row_hi == row_lo == 0 and len(ls) == 0
                continue
            local = ev[row_lo:row_hi]                                  # what this rank uploads
            assert (ls >= 0).all() and (ls + lc <= local.shape[0]).all() and lc.dtype == np.int32
            parts.append(wo.build_windows(local, ls, lc, idx[lo:hi], "stream"))
        assert np.array_equal(np.concatenate(parts), whole, equal_nan=True)
    with pytest.raises(ValueError):
        sharding.shard_windows([0, 1], [5], 0, 1)


# ------------------------------------------------------------------ training step: bucketed gradient exchange ----
def _mlp():
    torch.manual_seed(0)          # same initial weights on every rank, like the replicas of nn.DataParallel
    return torch.nn.Sequential(torch.nn.Linear(6, 16), torch.nn.ReLU(), torch.nn.Linear(16, 16), torch.nn.ReLU(),
                               torch.nn.Linear(16, 3), torch.nn.ReLU(), torch.nn.Linear(3, 2, bias=False))


def _train_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ev2hands_b200.trainer import BucketedGradReducer
    torch.set_num_threads(1)
    net = _mlp()
    net[6].weight.requires_grad_(True)
    red = BucketedGradReducer(net.parameters(), n_buckets=3)
    assert len(red.buckets) >= 2 and red.buckets[0][0] == 0 and red.buckets[-1][1] == red.flat.numel()
    opt = torch.optim.Adam(net.parameters(), lr=1e-2)
    x_all = torch.from_numpy(np.random.RandomState(5).randn(8, 6).astype(np.float32))
    (x,) = sharding.shard((x_all,), rank, world)
    for step in range(3):
        red.zero_grad()
        out = net(x)
        # step 1 leaves the last layer out of the loss: its bucket gets no hook and must be flushed by finish()
        loss = out.pow(2).mean() if step != 1 else net[:6](x).pow(2).mean()
        loss.backward()
        for p in net.parameters():                        # gradients live in the flat buffer
            assert p.grad.data_ptr() >= red.flat.data_ptr() and p.grad.data_ptr() < red.flat.data_ptr() + red.bytes
        red.finish()
        opt.step()
    torch.save([p.detach().clone() for p in net.parameters()], os.path.join(out_dir, "params_%d.pt" % rank))
    dist.barrier()
    dist.destroy_process_group()


def test_bucketed_gradient_exchange_equals_full_batch_training(tmp_path):
    """world_size-2 gloo: per-rank shards + bucketed all-reduce (launched from backward hooks, AVG) must reproduce
    single-process training on the whole batch (equal shard sizes: mean of shard means = batch mean), and keep
    the replicas identical."""
    world = 2
    mp.spawn(_train_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    got = [torch.load(os.path.join(str(tmp_path), "params_%d.pt" % r)) for r in range(world)]
    for a, b in zip(*got):
        assert torch.equal(a, b)
    net = _mlp()
    opt = torch.optim.Adam(net.parameters(), lr=1e-2)
    x_all = torch.from_numpy(np.random.RandomState(5).randn(8, 6).astype(np.float32))
    for step in range(3):
        opt.zero_grad(set_to_none=False)
        for p in net.parameters():
            if p.grad is None:
                p.grad = torch.zeros_like(p)
        if step != 1:
            loss = 0.5 * (net(x_all[:4]).pow(2).mean() + net(x_all[4:]).pow(2).mean())
        else:
            loss = 0.5 * (net[:6](x_all[:4]).pow(2).mean() + net[:6](x_all[4:]).pow(2).mean())
        loss.backward()
        opt.step()
    for a, b in zip(got[0], net.parameters()):
        assert torch.allclose(a, b.detach(), rtol=1e-5, atol=1e-7)
