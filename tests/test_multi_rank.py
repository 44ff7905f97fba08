"""N > 1 host logic on CPU (gloo, world_size 2): how bench.py / a deployment shards the path over ranks.

The data path has no exchange step: independent IQ streams are replicas (config 5), and one
stream with many VFOs is sharded by VFO group after a broadcast of the raw chunk (config 4).  What is checked here
is the host-side plumbing -- partitioning, the broadcast of a chunk, the max-over-ranks timing reduction and the
aggregate -- with the oracle standing in as the per-rank checker (no GPU in this container)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from sdrplusplus_b200.sharding import partition_vfos, aggregate_throughput  # noqa: E402


def test_partition_vfos_round_robin_is_a_partition():
    for n_vfo in (1, 7, 8, 16, 64):
        for world in (1, 2, 4, 8):
            parts = [partition_vfos(n_vfo, world, r) for r in range(world)]
            flat = sorted(i for p in parts for i in p)
            assert flat == list(range(n_vfo))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def test_aggregate_throughput_uses_max_time():
    assert aggregate_throughput([1e6, 1e6], [1.0, 2.0]) == pytest.approx(2e6 / 2.0)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle.oracle import Oracle
    from util import noise_iq, fm_carrier
    fs, n, chunk = 2.4e6, 48000, 12000
    offs = [300e3, -650e3, 100e3, 900e3]
    # config-4 style: rank 0 owns the stream, every chunk is broadcast, each rank demodulates its VFO group
    if rank == 0:
        x = noise_iq(n, 5, 0.02).copy()
        for o in offs:
            x += fm_carrier(n, fs, o)
        buf = torch.from_numpy(x.view(np.float32).copy())
    else:
        buf = torch.empty(2 * n, dtype=torch.float32)
    mine = partition_vfos(len(offs), world, rank)
    o = Oracle("restatement")
    chains = {i: (o.rxvfo(fs, 250e3, 150e3, offs[i]), o.wfm(75e3, 250e3)) for i in mine}
    sums = {}
    for c in range(0, n, chunk):
        seg = buf[2 * c: 2 * (c + chunk)].clone()
        dist.broadcast(seg, src=0)
        for i, (v, d) in chains.items():
            y = d.process(v.process(seg.numpy()))
            sums[i] = sums.get(i, 0.0) + float(np.sum(np.abs(y.astype(np.float64))))
    # every rank reports (samples, seconds); the job figure is total samples / max seconds
    t = torch.tensor([float(n), 1.0 + rank], dtype=torch.float64)
    tmax = t.clone()
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    gathered = [None] * world
    dist.all_gather_object(gathered, sums)
    if rank == 0:
        q.put((gathered, float(tmax[1])))
    dist.destroy_process_group()


def test_broadcast_and_vfo_group_sharding_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    gathered, tmax = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert tmax == 2.0
    merged = {}
    for g in gathered:
        merged.update(g)
    assert sorted(merged) == [0, 1, 2, 3]
    # the sharded result equals the single-process result
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle.oracle import Oracle
    from util import noise_iq, fm_carrier
    fs, n, chunk = 2.4e6, 48000, 12000
    offs = [300e3, -650e3, 100e3, 900e3]
    x = noise_iq(n, 5, 0.02).copy()
    for o_ in offs:
        x += fm_carrier(n, fs, o_)
    o = Oracle("restatement")
    for i, off in enumerate(offs):
        v, d = o.rxvfo(fs, 250e3, 150e3, off), o.wfm(75e3, 250e3)
        s = 0.0
        for c in range(0, n, chunk):
            s += float(np.sum(np.abs(d.process(v.process(x.view(np.float32)[2 * c: 2 * (c + chunk)])).astype(np.float64))))
        assert merged[i] == pytest.approx(s, rel=1e-12)
