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


# ---------------------------------------------------------------------------------------------- on the GPU
def _gpu_worker(rank, world, port, outdir):
    """config-4 sharding with the real kernels: rank 0 owns the stream, every rank runs the CUDA front end for its VFO
    group.  With one GPU per rank this is the product's sharded front end (b200_shard_*: the library broadcasts each raw
    chunk over NCCL on its communication stream and processes it from device memory); when the ranks have to share
    cuda:0 (the single-GPU test box) NCCL cannot run, the chunk travels over gloo and the plain front end processes it."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    nccl = torch.cuda.device_count() >= world
    dev = rank if nccl else 0
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl" if nccl else "gloo", rank=rank, world_size=world)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import sdrplusplus_b200 as sb
    from sdrplusplus_b200 import lib as L
    from test_multi_rank import _sharding_case
    assert L.load().b200_init(dev) == 0
    fs, n, chunk, cfgs, x = _sharding_case(sb, L)
    buf = torch.from_numpy(x.view(np.float32).copy()) if rank == 0 else torch.empty(2 * n, dtype=torch.float32)
    if nccl:
        buf = buf.cuda()
    mine = partition_vfos(len(cfgs), world, rank)
    fe = sb.FrontEnd(fs, chunk)
    if rank == 0:
        fe.set_fft(65536, 20.0, 2)              # rank 0 keeps the spectrum branch
    ids = {i: fe.add_vfo(cfgs[i]) for i in mine}
    outs = {i: [] for i in mine}
    lines = []
    sh = None
    if nccl:
        # the product's own sharded path (b200_shard_*): rank 0 hands the chunk in, the library broadcasts it over NCCL
        from sdrplusplus_b200.sharding import ShardedFrontEnd, make_unique_id
        uid = [make_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        sh = ShardedFrontEnd(fe, rank, world, uid[0])
    for c in range(0, n, chunk):
        if nccl:
            # device in, device out: the path a multi-GPU deployment runs
            o = L.Outputs()
            keep = {}
            for i, vid in ids.items():
                cap = fe.vfo_max_out(vid, chunk)
                keep[i] = torch.empty(2 * cap, device="cuda", dtype=torch.float32)
                o.vfo_out[vid] = keep[i].data_ptr(); o.vfo_cap[vid] = cap
            nl = max(1, fe.fft_max_lines(chunk))
            lt = torch.empty(nl * 65536, device="cuda", dtype=torch.float32)
            o.fft_out = lt.data_ptr(); o.fft_cap_lines = nl; o.out_mem = L.MEM_DEVICE
            seg = buf[2 * c: 2 * (c + chunk)] if rank == 0 else None
            torch.cuda.synchronize()
            sh.submit_ptr(seg.data_ptr() if rank == 0 else 0, chunk, L.FMT_CF32, L.MEM_DEVICE, o)
            sh.wait()
            for i, vid in ids.items():
                y = keep[i][: 2 * o.vfo_count[vid]].cpu().numpy()
                outs[i].append(y.view(np.complex64) if cfgs[i].demod == L.DEMOD_RAW else y.reshape(-1, 2))
            if o.fft_lines:
                lines.append(lt[: o.fft_lines * 65536].cpu().numpy().reshape(-1, 65536))
        else:
            seg = buf[2 * c: 2 * (c + chunk)].clone()
            dist.broadcast(seg, src=0)
            o, ln = fe.process(seg.numpy().view(np.complex64))
            for i, vid in ids.items():
                outs[i].append(o[vid])
            if ln.size:
                lines.append(ln)
    if sh is not None:
        assert sh.bytes_broadcast() == 8 * n
        sh.close()
    fe.close()
    np.savez(os.path.join(outdir, "rank%d.npz" % rank), lines=np.concatenate(lines) if lines else np.empty((0, 0), np.float32),
             transport=np.array([1 if nccl else 0]), **{"vfo%d" % i: np.concatenate(v) for i, v in outs.items()})
    dist.barrier()
    dist.destroy_process_group()


def _sharding_case(sb, L):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from util import noise_iq, fm_carrier, am_carrier, ssb_tone
    fs, n, chunk = 2.4e6, 120000, 12000
    x = noise_iq(n, 31, 0.002).copy()
    x += am_carrier(n, fs, -400e3) + fm_carrier(n, fs, 200e3, dev=5000.0, tones=((1000.0, 0.7),)) + ssb_tone(n, fs, 600e3, 1000.0)
    x += fm_carrier(n, fs, 900e3) + fm_carrier(n, fs, -300e3) + fm_carrier(n, fs, 300e3)
    cfgs = [sb.VfoConfig.am(-400e3), sb.VfoConfig.nfm(200e3), sb.VfoConfig.ssb(600e3, L.DEMOD_USB), sb.VfoConfig.wfm(900e3),
            sb.VfoConfig.wfm(-300e3), sb.VfoConfig.wfm(300e3), sb.VfoConfig.raw(900e3, 250e3, 150e3)]
    return fs, n, chunk, cfgs, x.astype(np.complex64)


def _error_vs_oracle(oracle, L, cfg, y, x, fs, chunk):
    """CUDA output of one VFO against the CPU oracle, same metric as tests/test_gpu_parity.py::test_frontend_mixed_modes:
    RMS-normalised relative error after the settling prefix; SSB against the exact-phase oracle mode (the reference's own
    fp32 phase recurrence walks); the RAW VFO through its phase increments (FM carrier)."""
    from test_gpu_parity import _oracle_chain
    from util import rel_rms
    ya = _oracle_chain(oracle, x, fs, chunk, cfg)
    if cfg.demod == L.DEMOD_RAW:
        ya = ya.view(np.complex64)
        assert y.shape == ya.shape
        d, do = np.angle(y[1:] * np.conj(y[:-1])), np.angle(ya[1:] * np.conj(ya[:-1]))
        return rel_rms(d[2000:], do[2000:])
    if cfg.demod in (L.DEMOD_USB, L.DEMOD_LSB, L.DEMOD_DSB):
        oracle.set_rotator_mode(1)
        try:
            ya = _oracle_chain(oracle, x, fs, chunk, cfg)
        finally:
            oracle.set_rotator_mode(0)
    ya = ya.reshape(-1, 2)
    assert y.shape == ya.shape, (y.shape, ya.shape)
    skip = max(y.shape[0] // 4, 64)               # AGC / DC-block settling, FM discriminator start-up
    return rel_rms(y[skip:], ya[skip:])


@pytest.mark.gpu
def test_vfo_group_sharding_matches_oracle(tmp_path, oracle, report):
    """BASELINE config 4 as a parity case: 7 mixed AM / NFM / USB / WFM / RAW VFOs split over two ranks after a broadcast
    of every raw chunk.  EVERY rank's VFO is gated against the CPU oracle at the north-star tolerance, and so is the
    single-process run of the same 7 VFOs: which stage-1 kernel form a VFO gets (filter bank, conjugate pair, tap-block
    alignment) depends on its group, so both groupings have to hold on their own."""
    import sdrplusplus_b200 as sb
    from sdrplusplus_b200 import lib as L
    assert L.load().b200_init(0) == 0
    fs, n, chunk, cfgs, x = _sharding_case(sb, L)
    fe = sb.FrontEnd(fs, chunk)
    fe.set_fft(65536, 20.0, 2)
    ids = [fe.add_vfo(c) for c in cfgs]
    ref, ref_lines = fe.process_chunks(x, chunk)
    fe.close()
    world = 2
    mp.spawn(_gpu_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    got = {}
    lines = None
    for r in range(world):
        z = np.load(os.path.join(str(tmp_path), "rank%d.npz" % r))
        for k in z.files:
            if k == "lines":
                if r == 0:
                    lines = z[k]
            elif k == "transport":
                pass
            else:
                got[int(k[3:])] = z[k]
    assert sorted(got) == list(range(len(cfgs)))
    errs = {}
    for i, vid in enumerate(ids):
        e_single = _error_vs_oracle(oracle, L, cfgs[i], ref[vid], x, fs, chunk)
        e_shard = _error_vs_oracle(oracle, L, cfgs[i], got[i], x, fs, chunk)
        errs["vfo%d_demod%d" % (i, cfgs[i].demod)] = {"single_process": e_single, "sharded": e_shard}
    report["sharding_vs_oracle"] = errs
    for k, e in errs.items():
        assert e["single_process"] < 1e-5 and e["sharded"] < 1e-5, (k, e)
    assert lines is not None and np.array_equal(lines, ref_lines)
