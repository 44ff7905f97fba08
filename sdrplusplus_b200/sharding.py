"""Multi-GPU sharding of the path: there is no reduction or exchange step.

* independent IQ streams  -> replicas: one full front end per GPU, aggregate = sum (bench.py --gpus N default)
* one stream, many VFOs   -> VFO groups per GPU (BASELINE config 4): rank 0 ingests the IQ and the library broadcasts each
                             raw chunk over NCCL on a communication stream, one chunk ahead of the compute
                             (b200_shard_* in include/b200dsp.h); rank 0 also keeps the FFT branch.  ShardedFrontEnd below
                             is the host-side mirror: one process per GPU, launched e.g. by torchrun.
"""
import ctypes as C

from . import lib as L


def partition_vfos(n_vfo, world, rank):
    """Round-robin VFO ids owned by `rank` (balanced to within one VFO)."""
    return [i for i in range(n_vfo) if i % world == rank]


def aggregate_throughput(samples_per_rank, seconds_per_rank):
    """Whole-job throughput: all samples processed / the slowest rank's time."""
    return float(sum(samples_per_rank)) / float(max(seconds_per_rank))


def make_unique_id():
    """128-byte NCCL id made on the calling rank (rank 0); hand it to the other ranks by any side channel."""
    buf = (C.c_ubyte * 128)()
    L.check(L.load().b200_shard_unique_id(C.cast(buf, C.c_void_p)))
    return bytes(buf)


class ShardedFrontEnd:
    """A FrontEnd (this rank's VFO group) fed by the ingest rank's broadcast.  Every rank calls submit_ptr with the same
    count / format per chunk; `ptr` is read on rank 0 only."""

    def __init__(self, frontend, rank, world, unique_id):
        self._l = L.load()
        self.fe = frontend
        self.rank, self.world = rank, world
        idb = (C.c_ubyte * 128).from_buffer_copy(unique_id)
        self._h = L.check_ptr(self._l.b200_shard_create(frontend._h, rank, world, C.cast(idb, C.c_void_p)))

    def submit_ptr(self, ptr, count, fmt, mem, outputs):
        L.check(self._l.b200_shard_submit(self._h, C.c_void_p(ptr if ptr else 0), count, fmt, mem, C.byref(outputs)))

    def wait(self):
        L.check(self._l.b200_shard_wait(self._h))

    def bytes_broadcast(self):
        return self._l.b200_shard_bytes_broadcast(self._h)

    def close(self):
        if getattr(self, "_h", None):
            self._l.b200_shard_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
