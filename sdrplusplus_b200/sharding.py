"""Multi-GPU sharding of the path: there is no reduction or exchange step.

* independent IQ streams  -> replicas: one full front end per GPU, aggregate = sum (bench.py --gpus N default)
* one stream, many VFOs   -> VFO groups per GPU; rank 0 ingests the IQ and broadcasts each raw chunk (8 B/sample,
                             <1 % of an NVLink 5 port at 1 GS/s); rank 0 also keeps the FFT branch
"""


def partition_vfos(n_vfo, world, rank):
    """Round-robin VFO ids owned by `rank` (balanced to within one VFO)."""
    return [i for i in range(n_vfo) if i % world == rank]


def aggregate_throughput(samples_per_rank, seconds_per_rank):
    """Whole-job throughput: all samples processed / the slowest rank's time."""
    return float(sum(samples_per_rank)) / float(max(seconds_per_rank))
