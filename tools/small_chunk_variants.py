#!/usr/bin/env python
"""Small chunks (the reference's own sizes): the chain behind stage 1 as register kernels (default) against the one-launch
fused tail (ft_regall = 0), with and without graph replay.  Device-resident cf32 and pinned int16.
    python tools/small_chunk_variants.py > gpurun_out/small_chunk_variants.json"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from small_chunk_probe import run                               # noqa: E402
from sdrplusplus_b200 import lib                                # noqa: E402


def main():
    from bench import bind_to_gpu_numa
    bind_to_gpu_numa(0)
    L = lib.load()
    assert L.b200_init(0) == 0
    res = []
    variants = (("register chain + graph (default)", {}),
                ("fused tail, one launch", {"ft_regall": 0, "ft_prereg": 0}),
                ("fused tail, one launch, no graph", {"ft_regall": 0, "ft_prereg": 0, "graph": 0}),
                ("fused tail behind 1 register stage", {"ft_regall": 0, "ft_prereg": 1}),
                ("fused tail behind 2 register stages", {"ft_regall": 0, "ft_prereg": 2}))
    for csz, n in ((500000, 600), (1000000, 400), (1 << 22, 150)):
        for name, opts in variants:
            for host in (False, True):
                if host and csz > 1000000:
                    continue
                try:
                    r = run(csz, n, dict(opts), True, host, 4 if host else 2)
                    r["variant"] = name
                except Exception as ex:      # noqa: BLE001
                    r = {"chunk": csz, "variant": name, "error": repr(ex)}
                res.append(r)
                print(json.dumps(r), file=sys.stderr)
    print(json.dumps([{k: r.get(k) for k in ("variant", "chunk", "mem", "GS_per_s", "us_per_chunk", "launches_per_chunk", "host_submit_us", "device_groups", "error")} for r in res], indent=1))


if __name__ == "__main__":
    main()
