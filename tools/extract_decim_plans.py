#!/usr/bin/env python
"""Dump the reference's power-of-two decimation plans (stage decimations + FIR coefficient tables,
core/src/dsp/multirate/decim/plans.h:24-140 and decim/taps/*.h) into one small binary file.

The coefficients are numeric data the hot path needs in order to produce the reference's results; they are
read here through oracle/_ref/libsdrpp_ref.so (the reference's own headers compiled in this container) and
written as a flat table -- no reference source is copied.  A drop-in integration does not need this file:
the host adapter hands the reference's own `decim::plans` to b200_register_decim_plan().

Layout (little endian): 8-byte magic "SDRPPDP1", int32 nplans, then per plan:
int32 ratio, int32 nstages, per stage: int32 decimation, int32 tapcount, float32[tapcount].
"""
import os
import struct
import sys

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from oracle.oracle import Oracle  # noqa: E402


def main():
    o = Oracle("reference")
    out = os.path.join(os.path.dirname(__file__), "..", "sdrplusplus_b200", "data", "decim_plans.bin")
    blob = bytearray(b"SDRPPDP1")
    ratios = [1 << k for k in range(1, 14)]
    blob += struct.pack("<i", len(ratios))
    for r in ratios:
        plan = o.decim_plan(r)
        blob += struct.pack("<ii", r, len(plan))
        for s, (d, t) in enumerate(plan):
            taps = o.decim_taps(r, s)
            assert taps.size == t
            blob += struct.pack("<ii", d, t) + taps.astype("<f4").tobytes()
    with open(out, "wb") as f:
        f.write(blob)
    print(f"wrote {out}: {len(blob)} bytes")


if __name__ == "__main__":
    main()
