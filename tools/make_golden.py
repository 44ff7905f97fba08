#!/usr/bin/env python
"""Generate tests/golden/*.npz from the REFERENCE's own dsp headers (oracle/_ref/libsdrpp_ref.so, built from
/root/reference by oracle/Makefile).  Runs only in the authoring container; the fixtures are committed so that
the oracle restatement (and through it the CUDA path) stays pinned to the reference's control flow on machines
where /root/reference does not exist.  Inputs are regenerated from the seeds stored in each file, so the fixtures
hold outputs only (kept small: heads/tails + order-sensitive checksums for the long streams)."""
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle.oracle import Oracle  # noqa: E402
from golden_cases import run_cases  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def main():
    os.makedirs(OUT, exist_ok=True)
    R = Oracle("reference")
    assert R.impl() == "reference-headers"
    g = run_cases(R)
    np.savez_compressed(os.path.join(OUT, "reference_path.npz"), **g)
    sz = os.path.getsize(os.path.join(OUT, "reference_path.npz"))
    print("wrote tests/golden/reference_path.npz: %d arrays, %d bytes" % (len(g), sz))


if __name__ == "__main__":
    main()
