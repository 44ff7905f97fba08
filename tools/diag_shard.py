"""Which configuration of the 7-VFO sharding case disagrees with the oracle for VFO 5 (WFM +300 kHz)?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import sdrplusplus_b200 as sb
from sdrplusplus_b200 import lib as L
from oracle.oracle import Oracle
from test_multi_rank import _sharding_case
from util import rel_rms
assert L.load().b200_init(0) == 0
fs, n, chunk, cfgs, x = _sharding_case(sb, L)
o = Oracle("restatement")
v = o.rxvfo(fs, 250e3, 150e3, 300e3); d = o.wfm(75e3, 250e3, False, True)
xf = x.view(np.float32)
ref = np.concatenate([d.process(v.process(xf[2 * i: 2 * (i + chunk)])) for i in range(0, n, chunk)]).reshape(-1, 2)
def run(sel, opts):
    fe = sb.FrontEnd(fs, chunk)
    for k, val in opts.items():
        fe.set_option(k, val)
    ids = {i: fe.add_vfo(cfgs[i]) for i in sel}
    outs, _ = fe.process_chunks(x, chunk)
    fe.close()
    return outs[ids[5]]
for sel in ([0, 1, 2, 3, 4, 5, 6], [1, 3, 5], [5], [3, 4, 5, 6], [4, 5], [3, 5, 6], [3, 4, 5]):
    for opts in ({}, {"s1": 6}, {"tails": 1}, {"pair": 0}):
        y = run(sel, opts)
        print(sel, opts, "vs oracle %.3g" % rel_rms(y[1000:], ref[1000:]), flush=True)
