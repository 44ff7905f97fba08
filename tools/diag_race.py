"""Small front-end run (config-1 geometry) for compute-sanitizer: options as key=value, prints the error against the oracle."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import sdrplusplus_b200 as sb
from sdrplusplus_b200 import lib as L
from oracle.oracle import Oracle
from util import rel_rms, noise_iq, fm_carrier
assert L.load().b200_init(0) == 0
fs, n, chunk = 2.4e6, 60000, 12000
x = (noise_iq(n, 7, 0.01) + fm_carrier(n, fs, 300e3) + fm_carrier(n, fs, -300e3)).astype(np.complex64)
o = Oracle("restatement")
fe = sb.FrontEnd(fs, chunk)
for a in sys.argv[1:]:
    k, v = a.split("=")
    fe.set_option(k, int(v))
fe.set_fft(65536, 20.0, 2)
ids = [fe.add_vfo(sb.VfoConfig.wfm(300e3)), fe.add_vfo(sb.VfoConfig.wfm(-300e3))]
outs, lines = fe.process_chunks(x, chunk)
fe.close()
xf = x.view(np.float32)
for vid, off in zip(ids, (300e3, -300e3)):
    v = o.rxvfo(fs, 250e3, 150e3, off); d = o.wfm(75e3, 250e3, False, True)
    ref = np.concatenate([d.process(v.process(xf[2 * i: 2 * (i + chunk)])) for i in range(0, n, chunk)]).reshape(-1, 2)
    y = outs[vid]
    bad = np.nonzero(np.abs(y[:, 0] - ref[:, 0]) > 1e-3)[0]
    print(sys.argv[1:], "vfo", vid, "rel rms %.3g" % rel_rms(y[1000:], ref[1000:]), "bad samples %d of %d" % (bad.size, y.shape[0]),
          ("first %d last %d, per 1250-sample chunk: %s" % (bad[0], bad[-1], np.bincount(bad // 1250, minlength=y.shape[0] // 1250).tolist())) if bad.size else "",
          "zeros %d nan %d" % (int(np.sum(y[:, 0] == 0)), int(np.sum(~np.isfinite(y[:, 0])))), flush=True)
