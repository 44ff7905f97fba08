"""One RDS demodulator launch on 5000 samples (one second of the 5 kS/s stream), for ncu."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import sdrplusplus_b200 as sb
from sdrplusplus_b200 import lib
from util import rds_baseband
L = lib.load(); lib.check(L.b200_init(0))
x, _ = rds_baseband(2400, 5)
d = sb.RdsDemod()
for i in range(0, 10000, 5000):
    s, h = d.process(x[i:i + 5000])
print(s.size, int(h.sum()))
