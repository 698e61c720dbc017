"""Compares the /tmp/ws_<tag>.npz dumps written by tools/ws_check.py (run on the same box): max |diff| and unequal cells per output."""
import sys, numpy as np
a = np.load("/tmp/ws_%s.npz" % sys.argv[1])
for t in sys.argv[2:]:
    b = np.load("/tmp/ws_%s.npz" % t)
    for k in a.files:
        x, y = a[k].astype(np.float64), b[k].astype(np.float64)
        d = np.abs(x - y)
        print(t, k, "max|d| %.3g" % d.max(), "rel %.3g" % (d / np.maximum(np.abs(x), 1e-30)).max() if k == "stats" else "", "neq", int((x != y).sum()), "of", x.size)
