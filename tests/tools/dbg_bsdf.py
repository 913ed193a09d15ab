import sys
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from bsdf_cases import make_cases
from fredholm_b200 import api
from oracle import binding as ob
cases, labels = make_cases(256)
g = api.bsdf_eval_sample(cases)
o = ob.bsdf_eval_sample(cases)
labels = np.array(labels)
names = ["f.x", "f.y", "f.z", "pdf", "wi.x", "wi.y", "wi.z", "fs.x", "fs.y", "fs.z", "pdfs"]
for cls in dict.fromkeys(labels):
    m = labels == cls
    gg, oo = g[m].astype(np.float64), o[m].astype(np.float64)
    both_nan = np.isnan(gg) & np.isnan(oo)
    err = np.abs(gg - oo) / (np.abs(oo) + 1e-6)
    err[both_nan] = 0
    nanmis = (np.isnan(gg) != np.isnan(oo)).sum()
    worst = np.nanmax(err, axis=0)
    print("%-12s nan-mismatch %d  worst rel err per column: %s" % (cls, nanmis, " ".join("%s=%.1e" % (n, w) for n, w in zip(names, worst))))
    bad = np.nonzero(np.nanmax(err, axis=1) > 1e-3)[0][:3]
    for b in bad:
        print("   case", b, "entering", cases[m][b, 33], "wo", cases[m][b, 30:33], "wi", cases[m][b, 34:37], "u", cases[m][b, 37:40])
        print("     gpu", gg[b]); print("     ref", oo[b])
