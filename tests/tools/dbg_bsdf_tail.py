"""Tail of the BSDF golden comparison (GPU box): per material class, max / quantiles of the relative error and
the inputs of the worst cases -- the data behind the bounds in tests/test_gpu_golden.py::test_bsdf_golden."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from fredholm_b200 import api
g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "golden", "bsdf.npz"))
cases, want, labels = g["cases"], g["out"].astype(np.float64), g["labels"]
got = api.bsdf_eval_sample(cases).astype(np.float64)
np.set_printoptions(precision=5, suppress=True, linewidth=220)
nanm = np.isnan(got) != np.isnan(want)
print("nan mismatches:", int(nanm.sum()), "of", nanm.size, "rows", np.where(nanm.any(axis=1))[0][:20])
err = np.abs(got - want) / (np.abs(want) + 1e-3)
err[np.isnan(err)] = 0
for cls in dict.fromkeys(labels.tolist()):
    s = labels == cls
    e = err[s]
    print("%-12s eval max %.2e  sample-dir max %.2e  sample-f/pdf max %.2e  q99 %.2e  q999 %.2e" % (
        cls, e[:, :4].max(), e[:, 4:7].max(), e[:, 7:].max(), np.quantile(e, 0.99), np.quantile(e, 0.999)))
worst = np.argsort(err.max(axis=1))[-12:]
for i in worst:
    c = cases[i]
    print(labels[i], "row", i, "wo", c[30:33], "entering", c[33], "wi", c[34:37], "u", c[37], "v", c[38:40])
    print("    got ", got[i])
    print("    want", want[i])
