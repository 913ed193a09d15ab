"""Host comparison of csrc/bsdf.cuh (tests/tools/bsdf_host_harness.cpp) with the reference BSDF (oracle) on
cases the GPU golden table does not hold: wo anywhere on the sphere (normal maps put the viewer below the
shading horizon), rough / smooth, all material classes.  Development tool."""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from bsdf_cases import MATERIAL_CLASSES
from fredholm_b200.types import shading_params
from oracle import binding as ob

L = C.CDLL(os.path.join(ROOT, "tests", "tools", "_build", "libbsdf_host.so"))

def host(cases):
    c = np.ascontiguousarray(cases, np.float32)
    out = np.zeros((len(c), 11), np.float32)
    L.host_bsdf_eval_sample(c.ctypes.data_as(C.c_void_p), C.c_uint32(len(c)), out.ctypes.data_as(C.c_void_p))
    return out

def make(n_per_class, seed, wo_sign):
    rng = np.random.default_rng(seed)
    rows, labels = [], []
    for name, kw in MATERIAL_CLASSES.items():
        sp = shading_params(**kw)
        for i in range(n_per_class):
            wo = rng.normal(size=3); wo /= np.linalg.norm(wo)
            if wo_sign > 0: wo[1] = abs(wo[1])
            if wo_sign < 0: wo[1] = -abs(wo[1])
            wi = rng.normal(size=3); wi /= np.linalg.norm(wi)
            rows.append(np.concatenate([sp, wo, [1.0 if i % 4 != 3 else 0.0], wi, [rng.uniform()], rng.uniform(size=2)]))
            labels.append(name)
    return np.asarray(rows, np.float32), np.array(labels)

for sign, what in ((1, "wo.y > 0"), (-1, "wo.y < 0")):
    cases, labels = make(400, 11, sign)
    a = host(cases).astype(np.float64); b = ob.bsdf_eval_sample(cases).astype(np.float64)
    print("==", what)
    for cls in MATERIAL_CLASSES:
        s = labels == cls
        nanm = (np.isnan(a[s]) != np.isnan(b[s]))
        err = np.abs(a[s] - b[s]) / (np.abs(b[s]) + 1e-3); err[np.isnan(err)] = 0
        print("%-12s nan mismatch %s  max err per column %s" % (cls, nanm.sum(axis=0), np.array2string(err.max(axis=0), precision=2)))

if len(sys.argv) > 1:
    cls = sys.argv[1]
    cases, labels = make(400, 11, -1)
    s = labels == cls
    a = host(cases[s]).astype(np.float64); b = ob.bsdf_eval_sample(cases[s]).astype(np.float64)
    err = np.abs(a - b) / (np.abs(b) + 1e-3); err[np.isnan(err)] = 0
    worst = np.argsort(err[:, 4:7].max(axis=1))[-6:]
    np.set_printoptions(precision=6, suppress=True, linewidth=200)
    for i in worst:
        c = cases[s][i]
        print("wo", c[30:33], "entering", c[33], "u", c[37], "v", c[38:40])
        print("   ours wi", a[i, 4:7], "f", a[i, 7:10], "pdf", a[i, 10])
        print("   ref  wi", b[i, 4:7], "f", b[i, 7:10], "pdf", b[i, 10])

if len(sys.argv) > 2 and sys.argv[2] == "nan":
    cls = sys.argv[1]
    cases, labels = make(400, 11, -1)
    s = labels == cls
    a = host(cases[s]); b = ob.bsdf_eval_sample(cases[s])
    bad = np.where((np.isnan(a) != np.isnan(b)).any(axis=1))[0]
    for i in bad:
        c = cases[s][i]
        print("wo", c[30:33], "entering", c[33], "u", c[37], "v", c[38:40])
        print("   ours", a[i, 4:]); print("   ref ", b[i, 4:])
