import numpy as np, sys
sys.path.insert(0, ".")
from fredholm_b200 import api
from oracle import binding as ob
kinds = "221" + "2" + "12" * 3 + "1" + "2212121" * 6
pos = []
for k in kinds:
    pos += [k] * (1 if k == "1" else 2)
for (w, h, idx, spp) in [(256, 256, 0, 0), (256, 256, 12345, 7), (1920, 1080, 2073599, 63), (64, 64, 77, 16)]:
    a = api.sampler_sequence(w, h, 1, idx, spp, kinds)
    b = ob.sampler_sequence(w, h, 1, idx, spp, kinds)
    bad = np.nonzero(a.view(np.uint32) != b.view(np.uint32))[0]
    print((w, h, idx, spp), "mismatch", len(bad), [(int(i), pos[i], float(a[i]), float(b[i]), hex(a.view(np.uint32)[i]), hex(b.view(np.uint32)[i])) for i in bad[:8]])
