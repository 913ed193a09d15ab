#!/usr/bin/env python3
"""Static SASS instruction counts of the traversal phases, per kernel, from the built cubin.

The issue-slot roofline of the traversal kernels (bench.py `roofline.sm_issue`) is
    useful warp-instructions / s = rays/s x (N_node c_node + N_tri c_tri) / 32
with N_node / N_tri measured live by the counting kernels and c_node / c_tri = the SASS instructions of one node
test / one triangle test.  This tool derives c_node and c_tri from `nvdisasm -g` line attribution: every SASS
instruction of a kernel whose innermost source line lies inside Traverser::node_phase (incl. the byte decode
helpers and the stack push / pop) counts towards c_node, inside triangle_phase / watertight_hit / shear_dot
towards c_tri (the double-precision fallback for exactly-zero edge functions and the alpha-test functor are
listed separately: they are off the common path).  Writes JSON with the cubin's sha256 so that a number in a
bench line can be tied to the binary it was measured on.

usage: tools/sass_counts.py [trace.o] [out.json]"""
import hashlib, json, os, re, subprocess, sys, tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
obj = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "fredholm_b200", "csrc", "build", "trace.o")
out_path = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "fredholm_b200", "sass_counts.json")
src = os.path.join(ROOT, "fredholm_b200", "csrc", "bvh.cuh")


def function_ranges(path):
    """{name: (first line, last line)} of the FR_D functions / methods of bvh.cuh (brace matching)."""
    lines = open(path).read().split("\n")
    out = {}
    i = 0
    while i < len(lines):
        m = re.match(r"\s*(?:template <[^>]*>\s*)?FR_D\s+[\w:<>&\s\*]+?\b(\w+)\(", lines[i])
        if m:
            name, depth, j, seen = m.group(1), 0, i, False
            while j < len(lines):
                depth += lines[j].count("{") - lines[j].count("}")
                seen = seen or "{" in lines[j]
                if seen and depth <= 0:
                    break
                j += 1
            out[name] = (i + 1, j + 1)
            i = j
        i += 1
    return out


fr = function_ranges(src)
NODE = [fr[k] for k in ("node_phase", "byte_f", "sign_extend_s8x4", "push", "pop")]
TRI = [fr[k] for k in ("triangle_phase", "watertight_hit", "shear_dot")]
text = open(src).read().split("\n")
# the zero-edge double fallback inside watertight_hit
dbl = [i + 1 for i, l in enumerate(text) if "__dsub_rn" in l or "(U == 0.0f || V == 0.0f || W == 0.0f)" in l]
DOUBLE = (min(dbl), max(dbl) + 1) if dbl else (0, -1)
# the divided tail of watertight_hit (1 / det, t, u, v): any-hit kernels run it only for alpha-tested triangles
rcp = [i + 1 for i, l in enumerate(text) if "const float rcp = " in l]
UV_TAIL = (rcp[0], next(i + 1 for i in range(rcp[0], len(text)) if text[i].startswith("}"))) if rcp else (0, -1)

with tempfile.TemporaryDirectory() as tmp:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, capture_output=True)
    cubins = [f for f in os.listdir(tmp) if f.endswith(".cubin")]
    assert len(cubins) == 1, cubins
    cubin = os.path.join(tmp, cubins[0])
    dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True, check=True).stdout

# Identity of the build = sha256 of the traversal kernels' instruction stream (kernel names and instruction text, without
# addresses and encodings).  NOT of the cubin file: nvcc puts a fresh unique id into the names of internal-linkage
# symbols at every compilation, so two builds of the same source differ as files while their code is identical.
UNIQUE = re.compile(r"_GLOBAL__N__[0-9a-f]+_\d+_[A-Za-z0-9_]+?_cu_[0-9a-f]+")
hash_lines = []
result = {"cubin": os.path.basename(obj), "cubin_sha256": None, "hash_of": "instruction stream of the k_trace_* kernels (names + SASS text)",
          "source": "nvdisasm -g line attribution, tools/sass_counts.py",
          "line_ranges": {"node": NODE, "triangle": TRI, "double_fallback": DOUBLE, "uv_tail": UV_TAIL}, "kernels": {}}
cur, file_, line = None, None, 0
for l in dis.split("\n"):
    m = re.match(r"\.text\.(\S+):", l)
    if m:
        sym = m.group(1)
        dm = subprocess.run(["c++filt", sym], capture_output=True, text=True).stdout.strip()
        km = re.search(r"(k_trace_\w+?)<([^>]*)>", dm) or re.search(r"(k_trace_\w+)", dm)
        cur = None
        if km:
            # template arguments: <COUNT, TWO> (stage kernels) or <TWO> (k_trace_batch, always counting)
            targs = [a.strip() in ("true", "(bool)1") for a in km.group(2).split(",")] if (km.lastindex or 0) >= 2 else []
            counting = len(targs) == 2 and targs[0]
            two = targs[-1] if targs else False
            name = km.group(1) + ("_count" if counting else "") + ("_two_level" if two else "")
            cur = result["kernels"].setdefault(name, {"total": 0, "c_node": 0, "c_tri": 0, "c_tri_double_fallback": 0,
                                                      "c_tri_alpha_only": 0, "any_hit": "shadow" in name})
            hash_lines.append("== " + name)
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        file_, line = os.path.basename(m.group(1)), int(m.group(2))
        continue
    if cur is None or not re.match(r"\s*/\*[0-9a-f]{4,}\*/", l):
        continue
    cur["total"] += 1
    hash_lines.append(UNIQUE.sub("_anon_", re.sub(r"/\*[0-9a-f]{4,}\*/", "", l)).strip())
    if file_ == "bvh.cuh":
        if DOUBLE[0] <= line <= DOUBLE[1]:
            cur["c_tri_double_fallback"] += 1
        elif cur["any_hit"] and UV_TAIL[0] <= line <= UV_TAIL[1]:
            cur["c_tri_alpha_only"] += 1  # not part of c_tri: off the common path of an any-hit kernel
        elif any(a <= line <= b for a, b in TRI):
            cur["c_tri"] += 1
        elif any(a <= line <= b for a, b in NODE):
            cur["c_node"] += 1
result["cubin_sha256"] = hashlib.sha256("\n".join(hash_lines).encode()).hexdigest()
json.dump(result, open(out_path, "w"), indent=1)
print(json.dumps(result["kernels"]))
