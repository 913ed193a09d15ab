"""The C-ABI boundary (include/fredholm_b200.h): the shared library loads, exports every
symbol the header declares, the ctypes signature table covers the header, and compute
entry points fail loudly (no CPU fallback) when there is no CUDA device."""
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from fredholm_b200 import api

HEADER = os.path.join(ROOT, "include", "fredholm_b200.h")


def header_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fr_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_something():
    syms = header_symbols()
    assert len(syms) >= 60
    for must in ("fr_renderer_create", "fr_load_scene", "fr_build_accel", "fr_render", "fr_wait",
                 "fr_render_frame_host", "fr_post_process", "fr_tone_mapping", "fr_copy_to_host"):
        assert must in syms


def test_library_exports_every_header_symbol():
    assert os.path.exists(api.LIB_PATH), "libfredholm_b200.so not built (python -c 'import __graft_entry__ as g; g.build()')"
    out = subprocess.run(["nm", "-D", "--defined-only", api.LIB_PATH], check=True, capture_output=True, text=True).stdout
    exported = {line.split()[-1] for line in out.splitlines() if line.strip()}
    missing = [s for s in header_symbols() if s not in exported]
    assert not missing, missing


def test_ctypes_table_matches_header():
    syms = set(header_symbols())
    table = set(api.SIGNATURES)
    assert syms == table, (sorted(syms - table), sorted(table - syms))
    L = api.lib()   # binds every entry of the table (raises AttributeError if one is missing)
    assert L.fr_version().decode().startswith("fredholm_b200 ")
    assert "sm_100a" in L.fr_version().decode()


def test_library_is_sm100a_only():
    out = subprocess.run(["cuobjdump", "-lelf", api.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_no_cpu_fallback():
    """Without a CUDA device the product refuses to run (it never routes to the oracle)."""
    if api.lib().fr_device_count() > 0:
        pytest.skip("CUDA device present")
    with pytest.raises(api.FredholmError):
        api.Renderer(0)
    assert api.lib().fr_renderer_create(0) is None
    assert b"CUDA" in api.lib().fr_last_error() or b"cuda" in api.lib().fr_last_error()
    with pytest.raises(api.FredholmError):
        api.sampler_sequence(64, 64, 1, 0, 0, "12")
    with pytest.raises(api.FredholmError):
        api.bsdf_eval_sample(np.zeros((1, 40), np.float32))


def test_product_does_not_import_oracle():
    """oracle/ is test infrastructure: nothing in the package or the C++ sources refers to it."""
    sources = []
    for top in ("fredholm_b200", "tools", "examples", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, top)):
            if any(d.startswith("build") for d in dirpath.split(os.sep)):
                continue
            sources += [os.path.join(dirpath, f) for f in files
                        if f.endswith((".py", ".cpp", ".cu", ".h", ".cuh", ".sh"))]
    assert len(sources) > 30
    for path in sources:
        text = open(path, errors="replace").read()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), path
        assert "libfredholm_oracle" not in text, path
