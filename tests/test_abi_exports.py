"""The C-ABI library loads on a CPU-only box and exports every symbol include/ngp.h declares.
No compute call is made here (that needs a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "ngp.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ngp_[a-z_0-9]+)\s*\(", src)))


def test_header_symbols_are_exported():
    from neuro_genetic_pong_self_play_b200 import _lib, build
    path = build.build()
    L = ctypes.CDLL(path)
    names = _declared()
    assert len(names) >= 16
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/ngp.h but not exported by libngp.so"
    assert set(_lib.EXPORTS) == set(names)


def test_default_config_mirrors_reference_config_py():
    from neuro_genetic_pong_self_play_b200 import Config, _lib
    L = _lib.load()
    c = _lib.NgpConfig()
    L.ngp_default_config(ctypes.byref(c), 64)
    py = Config().to_c()
    for name, _ in _lib.NgpConfig._fields_:
        a, b = getattr(c, name), getattr(py, name)
        if hasattr(a, "__len__"):
            assert list(a) == list(b), name
        else:
            assert a == b, name
    assert c.tournament_size == 16 and list(c.nodes)[:3] == [6, 2, 2] and c.games_to_play == 6
    assert L.ngp_version().startswith(b"ngp")


def test_engine_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from neuro_genetic_pong_self_play_b200 import Engine, NgpError
    with pytest.raises(NgpError):
        Engine()


def test_product_does_not_import_oracle():
    """The product package must never route through oracle/ (or any CPU fallback)."""
    pkg = os.path.join(ROOT, "neuro_genetic_pong_self_play_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", text, flags=re.M), f
                assert "oracle/" not in text or f == "build.py", f


def test_generated_core_is_up_to_date(tmp_path):
    """csrc/generated/pong_core.inc is exactly what tools/gen_rom_core.py emits for the bundled cartridge."""
    import subprocess
    import sys
    out = tmp_path / "pong_core.inc"
    subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "gen_rom_core.py"), str(out)])
    committed = open(os.path.join(ROOT, "neuro_genetic_pong_self_play_b200", "csrc", "generated", "pong_core.inc")).read()
    assert out.read_text() == committed


def test_integration_md_struct_matches_header():
    """The ctypes struct INTEGRATION.md tells a maintainer to paste must be the library's ngp_config field for field."""
    import ctypes
    import re
    from neuro_genetic_pong_self_play_b200 import _lib
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    m = re.search(r"class NgpConfig\(ctypes\.Structure\):.*?\n    _fields_ = \[(.*?)\]\n", text, re.S)
    assert m, "INTEGRATION.md lost its NgpConfig definition"
    ns = {"ctypes": ctypes}
    exec("class NgpConfig(ctypes.Structure):\n    _fields_ = [" + m.group(1) + "]\n", ns)
    doc, real = ns["NgpConfig"], _lib.NgpConfig
    assert [f[0] for f in doc._fields_] == [f[0] for f in real._fields_]
    assert ctypes.sizeof(doc) == ctypes.sizeof(real)
    for (name, _), (_, _) in zip(doc._fields_, real._fields_):
        assert getattr(doc, name).offset == getattr(real, name).offset and getattr(doc, name).size == getattr(real, name).size, name
    # and the header itself: every field name of the binding appears, in order, in include/ngp.h's struct
    hdr = open(os.path.join(ROOT, "include", "ngp.h")).read()
    body = hdr[hdr.index("typedef struct {"):hdr.index("} ngp_config;")]
    pos = -1
    for name, _ in real._fields_:
        nxt = body.find(name, pos + 1)
        assert nxt > pos, f"{name} missing or out of order in include/ngp.h"
        pos = nxt
    if os.path.exists(_lib.lib_path()):
        L = ctypes.CDLL(_lib.lib_path())
        L.ngp_config_size.restype = ctypes.c_int32
        assert L.ngp_config_size() == ctypes.sizeof(doc)
