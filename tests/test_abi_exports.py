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
