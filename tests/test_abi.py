"""The C-ABI library loads and exports every symbol include/fortuna_b200.h declares (no GPU needed)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as ge

    ge.build()
    from fortuna_b200 import _lib

    return _lib


def _declared():
    text = open(os.path.join(ROOT, "include", "fortuna_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fb200_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_all_exported(built):
    handle = ctypes.CDLL(built.LIB_PATH)
    names = _declared()
    assert len(names) >= 26
    for n in names:
        assert hasattr(handle, n), f"{n} declared in the header but not exported"
    assert sorted(built.SIGNATURES) == names, "ctypes binding table and header disagree"


def test_version_and_error_strings(built):
    lib = built.lib()
    assert lib.fb200_abi_version() == 1
    assert built.error_string(0) == "ok"
    for code in (-1, -2, -3, -4, -5):
        assert "fortuna_b200" in built.error_string(code)


def test_host_tile_table(built):
    lib = built.lib()
    leaves = (built.Leaf * 3)()
    for i, (o, n) in enumerate([(0, 5), (8, 4096), (4104, 1)]):
        leaves[i].offset, leaves[i].len = o, n
    nt = lib.fb200_sgmcmc_num_tiles_host(leaves, 3)
    assert nt == 1 + 2 + 1  # ceil(3/1024) + ceil(2048/1024) + ceil(1/1024)
    tiles = (built.Tile * nt)()
    assert lib.fb200_sgmcmc_build_tiles_host(leaves, 3, tiles, nt) == 0
    assert [(t.leaf, t.pair_start) for t in tiles] == [(0, 0), (1, 0), (1, 1024), (2, 0)]
    leaves[0].len = 0
    assert lib.fb200_sgmcmc_num_tiles_host(leaves, 3) == built.lib().fb200_sgmcmc_num_tiles_host(leaves, 3) < 0
    assert lib.fb200_sort_workspace_bytes(5000) == 8192 * 4


def test_argument_errors_without_gpu(built):
    lib = built.lib()
    assert lib.fb200_random_bits(None, 4, None, None) == -1
    assert lib.fb200_bma_reduce_cls(None, 0, None, None, 1, 1, 1, None, None) == -1
    assert lib.fb200_conformal_sets(None, 1, None, 1, 1, 0.5, None, None, None) == -1


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under fortuna_b200/ may reference it."""
    pkg = os.path.join(ROOT, "fortuna_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{f} imports the oracle"


def test_missing_library_fails_loudly(built, monkeypatch):
    monkeypatch.setattr(built, "_LIB", None)
    monkeypatch.setattr(built, "LIB_PATH", "/nonexistent/libfortuna_b200.so")
    with pytest.raises(built.FortunaB200LibraryError):
        built.lib()


def test_xla_ffi_shim_passes_the_syntax_check():
    """fortuna_b200/xla_ffi/fb200_xla_ffi.cc compiles (g++ -fsyntax-only) against the mock XLA FFI header, whose
    Binding::To() static_asserts every handler against its binding, and against the real C-ABI header; the Python-side
    target list names exactly the handlers the shim defines.  A syntax / signature check - jaxlib is not installable."""
    import subprocess

    r = subprocess.run(["bash", os.path.join(ROOT, "scripts", "check_xla_ffi_syntax.sh")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    src = open(os.path.join(ROOT, "fortuna_b200", "xla_ffi", "fb200_xla_ffi.cc")).read()
    handlers = re.findall(r"^XLA_FFI_DEFINE_HANDLER_SYMBOL\((\w+),", src, flags=re.M)
    assert len(handlers) >= 30 and len(set(handlers)) == len(handlers)
    py = open(os.path.join(ROOT, "fortuna_b200", "xla_ffi", "jax_bindings.py")).read()
    targets = re.findall(r'"(fb200_xla_\w+)"', py.split("_TARGETS = (")[1].split(")")[0])
    assert targets == handlers
    # every fb200_* entry point a handler forwards to is declared in the header
    called = set(re.findall(r"\b(fb200_[a-z0-9_]+)\s*\(", src)) - {"fb200_error_string"}
    assert called <= set(_declared()), called - set(_declared())
