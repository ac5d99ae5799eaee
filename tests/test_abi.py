"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol include/dlv3p.h
declares, and the ctypes table mirrors the header (argument counts)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "dlv3p.h")


def header_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(?:int|const char\*)\s+(dlv3p_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        args = m.group(2).strip()
        n = 0 if args in ("", "void") else len(args.split(","))
        out[m.group(1)] = n
    return out


def test_library_exports_every_declared_symbol():
    from deeplabv3plus_keras_b200 import _lib

    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    lib = _lib.load()
    funcs = header_functions()
    assert len(funcs) >= 40
    for name in funcs:
        assert hasattr(lib, name), f"{name} declared in include/dlv3p.h but not exported by libdlv3p.so"


def test_ctypes_table_matches_header():
    from deeplabv3plus_keras_b200 import _lib

    funcs = header_functions()
    for name, nargs in funcs.items():
        if name in ("dlv3p_last_error", "dlv3p_version", "dlv3p_device_arch", "dlv3p_set_pdl"):
            continue
        assert name in _lib.SIGNATURES, f"{name} missing from the ctypes table"
        assert len(_lib.SIGNATURES[name]) == nargs, f"{name}: header has {nargs} args, table {len(_lib.SIGNATURES[name])}"
    for name in _lib.SIGNATURES:
        assert name in funcs, f"{name} bound in ctypes but not declared in the header"


def test_missing_library_is_loud(monkeypatch, tmp_path):
    from deeplabv3plus_keras_b200 import _lib

    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError):
        _lib.load()


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: no module of the package may import / execute anything under oracle/ except
    smoke.py (the checker of __graft_entry__.smoke()), and nothing in the package may read /root/reference."""
    import ast

    pkg = os.path.join(ROOT, "deeplabv3plus_keras_b200")
    offenders = []
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if not f.endswith(".py"):
                continue
            path = os.path.join(dirpath, f)
            src = open(path).read()
            tree = ast.parse(src)
            docs = {id(n.value) for n in ast.walk(tree) if isinstance(n, ast.Expr) and isinstance(n.value, ast.Constant)}
            for node in ast.walk(tree):          # a path string in code (docstrings may cite the reference)
                if isinstance(node, ast.Constant) and isinstance(node.value, str) and id(node) not in docs \
                        and "/root/reference" in node.value:
                    offenders.append((os.path.relpath(path, ROOT), "/root/reference"))
            if f == "smoke.py":
                continue
            for node in ast.walk(tree):
                mods = []
                if isinstance(node, ast.Import):
                    mods = [a.name for a in node.names]
                elif isinstance(node, ast.ImportFrom) and node.level == 0:
                    mods = [node.module or ""]
                for m in mods:
                    if m == "oracle" or m.startswith("oracle.") or m == "tests" or m.startswith("tests."):
                        offenders.append((os.path.relpath(path, ROOT), m))
    assert not offenders, offenders


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    """No CPU fallback: a missing libdlv3p.so is a RuntimeError at load time, not a silent eager path."""
    from deeplabv3plus_keras_b200 import _lib

    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "absent.so"))
    monkeypatch.setattr(_lib, "_lib", None)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.load()
