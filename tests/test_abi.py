"""The C-ABI library loads and exports every symbol include/ast.h declares (no compute: no GPU needed)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "ast.h")).read()
    return sorted(set(re.findall(r"\b(ast_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported():
    from artist_style_transfer_b200 import _lib
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/ast.h but not exported by libast_b200.so"
    assert sorted(_lib.EXPORTS) == names, "ctypes binding and header disagree"
    assert lib.ast_abi_version() == 4
    assert lib.ast_instnorm_workspace_bytes(4, 128) > 0
    assert lib.ast_launch_count() >= 0


def test_argument_errors_are_loud():
    from artist_style_transfer_b200 import _lib
    lib = _lib.load()
    rc = lib.ast_pack_weights(None, None, 1, 1, 1, 1, 1, None, 0, None)
    assert rc < 0 and b"null" in lib.ast_last_error()


def test_no_cpu_fallback():
    import torch
    import artist_style_transfer_b200 as ast
    with pytest.raises(RuntimeError, match="CUDA"):
        ast.gram(torch.zeros(1, 4, 2, 2))
    net = ast.StyleTransfer(device="cpu")
    with pytest.raises(RuntimeError, match="CUDA"):
        net(torch.zeros(1, 3, 8, 8))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "artist_style_transfer_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), f"{f} imports the oracle"
                assert "/root/reference" not in src or f.endswith(".py"), f
