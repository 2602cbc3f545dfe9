"""CPU-only: the C-ABI library builds, loads and exports every symbol include/mtam.h declares; the
planner and argument validation work without a GPU (no compute calls)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "mtam.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mtam_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(lib_built):
    from mtamrecommender_b200 import _lib
    lib = C.CDLL(lib_built)
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/mtam.h but not exported"
    assert set(names) == set(_lib.SIGNATURES), set(names) ^ set(_lib.SIGNATURES)


def test_plan_sizes_and_validation(lib_built):
    from mtamrecommender_b200 import _lib
    from mtamrecommender_b200.engine import ModelConfig
    lib = _lib.load()
    s = _lib.Sizes()
    c = ModelConfig(kind="MTAM", max_batch=1024, L=50, D=64, H=1, N=6, user_count=1_000_000, item_count=100_000,
                    category_count=1000).to_c()
    assert lib.mtam_plan(C.byref(c), C.byref(s)) == 0
    tables = (1_000_003 + 100_003 + 1003 + 53) * 64
    assert s.param_floats > tables and s.param_floats < tables + 2_000_000
    assert s.workspace_bytes > 100e6
    bad = ModelConfig(kind="MTAM", D=48).to_c()
    assert lib.mtam_plan(C.byref(bad), C.byref(s)) == -1
    assert b"num_units" in lib.mtam_last_error(None)
    bad = ModelConfig(kind="MTAM", D=64, H=3).to_c()
    assert lib.mtam_plan(C.byref(bad), C.byref(s)) == -1
    c.abi_version = 99
    assert lib.mtam_plan(C.byref(c), C.byref(s)) == -1


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from mtamrecommender_b200 import _lib
    from mtamrecommender_b200.engine import Engine, ModelConfig
    with pytest.raises(_lib.MtamError):
        Engine(ModelConfig(kind="MTAM", D=64, L=5, N=1, user_count=3, item_count=9, category_count=2))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "mtamrecommender_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith(".py"):
                txt = open(os.path.join(dp, f)).read()
                assert "oracle" not in re.sub(r'""".*?"""', "", txt, flags=re.S).replace("#", "\n#").split("\n#")[0] \
                    or "import oracle" not in txt and "from oracle" not in txt, f
                assert "from oracle" not in txt and "import oracle" not in txt, f
