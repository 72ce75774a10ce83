"""The C-ABI library loads, exports every symbol include/cpz.h declares, agrees on struct sizes, and fails loudly
(no CPU fallback) when no GPU is present. No compute calls here."""
import ctypes as C
import os
import re

import pytest

from cpz_b200 import engine
from cpz_b200.desc import CClosureUvtDesc, CClosureDesc, CModelDesc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "cpz.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cpz_[a-z0-9_]+)\s*\(", src)) - {"cpz_allreduce_fn"})


def test_library_exports_every_declared_symbol():
    L = engine.lib()
    names = _declared()
    assert len(names) >= 28
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/cpz.h but not exported by libcpz.so"
    assert sorted(engine.EXPORTS) == names


def test_struct_sizes_match_the_library():
    L = engine.lib()
    assert L.cpz_sizeof_model_desc() == C.sizeof(CModelDesc)
    assert L.cpz_sizeof_closure_desc() == C.sizeof(CClosureDesc)
    assert L.cpz_sizeof_closure_uvt_desc() == C.sizeof(CClosureUvtDesc)
    assert L.cpz_version() == 1


def test_null_handles_are_rejected_with_a_message():
    L = engine.lib()
    out = C.c_void_p()
    assert L.cpz_model_create(None, None, C.byref(out)) == -1
    assert b"null" in L.cpz_last_error()
    assert L.cpz_solve(None, None, None, None, None, 4) == -1
    assert L.cpz_ctx_synchronize(None) == -1
    assert L.cpz_model_destroy(None) == 0 and L.cpz_ctx_destroy(None) == 0


def test_no_cpu_fallback_without_a_gpu():
    if engine.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(engine.CpzError) as ei:
        engine.Context(0)
    assert ei.value.code == -2 and "no CPU fallback" in str(ei.value)
