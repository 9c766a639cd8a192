"""The C-ABI library loads without a GPU and exports every symbol include/bp4.h declares;
without a device the calls fail loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from mf_data_locality_b200 import build, capi
    build.build_cuda()
    return capi.lib()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "bp4.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bp4_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(lib):
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/bp4.h but not exported by libbp4.so"
    from mf_data_locality_b200 import capi
    assert set(capi.EXPORTS) == set(names)


def test_host_libraries_export_plugin():
    from mf_data_locality_b200 import build, host
    build.build_all()
    assert host.lib("plain").bp4h_plugin() == b"benchmark_precond"
    assert host.lib("merged").bp4h_plugin() == b"benchmark_precond_merged"


def test_no_cpu_fallback(lib):
    import numpy as np
    from mf_data_locality_b200 import capi
    n = C.c_int(-1)
    rc = lib.bp4_device_count(C.byref(n))
    if rc == 0 and n.value > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(capi.Bp4Error):
        capi.Context(3, np.zeros((1, 27), np.uint32), np.zeros((1, 8, 3)), 3)


def test_argument_errors(lib):
    assert lib.bp4_ctx_create(None, None) == -1
    assert b"null" in lib.bp4_last_error()
    assert lib.bp4_vec_size(None, None) == -1


def test_product_does_not_import_oracle():
    """the oracle is test infrastructure: nothing under mf_data_locality_b200/ may reference it"""
    pkg = os.path.join(ROOT, "mf_data_locality_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".h", ".cc", ".cu", ".cuh")):
                text = open(os.path.join(d, f), errors="ignore").read()
                assert "oracle" not in text.lower().replace("# oracle-free", ""), os.path.join(d, f)
