"""CPU: the C-ABI shared library loads and exports every symbol include/neddf_b200.h declares
(no compute calls - there is no GPU in the build container)."""
import ctypes
import os
import re

import pytest


@pytest.fixture(scope="module")
def built(repo_root):
    import __graft_entry__ as ge
    ge.build()
    return os.path.join(repo_root, "neddf_b200", "libneddf_b200.so")


def _declared(repo_root):
    src = open(os.path.join(repo_root, "include", "neddf_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(neddf_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(repo_root, built):
    lib = ctypes.CDLL(built)
    names = _declared(repo_root)
    assert len(names) >= 17
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/neddf_b200.h but not exported"


def test_python_binding_covers_header(repo_root, built):
    from neddf_b200 import _lib as L
    assert L.exported_symbols() == _declared(repo_root)
    lib = L.lib()
    assert lib.neddf_abi_version() == 1
    assert lib.neddf_launch_count() == 0


def test_layer_shapes_and_validation(built):
    import ctypes as C

    import neddf_b200
    from neddf_b200 import _lib as L
    from oracle import neddf_oracle as orc
    lib = L.lib()
    for kw in (dict(col_layer_count=4), dict(), dict(ddf_layer_count=6, skips=[2]), dict(skips=[1, 4])):
        net = neddf_b200.NeDDF(**kw)
        cfg = net._config_struct()
        buf = (C.c_int32 * 96)()
        n = lib.neddf_field_layer_shapes(C.byref(cfg), buf, 48)
        ref = orc.layer_shapes(orc.FieldConfig(**kw))
        assert n == len(ref)
        assert [(buf[2 * i], buf[2 * i + 1]) for i in range(n)] == [(a, b) for _, a, b in ref]
        # module parameter shapes are the reference's (weight stored [in,out])
        sd = net.state_dict()
        for name, a, b in ref:
            assert tuple(sd[name + ".weight"].shape) == (a, b) and tuple(sd[name + ".bias"].shape) == (b,)
    bad = neddf_b200.NeDDF(ddf_layer_width=128, col_layer_width=128)
    cfg = bad._config_struct()
    assert lib.neddf_field_layer_shapes(C.byref(cfg), None, 0) == -3  # NEDDF_E_UNSUPPORTED
    assert b"256" in lib.neddf_last_error()
