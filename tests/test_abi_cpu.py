"""CPU: the C-ABI library loads, exports every symbol include/mg_b200.h declares, fails loudly without a GPU,
and its host-only integer pieces (bucket LUTs) agree bit-exactly with the stock transformers function."""
import ctypes
import os
import re

import pytest
import torch

from markushgrapher_b200 import _lib
from markushgrapher_b200.configuration import MarkushgrapherConfig, expected_weights

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "mg_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mg_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = _lib.lib()
    syms = declared_symbols()
    assert "mg_generate" in syms and "mg_generate_host" in syms and "mg_encode" in syms
    for s in syms:
        assert hasattr(lib, s), f"{s} is declared in include/mg_b200.h but not exported"


@pytest.mark.parametrize("bidir,nb,maxd", [(1, 32, 128), (1, 32, 100), (0, 32, 128)])
def test_bucket_lut_matches_transformers(bidir, nb, maxd):
    from transformers.models.udop.modeling_udop import UdopAttention

    n = 3000
    lut = (ctypes.c_int32 * n)()
    rc = _lib.lib().mg_rel_bucket_lut(bidir, nb, maxd, n, lut)
    _lib.check(rc, "mg_rel_bucket_lut")
    mine = torch.tensor(list(lut), dtype=torch.long)
    rel = torch.arange(n, dtype=torch.long)
    if bidir:
        # positive relative positions get +nb/2, negative ones the plain bucket of |rel|
        ref_neg = UdopAttention._relative_position_bucket(-rel, bidirectional=True, num_buckets=nb, max_distance=maxd)
        ref_pos = UdopAttention._relative_position_bucket(rel, bidirectional=True, num_buckets=nb, max_distance=maxd)
        assert torch.equal(mine, ref_neg)
        assert torch.equal(mine[1:] + nb // 2, ref_pos[1:])
    else:
        ref = UdopAttention._relative_position_bucket(-rel, bidirectional=False, num_buckets=nb, max_distance=maxd)
        assert torch.equal(mine, ref)


def test_no_gpu_fails_loudly():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from markushgrapher_b200.engine import MGEngine

    assert _lib.lib().mg_device_available() == 0
    with pytest.raises(_lib.MgError):
        MGEngine(MarkushgrapherConfig(), {})


def test_create_and_error_reporting_without_gpu():
    from markushgrapher_b200.engine import make_c_config, mg_config

    lib = _lib.lib()
    lib.mg_create.argtypes = [ctypes.POINTER(mg_config), ctypes.POINTER(ctypes.c_void_p)]
    lib.mg_destroy.argtypes = [ctypes.c_void_p]
    lib.mg_destroy.restype = None
    c = make_c_config(MarkushgrapherConfig())
    h = ctypes.c_void_p()
    assert lib.mg_create(ctypes.byref(c), ctypes.byref(h)) == 0
    assert lib.mg_create(None, ctypes.byref(h)) < 0
    assert b"null" in lib.mg_last_error()
    lib.mg_destroy(h)


def test_expected_weights_match_oracle_names():
    """the product-side parameter inventory is exactly what the oracle (stock HF modules) exports"""
    from oracle import mg_oracle as O

    for cfg in (O.MGConfig.tiny(), O.MGConfig.small()):
        state = O.MGOracle(cfg).export_state()
        want = expected_weights(MarkushgrapherConfig.from_dims(cfg))
        for name, shp in want.items():
            assert name in state, name
            assert tuple(state[name].shape) == tuple(shp), (name, state[name].shape, shp)


def test_full_config_parameter_count():
    """831 M parameters (reference README.md:217) = UDOP-large + Swin-B + projector (+ the untied LM head copy)"""
    w = expected_weights(MarkushgrapherConfig())
    n = 0
    for name, shp in w.items():
        k = 1
        for s in shp:
            k *= s
        n += k
    n_tied = n - 33201 * 1024  # the reference ties lm_head to the shared embedding
    assert abs(n_tied - 831.4e6) / 831.4e6 < 0.01, n_tied
