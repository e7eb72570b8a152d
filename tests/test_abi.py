"""CPU tests of the drop-in boundary: the C-ABI library loads without a GPU, exports every
symbol include/pzgpu.h declares, fills descriptors with the reference's constants, and
fails loudly (no CPU fallback) when asked to compute without a device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import pzhelpers as H


def test_exports_every_declared_symbol(pz):
    hdr = open(os.path.join(H.ROOT, "include", "pzgpu.h")).read()
    declared = set(re.findall(r"PZ_API [\w\s\*]+?\b(pz_\w+)\(", hdr))
    assert declared, "header parse failed"
    L = pz.lib()
    for sym in declared:
        assert hasattr(L, sym), f"libpzgpu.so does not export {sym}"
    assert declared == set(pz.exported_symbols()) | {"pz_abi_version", "pz_last_error"} or declared >= set(pz.exported_symbols())
    assert L.pz_abi_version() == 1


def test_struct_layout_matches_header(pz):
    # the ctypes mirrors used by the host side and by the oracle binding must agree
    assert C.sizeof(pz.ModelDesc) == C.sizeof(H.ModelDesc) == 8 * 4 + 8 * (36 + 36 + 18 + 9 + 6) + C.sizeof(H.MttParams)
    assert C.sizeof(pz.StepSummary) == C.sizeof(H.StepSummary) == 8 + 8 + 8 + 48 + 48 + 8 + 4 + 4 + 8


def test_descriptor_constants_match_reference(pz):
    for mine, ref in ((pz.rw1d(), H.model_rw1d()), (pz.stt(), H.model_stt()), (pz.mtt_track(), H.model_mtt_track())):
        assert (mine.kind, mine.nx, mine.ny, mine.scalar_ll_2x) == (ref.kind, ref.nx, ref.ny, ref.scalar_ll_2x)
        for a, b in zip(H.desc_arrays(mine), H.desc_arrays(ref)):
            assert np.array_equal(a, b)
    m = pz.mtt()
    assert m.mtt.clutter_lambda == 3 and m.mtt.pd == 0.8 and abs(m.mtt.survival_prob - np.exp(-0.02)) < 1e-15
    assert m.mtt.k_max == 16


def test_no_cpu_fallback(pz):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(pz.PzError) as e:
        pz.zparticles(1000, pz.rw1d())
    assert e.value.code == -2
    with pytest.raises(pz.PzError):
        pz.resample_systematic(np.zeros(10), 0.5)


def test_argument_validation(pz):
    d = pz.rw1d()
    d.nx = 9
    h = C.c_void_p()
    assert pz.lib().pz_filter_create(C.byref(d), 100, 0, 0, C.byref(h)) == -1
    assert b"nx=9" in pz.lib().pz_last_error()
    assert pz.lib().pz_model_rw1d(None, 1.0, 3.0) == -1


def test_delayed_descriptor_validation(pz):
    """model.delayed = 1 (delay=True): one GPU only (every particle is the same marginal), and the posterior
    read-out refuses null pointers; both are argument errors, reported before any device work."""
    d = pz.stt(delay=True)
    assert d.delayed == 1 and pz.stt().delayed == 0
    h = C.c_void_p()
    ident = (C.c_char * 128)()
    rc = pz.lib().pz_filter_create_sharded(C.byref(d), 2 * 256 * 4, 0, 0, 0, 2, ident, C.byref(h))
    assert rc == -1 and b"delayed" in pz.lib().pz_last_error()
    assert pz.lib().pz_filter_read_posterior(None, None, None) == -1
