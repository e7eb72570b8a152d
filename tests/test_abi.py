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
    assert L.pz_abi_version() == 2


def test_struct_layout_matches_header(pz):
    # the ctypes mirrors used by the host side and by the oracle binding must agree
    assert C.sizeof(pz.ModelDesc) == C.sizeof(H.ModelDesc) == 8 * 4 + 8 * (36 + 36 + 18 + 9 + 6) + C.sizeof(H.MttParams) + 8 + 4 + 4 + C.sizeof(H.WalkerParams) + 16
    assert C.sizeof(pz.WalkerParams) == C.sizeof(H.WalkerParams) == 8 * (3 + 3 + 9 + 3 + 3) + 8
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


def test_every_argument_error_is_reported_before_device_work(pz):
    """Error behaviour of the boundary: bad descriptors, sizes, ranks and null
    pointers are PZ_EINVAL (-1) with a message, on a machine with or without a GPU; nothing is created."""
    L = pz.lib()
    ident = (C.c_char * 128)()

    def create(d, n, rank=0, world=1, nccl=ident):
        h = C.c_void_p()
        if world == 1:
            rc = L.pz_filter_create(C.byref(d), n, 0, 0, C.byref(h))
        else:
            rc = L.pz_filter_create_sharded(C.byref(d), n, 0, 0, rank, world, nccl, C.byref(h))
        assert not h.value
        return rc, L.pz_last_error()

    d = pz.rw1d(); d.kind = 77
    assert create(d, 1000) == (-1, b"unknown model kind 77")
    d = pz.rw1d(); d.nx, d.ny = 2, 1
    rc, msg = create(d, 1000)
    assert rc == -1 and b"no kernel instantiated for nx=2 ny=1" in msg
    rc, msg = create(pz.rw1d(), 0)
    assert rc == -1 and b"n_particles" in msg
    rc, msg = create(pz.rw1d(), (1 << 30) + 1)
    assert rc == -1 and b"2^30" in msg
    # sharded: rank / world, whole resampling blocks per rank, systematic resampler only, an id is required
    rc, msg = create(pz.rw1d(), 2 * 256 * 8, rank=2, world=2)
    assert rc == -1 and b"bad rank 2 / world 2" in msg
    rc, msg = create(pz.rw1d(), 2 * 256 * 8 + 256, rank=0, world=2)
    assert rc == -1 and b"divisible by world * 256" in msg
    rc, msg = create(pz.rw1d(resampler=pz.PZ_RESAMPLE_MULTINOMIAL_REF), 2 * 256 * 8, rank=0, world=2)
    assert rc == -1 and b"systematic" in msg
    rc, msg = create(pz.rw1d(), 2 * 256 * 8, rank=0, world=2, nccl=None)
    assert rc == -1 and b"nccl_unique_id" in msg
    rc, msg = create(pz.rw1d(), 33 * 256, rank=0, world=33)
    assert rc == -1 and b"at most 32 ranks" in msg
    # tracker: fixed capacities, one GPU
    m = pz.mtt(); m.mtt.k_max = 8
    rc, msg = create(m, 1000)
    assert rc == -1 and b"k_max" in msg
    m = pz.mtt(); m.mtt.m_max = 33
    rc, msg = create(m, 1000)
    assert rc == -1 and b"m_max" in msg
    rc, msg = create(pz.mtt(), (1 << 28) + 1)
    assert rc == -1 and b"2^28" in msg
    rc, msg = create(pz.mtt(), 2 * 256 * 8, rank=0, world=2)
    assert rc == -1 and b"pz_filter_create_multi" in msg   # a tracker over several GPUs is a single-process multi-device filter
    # multi-device (one process): device list and per-device share are checked before any device is touched
    def create_multi(d, n, ids):
        h = C.c_void_p()
        arr = (C.c_int * len(ids))(*ids)
        rc = L.pz_filter_create_multi(C.byref(d), n, 0, len(ids), arr, C.byref(h))
        assert not h.value
        return rc, L.pz_last_error()
    rc, msg = create_multi(pz.mtt(), 100000, [0, 0])
    assert rc == -1 and b"listed twice" in msg
    rc, msg = create_multi(pz.mtt(), 100000, list(range(9)))
    assert rc == -1 and b"up to 8 devices" in msg
    monkey = os.environ.get("PZ_MULTI_SAME_DEVICE")
    os.environ["PZ_MULTI_SAME_DEVICE"] = "1"
    try:
        rc, msg = create_multi(pz.mtt(), 3000, [0, 0, 0])
        assert rc == -1 and b"too few for 3 devices" in msg   # whole 2048-slot resampling tiles per device
        rc, msg = create_multi(pz.rw1d(), 256 * 8, [0, 0])
        assert rc == -1 and b"listed twice" in msg             # the hook applies to the trackers only
    finally:
        if monkey is None:
            del os.environ["PZ_MULTI_SAME_DEVICE"]
        else:
            os.environ["PZ_MULTI_SAME_DEVICE"] = monkey
    mc = pz.multicam(); mc.F[1] = 0.5
    rc, msg = create(mc, 1000)
    assert rc == -1 and b"F = I" in msg
    mc = pz.multicam(); mc.mtt.pd = 0.9
    rc, msg = create(mc, 1000)
    assert rc == -1 and b"pd = 1" in msg
    # null handles / buffers
    s = pz.StepSummary()
    y = (C.c_double * 1)(0.0)
    assert L.pz_filter_step(None, y, 1, 1, C.byref(s)) == -1
    assert L.pz_filter_run(None, y, 1, 1, None) == -1
    assert L.pz_filter_read_particles(None, 1, None, 0) == -1
    assert L.pz_filter_read_ancestors(None, None, 0) == -1
    assert L.pz_filter_read_state(None, None, 0) == -1
    assert L.pz_filter_last_timing(None, None, None) == -1
    assert L.pz_filter_n_local(None) == 0
    assert L.pz_comm_unique_id(None) == -1
    L.pz_filter_destroy(None)   # a no-op, like free(NULL)
    lo, hi = C.c_int64(), C.c_int64()


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


def _header_layout(tmp_path):
    """sizeof / offsetof as a C compiler sees include/pzgpu.h (also proves the header is plain C99)."""
    import subprocess
    src = tmp_path / "layout.c"
    src.write_text(
        '#include <stddef.h>\n#include <stdio.h>\n#include "pzgpu.h"\n'
        "#define O(s, f) printf(#s \".\" #f \" %zu\\n\", offsetof(s, f))\n"
        "int main(void) {\n"
        '  printf("pz_model_desc %zu\\n", sizeof(pz_model_desc));\n'
        '  printf("pz_step_summary %zu\\n", sizeof(pz_step_summary));\n'
        '  printf("pz_mtt_params %zu\\n", sizeof(pz_mtt_params));\n'
        "  O(pz_model_desc, kind); O(pz_model_desc, nx); O(pz_model_desc, ny); O(pz_model_desc, scalar_ll_2x);\n"
        "  O(pz_model_desc, resampler); O(pz_model_desc, delayed); O(pz_model_desc, F); O(pz_model_desc, Q);\n"
        "  O(pz_model_desc, H); O(pz_model_desc, R); O(pz_model_desc, x0); O(pz_model_desc, mtt);\n"
        "  O(pz_step_summary, step); O(pz_step_summary, log_evidence_inc); O(pz_step_summary, ess);\n"
        "  O(pz_step_summary, mean); O(pz_step_summary, var); O(pz_step_summary, n_migrated);\n"
        "  O(pz_step_summary, e_max); O(pz_step_summary, total_weight);\n"
        "  return 0;\n}\n")
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", os.path.join(H.ROOT, "include"), str(src), "-o", str(exe)],
                   check=True, capture_output=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    return {k: int(v) for k, v in (line.split() for line in out.splitlines())}


def test_every_binding_agrees_with_the_compiled_header(pz, tmp_path):
    """The three bindings of the boundary (ctypes host mirror, the tests' oracle mirror, the Haskell
    FFI module) use the sizes and byte offsets a C compiler derives from include/pzgpu.h."""
    lay = _header_layout(tmp_path)
    for mirror in (pz, H):
        assert C.sizeof(mirror.ModelDesc) == lay["pz_model_desc"]
        assert C.sizeof(mirror.StepSummary) == lay["pz_step_summary"]
        for f in ("kind", "nx", "ny", "scalar_ll_2x", "resampler", "delayed", "F", "Q", "H", "R", "x0", "mtt"):
            assert getattr(mirror.ModelDesc, f).offset == lay[f"pz_model_desc.{f}"], f
        for f in ("step", "log_evidence_inc", "ess", "mean", "var", "n_migrated", "e_max", "total_weight"):
            assert getattr(mirror.StepSummary, f).offset == lay[f"pz_step_summary.{f}"], f
    hs = open(os.path.join(H.ROOT, "haskell", "PzGpu.hs")).read()
    assert int(re.search(r"sizeofModelDesc = (\d+)", hs).group(1)) == lay["pz_model_desc"]
    assert int(re.search(r"sizeofSummary = (\d+)", hs).group(1)) == lay["pz_step_summary"]
    assert f"pokeByteOff p {lay['pz_model_desc.delayed']} (1 :: Int32)" in hs
    for field, ty in (("step", "Int64"), ("log_evidence_inc", "CDouble"), ("ess", "CDouble")):
        assert re.search(rf"peekByteOff ps {lay['pz_step_summary.' + field]} :: IO {ty}", hs), field
    assert f"ps `plusPtr` {lay['pz_step_summary.mean']}" in hs and f"ps `plusPtr` {lay['pz_step_summary.var']}" in hs
    # every symbol the Haskell module imports is declared by the header
    hdr = open(os.path.join(H.ROOT, "include", "pzgpu.h")).read()
    for sym in re.findall(r'foreign import ccall \w+ "&?(pz_\w+)"', hs):
        assert re.search(rf"\b{sym}\(", hdr), f"PzGpu.hs imports {sym}, which include/pzgpu.h does not declare"
