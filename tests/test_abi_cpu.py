"""-m "not gpu": the C-ABI library loads, exports every symbol include/sfmloss.h declares, and
validates arguments (no compute without a GPU) -- plus host-side checks of the Python mirror."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def lib():
    import __graft_entry__ as ge
    ge.build()
    from sfm_learner_chainer_b200 import lib as L
    return L


def test_every_declared_symbol_is_exported(lib):
    hdr = open(os.path.join(ROOT, 'include', 'sfmloss.h')).read()
    hdr = re.sub(r'/\*.*?\*/', '', hdr, flags=re.S)
    declared = set(re.findall(r'\b(sfm_[a-z_0-9]+)\s*\(', hdr))
    assert declared == set(lib.SYMBOLS), declared ^ set(lib.SYMBOLS)
    so = lib.load()
    for name in declared:
        assert getattr(so, name) is not None
    assert so.sfm_version() == 102


def test_struct_layout_matches_header(lib):
    assert C.sizeof(lib.SfmDesc) == 48      # 10 x 4 bytes + raw_disp_scales + raw_pose_hw
    assert C.sizeof(lib.SfmInputs) == 8 * (3 + 4 + 1 + 4 + 2)
    assert C.sizeof(lib.SfmGrads) == 8 * 9
    assert C.sizeof(lib.SfmDebug) == 8 * 16


def test_workspace_and_validation(lib):
    so = lib.load()
    d = lib.SfmDesc(4, 2, 128, 416, 4, 0, 0.1, 0.0, 0.15, 0)
    n = so.sfm_workspace_bytes(C.byref(d))
    pyr = 4 * 3 * (70720 - 128 * 416) * 12      # planar fp32 pyramid of the scales >= 1 (scale 0 is never copied)
    assert pyr <= n <= pyr + 64 * 1024
    for bad, code in [(lib.SfmDesc(0, 2, 128, 416, 4, 0, 0, 0, 0, 0), 'B=0'),
                      (lib.SfmDesc(4, 9, 128, 416, 4, 0, 0, 0, 0, 0), 'S=9'),
                      (lib.SfmDesc(4, 2, 16, 416, 4, 0, 0, 0, 0, 0), '2x52'),
                      (lib.SfmDesc(4, 2, 128, 416, 5, 0, 0, 0, 0, 0), 'n_scales=5'),
                      (lib.SfmDesc(4, 2, 128, 416, 4, 2, 0, 0, 0, 0), 'B_global=2'),
                      (lib.SfmDesc(4, 2, 128, 416, 4, 0, 0, 0, 0, 0, 0x10, 0), 'raw_disp_scales'),
                      (lib.SfmDesc(4, 2, 128, 416, 4, 0, 0, 0, 0, 0, 0, 129), 'raw_pose_hw')]:
        assert so.sfm_workspace_bytes(C.byref(bad)) == 0
        assert code in so.sfm_last_error().decode()
    # argument errors come back as negative codes with a message, before any device work
    inp = lib.SfmInputs()
    rc = so.sfm_loss_forward(C.byref(d), C.byref(inp), C.c_void_p(16), None, C.c_void_p(256), None)
    assert rc == lib.SFM_E_NULL_POINTER and b'NULL' in so.sfm_last_error()
    rc = so.sfm_loss_forward(C.byref(d), C.byref(inp), None, None, None, None)
    assert rc == lib.SFM_E_NULL_POINTER
    assert so.sfm_warp_forward(0, 8, 8, None, None, None, None, None, None, None, None, None, None, None) == lib.SFM_E_INVALID_SHAPE
    assert so.sfm_sampler_interp_forward(1, 3, 8, 8, 0, 4, None, None, None, None) == lib.SFM_E_INVALID_SHAPE
    with pytest.raises(lib.SfmError):
        lib.check(so.sfm_warp_backward(1, 8, 8, None, None, None, None, None, None, None, None, None, None))


def test_comm_entry_points_without_a_device(lib):
    """sfm_comm_* / sfm_allreduce_partials: argument errors and the no-device error come back as codes;
    NCCL itself is bound at run time (the copy torch ships is found here)."""
    import torch  # noqa: F401  -- maps torch's bundled libnccl.so.2 into the process
    so = lib.load()
    assert so.sfm_comm_unique_id(None) == lib.SFM_E_NULL_POINTER
    assert so.sfm_allreduce_partials(None, None, 5, None) == lib.SFM_E_NULL_POINTER
    ident = C.create_string_buffer(lib.SFM_NCCL_UNIQUE_ID_BYTES)
    comm = C.c_void_p()
    assert so.sfm_comm_create(C.cast(ident, C.c_void_p), 2, 2, C.byref(comm)) == lib.SFM_E_INVALID_DESC
    if not torch.cuda.is_available():
        assert so.sfm_comm_create(C.cast(ident, C.c_void_p), 1, 0, C.byref(comm)) == lib.SFM_E_NO_DEVICE
        assert b'no CPU fallback' in so.sfm_last_error()
    v = so.sfm_nccl_version()
    assert v == 0 or v >= 21000, v
    assert so.sfm_comm_destroy(None) == 0


def test_no_cpu_fallback(lib):
    """Host arrays are rejected: the product path never computes on the CPU."""
    from sfm_learner_chainer_b200 import ViewSynthesisLoss, projective_inverse_warp, SpatialTransformerSamplerInterp
    from sfm_learner_chainer_b200.synthetic import make_snippets
    d = make_snippets(1, 2, 32, 104)
    op = ViewSynthesisLoss(0.1, 0.0, 0.15)
    with pytest.raises(TypeError, match='no CPU fallback'):
        op.forward_backward(d['tgt'], d['src'], d['intrinsics'], d['disps'], d['poses'], d['logits'])
    with pytest.raises(TypeError):
        projective_inverse_warp(d['src'][:, 0], np.ones((1, 32 * 104), np.float32), d['poses'][:, 0], d['intrinsics'][:, 0])
    with pytest.raises(RuntimeError, match='no CPU path'):
        SpatialTransformerSamplerInterp().forward_cpu((d['tgt'], d['tgt']))


def test_product_package_does_not_import_oracle():
    import subprocess, sys
    code = ("import sys; sys.path.insert(0, %r); import sfm_learner_chainer_b200, sfm_learner_chainer_b200.torch_adapter, "
            "sfm_learner_chainer_b200.chainer_adapter; "
            "assert not [m for m in sys.modules if m.startswith('oracle')], 'oracle imported by product'" % ROOT)
    subprocess.check_call([sys.executable, '-c', code])
    for dirpath, _, files in os.walk(os.path.join(ROOT, 'sfm_learner_chainer_b200')):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                src = open(os.path.join(dirpath, f)).read()
                assert 'import oracle' not in src and 'from oracle' not in src, f
