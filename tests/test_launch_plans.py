"""Host-side launch planning (b200dn_igemm_plan: no GPU touched): every layer shape of the three network widths, at the
BASELINE batch sizes and at batch 1-2, plus a sweep of odd shapes, must get a configuration that fits the kernels'
shared-memory and TMEM budgets.  The dispatch rules DESIGN.md states are pinned for the headline layers."""
import ctypes as C
import itertools

import pytest

from vub_image_denoising_b200 import _lib

SMS = 148
DUMMY = 0x1000   # plan-only calls check pointers for NULL, nothing more


def plan(mode, B, H, W, cin, cout, prec=_lib.PREC_BF16, in_ctot=None, out_ctot=None, coff=0, nchw=False, impl=0,
         block_n=0, m_tiles=0, max_ctas=0, sms=SMS):
    a = _lib.IgemmArgs()
    a.mode, a.prec, a.B, a.H, a.W, a.cin, a.cout = mode, prec, B, H, W, cin, cout
    a.in_[0] = DUMMY
    a.in_[1] = DUMMY if prec in _lib.TWO_PLANE_PRECS else None
    a.in_ctot = in_ctot or ((cin + 7) // 8 * 8)
    a.wpacked, a.bias, a.slope = DUMMY, DUMMY, DUMMY
    if nchw:
        a.out_kind = _lib.OUT_NCHW32
        a.out_nchw, a.res_nchw, a.res_bmod = DUMMY, DUMMY, B
    else:
        a.out_kind = _lib.OUT_NHWC16
        a.out[0] = DUMMY
        a.out[1] = DUMMY if prec in _lib.TWO_PLANE_PRECS else None
        a.out_ctot, a.out_coff = out_ctot or (coff + cout + 7) // 8 * 8, coff
    a.impl, a.block_n, a.m_tiles, a.max_ctas = impl, block_n, m_tiles, max_ctas
    info = _lib.IgemmPlanInfo()
    rc = _lib.lib().b200dn_igemm_plan(C.byref(a), sms, C.byref(info))
    _lib.check(rc, "igemm_plan")
    return info


def check_budgets(i, what):
    assert i.data_bytes_used <= i.data_bytes_budget, f"{what}: {i.data_bytes_used} B of rings > {i.data_bytes_budget}"
    assert 32 <= i.tmem_cols <= 512 and i.tmem_cols & (i.tmem_cols - 1) == 0, f"{what}: tmem_cols {i.tmem_cols}"
    assert i.tmem_cols >= 2 * i.mt * i.block_n, f"{what}: two accumulator stages do not fit {i.tmem_cols} columns"
    assert i.block_n % 16 == 0 and 16 <= i.block_n <= 256, what
    assert i.mt in (1, 2) and i.num_tiles >= 1 and 1 <= i.grid <= SMS, f"{what}: grid {i.grid}"
    if i.kernel == 2:
        assert i.grid % 2 == 0 and not i.wres and i.block_n >= 32, what
    if i.kernel in (1, 2):
        assert i.num_slabs >= 2 or (i.wres and i.num_slabs >= 1), f"{what}: {i.num_slabs} slab(s)"
        assert i.num_slabs <= 6 and i.slab_bytes % 1024 == 0, what
        if not i.wres:
            assert 2 <= i.num_stages <= 8 and i.stage_bytes % 1024 == 0, f"{what}: W ring {i.num_stages} x {i.stage_bytes}"
    else:
        assert 2 <= i.num_stages <= 8, f"{what}: ring of {i.num_stages}"


def network_layers(F, B, S=256):
    """(mode, B, H, W, cin, cout, in_ctot, out_ctot, coff, nchw) of the 68 tensor-core launches (rdunet.ForwardPlan)."""
    L = []
    ch = [F << l for l in range(4)]
    hs = [S >> l for l in range(4)]

    def dense(l, dst_ctot):
        Cc, g = ch[l], ch[l] // 2
        for k in range(3):
            L.append((0, B, hs[l], hs[l], Cc + k * g, g, Cc * 5 // 2, Cc * 5 // 2, Cc + k * g, False))
        L.append((0, B, hs[l], hs[l], Cc + 3 * g, Cc, Cc * 5 // 2, dst_ctot, 0, False))

    L.append((0, B, S, S, F, F, F, F * 5 // 2, 0, False))
    for l in range(4):
        dense(l, ch[l] * 5 // 2)
        dense(l, ch[l] * 3 if l < 3 else ch[l] * 5 // 2)
        if l < 3:
            L.append((1, B, hs[l], hs[l], ch[l], 2 * ch[l], 3 * ch[l], ch[l + 1] * 5 // 2, 0, False))
    for l in (2, 1, 0):
        L.append((2, B, hs[l + 1], hs[l + 1], ch[l + 1], ch[l + 1], ch[l + 1] * 5 // 2, 3 * ch[l], ch[l], False))
        L.append((0, B, hs[l], hs[l], 3 * ch[l], ch[l], 3 * ch[l], ch[l] * 5 // 2, 0, False))
        dense(l, ch[l] * 5 // 2)
        dense(l, ch[l] * 5 // 2)
    L.append((0, B, S, S, F, F, F * 5 // 2, F, 0, False))
    L.append((0, B, S, S, F, 3, F, 0, 0, True))
    return L


@pytest.mark.parametrize("F,B", [(128, 64), (128, 1), (128, 2), (64, 64), (64, 1), (32, 32), (32, 2), (32, 4), (16, 3), (48, 5)])
def test_every_network_layer_fits(F, B, built_lib):
    layers = network_layers(F, B)
    assert len(layers) == 68
    for prec in (_lib.PREC_BF16, _lib.PREC_FP16X2, _lib.PREC_BF16X3):
        for (mode, b, h, w, cin, cout, ictot, octot, coff, nchw) in layers:
            i = plan(mode, b, h, w, cin, cout, prec=prec, in_ctot=ictot, out_ctot=octot, coff=coff, nchw=nchw)
            check_budgets(i, f"F={F} B={B} prec={prec} mode={mode} {h}x{w} {cin}->{cout}")


def test_shape_sweep_fits(built_lib):
    n = 0
    for cin, cout, hw, B, impl in itertools.product((8, 16, 24, 48, 80, 96, 160, 200, 320, 640, 1000, 2560),
                                                    (8, 16, 24, 40, 64, 72, 128, 200, 256, 512, 1024),
                                                    ((8, 8), (16, 24), (40, 56), (128, 128), (256, 256)), (1, 3, 64), (0, 1, 2, 3)):
        i = plan(0, B, hw[0], hw[1], cin, cout, impl=impl)
        check_budgets(i, f"conv3x3 impl={impl} B={B} {hw} {cin}->{cout}")
        n += 1
    for cin, hw, B in itertools.product((16, 64, 208, 512, 1024), ((8, 8), (24, 40), (128, 128)), (1, 5, 64)):
        check_budgets(plan(1, B, hw[0], hw[1], cin, 2 * cin), f"down {cin}")
        check_budgets(plan(2, B, hw[0], hw[1], cin, cin, out_ctot=cin // 2 + cin, coff=cin // 2), f"up {cin}")
        n += 2
    assert n > 7000


def test_dispatch_rules_of_the_headline_layers(built_lib):
    # BASELINE config 2 (F = 128, B = 64): growth convs (N = 64) and conv_3 (N = 128) at level 0 run on CTA pairs with
    # two sub-tiles per CTA and a filter row per W stage for N = 64; level-3 convs (N = 256 tiles) with one sub-tile
    i = plan(0, 64, 256, 256, 128, 64, in_ctot=320, out_ctot=320, coff=128)
    assert (i.kernel, i.mt, i.block_n, i.w_taps, i.grid) == (2, 2, 64, 3, 148)
    i = plan(0, 64, 256, 256, 320, 128, in_ctot=320, out_ctot=320)
    assert (i.kernel, i.mt, i.block_n, i.w_taps) == (2, 2, 128, 1)
    i = plan(0, 64, 32, 32, 2560, 1024, in_ctot=2560, out_ctot=2560)
    assert (i.kernel, i.mt, i.block_n, i.num_n_tiles) == (2, 1, 256, 4)
    # the F = 32 network's level-0 growth convs keep their weights resident; the output conv pads N 3 -> 16
    i = plan(0, 32, 256, 256, 32, 16, prec=_lib.PREC_FP16, in_ctot=80, out_ctot=80, coff=32)
    assert (i.kernel, i.wres, i.mt, i.block_n) == (1, 1, 2, 16) and i.num_slabs >= 3
    i = plan(0, 64, 256, 256, 128, 3, in_ctot=128, nchw=True)
    assert (i.kernel, i.wres, i.block_n) == (1, 1, 16)
    # batch 1: the level-3 conv splits N down to 64 and still uses CTA pairs (under-filled grid)
    i = plan(0, 1, 32, 32, 2560, 1024, in_ctot=2560, out_ctot=2560)
    assert (i.kernel, i.block_n, i.num_n_tiles) == (2, 64, 16) and i.grid == 2 * 4 * 16
    # explicit implementations are honoured; a layer with one pixel tile cannot pair
    assert plan(0, 2, 32, 32, 128, 64, impl=1).kernel == 0
    assert plan(0, 2, 32, 32, 128, 64, impl=2).kernel == 1
    assert plan(0, 2, 32, 32, 128, 64, impl=3).kernel == 2
    assert plan(0, 1, 16, 8, 128, 64).kernel == 1
    # down / up convs run on the per-tap kernel
    assert plan(1, 64, 256, 256, 128, 256, in_ctot=384).kernel == 0
    assert plan(2, 64, 128, 128, 256, 256, in_ctot=640, out_ctot=384, coff=128).kernel == 0


def test_plan_rejects_what_launch_rejects(built_lib):
    with pytest.raises(RuntimeError, match="in_ctot"):
        plan(0, 1, 8, 8, 16, 16, in_ctot=20)
    with pytest.raises(RuntimeError, match="even H, W"):
        plan(1, 1, 9, 8, 16, 32)
    with pytest.raises(RuntimeError, match="8-channel aligned"):
        plan(0, 1, 8, 8, 16, 12, out_ctot=16)
    with pytest.raises(RuntimeError, match="sm_count"):
        plan(0, 1, 8, 8, 16, 16, sms=0)
