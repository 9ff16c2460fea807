"""Parity of the tcgen05/TMEM/TMA implicit-GEMM convolution (csrc/conv_tc.cu) through the C ABI
entry dbs_conv_tc, against torch's fp32 CPU convolution on the same bf16-rounded operands.

Floating point: products of bf16 values are exact in fp32, so only the accumulation order and the
final bf16 rounding differ: tolerance = 1 bf16 ulp of the value (2^-8 relative) + 2e-3 absolute.
GroupNorm partial sums (fp32 atomics) are compared at 2e-3 relative."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _bf(x):
    return x.to(torch.bfloat16)


def _run(kind, ksize, B, H, W, C0, C1, Cout, seed=0, stats=True):
    from dbs_b200 import abi, runtime
    torch.manual_seed(seed)
    Cin = C0 + C1
    x0 = _bf(torch.randn(B, H, W, C0))
    x1 = _bf(torch.randn(B, H, W, C1)) if C1 else None
    taps = 16 if kind == 2 else ksize * ksize
    w = _bf(torch.randn(taps, Cout, Cin) / (taps * Cin) ** 0.5)
    bias = torch.randn(Cout)
    # torch reference on the rounded values, NCHW fp32
    xin = torch.cat([x0] + ([x1] if C1 else []), dim=-1).float().permute(0, 3, 1, 2)
    if kind == 2:
        wt = w.float().reshape(4, 4, Cout, Cin).permute(3, 2, 0, 1).contiguous()       # [ci][co][kh][kw]
        ref = F.conv_transpose2d(xin, wt, bias, stride=2, padding=1)
    else:
        wt = w.float().reshape(ksize, ksize, Cout, Cin).permute(2, 3, 0, 1).contiguous()  # [co][ci][kh][kw]
        ref = F.conv2d(xin, wt, bias, stride=2 if kind == 1 else 1, padding=ksize // 2)
    Ho, Wo = ref.shape[-2:]
    out = torch.zeros(B, Ho, Wo, Cout, dtype=torch.bfloat16, device="cuda")
    G = min(8, Cout)
    st = torch.zeros(B, G, 2, device="cuda") if stats else None
    d0, d1, dw, db = x0.cuda(), (x1.cuda() if C1 else None), w.cuda(), bias.cuda()
    abi.check(abi.lib.dbs_conv_tc(d0.data_ptr(), C0, runtime.ptr(d1), C1, dw.data_ptr(), db.data_ptr(), out.data_ptr(),
                                  runtime.ptr(st), B, H, W, Cout, kind, ksize, runtime.current_stream()), "conv_tc")
    torch.cuda.synchronize()
    got = out.float().cpu().permute(0, 3, 1, 2)
    tol = 2e-3 + np.abs(ref.numpy()) * 2.0 ** -8
    assert (np.abs(got.numpy() - ref.numpy()) <= tol).all(), float(np.abs(got.numpy() - ref.numpy()).max())
    if stats:
        r = ref.reshape(B, G, -1)
        want = torch.stack([r.sum(-1), (r * r).sum(-1)], dim=-1)
        np.testing.assert_allclose(st.cpu().numpy(), want.numpy(), rtol=2e-3, atol=2e-2)


# (kind, ksize, B, H, W, C0, C1, Cout): every layer shape of the default net at 64x64 / 32x32 plus edge cases
CASES = [
    (0, 3, 2, 64, 64, 32, 0, 64),     # encoder.0.conv1           KC=32 (64B swizzle)
    (0, 3, 2, 64, 64, 64, 0, 64),     # encoder.0.conv2           KC=64 (128B swizzle)
    (0, 1, 2, 64, 64, 32, 0, 64),     # encoder.0.residual 1x1
    (1, 3, 2, 64, 64, 64, 0, 64),     # downsample.0 (space-to-depth view)
    (0, 3, 3, 32, 32, 64, 0, 128),    # encoder.1.conv1
    (0, 3, 3, 16, 16, 256, 0, 256),   # encoder.2.conv2
    (1, 3, 3, 16, 16, 256, 0, 256),   # downsample.2
    (0, 3, 5, 8, 8, 256, 0, 256),     # middle: two images per tile, batch not a multiple of Nt
    (2, 4, 3, 8, 8, 256, 0, 256),     # upsample.0 (four phase convs)
    (0, 3, 2, 16, 16, 256, 256, 128), # decoder.0.conv1 over the skip concat
    (0, 1, 2, 16, 16, 256, 256, 128), # decoder.0.residual 1x1 over the concat
    (2, 4, 2, 32, 32, 64, 0, 64),     # upsample.2
    (0, 3, 2, 64, 64, 64, 64, 32),    # decoder.2.conv1           Nb=32
    (0, 3, 2, 64, 64, 32, 0, 32),     # decoder.2.conv2
    (0, 3, 7, 4, 4, 128, 0, 128),     # 32x32 lattices: 4x4 middle, 8 images per tile
    (0, 3, 9, 2, 2, 64, 0, 64),       # 2x2: 32 images per tile, per-thread stats path
    (0, 3, 3, 16, 16, 16, 16, 16),    # demo net widths: KC=16 (32B swizzle), Nb=16
    (0, 3, 1, 128, 128, 32, 0, 32),   # API maximum lattice (W = tile width)
    (0, 3, 2, 16, 32, 64, 0, 64),     # non-square
    (0, 3, 300, 8, 8, 64, 0, 64),     # more tiles than SMs: persistent loop + TMEM double buffering
    # strip kernel (conv_strip.cu) specifics
    (0, 3, 3, 16, 16, 256, 0, 64),    # decoder.1.conv1-like: weights too large to stay resident (streamed slabs)
    (0, 3, 2, 16, 16, 128, 128, 64),  # decoder.1.conv1 over the concat, streamed
    (0, 3, 5, 16, 8, 64, 0, 32),      # W = 8: 16 rows per tile
    (0, 3, 3, 64, 64, 64, 0, 16),     # Nb = 16, rows span two warps (cross-warp neighbour exchange)
    (2, 4, 2, 64, 64, 32, 0, 32),     # transposed-conv phases with W = 64
    (0, 3, 2, 128, 128, 64, 0, 64),   # one image row per tile, rows span four warps
    (0, 3, 150, 32, 32, 32, 0, 32),   # 1200 tiles: persistent loop, stats double buffering
    (0, 1, 3, 32, 32, 128, 0, 32),    # decoder.2.residual 1x1
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: "k%d_%dx%d_B%d_%dx%d_c%d+%d_o%d" % (c[0], c[1], c[1], c[2], c[3], c[4], c[5], c[6], c[7]))
def test_conv_tc_vs_torch(case):
    _run(*case)


@pytest.mark.parametrize("B,H,W,C", [(3, 8, 8, 256), (5, 4, 4, 256), (40, 4, 4, 128), (2, 8, 16, 128), (3, 8, 8, 512)])
def test_conv_tc_transposed_phases_merged(B, H, W, C):
    """ConvTranspose2d(4, 2, 1) below the strip kernel's resolutions, no statistics: the tile kernel runs the four
    output phases in ONE persistent launch (ConvGeom::n_phases = 4); B = 40 at 4x4 gives more (tile, phase) units
    than a wave's worth per phase and a batch that is not a multiple of the images per tile."""
    _run(2, 4, B, H, W, C, 0, C, stats=False)


@pytest.mark.parametrize("B,H,W,C", [(2, 32, 32, 64), (3, 16, 16, 64), (2, 16, 16, 128), (2, 64, 64, 32), (1, 128, 128, 32),
                                     (150, 16, 16, 128), (3, 16, 8, 32)])
def test_conv_tc_transposed_phase_pairs_strip(B, H, W, C):
    """ConvTranspose2d(4, 2, 1) at the strip kernel's resolutions, no statistics: two launches, each the pair of
    horizontal phases of one vertical phase (ConvGeom::n_phases = 2: four tap slabs per MMA, two outputs per tile).
    Covers the cross-warp neighbour exchange (W >= 64), N-tiled layers in pass mode (128 couts, B = 150) and Nb = 32."""
    _run(2, 4, B, H, W, C, 0, C, stats=False)


def test_conv_tc_wide_cout_two_n_tiles():
    _run(0, 3, 2, 8, 8, 64, 0, 512, stats=True)


def test_conv_tc_rejects_ineligible_shapes():
    from dbs_b200 import abi, runtime
    x = torch.zeros(1, 12, 12, 32, dtype=torch.bfloat16, device="cuda")
    w = torch.zeros(9, 32, 32, dtype=torch.bfloat16, device="cuda")
    b = torch.zeros(32, device="cuda")
    o = torch.zeros(1, 12, 12, 32, dtype=torch.bfloat16, device="cuda")
    rc = abi.lib.dbs_conv_tc(x.data_ptr(), 32, 0, 0, w.data_ptr(), b.data_ptr(), o.data_ptr(), 0, 1, 12, 12, 32, 0, 3,
                             runtime.current_stream())
    assert rc == abi.ERR_UNSUPPORTED and "not eligible" in abi.last_error()


@pytest.mark.parametrize("B,H,W,C0,C1,Cout", [(3, 32, 32, 32, 0, 64), (2, 64, 64, 32, 0, 64), (2, 32, 32, 64, 64, 32),
                                              (2, 16, 16, 128, 128, 64), (3, 16, 16, 64, 0, 128), (2, 16, 32, 64, 0, 32)])
def test_fused_conv3x3_plus_residual_1x1(B, H, W, C0, C1, Cout):
    """One strip-kernel pass = ConvBlock.conv1 (+GN sums) and ConvBlock.residual on the same input."""
    from dbs_b200 import abi, runtime
    torch.manual_seed(1)
    Cin = C0 + C1
    x0 = _bf(torch.randn(B, H, W, C0))
    x1 = _bf(torch.randn(B, H, W, C1)) if C1 else None
    w3 = _bf(torch.randn(9, Cout, Cin) / (9 * Cin) ** 0.5)
    w1 = _bf(torch.randn(1, Cout, Cin) / Cin ** 0.5)
    b3, b1 = torch.randn(Cout), torch.randn(Cout)
    xin = torch.cat([x0] + ([x1] if C1 else []), dim=-1).float().permute(0, 3, 1, 2)
    ref3 = F.conv2d(xin, w3.float().reshape(3, 3, Cout, Cin).permute(2, 3, 0, 1).contiguous(), b3, padding=1)
    ref1 = F.conv2d(xin, w1.float().reshape(1, 1, Cout, Cin).permute(2, 3, 0, 1).contiguous(), b1)
    o3 = torch.zeros(B, H, W, Cout, dtype=torch.bfloat16, device="cuda")
    o1 = torch.zeros_like(o3)
    st = torch.zeros(B, 8, 2, device="cuda")
    d0, d1 = x0.cuda(), (x1.cuda() if C1 else None)
    dw3, dw1, db3, db1 = w3.cuda(), w1.cuda(), b3.cuda(), b1.cuda()
    abi.check(abi.lib.dbs_conv_block_tc(d0.data_ptr(), C0, runtime.ptr(d1), C1, dw3.data_ptr(), db3.data_ptr(),
                                        dw1.data_ptr(), db1.data_ptr(), o3.data_ptr(), o1.data_ptr(), st.data_ptr(), B, H, W,
                                        Cout, runtime.current_stream()), "conv_block_tc")
    torch.cuda.synchronize()
    for got, ref in ((o3, ref3), (o1, ref1)):
        g_ = got.float().cpu().permute(0, 3, 1, 2).numpy()
        tol = 2e-3 + np.abs(ref.numpy()) * 2.0 ** -8
        assert (np.abs(g_ - ref.numpy()) <= tol).all(), float(np.abs(g_ - ref.numpy()).max())
    r = ref3.reshape(B, 8, -1)
    want = torch.stack([r.sum(-1), (r * r).sum(-1)], dim=-1)
    np.testing.assert_allclose(st.cpu().numpy(), want.numpy(), rtol=2e-3, atol=2e-2)


def test_conv_tc_repeated_runs_are_identical_and_correct():
    """Hazard check for the persistent pipelines (pass-resident weights reloaded behind the `weights free` barrier,
    merged / paired transposed-conv phases, TMEM double buffering): DBS_CONV_STRESS (default 3) fresh random problems
    per regime, each launched twice (without and, where defined, with GroupNorm statistics) and checked against torch;
    run with DBS_CONV_STRESS=12 after touching the barriers."""
    import os
    reps = int(os.environ.get("DBS_CONV_STRESS", "3"))
    regimes = [(0, 3, 300, 16, 16, 128, 0, 128), (0, 3, 300, 16, 16, 64, 0, 128), (2, 4, 150, 16, 16, 128, 0, 128),
               (2, 4, 300, 8, 8, 128, 0, 128), (2, 4, 160, 32, 32, 64, 0, 64), (0, 3, 160, 32, 32, 64, 64, 32)]
    for r in range(reps):
        for case in regimes:
            _run(*case, seed=1000 + r, stats=False)
            _run(*case, seed=1000 + r, stats=case[0] != 2)
