"""CPU model of the DIRECT fp16-split convolution kernel (csrc/conv_gemm_tc_f16d.cu, conv_gemm_tc4h_kernel).

The kernel cannot run here; its index arithmetic can.  The batch is one tall image of n_img*H rows of W pixels;
pixel p = (img*H + ih)*W + iw owns slab row p - s0 of the tile that starts at slot s0; the A operand of filter tap
(kh, kw) for the 128 slots of sub-tile t is slab rows t*128 + kh*W + kw .. + 127 (a shifted window, read through the
128-byte swizzle: 16-byte chunk j of slab row r lives at chunk j ^ (r & 7)); slot s = (img*H + oh)*W + ow is a real
output iff oh < OH and ow < OW.  Checked against the definition of a valid stride-1 convolution, including tiles that
straddle images and the zero-filled tail."""
import numpy as np
import pytest

DT, HBK = 2, 64


def direct_conv(x, w):
    """x [n, H, W, 64] float64, w [KH, KW, 64, N] -> y [n, OH, OW, N] through the kernel's addressing."""
    n, H, W, C = x.shape
    KH, KW, _, N = w.shape
    assert C == HBK
    OH, OW = H - KH + 1, W - KW + 1
    npix = (DT * 128 + (KH - 1) * W + KW - 1 + 7) & ~7
    total_pix = n * H * W
    total_slots = ((n - 1) * H + OH - 1) * W + OW
    n_tiles = (total_slots + DT * 128 - 1) // (DT * 128)
    flat = x.reshape(total_pix, C)
    y = np.full((n, OH, OW, N), np.nan)
    wk = w.reshape(KH * KW, C, N)
    for tile in range(n_tiles):
        s0 = tile * DT * 128
        # fill: thread (pl, j) stores channels 8j..8j+7 of pixel s0 + pl at chunk j ^ (pl & 7) of slab row pl
        slab = np.zeros((npix, 8, 8))
        for pl in range(npix):
            gp = s0 + pl
            row = flat[gp] if gp < total_pix else np.zeros(C)
            for j in range(8):
                slab[pl, j ^ (pl & 7)] = row[8 * j:8 * j + 8]
        for t in range(DT):
            acc = np.zeros((128, N))
            for kb in range(KH * KW):
                kh, kw = divmod(kb, KW)
                first = t * 128 + kh * W + kw
                assert first + 128 <= npix, 'window leaves the slab'
                a = np.empty((128, C))
                for m in range(128):                      # the tensor core un-swizzles with the ABSOLUTE row
                    r = first + m
                    for j in range(8):
                        a[m, 8 * j:8 * j + 8] = slab[r, j ^ (r & 7)]
                acc += a @ wk[kb]
            for m in range(128):                          # epilogue: lane = slot
                slot = s0 + t * 128 + m
                img, rem = divmod(slot, H * W)
                oh, ow = divmod(rem, W)
                if img < n and oh < OH and ow < OW:
                    assert np.isnan(y[img, oh, ow, 0]), 'slot written twice'
                    y[img, oh, ow] = acc[m]
    return y


@pytest.mark.parametrize('n,H,W,KH,KW', [(3, 65, 17, 5, 4), (5, 9, 7, 3, 3), (2, 30, 7, 3, 3), (7, 6, 5, 2, 4), (1, 65, 20, 5, 4)])
def test_direct_addressing_is_a_valid_convolution(n, H, W, KH, KW):
    rng = np.random.default_rng(n * 1000 + H)
    x = rng.standard_normal((n, H, W, HBK))
    w = rng.standard_normal((KH, KW, HBK, 8))
    y = direct_conv(x, w)
    assert not np.isnan(y).any(), 'an output was never produced'
    OH, OW = H - KH + 1, W - KW + 1
    ref = np.zeros((n, OH, OW, 8))
    for kh in range(KH):
        for kw in range(KW):
            ref += np.einsum('nhwc,co->nhwo', x[:, kh:kh + OH, kw:kw + OW], w[kh, kw])
    np.testing.assert_allclose(y, ref, rtol=1e-12, atol=1e-12)


def test_useful_row_fraction_of_the_segmenter_layer():
    H, W, KH, KW = 65, 17, 5, 4
    assert abs(((H - KH + 1) / H) * ((W - KW + 1) / W) - 0.773) < 1e-3


def direct_conv_same3x3(x, w):
    """'same' 3x3 / stride-1 / padding-1 convolution through the kernel's DIN_PAD addressing: the tall image is the padded
    one (H + 1 rows of W + 1 pixels per image, row 0 and column 0 zero; the next row's zero column and the next image's
    zero row are this row's right and this image's bottom padding), the tap windows stay plain row shifts, slot
    s = (img * (H + 1) + oh) * (W + 1) + ow is a real output iff oh < H and ow < W."""
    n, H, W, C = x.shape
    N = w.shape[3]
    assert C == HBK and w.shape[:2] == (3, 3)
    Hp, Wp = H + 1, W + 1
    npix = (DT * 128 + 2 * Wp + 2 + 7) & ~7
    total_pix = n * Hp * Wp
    total_slots = ((n - 1) * Hp + H - 1) * Wp + W
    n_tiles = (total_slots + DT * 128 - 1) // (DT * 128)
    flat = x.reshape(n * H * W, C)
    y = np.full((n, H, W, N), np.nan)
    wk = w.reshape(9, C, N)
    for tile in range(n_tiles):
        s0 = tile * DT * 128
        slab = np.zeros((npix, 8, 8))
        for pl in range(npix):                                # the DIN_PAD fill
            gp = s0 + pl
            row = np.zeros(C)
            if gp < total_pix:
                img, rem = divmod(gp, Hp * Wp)
                r, cc = divmod(rem, Wp)
                if r >= 1 and cc >= 1:
                    row = flat[(img * H + r - 1) * W + cc - 1]
            for j in range(8):
                slab[pl, j ^ (pl & 7)] = row[8 * j:8 * j + 8]
        for t in range(DT):
            acc = np.zeros((128, N))
            for kb in range(9):
                kh, kw = divmod(kb, 3)
                first = t * 128 + kh * Wp + kw
                assert first + 128 <= npix, 'window leaves the slab'
                a = np.empty((128, C))
                for m in range(128):
                    r = first + m
                    for j in range(8):
                        a[m, 8 * j:8 * j + 8] = slab[r, j ^ (r & 7)]
                acc += a @ wk[kb]
            for m in range(128):
                slot = s0 + t * 128 + m
                img, rem = divmod(slot, Hp * Wp)
                oh, ow = divmod(rem, Wp)
                if img < n and oh < H and ow < W:
                    assert np.isnan(y[img, oh, ow, 0]), 'slot written twice'
                    y[img, oh, ow] = acc[m]
    return y


@pytest.mark.parametrize('n,H,W', [(3, 16, 36), (5, 8, 18), (2, 5, 3), (1, 1, 1), (4, 7, 40)])
def test_padded_direct_addressing_is_a_same_convolution(n, H, W):
    rng = np.random.default_rng(n * 100 + H)
    x = rng.standard_normal((n, H, W, HBK))
    w = rng.standard_normal((3, 3, HBK, 8))
    y = direct_conv_same3x3(x, w)
    assert not np.isnan(y).any(), 'an output was never produced'
    xp = np.zeros((n, H + 2, W + 2, HBK))
    xp[:, 1:-1, 1:-1] = x
    ref = np.zeros((n, H, W, 8))
    for kh in range(3):
        for kw in range(3):
            ref += np.einsum('nhwc,co->nhwo', xp[:, kh:kh + H, kw:kw + W], w[kh, kw])
    np.testing.assert_allclose(y, ref, rtol=1e-12, atol=1e-12)
    assert abs((16 / 17) * (36 / 37) - 0.916) < 1e-3          # useful slots on ResNet101's stage-3 maps
