"""CPU model of the addressing of the slab convolution kernel (csrc/conv_gemm_tc.cu, conv_gemm_tc3_kernel).

The kernel cannot run here, but its index arithmetic can be restated and checked against the definition of
a valid (un-padded, stride-1) convolution for arbitrary shapes, including tiles that straddle images:

  tile t          : output rows q in [t*R, t*R + R) of the global sequence q = img*OH + oh, R = 128 // OW
  slab            : input rows g(q0) .. of the NHWC tensor, g(q) = q + (q // OH) * (KH - 1)
  GEMM row r      : (q0 + r // OW, r % OW); slab pixel pix0 = (dq + (KH-1) * (img(q) - img(q0))) * W + ow
  k-block walk    : c0 += 32; at c0 == C: c0 = 0, pixel offset += 1; after KW taps: += W - KW
  swizzle         : 16-byte chunk j of slab pixel p lives at chunk (j & ~7) | ((j ^ key(p)) & 7)

It also counts shared-memory bank conflicts of the lane = GEMM-row reads for the current swizzle key (p & 7) and for
the row-wrap-free key proposed in DESIGN.md 4.1, which is what the ncu capture shows (34 % extra wavefronts)."""
import numpy as np
import pytest


def tile_plan(n_img, H, W, KH, KW):
    OH, OW = H - KH + 1, W - KW + 1
    R = 128 // OW
    Q = n_img * OH
    return OH, OW, R, Q


def emulate_tile(x, t, KH, KW, key):
    """x: [n_img, H, W, C] input.  Yields for tile t the im2col rows the kernel would feed to the MMA
    (shape [valid, KH*KW*C]) read THROUGH the swizzled slab, plus the output offsets."""
    n_img, H, W, C = x.shape
    OH, OW, R, Q = tile_plan(n_img, H, W, KH, KW)
    q0 = t * R
    nq = min(R, Q - q0)
    valid = nq * OW
    img0 = q0 // OH
    g0 = q0 + img0 * (KH - 1)
    img1 = (q0 + nq - 1) // OH
    rows = nq + (KH - 1) * (img1 - img0 + 1)
    flat = x.reshape(-1)
    cpp = C // 4
    # ---- fill: chunk q of the contiguous range -> swizzled slab chunk
    slab = np.full((rows * W * cpp, 4), np.nan, dtype=np.float32)
    first = g0 * W * C
    seen = set()
    for q in range(rows * W * cpp):
        p, j = divmod(q, cpp)
        phys = p * cpp + ((j & ~7) | ((j ^ key(p, W, OW)) & 7))
        assert phys not in seen
        seen.add(phys)
        src = first + q * 4
        slab[phys] = flat[src:src + 4] if src + 4 <= flat.size else 0.0
    # ---- reads: per GEMM row, walk the k-blocks as the producer does
    out = np.zeros((valid, KH * KW * C), dtype=np.float32)
    for r in range(valid):
        dq, ow = divmod(r, OW)
        pix0 = (dq + (KH - 1) * ((q0 + dq) // OH - img0)) * W + ow
        srow = dq + (KH - 1) * ((q0 + dq) // OH - img0)
        key0, is_key = ow + srow * OW, 0                      # the V2 kernel carries the key incrementally
        c0 = ss = poff = 0
        for kb in range(KH * KW * C // 32):
            p = pix0 + poff
            kx = key(p, W, OW) & 7
            if key is key_rowwrap_free:
                assert kx == (key0 + is_key) & 7
            for jj in range(8):
                phys = p * cpp + c0 // 4 + (jj ^ kx)
                out[r, kb * 32 + jj * 4:kb * 32 + jj * 4 + 4] = slab[phys]
            c0 += 32
            if c0 == C:
                c0 = 0
                poff += 1
                is_key += 1
                ss += 1
                if ss == KW:
                    ss = 0
                    poff += W - KW
                    is_key += OW - KW
    return q0 * OW, out


def key_current(p, W, OW):
    return p & 7


def key_rowwrap_free(p, W, OW):
    row, xx = divmod(p, W)
    return (xx + row * OW) & 7


@pytest.mark.parametrize('shape', [(3, 9, 8, 32, 3, 3), (2, 12, 17, 64, 5, 4), (5, 6, 7, 32, 3, 3), (4, 5, 5, 64, 3, 3), (1, 68, 9, 32, 4, 5)])
@pytest.mark.parametrize('key', [key_current, key_rowwrap_free])
def test_slab_tiles_reproduce_im2col(shape, key):
    n_img, H, W, C, KH, KW = shape
    rng = np.random.default_rng(0)
    x = rng.standard_normal((n_img, H, W, C)).astype(np.float32)
    OH, OW, R, Q = tile_plan(n_img, H, W, KH, KW)
    want = np.zeros((n_img * OH * OW, KH * KW * C), dtype=np.float32)        # the definition
    m = 0
    for n in range(n_img):
        for oh in range(OH):
            for ow in range(OW):
                want[m] = x[n, oh:oh + KH, ow:ow + KW, :].reshape(-1)
                m += 1
    covered = 0
    for t in range((Q + R - 1) // R):
        m0, rows = emulate_tile(x, t, KH, KW, key)
        assert m0 == covered
        assert np.array_equal(rows, want[m0:m0 + len(rows)])
        covered += len(rows)
    assert covered == len(want)


def _conflict_ratio(H, W, C, KH, KW, key, n_img=3):
    """extra wavefronts / ideal wavefronts of the LDS.128 reads (8 lanes per wavefront, 32 banks x 4 B)."""
    OH, OW, R, Q = tile_plan(n_img, H, W, KH, KW)
    cpp = C // 4
    ideal = extra = 0
    for t in range((Q + R - 1) // R):
        q0 = t * R
        nq = min(R, Q - q0)
        img0 = q0 // OH
        pix = []
        for r in range(128):
            dq, ow = divmod(r, OW)
            pix.append((dq + (KH - 1) * ((q0 + dq) // OH - img0)) * W + ow if r < nq * OW else 0)
        for tap_off in [kh * W + kw for kh in range(KH) for kw in range(KW)]:
            for jj in range(8):
                for w0 in range(0, 128, 8):                                  # one quarter-warp = one wavefront if conflict-free
                    banks = {}
                    for r in range(w0, w0 + 8):
                        p = pix[r] + tap_off
                        chunk = p * cpp + (jj ^ (key(p, W, OW) & 7))
                        banks.setdefault((chunk * 4) % 32, set()).add(chunk)
                    ways = max(len(v) for v in banks.values())
                    ideal += 1
                    extra += ways - 1
    return extra / ideal


def test_bank_conflicts_of_the_two_keys():
    cur = _conflict_ratio(65, 17, 64, 5, 4, key_current)          # the dominant layer (VAD 64->64 5x4 on 65x17)
    new = _conflict_ratio(65, 17, 64, 5, 4, key_rowwrap_free)
    print('extra LDS wavefronts per ideal wavefront: key p&7 = %.3f, row-wrap-free key = %.3f' % (cur, new))
    assert 0.15 < cur < 0.6             # ncu: 43.6 M conflicts / 126 M wavefronts for the whole kernel
    assert new <= 0.25 * cur
