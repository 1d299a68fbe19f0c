"""The integer identities behind the episode kernel's cell-distance instructions (csrc/episode.cu, Warp::place), checked
exhaustively on the CPU with the instruction semantics written out.

The placement search of the reference (gpu/metal_location_search.rs:139-156) penalises a candidate site by distance / radius for
every plant inside the radius. Sites and plants sit on an integer grid, so the squared cell distance d2 = (si-gi)^2 + (sj-gj)^2
selects the factor. The kernel computes d2 as |s|^2 + |g|^2 - 2 s.g in one dot-product instruction per (site, plant):

  compact maps (at most 64 sites per axis)   IDP.4A.S8.S8   site bytes (-2sj, -2si, 1, 64) . plant bytes (gj, gi, q & 63, q >> 6) + |s|^2
  medium maps (at most 181 sites per axis)   IDP.2A.LO.S16.U8 + add   (-2sj, -2si as s16) . (gj, gi as u8) + |s|^2, plus q from the word's high half

with q = gi^2 + gj^2. These tests hold the packing to the plain formula; the GPU parity suites hold the kernel to the oracle.
"""
import numpy as np


def _s8(x):
    x = np.asarray(x, dtype=np.int64) & 0xFF
    return np.where(x >= 128, x - 256, x)


def _s16(x):
    x = np.asarray(x, dtype=np.int64) & 0xFFFF
    return np.where(x >= 32768, x - 65536, x)


def dp4a_s8s8(a, b, c):
    """dp4a.s32.s32: four signed byte products accumulated onto c."""
    a, b = np.asarray(a, dtype=np.int64), np.asarray(b, dtype=np.int64)
    return sum(_s8(a >> (8 * i)) * _s8(b >> (8 * i)) for i in range(4)) + c


def dp2a_lo_s16u8(a, b, c):
    """dp2a.lo.s32.u32: the two signed 16-bit halves of a times the two low unsigned bytes of b, accumulated onto c."""
    a, b = np.asarray(a, dtype=np.int64), np.asarray(b, dtype=np.int64)
    return _s16(a) * (b & 0xFF) + _s16(a >> 16) * ((b >> 8) & 0xFF) + c


def test_compact_form_one_dp4a_per_site_and_plant():
    n = 64
    si, sj, gi, gj = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), np.arange(n), indexing="ij")
    packed = (si << 8) | sj                                       # the site word of the walk list
    sa = (((0x8080 - 2 * packed) & 0xFFFFFFFF) ^ 0x8080) | 0x40010000   # bytes (-2sj, -2si, 1, 64)
    assert np.array_equal(_s8(sa), -2 * sj) and np.array_equal(_s8(sa >> 8), -2 * si)
    sq = dp4a_s8s8(packed, packed, 0)                             # |s|^2 by the same instruction
    assert np.array_equal(sq, si * si + sj * sj)
    q = gi * gi + gj * gj
    assert (q >> 6).max() <= 127                                  # fits a signed byte
    word = (gi << 8) | gj | ((q & 63) << 16) | ((q >> 6) << 24)   # the plant word written by add_generator
    d2 = dp4a_s8s8(sa, word, sq)
    assert np.array_equal(d2, (si - gi) ** 2 + (sj - gj) ** 2)
    # the word that pads a plant list to a multiple of four is out of range of every site (tables of at most 2048 entries)
    sentinel = (63 << 16) | (127 << 24)
    assert (dp4a_s8s8(sa[:, :, 0, 0], sentinel, sq[:, :, 0, 0]) >= 8191).all()


def test_medium_form_dp2a_plus_norm():
    n = 182
    rng = np.random.default_rng(5)
    si, sj, gi, gj = (rng.integers(0, n, 400_000) for _ in range(4))
    # corners and edges explicitly
    edge = np.array([0, 1, 63, 64, 127, 128, 180, 181])
    e = np.array(np.meshgrid(edge, edge, edge, edge, indexing="ij")).reshape(4, -1)
    si, sj, gi, gj = (np.concatenate([a, b]) for a, b in zip((si, sj, gi, gj), e))
    sa = ((-2 * sj) & 0xFFFF) | (((-2 * si) & 0xFFFF) << 16)
    sq = si * si + sj * sj
    q = gi * gi + gj * gj
    assert q.max() < 65536                                        # 181^2 * 2 = 65,522: the bound of the medium form
    word = (gi << 8) | gj | (q << 16)
    d2 = dp2a_lo_s16u8(sa, word, sq) + (word >> 16)
    assert np.array_equal(d2, (si - gi) ** 2 + (sj - gj) ** 2)
    assert (dp2a_lo_s16u8(sa, 0xFFFF0000, sq) + 0xFFFF >= 65535).all()   # padding word: out of range of every site


def test_lookup_without_a_range_test():
    # an out-of-range plant reads the 1.0 stored after the class's factors: the clamp of the index replaces the comparison,
    # and multiplying by 1.0 leaves every double unchanged bit for bit
    r2lim = 49
    table = np.concatenate([np.sqrt(np.arange(r2lim)) * 1000.0 / 7000.0, [1.0]])
    d2 = np.arange(0, 6000)
    f = table[np.minimum(d2, r2lim)]
    assert np.array_equal(f[:r2lim], table[:r2lim]) and (f[r2lim:] == 1.0).all()
    x = np.random.default_rng(1).random(1000) * np.logspace(-300, 300, 1000)
    assert np.array_equal((x * 1.0).view(np.uint64), x.view(np.uint64))
    tiny = np.array([5e-324, 2.2250738585072014e-308, 0.0, -0.0])
    assert np.array_equal((tiny * 1.0).view(np.uint64), tiny.view(np.uint64))
