"""TEST INFRASTRUCTURE (oracle) — never imported by the product path (stim_b200/).

numpy restatement of the two random primitives the device noise generator is specified with
(DESIGN.md "RNG addressing"):

  * Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11).
    Pinned against the Random123 known-answer vectors in tests/test_oracle_philox.py.
  * exp_draw: Exp(1) variate from a uniform u32 using only IEEE-754 double + - * / in a fixed order,
    so the CUDA kernel (stim_b200/csrc/kernels.cu: exp_draw) reproduces it bit for bit.

These replace, in distribution, the reference's std::mt19937_64 + std::geometric_distribution
(/root/reference/src/stim/util_bot/probability_util.cc:23-43): floor(Exp(1)/lambda) with
lambda = -log1p(-p) is exactly Geometric(p).
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK32 = np.uint64(0xFFFFFFFF)

TAG_COLLAPSE = 0x434F4C4C


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """All arguments broadcastable integer arrays (values < 2**32). Returns 4 uint32 arrays."""
    if all(isinstance(a, (int, np.integer)) for a in (c0, c1, c2, c3)):
        return _philox_scalar(int(c0), int(c1), int(c2), int(c3), int(k0) & 0xFFFFFFFF, int(k1) & 0xFFFFFFFF)
    c0, c1, c2, c3 = np.broadcast_arrays(*(np.asarray(a, dtype=np.uint64) for a in (c0, c1, c2, c3)))
    c0 = c0.copy()
    c1 = c1.copy()
    c2 = c2.copy()
    c3 = c3.copy()
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK32
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK32
        n0 = hi1 ^ c1 ^ np.uint64(k0)
        n2 = hi0 ^ c3 ^ np.uint64(k1)
        c0, c1, c2, c3 = n0, lo1, n2, lo0
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return tuple(a.astype(np.uint32) for a in (c0, c1, c2, c3))


def _philox_scalar(c0, c1, c2, c3, k0, k1):
    """The same ten rounds on Python integers (one counter): the serial event walks of the oracles call this ~10^6 times."""
    for _ in range(10):
        p0 = 0xD2511F53 * c0
        p1 = 0xCD9E8D57 * c2
        c0, c1, c2, c3 = (p1 >> 32) ^ c1 ^ k0, p1 & 0xFFFFFFFF, (p0 >> 32) ^ c3 ^ k1, p0 & 0xFFFFFFFF
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return np.uint32(c0), np.uint32(c1), np.uint32(c2), np.uint32(c3)


_COEFS = [1.0 / 19.0, 1.0 / 17.0, 1.0 / 15.0, 1.0 / 13.0, 1.0 / 11.0, 1.0 / 9.0, 1.0 / 7.0, 1.0 / 5.0, 1.0 / 3.0, 1.0]
_LN2 = 0.6931471805599453
_SQRT2 = 1.4142135623730951


def exp_draw(r):
    """-ln((r + 1/2) / 2**32) for uint32 r, evaluated exactly like the device code."""
    r = np.asarray(r, dtype=np.uint64)
    v = np.uint64(2) * r + np.uint64(1)  # odd, < 2**33, exact in float64
    vf = v.astype(np.float64)
    # t = floor(log2 v): frexp gives vf = mant * 2**e with mant in [0.5, 1) -> t = e - 1 (exact for integers < 2**53)
    _, e = np.frexp(vf)
    t = e.astype(np.int64) - 1
    m = vf * np.ldexp(1.0, -t)  # exact scaling into [1, 2)
    big = m > _SQRT2
    m = np.where(big, m * 0.5, m)
    t = np.where(big, t + 1, t)
    s = (m - 1.0) / (m + 1.0)
    s2 = s * s
    poly = np.full_like(s, 1.0 / 21.0)
    for c in _COEFS:
        poly = poly * s2 + c
    lnm = (2.0 * s) * poly
    lnx = lnm + (t - 33).astype(np.float64) * _LN2
    return -lnx


# ---------------------------------------------------------------------------------------------
# Fixed-point exponential clock (spec v2, DESIGN.md "RNG addressing"): all integer arithmetic, so the
# device (stim_b200/csrc/kernels.cu: exp_draw_fx) and this oracle agree bit for bit by construction.
# Unit = 2**-56 nat.
# ---------------------------------------------------------------------------------------------
from .log2_table import LN2_Q24, LOG2_T  # noqa: E402

FX_SHIFT = 56
LAM_MAX = 1 << 62
REM_SAT = 1 << 63


def exp_draw_fx(r: int) -> int:
    """-ln((r + 1/2) / 2**32) in units of 2**-56, via a 256-entry log2 table with linear interpolation."""
    v = 2 * int(r) + 1
    t = v.bit_length() - 1
    vn = v << (32 - t)
    frac = vn & 0xFFFFFFFF
    i, f = frac >> 24, frac & 0xFFFFFF
    log2m = LOG2_T[i] + (((LOG2_T[i + 1] - LOG2_T[i]) * f) >> 24)
    return ((33 << 32) - ((t << 32) + log2m)) * LN2_Q24


def lam_fx(p: float) -> int:
    """Per-shot event rate of probability p (narrowed to float32 like the reference) in clock units."""
    import math

    f = float(np.float32(p))
    if not f > 0:
        return 0
    if f >= 1:
        return LAM_MAX
    return min(int(math.ldexp(-math.log1p(-f), FX_SHIFT)), LAM_MAX)


# ---------------------------------------------------------------------------------------------
# 32-bit gap arithmetic of the detector-error-model sampler (stim_b200/csrc/dem.cu, header comment): all integer, so the
# device and this restatement agree bit for bit by construction.
# ---------------------------------------------------------------------------------------------
from .log2_table import LN2_Q32, LOG2_T26  # noqa: E402


def exp_draw_q26(r: int) -> int:
    """-ln(v / 2**32), v = r | 1, in units of 2**-26 nat: 256-entry log2 table (Q26), 13-bit linear interpolation,
    multiply-high by ln 2 (Q32)."""
    v = int(r) | 1
    t = v.bit_length() - 1
    frac = (v << (32 - t)) & 0xFFFFFFFF  # bits below the leading one, left aligned
    i, f = frac >> 24, (frac >> 11) & 0x1FFF
    log2v = (t << 26) + LOG2_T26[i] + (((LOG2_T26[i + 1] - LOG2_T26[i]) * f) >> 13)
    return (((1 << 31) - log2v) * LN2_Q32) >> 32


def rate_of(p: float):
    """Probability (narrowed to float32 like the reference) -> (INV, SH) of the gap arithmetic, or None if it never fires."""
    import math

    f = float(np.float32(p))
    if not f > 0:
        return None
    if f >= 1:
        return (0, 0)
    lam = -math.log1p(-f)
    m, e = math.frexp(1.0 / lam)
    sh = 58 - e
    if sh < 0:
        return None
    return (int(math.floor(math.ldexp(m, 32))), sh)


def gap_of(gap_word: int, rate) -> int:
    """floor(Exp(1) / lambda) in the fixed-point arithmetic of the device."""
    return (exp_draw_q26(gap_word) * rate[0]) >> rate[1]
