"""The ALGORITHM of crc_small_kernel (curvine_b200/csrc/kernels.cu), restated thread by thread in Python and checked against the
oracle on CPU -- chunk geometry (1024 right-aligned chunks of S bytes), the init folded into the first four message bytes, the
lane / warp combine trees with a multiplier that is squared per level starting at x^(8S).  It does not execute the CUDA code (the
GPU tests do); it pins the arithmetic the kernel is written to, so that a length or alignment the GPU suite has not met yet
cannot hide a flaw in the scheme itself."""
import numpy as np
import pytest

from oracle import crc as C


def _emulate(data: bytes, poly: int) -> int:
    n = len(data)
    if n == 0:
        return 0
    if n < 4:
        return C.crc_table(data, poly)
    S = ((n + 1023) // 1024 + 15) & ~15
    assert S // 16 < 64, "xp128 table bound"
    folded = bytes(b ^ 0xFF for b in data[:4]) + data[4:]
    c = []
    for t in range(1024):
        end = n - (1023 - t) * S
        beg = max(end - S, 0)
        c.append(C.crc_raw(folded[beg:end], poly) if end > 0 else 0)
    m = C.gf_xpow(8 * S, poly)
    # lanes: level d combines lane L with lane L + d for L % (2d) == 0
    d = 1
    while d < 32:
        for w in range(32):
            for lane in range(0, 32, 2 * d):
                c[w * 32 + lane] = C.gf_mul(c[w * 32 + lane], m, poly) ^ c[w * 32 + lane + d]
        m = C.gf_mul(m, m, poly)
        d *= 2
    part = [c[w * 32] for w in range(32)]
    d = 1
    while d < 32:
        for lane in range(0, 32, 2 * d):
            part[lane] = C.gf_mul(part[lane], m, poly) ^ part[lane + d]
        m = C.gf_mul(m, m, poly)
        d *= 2
    return (~part[0]) & 0xFFFFFFFF


@pytest.mark.parametrize("poly", [C.POLY_IEEE, C.POLY_CASTAGNOLI])
def test_small_kernel_scheme_equals_the_crc_definition(poly):
    rng = np.random.default_rng(5)
    data = rng.integers(0, 256, size=1024 * 1008, dtype=np.uint8).tobytes()
    for n in [1, 2, 3, 4, 5, 7, 8, 9, 15, 16, 17, 100, 1023, 1024, 1025, 4097, 12345, 16385, 65543, 200000, 204099, 262144, 262149]:
        assert _emulate(data[:n], poly) == C.crc_table(data[:n], poly), n
    for n in [1024 * 1008 - 1, 1024 * 1008]:  # the largest block the kernel takes (S = 1008)
        assert _emulate(data[:n], poly) == C.crc_table(data[:n], poly), n
    assert _emulate(b"123456789", C.POLY_IEEE) == 0xCBF43926 and _emulate(b"123456789", C.POLY_CASTAGNOLI) == 0xE3069283
