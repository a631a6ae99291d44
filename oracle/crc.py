"""CRC-32 (ISO-HDLC, zlib) and CRC-32C (Castagnoli): bitwise reference + GF(2) helpers.

Oracle / test infrastructure only (see oracle/__init__.py).

Reference semantics: ``Utils::crc32(buf) = crc32fast::hash(buf)``
(orpc/src/common/utils.rs:73-75) = CRC-32/ISO-HDLC: reflected polynomial
0xEDB88320, init 0xFFFFFFFF, xorout 0xFFFFFFFF == ``zlib.crc32``.
The bench/test checksum is ``sum_u64(crc32(buf))`` over read buffers
(curvine-tests/src/curvine_bench.rs:37-48,222-231).
"""
import zlib

POLY_IEEE = 0xEDB88320  # reflected 0x04C11DB7
POLY_CASTAGNOLI = 0x82F63B78  # reflected 0x1EDC6F41
POLYS = {0: POLY_IEEE, 1: POLY_CASTAGNOLI}

CHECK_INPUT = b"123456789"
CHECK_IEEE = 0xCBF43926
CHECK_CASTAGNOLI = 0xE3069283


def crc_bitwise(data: bytes, poly: int = POLY_IEEE, crc: int = 0) -> int:
    """Bit-at-a-time reflected CRC with init/xorout 0xFFFFFFFF (slow; small inputs)."""
    r = crc ^ 0xFFFFFFFF
    for b in data:
        r ^= b
        for _ in range(8):
            r = (r >> 1) ^ (poly if r & 1 else 0)
    return r ^ 0xFFFFFFFF


_TABLES = {}


def table(poly: int):
    t = _TABLES.get(poly)
    if t is None:
        t = []
        for i in range(256):
            r = i
            for _ in range(8):
                r = (r >> 1) ^ (poly if r & 1 else 0)
            t.append(r)
        _TABLES[poly] = t
    return t


def crc_table(data: bytes, poly: int = POLY_IEEE, crc: int = 0) -> int:
    """Byte-at-a-time table CRC (Sarwate); any polynomial."""
    t = table(poly)
    r = crc ^ 0xFFFFFFFF
    for b in data:
        r = t[(r ^ b) & 0xFF] ^ (r >> 8)
    return r ^ 0xFFFFFFFF


def crc32(data, crc: int = 0) -> int:
    """The reference's Utils::crc32 (== zlib.crc32)."""
    return zlib.crc32(data, crc) & 0xFFFFFFFF


def crc32c(data, crc: int = 0) -> int:
    return crc_table(bytes(data), POLY_CASTAGNOLI, crc)


def crc(data, poly_id: int) -> int:
    return crc32(data) if poly_id == 0 else crc32c(data)


def bench_checksum(data: bytes, buf_size: int, stale_tail: bool = True) -> int:
    """``curvine-bench --checksum true`` read-side figure.

    curvine_bench.rs:222-231: ``loop { n = read_full(&mut buf); if n == 0 {break};
    update_ck(&buf) }`` -- note it checksums the WHOLE buffer (``&buf``), not
    ``&buf[..n]``, so a short final read carries the previous iteration's tail
    bytes.  ``stale_tail=False`` gives the "intended" sum over ``buf[..n]``.
    Result is a u64 wrapping sum of per-buffer CRC-32 values.
    """
    total = 0
    buf = bytearray(buf_size)
    pos = 0
    while pos < len(data):
        n = min(buf_size, len(data) - pos)
        buf[:n] = data[pos:pos + n]
        total += crc32(bytes(buf) if stale_tail else bytes(buf[:n]))
        pos += n
    return total & 0xFFFFFFFFFFFFFFFF


# ---------------------------------------------------------------------------
# GF(2)[x] mod P helpers in the *reflected* representation used by the kernels:
# bit i of a 32-bit value is the coefficient of x^(31-i).
# ---------------------------------------------------------------------------

def gf_mulx(a: int, poly: int) -> int:
    """a * x mod P."""
    return (a >> 1) ^ (poly if a & 1 else 0)


def gf_mul(a: int, b: int, poly: int) -> int:
    """a * b mod P (reflected; 0x80000000 is the polynomial 1)."""
    r = 0
    for i in range(32):
        if b & (0x80000000 >> i):  # coefficient of x^i in b
            r ^= a
        a = gf_mulx(a, poly)
    return r


def gf_xpow(n: int, poly: int) -> int:
    """x^n mod P."""
    r = 0x80000000
    base = 0x40000000  # x^1
    while n:
        if n & 1:
            r = gf_mul(r, base, poly)
        base = gf_mul(base, base, poly)
        n >>= 1
    return r


def crc_raw(data: bytes, poly: int) -> int:
    """Linear part: init 0, no xorout  (= M(x) * x^32 mod P)."""
    t = table(poly)
    r = 0
    for b in data:
        r = t[(r ^ b) & 0xFF] ^ (r >> 8)
    return r


def python_bench_checksum(thread_streams) -> int:
    """The reference's PYTHON bench figure (curvine-libsdk/python/test/curvineBench.py:30-52): every thread keeps ONE running
    ``zlib.crc32(data, running)`` over all the buffers it moves (i.e. the CRC-32 of the concatenation of its buffers), and the
    per-thread values are folded in thread order with ``zlib.crc32(value.to_bytes(4, "big"), acc)``.
    ``thread_streams``: one iterable of buffers per thread."""
    acc = 0
    for bufs in thread_streams:
        run = 0
        for b in bufs:
            run = zlib.crc32(b, run)
        acc = zlib.crc32((run & 0xFFFFFFFF).to_bytes(4, "big"), acc)
    return acc & 0xFFFFFFFF


def crc_combine(crc_a: int, crc_b: int, len_b: int, poly: int) -> int:
    """CRC(A||B) from CRC(A), CRC(B), len(B) -- zlib's crc32_combine restated."""
    return gf_mul(crc_a, gf_xpow(8 * len_b, poly), poly) ^ crc_b
