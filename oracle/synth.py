"""Synthetic file content -- independent restatement of the generator.

Oracle / test infrastructure only (see oracle/__init__.py).  The product ships its
own generator (curvine_b200/csrc/synth.cc); tests check the two agree byte for byte.

Block ``b`` of file ``file_id`` = first ``len`` bytes of a xoshiro256** stream
(little-endian u64 outputs) whose 4-word state is four successive splitmix64
outputs from seed ``0xC0FFEEB200 ^ (file_id << 32) ^ b``.
Mode "az": one lowercase a-z buffer of ``buf_size`` bytes repeated
(curvine-tests/src/bench_action.rs:100-102 / orpc utils.rs:77-85 shape), seed 42.
"""
import numpy as np

M64 = 0xFFFFFFFFFFFFFFFF
SEED_BASE = 0xC0FFEEB200


def splitmix64(x: int):
    x = (x + 0x9E3779B97F4A7C15) & M64
    z = x
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
    return x, z ^ (z >> 31)


def _rotl(x, k):
    return ((x << k) | (x >> (64 - k))) & M64


def block_bytes(file_id: int, block_index: int, length: int) -> bytes:
    x = (SEED_BASE ^ (file_id << 32) ^ block_index) & M64
    s = []
    for _ in range(4):
        x, v = splitmix64(x)
        s.append(v)
    n = (length + 7) // 8
    out = np.empty(n, dtype="<u8")
    s0, s1, s2, s3 = s
    for i in range(n):
        out[i] = (_rotl((s1 * 5) & M64, 7) * 9) & M64
        t = (s1 << 17) & M64
        s2 ^= s0
        s3 ^= s1
        s1 ^= s2
        s0 ^= s3
        s2 ^= t
        s3 = _rotl(s3, 45)
    return out.tobytes()[:length]


def file_bytes(file_id: int, length: int, block_size: int) -> bytes:
    parts, b, pos = [], 0, 0
    while pos < length:
        n = min(block_size, length - pos)
        parts.append(block_bytes(file_id, b, n))
        pos += n
        b += 1
    return b"".join(parts)


def az_buffer(buf_size: int, seed: int = 42) -> bytes:
    x = seed
    out = bytearray(buf_size)
    for i in range(buf_size):
        x, v = splitmix64(x)
        out[i] = ord("a") + v % 26
    return bytes(out)
