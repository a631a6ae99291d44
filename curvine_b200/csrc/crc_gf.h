// GF(2)[x] mod P arithmetic in the reflected representation (bit i <-> x^(31-i)) and the
// constant tables the CRC kernels use.  Shared by host (table generation, manifest CRCs,
// crc_combine) and device (tail folds).  P is a parameter: 0xEDB88320 for CRC-32/ISO-HDLC
// (what the reference's Utils::crc32 computes, orpc/src/common/utils.rs:73-75) and
// 0x82F63B78 for CRC-32C.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define CV_HD __host__ __device__ __forceinline__
#else
#define CV_HD inline
#endif

namespace cv {

constexpr uint32_t kPolyIeee = 0xEDB88320u;
constexpr uint32_t kPolyCastagnoli = 0x82F63B78u;
constexpr uint32_t kOne = 0x80000000u;  // the polynomial "1"

CV_HD uint32_t poly_of(int poly_id) { return poly_id == 0 ? kPolyIeee : kPolyCastagnoli; }

// a * x mod P
CV_HD uint32_t gf_mulx(uint32_t a, uint32_t poly) { return (a >> 1) ^ ((0u - (a & 1u)) & poly); }

// a * b mod P
CV_HD uint32_t gf_mul(uint32_t a, uint32_t b, uint32_t poly) {
    uint32_t r = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll 8
#endif
    for (int i = 0; i < 32; i++) {
        r ^= (0u - ((b >> (31 - i)) & 1u)) & a;
        a = gf_mulx(a, poly);
    }
    return r;
}

// x^n mod P
CV_HD uint32_t gf_xpow(uint64_t n, uint32_t poly) {
    uint32_t r = kOne, base = kOne >> 1;
    while (n) {
        if (n & 1) r = gf_mul(r, base, poly);
        base = gf_mul(base, base, poly);
        n >>= 1;
    }
    return r;
}

// Constant tables for one polynomial; lives in device global memory (one copy per device and polynomial).
struct CrcConsts {
    uint32_t t0[256];     // Sarwate byte table: t0[v] = v * x^8 ... (standard reflected table)
    uint32_t m[4][256];   // multiply-by-x^4096 (one 512-byte warp row), sliced by state byte
    uint32_t xp128[64];   // x^(128*e): weight of a 16-byte vector e vectors from the end
    uint32_t pw8[16];     // x^(8*n), n < 16: head/tail byte runs
    uint32_t poly;
    uint32_t pad_[3];
};

inline void build_consts(uint32_t poly, CrcConsts* c) {
    for (uint32_t v = 0; v < 256; v++) {
        uint32_t r = v;
        for (int k = 0; k < 8; k++) r = gf_mulx(r, poly);
        c->t0[v] = r;
    }
    const uint32_t x4096 = gf_xpow(4096, poly);
    for (int i = 0; i < 4; i++)
        for (uint32_t v = 0; v < 256; v++) c->m[i][v] = gf_mul(v << (8 * i), x4096, poly);
    for (int e = 0; e < 64; e++) c->xp128[e] = gf_xpow(128ull * e, poly);
    for (int n = 0; n < 16; n++) c->pw8[n] = gf_xpow(8ull * n, poly);
    c->poly = poly;
    c->pad_[0] = c->pad_[1] = c->pad_[2] = 0;
}

// CRC(A||B) from CRC(A), CRC(B), len(B)  (zlib crc32_combine semantics)
inline uint32_t crc_combine(uint32_t crc_a, uint32_t crc_b, uint64_t len_b, uint32_t poly) {
    return gf_mul(crc_a, gf_xpow(8 * len_b, poly), poly) ^ crc_b;
}

}  // namespace cv
