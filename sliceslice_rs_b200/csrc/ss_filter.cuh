// ss_filter.cuh -- the pure arithmetic of the scan: SWAR zero-byte tests, the two-anchor (+ extra
// anchor) filter word, and the register-window refinement of the hit path.  No memory access, no
// intrinsics that only exist on the device: every function is __host__ __device__, so the exact code
// the kernels compile is also emulated and checked on the CPU (tests/cpp/test_filter_host.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define SS_HD __host__ __device__ __forceinline__
#if defined(__CUDA_ARCH__)
#define SS_UNROLL _Pragma("unroll")
#else
#define SS_UNROLL // the host pass (emulation test, host halves of .cu files) has no use for it
#endif

// (hi:lo) >> s, low 32 bits; s in 0..31
SS_HD uint32_t ss_funnel_r(uint32_t lo, uint32_t hi, uint32_t s)
{
#ifdef __CUDA_ARCH__
    return __funnelshift_r(lo, hi, s);
#else
    return (uint32_t)(((((unsigned long long)hi) << 32) | lo) >> (s & 31u));
#endif
}

// SWAR "some byte of x is zero" accumulator term (bit 7 of each byte, plus
// possible false positives above a true zero byte).
SS_HD uint32_t swar_zero_term(uint32_t x) { return (x - 0x01010101u) & ~x; }

// Exact: 0x80 in every byte of x that is zero, 0 elsewhere.
SS_HD uint32_t swar_zero_exact(uint32_t x)
{
    return ~(((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x | 0x7F7F7F7Fu);
}

// Word j (0..3) of the 16 bytes that start at byte R of the 32-byte window lo||hi, with the shift
// split as R = 4*WS + bs/8: the word offset WS is a template parameter (register
// selection must be static), the bit shift `bs` (8, 16 or 24) is a launch-uniform runtime value, and
// BSZ says bs == 0 (no funnel shift at all).  8 instantiations cover the 16 byte shifts.
template <int WS, bool BSZ>
SS_HD uint32_t window_word_rt(const uint4 &lo, const uint4 &hi, int j, uint32_t bs)
{
    const uint32_t v[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
    if (BSZ)
        return v[WS + j];
    return ss_funnel_r(v[WS + j], v[WS + j + 1], bs);
}

// The filter word for 4 start positions (word j of a chunk):
//   zero byte <=> hay[i] == needle[0] && hay[i+pos] == needle[pos] [&& extra anchors]
// The first two terms are the reference's two anchors (VectorHash first/last, src/lib.rs:166-178,
// :207-214).  Extra anchors (XK) only make the filter more selective -- fewer trips to the divergent
// verify path on natural text; a real match always passes, so results never change:
//   XK = 0  none
//   XK = 1  needle[4]             word-aligned with the first stream: one LOP3 per word, no shift
//   XK = 2  needle[4], needle[8]  two LOP3 per word
//   XK = 3  needle[xo], xo in 1..3 (short needles): one funnel shift + one LOP3 per word
//   av : haystack bytes [16c, 16c+16)      nx : the next 16 bytes [16c+16, 16c+32)
//   lo : haystack bytes [16(c+q), +16)     hi : the 16 after lo (lo/hi == av/nx when q == 0)
struct FilterConsts {
    uint32_t f4, l4, bs; // needle[0] x4, needle[pos] x4, 8 * (pos % 4)
    uint32_t e4[2];      // extra anchor bytes x4
    uint32_t xbs;        // XK == 3: 8 * xo
};

template <int WS, bool BSZ, bool K1, int XK>
SS_HD uint32_t filter_word(const uint4 &av, const uint4 &nx, const uint4 &lo, const uint4 &hi,
                                                int j, const FilterConsts &fc)
{
    const uint32_t w[8] = {av.x, av.y, av.z, av.w, nx.x, nx.y, nx.z, nx.w};
    uint32_t x = w[j] ^ fc.f4;
    if (!K1) {
        x |= window_word_rt<WS, BSZ>(lo, hi, j, fc.bs) ^ fc.l4;
        if (XK == 1 || XK == 2)
            x |= w[j + 1] ^ fc.e4[0];
        if (XK == 2)
            x |= w[j + 2] ^ fc.e4[1];
        if (XK == 3)
            x |= ss_funnel_r(w[j], w[j + 1], fc.xbs) ^ fc.e4[0];
    }
    return x;
}

// Candidate test for the 16 start positions of one chunk: non-zero iff some position MAY pass the filter.
template <int WS, bool BSZ, bool K1, int XK>
SS_HD uint32_t chunk_flag_x(const uint4 &av, const uint4 &nx, const uint4 &lo, const uint4 &hi,
                                                 const FilterConsts &fc)
{
    uint32_t acc = 0;
SS_UNROLL
    for (int j = 0; j < 4; j++)
        acc |= swar_zero_term(filter_word<WS, BSZ, K1, XK>(av, nx, lo, hi, j, fc));
    return acc & 0x80808080u;
}

// Exact match mask of the hit path, out of registers.  av||nx are the 32 haystack bytes from the chunk
// start, which cover needle bytes 0..16 of all 16 start positions of the chunk.  The XOR differences of
// every compared byte pair are OR-ed into one accumulator word per 4 positions -- acc byte == 0 <=>
// hay[i + j] == needle[j] for every j compared so far -- starting from the two anchors (filter_word) and
// adding needle bytes 1 .. min(k - 1, 16): one funnel shift + one LOP3 per word and needle byte, one exact
// zero-byte test at the end.  This is the analogue of the reference's constant-length memcmp arms
// (src/lib.rs:222-241) for all 16 positions at once.  z[t] receives 0x80 in the byte of every start
// position whose first min(k, 17) bytes equal the needle's; returns false when there is none (checked
// after needle bytes 1, 2, 4, 8 and 12, so the false candidates of natural text leave after one byte).
// `needle4_at(j)` yields needle byte j splatted over the four bytes of a word.
template <int WS, bool BSZ, bool K1, class Needle4At>
SS_HD bool exact_alive(const uint4 &av, const uint4 &nx, const uint4 &lo, const uint4 &hi, const FilterConsts &fc,
                       uint32_t k, Needle4At needle4_at, uint32_t (&z)[4])
{
    const uint32_t w[8] = {av.x, av.y, av.z, av.w, nx.x, nx.y, nx.z, nx.w};
    uint32_t acc[4];
SS_UNROLL
    for (int t = 0; t < 4; t++)
        acc[t] = filter_word<WS, BSZ, K1, 0>(av, nx, lo, hi, t, fc);
    if (!K1) {
        const uint32_t jmax = k - 1 < 16u ? k - 1 : 16u;
SS_UNROLL
        for (uint32_t j = 1; j <= 16u; j++) {
            if (j > jmax)
                break;
            const uint32_t n4 = needle4_at(j);
            const uint32_t wo = j >> 2, sh = 8u * (j & 3u); // window shifted by j bytes = wo words + sh bits
SS_UNROLL
            for (uint32_t t = 0; t < 4; t++) {
                const uint32_t x = sh ? ss_funnel_r(w[t + wo], w[t + wo + 1], sh) : w[t + wo];
                acc[t] |= x ^ n4;
            }
            if ((j == 1u || j == 2u || j == 4u || j == 8u || j == 12u) && j < jmax) {
                const uint32_t any = (swar_zero_term(acc[0]) | swar_zero_term(acc[1]) | swar_zero_term(acc[2]) |
                                      swar_zero_term(acc[3])) & 0x80808080u;
                if (!any)
                    return false;
            }
        }
    }
    uint32_t any = 0;
SS_UNROLL
    for (int t = 0; t < 4; t++) {
        z[t] = swar_zero_exact(acc[t]);
        any |= z[t];
    }
    return any != 0;
}

// The 0x80-per-alive-position words of one chunk (exact_alive) as 16 bits, bit p <-> start position p.
SS_HD uint32_t pack_alive16(const uint32_t (&z)[4])
{
    uint32_t b16 = 0;
SS_UNROLL
    for (int j = 0; j < 4; j++)
        b16 |= ((((z[j] >> 7) * 0x00204081u) >> 21) & 0xFu) << (4 * j); // bits 0, 8, 16, 24 -> bits 0..3
    return b16;
}

// Do the anchors of a launch's filter -- needle[0], needle[pos] and, for extra-anchor kind 3, needle[xo] --
// compare EVERY byte of a needle of length k?  Then a zero byte of the filter word is an occurrence, and
// count mode may count needles of up to three bytes straight from the filter words (scan_long.cuh).
SS_HD bool filter_covers_needle(uint32_t k, uint32_t pos, int xk, uint32_t xo)
{
    if (k == 0u || k > 3u)
        return false;
    uint32_t covered = 1u | (1u << (pos < 31u ? pos : 31u));
    if (xk == 3)
        covered |= 1u << (xo < 31u ? xo : 31u);
    return covered == (1u << k) - 1u;
}
