// test_filter_host.cu -- CPU emulation of the scan built from the SAME arithmetic the kernels compile
// (sliceslice_rs_b200/csrc/ss_filter.cuh: filter_word, chunk_flag_x, swar_zero_exact, exact_alive),
// checked against a naive search.  Runs without a GPU (host code only).
//
// For every random case (alphabet 2..4, haystack 0..400 bytes, needle 1..40 bytes, every head alignment
// class, random `position`) and every extra-anchor kind the needle offers, the emulation walks the
// 16-byte chunks exactly as a lane does -- clamped loads, second-anchor window at +q, register-window
// exact compare, range check -- and asserts:
//   * no false negatives: a chunk that holds the start of a real match always raises its filter flag,
//     with and without the extra anchors;
//   * the verified positions are exactly the naive match positions (so first offset and count agree).
#include "../../sliceslice_rs_b200/csrc/ss_filter.cuh"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

static long long g_exact_filter_chunks = 0; // chunks whose occurrences were counted straight from the filter words

static uint4 load_chunk(const uint8_t *base, unsigned long long c)
{
    uint4 r;
    memcpy(&r, base + 16 * c, 16);
    return r;
}

struct Case {
    const uint8_t *base; // 16-byte aligned; haystack starts at base + head
    unsigned long long n, head;
    const uint8_t *needle;
    uint32_t k, pos;
    int xk;        // extra-anchor kind under test
    uint32_t e4[2], xbs;
};

template <int WS, bool BSZ, bool K1, int XK>
static bool run_case(const Case &t, const std::vector<long long> &truth)
{
    const uint8_t *hay = t.base + t.head;
    const unsigned long long end = t.n - t.k + 1;
    const unsigned long long n_chunks = (t.head + end + 15) / 16, last = (t.head + t.n - 1) / 16;
    const unsigned long long q = t.pos / 16;
    FilterConsts fc;
    fc.f4 = 0x01010101u * t.needle[0];
    fc.l4 = 0x01010101u * t.needle[t.pos];
    fc.bs = 8u * (t.pos % 4u);
    fc.e4[0] = t.e4[0];
    fc.e4[1] = t.e4[1];
    fc.xbs = t.xbs;
    std::vector<long long> got;
    for (unsigned long long c = 0; c < n_chunks; c++) {
        const uint4 av = load_chunk(t.base, c < last ? c : last);
        const uint4 nx = load_chunk(t.base, c + 1 < last ? c + 1 : last);
        const uint4 lo = (K1 || q == 0) ? av : load_chunk(t.base, c + q < last ? c + q : last);
        const uint4 hi = (K1 || q == 0) ? nx : load_chunk(t.base, c + q + 1 < last ? c + q + 1 : last);
        const uint32_t flag_plain = chunk_flag_x<WS, BSZ, K1, 0>(av, nx, lo, hi, fc);
        const uint32_t flag_extra = chunk_flag_x<WS, BSZ, K1, XK>(av, nx, lo, hi, fc);
        // hit path, as step_alive_mask + hit_tail
        std::vector<long long> here;
        uint32_t z[4] = {0, 0, 0, 0};
        const bool alive = exact_alive<WS, BSZ, K1>(av, nx, lo, hi, fc, t.k,
                                                    [&](uint32_t j) { return 0x01010101u * (uint32_t)t.needle[j]; }, z);
        if (!alive && (z[0] | z[1] | z[2] | z[3])) {
            printf("exact_alive returned false with live positions\n");
            return false;
        }
        const long long p0 = (long long)(c * 16) - (long long)t.head;
        // the packed form the kernels loop over: bit p <-> start position p of the chunk
        const uint32_t b16 = alive ? pack_alive16(z) : 0u;
        for (int p = 0; alive && p < 16; p++)
            if (((b16 >> p) & 1u) != ((z[p / 4] >> (8 * (p % 4) + 7)) & 1u)) {
                printf("pack_alive16 disagrees with the alive mask at position %d\n", p);
                return false;
            }
        // count mode, needles of up to three bytes: when the anchors cover the needle, the zero bytes of the
        // filter words ARE the occurrences (scan_long.cuh counts them without a hit path)
        if (filter_covers_needle(t.k, t.pos, XK, t.xbs / 8u) && (XK == 0 || XK == 3) && p0 >= 0 &&
            (unsigned long long)p0 + 16 <= end) {
            g_exact_filter_chunks++;
            uint32_t from_filter = 0, from_truth = 0;
            for (int j = 0; j < 4; j++) {
                uint32_t zz = swar_zero_exact(filter_word<WS, BSZ, K1, XK>(av, nx, lo, hi, j, fc));
                for (; zz; zz &= zz - 1)
                    from_filter++;
            }
            for (long long i : truth)
                from_truth += (i >= p0 && i < p0 + 16);
            if (from_filter != from_truth) {
                printf("count from filter words %u != %u occurrences in chunk %llu (k %u pos %u XK %d)\n", from_filter,
                       from_truth, c, t.k, t.pos, XK);
                return false;
            }
        }
        for (int j = 0; alive && j < 4; j++)
            for (int b = 0; b < 4; b++)
                if (z[j] & (0x80u << (8 * b))) {
                    const long long i = p0 + 4 * j + b;
                    if (i < 0 || (unsigned long long)i >= end)
                        continue;
                    if (t.k <= 17 || memcmp(hay + i + 17, t.needle + 17, t.k - 17) == 0)
                        here.push_back(i);
                }
        bool chunk_has_match = false;
        for (long long i : truth)
            if (i >= p0 && i < p0 + 16)
                chunk_has_match = true;
        if (chunk_has_match && (flag_plain == 0 || flag_extra == 0)) {
            printf("false negative: chunk %llu plain %x extra %x (XK %d)\n", c, flag_plain, flag_extra, XK);
            return false;
        }
        if (!here.empty() && flag_plain == 0) {
            printf("verified match in an unflagged chunk %llu\n", c);
            return false;
        }
        got.insert(got.end(), here.begin(), here.end());
    }
    if (got != truth) {
        printf("position sets differ: %zu emulated vs %zu naive\n", got.size(), truth.size());
        return false;
    }
    return true;
}

template <int XK>
static bool dispatch(const Case &t, const std::vector<long long> &truth)
{
    if (t.k == 1)
        return run_case<0, true, true, 0>(t, truth);
    const uint32_t r = t.pos % 16, ws = r / 4;
    const bool bsz = (r % 4) == 0;
#define SS_D(WS)                                                                                                     \
    case WS:                                                                                                         \
        return bsz ? run_case<WS, true, false, XK>(t, truth) : run_case<WS, false, false, XK>(t, truth);
    switch (ws) {
        SS_D(0) SS_D(1) SS_D(2)
    default:
        return bsz ? run_case<3, true, false, XK>(t, truth) : run_case<3, false, false, XK>(t, truth);
    }
#undef SS_D
}

int main(int argc, char **argv)
{
    const int cases = argc > 1 ? atoi(argv[1]) : 60000;
    std::mt19937_64 rng(12345);
    std::vector<uint8_t> buf(16 + 1024 + 64);
    long long checked = 0, with_match = 0;
    for (int it = 0; it < cases; it++) {
        const int alphabet = 2 + (int)(rng() % 3);
        const unsigned long long n = rng() % 5 ? rng() % 400 : rng() % 40;
        uint32_t k = 1 + (uint32_t)(rng() % (it % 3 ? 8 : 40));
        if (n < k)
            continue; // decided on the host by the library (src/x86.rs:357-359)
        const unsigned long long head = rng() % 16;
        uint8_t *base = buf.data() + ((16 - ((uintptr_t)buf.data() & 15)) & 15);
        uint8_t *hay = base + head;
        for (unsigned long long i = 0; i < n; i++)
            hay[i] = (uint8_t)(97 + rng() % alphabet);
        std::vector<uint8_t> needle(k);
        if (rng() % 2) {
            const unsigned long long st = rng() % (n - k + 1);
            memcpy(needle.data(), hay + st, k);
        } else {
            for (auto &c : needle)
                c = (uint8_t)(97 + rng() % alphabet);
        }
        // poison what surrounds the haystack with the first needle byte: an out-of-range read that
        // leaked into a result would show up as an extra position
        memset(base, needle[0], head);
        memset(hay + n, needle[0], 48);
        const uint32_t pos = k == 1 ? 0 : (uint32_t)(rng() % k);
        std::vector<long long> truth;
        for (unsigned long long i = 0; i + k <= n; i++)
            if (memcmp(hay + i, needle.data(), k) == 0)
                truth.push_back((long long)i);
        Case t;
        t.base = base;
        t.n = n;
        t.head = head;
        t.needle = needle.data();
        t.k = k;
        t.pos = pos;
        // every extra-anchor kind this needle can offer (as choose_extra_anchors in scan_long.cu would, and more)
        bool ok = true;
        t.xk = 0;
        t.e4[0] = t.e4[1] = t.xbs = 0;
        ok = ok && dispatch<0>(t, truth);
        if (k > 4) {
            t.e4[0] = 0x01010101u * needle[4];
            ok = ok && dispatch<1>(t, truth);
        }
        if (k > 8) {
            t.e4[0] = 0x01010101u * needle[4];
            t.e4[1] = 0x01010101u * needle[8];
            ok = ok && dispatch<2>(t, truth);
        }
        for (uint32_t o = 1; o <= 3 && o < k; o++) {
            t.e4[0] = 0x01010101u * needle[o];
            t.xbs = 8 * o;
            ok = ok && dispatch<3>(t, truth);
        }
        if (!ok) {
            printf("FAILED at case %d: n=%llu k=%u pos=%u head=%llu alphabet=%d\n", it, n, k, pos, head, alphabet);
            return 1;
        }
        checked++;
        with_match += !truth.empty();
    }
    printf("ok: %lld cases (%lld with a match), every extra-anchor kind; %lld chunks counted from filter words\n", checked,
           with_match, g_exact_filter_chunks);
    return 0;
}
