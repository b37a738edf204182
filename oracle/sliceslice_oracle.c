/*
 * oracle/sliceslice_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the one hot path of cloudflare/sliceslice-rs that this
 * repository accelerates: DynamicAvx2Searcher::{new, with_position, search_in}.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library, and only as the checker / the timed
 * CPU baseline.  Nothing under sliceslice_rs_b200/ links, imports or calls it.
 *
 * Parity status: PINNED.  The restatement is checked (tests/test_oracle.py)
 * against every known-answer vector the reference's own tests hold for this
 * path (src/lib.rs:303-331 memchr KATs, src/lib.rs:422-544 the 32 pairs x every
 * position, src/x86.rs:6-14 doctest, src/x86.rs:533-565 panic contract,
 * tests/i386.rs:46-70 corpus sweeps) and, in the build container, against the
 * reference's vendored C++ ancestor avx2_strstr_v2 compiled from
 * /root/reference into oracle/_ref (oracle/Makefile).
 *
 * All file:line citations are relative to the reference checkout.
 *
 * Every function returns the index at which the reference's scan stops with
 * `true` (fact: that index is always the leftmost occurrence, because chunks
 * are visited in ascending order, src/lib.rs:263-274, and bits in ascending
 * order, src/lib.rs:221) or SS_ORACLE_NPOS when the reference returns false.
 */
#define _GNU_SOURCE
#include <immintrin.h>
#include <pthread.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define SS_ORACLE_NPOS ((size_t)-1)

/* Construction outcomes, mirroring the reference's panics. */
#define SS_ORACLE_OK 0
#define SS_ORACLE_E_POSITION 1     /* assert!(position < size) src/x86.rs:300; assert_eq!(position,0) :473 */
#define SS_ORACLE_E_EMPTY_NEEDLE 2 /* Avx2Searcher::new([]) panics, src/x86.rs:285,300 */

/* ------------------------------------------------------------------------ */
/* Ground truth: leftmost occurrence, the oracle the reference's own tests use
 * (tests/i386.rs:6-10 windows().position(); src/lib.rs:371-373).             */
size_t ss_oracle_naive_find(const uint8_t *hay, size_t n, const uint8_t *needle, size_t k)
{
    if (k == 0)
        return 0; /* DynamicAvx2Searcher::N0 => true (src/x86.rs:470,500) */
    if (n < k)
        return SS_ORACLE_NPOS;
    for (size_t i = 0; i + k <= n; i++) {
        size_t j = 0;
        while (j < k && hay[i + j] == needle[j])
            j++;
        if (j == k)
            return i;
    }
    return SS_ORACLE_NPOS;
}

/* ------------------------------------------------------------------------ */
/* Constructor contract.
 *   dynamic=1: DynamicAvx2Searcher::with_position (src/x86.rs:468-493)
 *   dynamic=0: Avx2Searcher::with_position        (src/x86.rs:297-316)        */
int ss_oracle_check_ctor(size_t k, size_t position, int dynamic)
{
    if (dynamic) {
        if (k == 0)
            return SS_ORACLE_OK; /* [] => N0, position ignored (:470) */
        if (k == 1)
            return position == 0 ? SS_ORACLE_OK : SS_ORACLE_E_POSITION; /* :473 */
        return position < k ? SS_ORACLE_OK : SS_ORACLE_E_POSITION;      /* :300 via :476-491 */
    }
    if (k == 0)
        return SS_ORACLE_E_EMPTY_NEEDLE; /* position < 0 is impossible (:298-300) */
    return position < k ? SS_ORACLE_OK : SS_ORACLE_E_POSITION;
}

/* Default position of `new`: len.wrapping_sub(1) (src/x86.rs:457, :285). */
size_t ss_oracle_default_position(size_t k) { return k - 1; /* wraps for k==0, ignored there */ }

/* ------------------------------------------------------------------------ */
/* The `Vector` abstraction (src/lib.rs:144-159) instantiated the way
 * src/x86.rs:26-235 does it.  Each loader returns the bitmask of lanes where
 * hay[s+i]==first && hay[s+position+i]==last, already limited to LANES bits. */

static inline uint32_t block_mask_32(const uint8_t *s, size_t position, __m256i vf, __m256i vl)
{
    /* impl Vector for __m256i: loadu / cmpeq_epi8 / and / movemask (src/x86.rs:202-235) */
    __m256i a = _mm256_loadu_si256((const __m256i *)s);
    __m256i b = _mm256_loadu_si256((const __m256i *)(s + position));
    __m256i e = _mm256_and_si256(_mm256_cmpeq_epi8(vf, a), _mm256_cmpeq_epi8(vl, b));
    return (uint32_t)_mm256_movemask_epi8(e);
}

static inline uint32_t block_mask_16(const uint8_t *s, size_t position, __m128i vf, __m128i vl)
{
    /* impl Vector for __m128i (src/x86.rs:167-200) */
    __m128i a = _mm_loadu_si128((const __m128i *)s);
    __m128i b = _mm_loadu_si128((const __m128i *)(s + position));
    __m128i e = _mm_and_si128(_mm_cmpeq_epi8(vf, a), _mm_cmpeq_epi8(vl, b));
    return (uint32_t)_mm_movemask_epi8(e);
}

static inline uint32_t block_mask_8(const uint8_t *s, size_t position, __m128i vf, __m128i vl)
{
    /* __m64i: read_unaligned 8 bytes, set1_epi64x, movemask & 0xFF (src/x86.rs:120-165) */
    int64_t a8, b8;
    memcpy(&a8, s, 8);
    memcpy(&b8, s + position, 8);
    __m128i e = _mm_and_si128(_mm_cmpeq_epi8(vf, _mm_set1_epi64x(a8)), _mm_cmpeq_epi8(vl, _mm_set1_epi64x(b8)));
    return (uint32_t)_mm_movemask_epi8(e) & 0xFFu;
}

static inline uint32_t block_mask_4(const uint8_t *s, size_t position, __m128i vf, __m128i vl)
{
    /* __m32i: 4 bytes, set1_epi32, movemask & 0xF (src/x86.rs:73-118) */
    int32_t a4, b4;
    memcpy(&a4, s, 4);
    memcpy(&b4, s + position, 4);
    __m128i e = _mm_and_si128(_mm_cmpeq_epi8(vf, _mm_set1_epi32(a4)), _mm_cmpeq_epi8(vl, _mm_set1_epi32(b4)));
    return (uint32_t)_mm_movemask_epi8(e) & 0xFu;
}

static inline uint32_t block_mask_2(const uint8_t *s, size_t position, __m128i vf, __m128i vl)
{
    /* __m16i: 2 bytes, set1_epi16, movemask & 0x3 (src/x86.rs:26-71) */
    int16_t a2, b2;
    memcpy(&a2, s, 2);
    memcpy(&b2, s + position, 2);
    __m128i e = _mm_and_si128(_mm_cmpeq_epi8(vf, _mm_set1_epi16(a2)), _mm_cmpeq_epi8(vl, _mm_set1_epi16(b2)));
    return (uint32_t)_mm_movemask_epi8(e) & 0x3u;
}

/* vector_search_in_chunk (src/lib.rs:199-251): walk the set bits lowest first
 * (:221 trailing_zeros, :247 clear lowest), verify needle[1..] (:216-218,
 * memcmp! :190-197; the 16 literal-length arms :222-241 only change codegen),
 * stop at the first verified candidate (:242-244). Returns its index in the
 * haystack or NPOS.                                                          */
static inline size_t verify_block(uint32_t eq, const uint8_t *hay, size_t start, const uint8_t *needle, size_t k)
{
    while (eq) {
        size_t c = start + (size_t)__builtin_ctz(eq);
        if (memcmp(hay + c + 1, needle + 1, k - 1) == 0)
            return c;
        eq &= eq - 1;
    }
    return SS_ORACLE_NPOS;
}

/* vector_search_in (src/lib.rs:253-287), one copy per lane width:
 * chunks_exact(LANES) over haystack[..end] ascending (:263-274), then, when
 * end % LANES = rem > 0, one block at end-LANES with
 * mask = u32::MAX << (LANES - rem) (:276-284).                              */
#define DEFINE_VECTOR_SEARCH(NAME, LANES, VT, MASKFN)                                                        \
    static size_t NAME(const uint8_t *hay, size_t end, const uint8_t *needle, size_t k, size_t position,     \
                       VT vf, VT vl)                                                                         \
    {                                                                                                        \
        size_t full = end / (LANES);                                                                         \
        for (size_t b = 0; b < full; b++) {                                                                  \
            size_t s = b * (LANES);                                                                          \
            uint32_t eq = MASKFN(hay + s, position, vf, vl); /* & u32::MAX */                                \
            if (eq) {                                                                                        \
                size_t r = verify_block(eq, hay, s, needle, k);                                              \
                if (r != SS_ORACLE_NPOS)                                                                     \
                    return r;                                                                                \
            }                                                                                                \
        }                                                                                                    \
        size_t rem = end % (LANES);                                                                          \
        if (rem > 0) {                                                                                       \
            size_t s = end - (LANES);                                                                        \
            uint32_t mask = 0xFFFFFFFFu << ((LANES) - rem);                                                  \
            uint32_t eq = MASKFN(hay + s, position, vf, vl) & mask;                                          \
            if (eq) {                                                                                        \
                size_t r = verify_block(eq, hay, s, needle, k);                                              \
                if (r != SS_ORACLE_NPOS)                                                                     \
                    return r;                                                                                \
            }                                                                                                \
        }                                                                                                    \
        return SS_ORACLE_NPOS;                                                                               \
    }

DEFINE_VECTOR_SEARCH(vector_search_32, 32, __m256i, block_mask_32)
DEFINE_VECTOR_SEARCH(vector_search_16, 16, __m128i, block_mask_16)
DEFINE_VECTOR_SEARCH(vector_search_8, 8, __m128i, block_mask_8)
DEFINE_VECTOR_SEARCH(vector_search_4, 4, __m128i, block_mask_4)
DEFINE_VECTOR_SEARCH(vector_search_2, 2, __m128i, block_mask_2)

/* Avx2Searcher::inlined_search_in (src/x86.rs:356-376). Requires k >= 1 and
 * position < k (the constructor has already enforced it).                    */
static size_t avx2_searcher_find(const uint8_t *hay, size_t n, const uint8_t *needle, size_t k, size_t position)
{
    if (n <= k) /* :357-359  haystack == needle */
        return (n == k && memcmp(hay, needle, k) == 0) ? 0 : SS_ORACLE_NPOS;

    size_t end = n - k + 1; /* :361, >= 2 here */
    uint8_t f = needle[0], l = needle[position]; /* VectorHash::new(bytes[0], bytes[position]) :307-308 */
    __m128i xf = _mm_set1_epi8((char)f), xl = _mm_set1_epi8((char)l);

    /* lane ladder :363-375 */
    if (end < 4)
        return vector_search_2(hay, end, needle, k, position, xf, xl);
    if (end < 8)
        return vector_search_4(hay, end, needle, k, position, xf, xl);
    if (end < 16)
        return vector_search_8(hay, end, needle, k, position, xf, xl);
    if (end < 32)
        return vector_search_16(hay, end, needle, k, position, xf, xl);
    return vector_search_32(hay, end, needle, k, position, _mm256_set1_epi8((char)f), _mm256_set1_epi8((char)l));
}

/* DynamicAvx2Searcher::inlined_search_in (src/x86.rs:498-519) over the variant
 * chosen by with_position (src/x86.rs:468-493).
 * Returns SS_ORACLE_OK and writes *offset (NPOS = search_in() == false), or the
 * constructor error.                                                         */
int ss_oracle_dynamic_avx2_find(const uint8_t *hay, size_t n, const uint8_t *needle, size_t k, size_t position,
                                size_t *offset)
{
    int rc = ss_oracle_check_ctor(k, position, 1);
    if (rc != SS_ORACLE_OK)
        return rc;
    if (k == 0) { /* N0 => true, even for an empty haystack (:500) */
        *offset = 0;
        return SS_ORACLE_OK;
    }
    if (k == 1) { /* N1(MemchrSearcher): src/lib.rs:130-136 */
        if (n == 0) {
            *offset = SS_ORACLE_NPOS;
            return SS_ORACLE_OK;
        }
        const uint8_t *p = (const uint8_t *)memchr(hay, needle[0], n); /* memchr crate 2.x: first equal byte */
        *offset = p ? (size_t)(p - hay) : SS_ORACLE_NPOS;
        return SS_ORACLE_OK;
    }
    *offset = avx2_searcher_find(hay, n, needle, k, position);
    return SS_ORACLE_OK;
}

/* Avx2Searcher (non-dynamic) flavour: empty needle is a constructor error. */
int ss_oracle_avx2_find(const uint8_t *hay, size_t n, const uint8_t *needle, size_t k, size_t position,
                        size_t *offset)
{
    int rc = ss_oracle_check_ctor(k, position, 0);
    if (rc != SS_ORACLE_OK)
        return rc;
    *offset = avx2_searcher_find(hay, n, needle, k, position);
    return SS_ORACLE_OK;
}

/* Number of two-anchor filter candidates (set bits that reach the verify
 * step if no match stops the scan) -- used to pin SURVEY's "242 for ipsum". */
size_t ss_oracle_count_candidates(const uint8_t *hay, size_t n, const uint8_t *needle, size_t k, size_t position)
{
    if (k == 0 || n < k || position >= k)
        return 0;
    size_t end = n - k + 1, c = 0;
    for (size_t i = 0; i < end; i++)
        c += (hay[i] == needle[0] && hay[i + position] == needle[position]);
    return c;
}

/* ------------------------------------------------------------------------ */
/* Multi-threaded driver for the CPU baseline: each
 * thread runs the restated searcher over a contiguous slice of start
 * positions with a k-1 byte halo; first offsets are min-reduced.  The
 * reference itself is single-threaded; this is the "all host cores" figure. */
typedef struct {
    const uint8_t *hay;
    size_t lo, hi; /* start positions [lo, hi) */
    size_t n;
    const uint8_t *needle;
    size_t k, position;
    size_t result;
} mt_job;

static void *mt_worker(void *arg)
{
    mt_job *j = (mt_job *)arg;
    size_t len = (j->hi - j->lo) + j->k - 1; /* slice incl. halo */
    size_t off = SS_ORACLE_NPOS;
    ss_oracle_dynamic_avx2_find(j->hay + j->lo, len, j->needle, j->k, j->position, &off);
    j->result = (off == SS_ORACLE_NPOS) ? SS_ORACLE_NPOS : j->lo + off;
    return NULL;
}

int ss_oracle_dynamic_avx2_find_mt(const uint8_t *hay, size_t n, const uint8_t *needle, size_t k, size_t position,
                                   int nthreads, size_t *offset)
{
    int rc = ss_oracle_check_ctor(k, position, 1);
    if (rc != SS_ORACLE_OK)
        return rc;
    if (nthreads <= 1 || k == 0 || n < k + 4096) /* tiny inputs: not worth threads */
        return ss_oracle_dynamic_avx2_find(hay, n, needle, k, position, offset);
    size_t end = n - k + 1;
    if ((size_t)nthreads > end / 1024)
        nthreads = (int)(end / 1024);
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nthreads);
    mt_job *jobs = (mt_job *)malloc(sizeof(mt_job) * (size_t)nthreads);
    size_t per = (end + (size_t)nthreads - 1) / (size_t)nthreads;
    for (int t = 0; t < nthreads; t++) {
        size_t lo = per * (size_t)t, hi = lo + per;
        if (lo > end)
            lo = end;
        if (hi > end)
            hi = end;
        jobs[t] = (mt_job){hay, lo, hi, n, needle, k, position, SS_ORACLE_NPOS};
        if (hi > lo)
            pthread_create(&th[t], NULL, mt_worker, &jobs[t]);
    }
    size_t best = SS_ORACLE_NPOS;
    for (int t = 0; t < nthreads; t++) {
        if (jobs[t].hi > jobs[t].lo) {
            pthread_join(th[t], NULL);
            if (jobs[t].result < best)
                best = jobs[t].result;
        }
    }
    free(th);
    free(jobs);
    *offset = best;
    return SS_ORACLE_OK;
}

/* ------------------------------------------------------------------------ */
/* Corpus sweeps (workload definitions: bench/benches/i386.rs:246-257 long,
 * :118-131 short; tests/i386.rs:46-70).  Needles/haystacks arrive as a blob
 * plus n+1 offsets (CSR).  `use_naive` switches between the restatement and
 * the naive ground truth so tests can diff the two.                        */

typedef struct {
    const uint8_t *blob;
    const uint64_t *off;
    size_t first, last; /* needle index range */
    const uint8_t *hay;
    size_t n;
    int use_naive;
    uint64_t *out;
} long_job;

static void *long_worker(void *arg)
{
    long_job *j = (long_job *)arg;
    for (size_t w = j->first; w < j->last; w++) {
        const uint8_t *nd = j->blob + j->off[w];
        size_t k = (size_t)(j->off[w + 1] - j->off[w]);
        size_t r;
        if (j->use_naive)
            r = ss_oracle_naive_find(j->hay, j->n, nd, k);
        else
            ss_oracle_dynamic_avx2_find(j->hay, j->n, nd, k, k - 1, &r);
        j->out[w] = (uint64_t)r;
    }
    return NULL;
}

/* every needle over one haystack; out[w] = first offset or UINT64_MAX */
void ss_oracle_long_sweep(const uint8_t *blob, const uint64_t *off, size_t n_needles, const uint8_t *hay, size_t n,
                          int use_naive, int nthreads, uint64_t *out)
{
    if (nthreads < 1)
        nthreads = 1;
    if ((size_t)nthreads > n_needles)
        nthreads = (int)(n_needles ? n_needles : 1);
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nthreads);
    long_job *jobs = (long_job *)malloc(sizeof(long_job) * (size_t)nthreads);
    /* interleave-free contiguous split; early needles are not cheaper on average */
    size_t per = (n_needles + (size_t)nthreads - 1) / (size_t)nthreads;
    for (int t = 0; t < nthreads; t++) {
        size_t a = per * (size_t)t, b = a + per;
        if (a > n_needles)
            a = n_needles;
        if (b > n_needles)
            b = n_needles;
        jobs[t] = (long_job){blob, off, a, b, hay, n, use_naive, out};
        if (nthreads == 1)
            long_worker(&jobs[t]);
        else
            pthread_create(&th[t], NULL, long_worker, &jobs[t]);
    }
    if (nthreads > 1)
        for (int t = 0; t < nthreads; t++)
            pthread_join(th[t], NULL);
    free(th);
    free(jobs);
}

/* Triangular pair sweep over a length-sorted word list: needle i against every
 * haystack j >= i (bench/benches/i386.rs:124-129).  Pair (i,j) has linear
 * index  i*W - i*(i-1)/2 + (j-i); bit p of bitmap[] (LSB-first within each
 * u32) is the search_in() result.  Returns the number of matches.          */
uint64_t ss_oracle_short_sweep(const uint8_t *blob, const uint64_t *off, size_t n_words, int use_naive,
                               uint32_t *bitmap /* nullable */)
{
    uint64_t matches = 0, p = 0;
    for (size_t i = 0; i < n_words; i++) {
        const uint8_t *nd = blob + off[i];
        size_t k = (size_t)(off[i + 1] - off[i]);
        for (size_t j = i; j < n_words; j++, p++) {
            const uint8_t *hs = blob + off[j];
            size_t n = (size_t)(off[j + 1] - off[j]);
            size_t r;
            if (use_naive)
                r = ss_oracle_naive_find(hs, n, nd, k);
            else
                ss_oracle_dynamic_avx2_find(hs, n, nd, k, k - 1, &r);
            if (r != SS_ORACLE_NPOS) {
                matches++;
                if (bitmap)
                    bitmap[p >> 5] |= 1u << (p & 31);
            }
        }
    }
    return matches;
}

/* Arbitrary (needle, haystack) pair list; needles and haystacks are separate
 * CSR sets.  out[p] = first offset or UINT64_MAX.                          */
void ss_oracle_pairs(const uint8_t *nblob, const uint64_t *noff, const uint8_t *hblob, const uint64_t *hoff,
                     const uint32_t *pair_needle, const uint32_t *pair_hay, size_t n_pairs, int use_naive,
                     uint64_t *out)
{
    for (size_t p = 0; p < n_pairs; p++) {
        const uint8_t *nd = nblob + noff[pair_needle[p]];
        size_t k = (size_t)(noff[pair_needle[p] + 1] - noff[pair_needle[p]]);
        const uint8_t *hs = hblob + hoff[pair_hay[p]];
        size_t n = (size_t)(hoff[pair_hay[p] + 1] - hoff[pair_hay[p]]);
        size_t r;
        if (use_naive)
            r = ss_oracle_naive_find(hs, n, nd, k);
        else
            ss_oracle_dynamic_avx2_find(hs, n, nd, k, k - 1, &r);
        out[p] = (uint64_t)r;
    }
}

/* ------------------------------------------------------------------------ */
/* The synthetic generator of BASELINE configs 4/5:
 *   byte[i] = (splitmix64(seed ^ (i>>3)) >> (8*(i&7))) & 0xFF, 0xFF -> 0x00.
 * CPU copy so tests can regenerate any slice the device generated.          */
static inline uint64_t splitmix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

void ss_oracle_fill_random(uint8_t *dst, uint64_t global_start, size_t len, uint64_t seed)
{
    for (size_t t = 0; t < len; t++) {
        uint64_t i = global_start + t;
        uint8_t b = (uint8_t)(splitmix64(seed ^ (i >> 3)) >> (8 * (i & 7)));
        dst[t] = (b == 0xFF) ? 0x00 : b;
    }
}

/* Tile `src` (len m) periodically: dst[t] = src[(global_start + t) % m]. */
void ss_oracle_fill_tiled(uint8_t *dst, uint64_t global_start, size_t len, const uint8_t *src, size_t m)
{
    size_t ph = (size_t)(global_start % m);
    for (size_t t = 0; t < len;) {
        size_t c = m - ph;
        if (c > len - t)
            c = len - t;
        memcpy(dst + t, src + ph, c);
        t += c;
        ph = 0;
    }
}
