// hayset.cu -- a prepared, device-resident SET of haystacks for the many-haystack mode.
//
// "Construct once, search many times" (the reference's searcher pattern, e.g. tests/i386.rs:20-22,
// applied to the haystack side): the set keeps, next to the caller's blob and offsets, one lookup
// hint per 4 KiB of blob -- the index of the haystack holding that byte -- so that the scan's hit
// path finds the haystack of a match with one or two probes instead of a binary search over the whole
// offset table (20 dependent loads for a million haystacks).  Measured on 8 GiB / 1 048 053 haystacks:
// a needle present in 75 % of the haystacks 676 -> 1 376 GB/s, in 95 % of them 235 -> 629 GB/s; absent
// needles are unaffected (7.0 TB/s either way).
#include "capi_internal.h"

#include <new>

struct ss_b200_hayset {
    const uint8_t *blob = nullptr;
    const uint64_t *offsets = nullptr;
    size_t n = 0;
    size_t blob_len = 0;
    uint32_t *hint = nullptr; // n_gran entries, or nullptr when the set has 2^32 or more haystacks
    size_t n_gran = 0;
};

namespace {

// hint[g] = last h with off[h] <= g * SS_HINT_GRANULE (off[0] == 0, off[n] == blob_len > that byte)
__global__ void build_hints_kernel(const unsigned long long *__restrict__ off, unsigned long long n,
                                   unsigned long long n_gran, uint32_t *__restrict__ hint)
{
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long g = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; g < n_gran; g += stride) {
        const unsigned long long pos = g << SS_HINT_SHIFT;
        unsigned long long lo = 0, hi = n;
        while (hi - lo > 1) {
            const unsigned long long mid = (lo + hi) >> 1;
            if (__ldg(off + mid) <= pos)
                lo = mid;
            else
                hi = mid;
        }
        hint[g] = (uint32_t)lo;
    }
}

} // namespace

extern "C" int ss_b200_hayset_create(const void *d_blob, const uint64_t *d_offsets, size_t n_haystacks,
                                     size_t blob_len, void *stream, ss_b200_hayset **out)
{
    if (!out || !d_offsets || (blob_len && !d_blob))
        return SS_B200_E_ARG;
    *out = nullptr;
    ss_b200_hayset *hs = new (std::nothrow) ss_b200_hayset();
    if (!hs)
        return SS_B200_E_NOMEM;
    hs->blob = (const uint8_t *)d_blob;
    hs->offsets = d_offsets;
    hs->n = n_haystacks;
    hs->blob_len = blob_len;
    if (n_haystacks > 0 && n_haystacks < 0xFFFFFFFFull && blob_len > 0) {
        SsDeviceInfo dev;
        int rc = ss_capi_device_info(dev);
        if (rc != SS_B200_OK) {
            delete hs;
            return rc;
        }
        hs->n_gran = (blob_len + SS_HINT_GRANULE - 1) >> SS_HINT_SHIFT;
        cudaError_t e = cudaMalloc(&hs->hint, hs->n_gran * sizeof(uint32_t));
        if (e != cudaSuccess) {
            delete hs;
            return ss_capi_cuda_fail(e, "cudaMalloc(hayset hints)");
        }
        unsigned long long blocks = (hs->n_gran + 255) / 256;
        if (blocks > (unsigned long long)dev.sm_count * 16)
            blocks = (unsigned long long)dev.sm_count * 16;
        build_hints_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
            (const unsigned long long *)d_offsets, n_haystacks, hs->n_gran, hs->hint);
        ss_host_count_launch(1);
        e = cudaGetLastError();
        if (e != cudaSuccess) {
            cudaFree(hs->hint);
            delete hs;
            return ss_capi_cuda_fail(e, "build_hints_kernel");
        }
    }
    *out = hs;
    return SS_B200_OK;
}

extern "C" void ss_b200_hayset_free(ss_b200_hayset *hs)
{
    if (!hs)
        return;
    if (hs->hint)
        cudaFree(hs->hint);
    delete hs;
}

extern "C" size_t ss_b200_hayset_len(const ss_b200_hayset *hs) { return hs ? hs->n : 0; }

extern "C" int ss_b200_hayset_search_async(const ss_b200_searcher *s, const ss_b200_hayset *hs, uint8_t *d_flags,
                                           void *workspace, void *stream)
{
    if (!hs)
        return SS_B200_E_ARG;
    return ss_capi_search_many(s, hs->blob, hs->offsets, hs->n, hs->blob_len, d_flags, workspace, hs->hint,
                               hs->n_gran, stream);
}

// ---------------------------------------------------------------------------------------------
// Bit-packed flags for the cross-GPU OR.  Every haystack of a partitioned set lives on one rank, so the
// ranks' flags occupy disjoint bit ranges of one global bitmap: each rank packs its byte flags into its
// own range (all other bits of its copy zero) and the copies are combined with ncclAllReduce(ncclSum)
// on uint32 words -- with disjoint bits a sum is the bitwise OR NCCL lacks -- moving 1 bit per haystack
// instead of the 1 byte per haystack x world of a MAX-reduced byte array.

namespace {
// word w of the bitmap covers global haystacks [32w, 32w+32); this rank holds [bit0, bit0 + n)
__global__ void pack_flags_kernel(const uint8_t *__restrict__ flags, unsigned long long n, unsigned long long bit0,
                                  uint32_t *__restrict__ words)
{
    const unsigned long long w0 = bit0 >> 5, w1 = (bit0 + n + 31) >> 5;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long w = w0 + (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; w < w1; w += stride) {
        uint32_t v = 0;
        const long long g0 = (long long)(w << 5) - (long long)bit0; // local index of the word's bit 0
        if (g0 >= 0 && (unsigned long long)g0 + 32 <= n && ((reinterpret_cast<uintptr_t>(flags + g0) & 15) == 0)) {
            const uint4 a = *reinterpret_cast<const uint4 *>(flags + g0);
            const uint4 b = *reinterpret_cast<const uint4 *>(flags + g0 + 16);
            const uint32_t q[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
            for (int t = 0; t < 8; t++) {
                // four byte flags (any non-zero value counts) -> four bits
                const uint32_t nz = ~swar_zero_exact(q[t]) & 0x80808080u;
                v |= (((nz >> 7) * 0x00204081u) >> 21 & 0xFu) << (4 * t);
            }
        } else {
            for (int t = 0; t < 32; t++) {
                const long long i = g0 + t;
                if (i >= 0 && (unsigned long long)i < n && flags[i])
                    v |= 1u << t;
            }
        }
        words[w] = v;
    }
}
} // namespace

extern "C" int ss_b200_pack_flags_async(const uint8_t *d_flags, size_t n, size_t first_bit, uint32_t *d_words,
                                        size_t total_bits, void *stream)
{
    if (!d_words || (n && !d_flags) || first_bit + n > total_bits)
        return SS_B200_E_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t n_words = (total_bits + 31) / 32;
    SS_CUDA(cudaMemsetAsync(d_words, 0, n_words * 4, st));
    if (n == 0)
        return SS_B200_OK;
    const unsigned long long cnt = ((first_bit + n + 31) >> 5) - (first_bit >> 5);
    unsigned long long blocks = (cnt + 255) / 256;
    if (blocks > 148ull * 8)
        blocks = 148ull * 8;
    pack_flags_kernel<<<(unsigned)blocks, 256, 0, st>>>(d_flags, n, first_bit, d_words);
    ss_host_count_launch(1);
    SS_CUDA(cudaGetLastError());
    return SS_B200_OK;
}
