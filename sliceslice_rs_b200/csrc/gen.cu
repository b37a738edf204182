// gen.cu -- device-side generators for the synthetic haystacks of BASELINE configs 2'/4/5.
// Bit-identical CPU copies: oracle/sliceslice_oracle.c ss_oracle_fill_random / _fill_tiled.
#include "ss_host.h"

namespace {

__device__ __forceinline__ unsigned long long splitmix64(unsigned long long x)
{
    x += 0x9E3779B97F4A7C15ull;
    unsigned long long z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

// word g covers global byte indices [8g, 8g+8): byte j = (splitmix64(seed ^ g) >> 8j) & 0xFF, 0xFF -> 0x00
__device__ __forceinline__ unsigned long long random_word(unsigned long long g, unsigned long long seed)
{
    unsigned long long z = splitmix64(seed ^ g);
    const unsigned long long nz = ~z; // zero byte <=> source byte was 0xFF
    const unsigned long long lo7 = 0x7F7F7F7F7F7F7F7Full;
    const unsigned long long m = ~(((nz & lo7) + lo7) | nz | lo7); // 0x80 where byte == 0xFF
    const unsigned long long ff = (m >> 7) * 0xFFull;
    return z & ~ff;
}

__global__ void fill_random_kernel(uint8_t *dst, unsigned long long len, unsigned long long gs, unsigned long long seed)
{
    // thread t produces the generator word g = (gs >> 3) + t
    const unsigned long long g0 = gs >> 3;
    const unsigned long long n_words = ((gs + len + 7) >> 3) - g0;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    const bool fast = ((gs & 7) == 0) && ((reinterpret_cast<uintptr_t>(dst) & 7) == 0);
    for (unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; t < n_words; t += stride) {
        const unsigned long long w = random_word(g0 + t, seed);
        const unsigned long long i0 = (g0 + t) << 3; // global index of byte 0 of w
        if (fast && i0 + 8 <= gs + len) {
            *reinterpret_cast<unsigned long long *>(dst + (i0 - gs)) = w;
        } else {
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const unsigned long long i = i0 + j;
                if (i >= gs && i < gs + len)
                    dst[i - gs] = (uint8_t)(w >> (8 * j));
            }
        }
    }
}

__global__ void fill_tiled_kernel(uint8_t *dst, unsigned long long len, unsigned long long gs, const uint8_t *src,
                                  unsigned long long m)
{
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    const unsigned long long n_groups = (len + 15) >> 4;
    const bool aligned = (reinterpret_cast<uintptr_t>(dst) & 15) == 0;
    for (unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; t < n_groups; t += stride) {
        const unsigned long long t0 = t << 4;
        unsigned long long ph = (gs + t0) % m;
        uint32_t w[4] = {0, 0, 0, 0};
        const int cnt = (len - t0 >= 16) ? 16 : (int)(len - t0);
        for (int j = 0; j < cnt; j++) {
            w[j >> 2] |= (uint32_t)__ldg(src + ph) << (8 * (j & 3));
            if (++ph == m)
                ph = 0;
        }
        if (aligned && cnt == 16) {
            *reinterpret_cast<uint4 *>(dst + t0) = make_uint4(w[0], w[1], w[2], w[3]);
        } else {
            for (int j = 0; j < cnt; j++)
                dst[t0 + j] = (uint8_t)(w[j >> 2] >> (8 * (j & 3)));
        }
    }
}

} // namespace

cudaError_t ss_host_fill_random(void *d_dst, size_t len, uint64_t global_start, uint64_t seed, int sm_count,
                                cudaStream_t stream)
{
    if (len == 0)
        return cudaSuccess;
    fill_random_kernel<<<sm_count * 8, 256, 0, stream>>>((uint8_t *)d_dst, len, global_start, seed);
    ss_host_count_launch(1);
    return cudaGetLastError();
}

cudaError_t ss_host_fill_tiled(void *d_dst, size_t len, uint64_t global_start, const void *d_src, size_t src_len,
                               int sm_count, cudaStream_t stream)
{
    if (len == 0)
        return cudaSuccess;
    fill_tiled_kernel<<<sm_count * 8, 256, 0, stream>>>((uint8_t *)d_dst, len, global_start, (const uint8_t *)d_src,
                                                         src_len);
    ss_host_count_launch(1);
    return cudaGetLastError();
}
