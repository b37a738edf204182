// read_peak.cu -- read-only HBM stream rate of this GPU with next to no arithmetic (development tool):
// the ceiling for any scan that must read every haystack byte once.  MEASURED_PEAKS.json's hbm_gbs is a
// COPY (read + write) figure; a read-only stream runs faster, which is why roofline.frac can exceed 1.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o read_peak tools/read_peak.cu && ./read_peak [GiB]
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

template <int U>
__global__ void __launch_bounds__(256) read_kernel(const uint4 *__restrict__ p, size_t n16, unsigned int *sink)
{
    unsigned int acc = 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x * U;
    for (size_t i = (size_t)blockIdx.x * blockDim.x * U + threadIdx.x; i + (U - 1) * 256 < n16; i += stride) {
        uint4 v[U];
#pragma unroll
        for (int u = 0; u < U; u++)
            asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                         : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w)
                         : "l"(p + i + u * 256));
#pragma unroll
        for (int u = 0; u < U; u++)
            acc ^= v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
    }
    if (acc == 0x12345678u)
        atomicAdd(sink, 1u); // keeps the loads alive
}

int main(int argc, char **argv)
{
    const double gib = argc > 1 ? atof(argv[1]) : 8.0;
    const size_t bytes = (size_t)(gib * (1ull << 30));
    uint4 *d;
    unsigned int *sink;
    cudaMalloc(&d, bytes);
    cudaMalloc(&sink, 4);
    cudaMemset(d, 0x5a, bytes);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    for (int per_sm : {4, 6, 8}) {
        for (int rep = 0; rep < 2; rep++) {
            cudaEventRecord(e0);
            for (int i = 0; i < 10; i++)
                read_kernel<4><<<sms * per_sm, 256>>>(d, bytes / 16, sink);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms = 0;
            cudaEventElapsedTime(&ms, e0, e1);
            if (rep)
                printf("read-only stream, %d CTAs/SM x 256 thr x 4 x 16 B: %.1f GB/s\n", per_sm,
                       bytes * 10.0 / (ms * 1e-3) / 1e9);
        }
    }
    return 0;
}
