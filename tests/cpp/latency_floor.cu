// latency_floor.cu -- what a launch + completion signal costs on this box with no work at all
// (development tool): the floor under any one-kernel-per-search_in design.
#include <chrono>
#include <cstdio>
#include <cuda_runtime.h>

__global__ void signal_kernel(volatile unsigned long long *slot, unsigned long long v, unsigned int *done)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int prev;
        asm volatile("atom.add.acq_rel.gpu.global.u32 %0, [%1], 1;" : "=r"(prev) : "l"(done) : "memory");
        if (prev == gridDim.x - 1) {
            *done = 0;
            asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(slot), "l"(v) : "memory");
        }
    }
}

int main()
{
    unsigned long long *slot, *slot_dev;
    unsigned int *done;
    cudaHostAlloc((void **)&slot, 8, cudaHostAllocMapped);
    cudaHostGetDevicePointer((void **)&slot_dev, slot, 0);
    cudaMalloc(&done, 4);
    cudaMemset(done, 0, 4);
    cudaStream_t st;
    cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
    for (int grid : {1, 210, 840}) {
        for (int rep = 0; rep < 2; rep++) {
            auto t0 = std::chrono::steady_clock::now();
            const int iters = 5000;
            for (int i = 1; i <= iters; i++) {
                *(volatile unsigned long long *)slot = ~0ull;
                signal_kernel<<<grid, 256, 0, st>>>(slot_dev, (unsigned long long)i, done);
                while (*(volatile unsigned long long *)slot == ~0ull) {
                }
            }
            double us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count() / iters;
            if (rep)
                printf("grid %4d x 256: %.2f us per launch + mapped-flag completion\n", grid, us);
        }
    }
    // pipelined: launches queued back to back on one stream, one synchronisation at the end --
    // the launch-rate floor under any stream-ordered one-kernel-per-search loop
    for (int grid : {1, 210}) {
        for (int rep = 0; rep < 2; rep++) {
            const int iters = 5000;
            cudaStreamSynchronize(st);
            auto t0 = std::chrono::steady_clock::now();
            for (int i = 1; i <= iters; i++)
                signal_kernel<<<grid, 256, 0, st>>>(slot_dev, (unsigned long long)i, done);
            double enq = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count() / iters;
            cudaStreamSynchronize(st);
            double us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count() / iters;
            if (rep)
                printf("grid %4d x 256 pipelined: %.2f us per launch (host enqueue alone %.2f us)\n", grid, us, enq);
        }
    }
    return 0;
}
