// bench_scan.cpp -- the roofline scan timed at the C ABI, no Python in the loop (development tool;
// also the A/B harness for kernel changes: point LD_LIBRARY_PATH at another build of the library).
//   g++ -O2 -std=c++17 -Iinclude -I/usr/local/cuda/include tests/cpp/bench_scan.cpp \
//       -Lsliceslice_rs_b200 -lsliceslice_b200 -L/usr/local/cuda/lib64 -lcudart -o bench_scan
//   bench_scan <i386.txt> [GiB = 8] [steps = 100] [needle = ipsum] [mode = find | count] [tile_kib = 0] [stages = 0]
#include "sliceslice_b200.h"

#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iterator>
#include <string>

#define CK(x)                                                                                                        \
    do {                                                                                                             \
        if ((x) != 0) {                                                                                              \
            fprintf(stderr, "failed: %s (%s)\n", #x, ss_b200_last_error());                                          \
            return 1;                                                                                                \
        }                                                                                                            \
    } while (0)

int main(int argc, char **argv)
{
    if (argc < 2)
        return 2;
    std::ifstream f(argv[1], std::ios::binary);
    const std::string i386((std::istreambuf_iterator<char>(f)), {});
    const double gib = argc > 2 ? atof(argv[2]) : 8.0;
    const int steps = argc > 3 ? atoi(argv[3]) : 100;
    const std::string needle = argc > 4 ? argv[4] : "ipsum";
    const bool count_mode = argc > 5 && std::string(argv[5]) == "count";
    if (argc > 6)
        CK(ss_b200_set_scan_tuning(0, 0, atoi(argv[6]), argc > 7 ? atoi(argv[7]) : 0));
    const size_t n = (size_t)(gib * (1ull << 30));
    uint8_t *d_src = nullptr, *d_hay = nullptr;
    uint64_t *d_res = nullptr;
    void *d_ws = nullptr;
    CK(cudaMalloc(&d_src, i386.size()));
    CK(cudaMalloc(&d_hay, n + 16));
    CK(cudaMalloc(&d_res, 8));
    CK(cudaMalloc(&d_ws, 32));
    CK(cudaMemset(d_ws, 0, 32));
    CK(cudaMemcpy(d_src, i386.data(), i386.size(), cudaMemcpyHostToDevice));
    cudaStream_t st;
    CK(cudaStreamCreate(&st));
    CK(ss_b200_fill_tiled(d_hay, n, 0, d_src, i386.size(), st));
    ss_b200_searcher *s = nullptr;
    CK(ss_b200_searcher_new((const uint8_t *)needle.data(), needle.size(), &s));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    auto once = [&]() -> int {
        if (count_mode)
            return ss_b200_count_in_device_async(s, d_hay, n, SS_B200_NPOS, d_ws, d_res, st);
        return ss_b200_find_in_device_async(s, d_hay, n, 0, SS_B200_NPOS, d_ws, d_res, st);
    };
    for (int rep = 0; rep < 3; rep++) {
        for (int i = 0; i < 3; i++)
            CK(once());
        cudaEventRecord(e0, st);
        for (int i = 0; i < steps; i++)
            CK(once());
        cudaEventRecord(e1, st);
        CK(cudaStreamSynchronize(st));
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        uint64_t r = 0;
        CK(cudaMemcpy(&r, d_res, 8, cudaMemcpyDeviceToHost));
        if (count_mode)
            printf("count %s over %.3g GiB: %.4f ms/scan, %.1f GB/s, %llu occurrences\n", needle.c_str(), gib, ms / steps,
                   n / (ms / steps * 1e-3) / 1e9, (unsigned long long)r);
        else
            printf("%s over %.3g GiB: %.4f ms/scan, %.1f GB/s, result %s\n", needle.c_str(), gib, ms / steps,
                   n / (ms / steps * 1e-3) / 1e9, r == SS_B200_DEVICE_NONE ? "none" : "found");
    }
    return 0;
}
