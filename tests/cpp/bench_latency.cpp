// bench_latency.cpp -- per-call latency of the C-ABI entry points (development tool).
//   g++ -O2 -std=c++17 -Iinclude -I/usr/local/cuda/include tests/cpp/bench_latency.cpp \
//       -Lsliceslice_rs_b200 -lsliceslice_b200 -L/usr/local/cuda/lib64 -lcudart -Wl,-rpath,$PWD/sliceslice_rs_b200
//   bench_latency <i386.txt> <words.txt> [service = 1 | 0]   (0: one kernel launch per synchronous call)
#include "sliceslice_b200.hpp"

#include <cuda_runtime.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iterator>
#include <string>
#include <vector>

using namespace sliceslice_b200;
using clk = std::chrono::steady_clock;

static std::string read_file(const char *path)
{
    std::ifstream f(path, std::ios::binary);
    return std::string(std::istreambuf_iterator<char>(f), {});
}

int main(int argc, char **argv)
{
    if (argc < 3)
        return 2;
    const std::string i386 = read_file(argv[1]);
    std::vector<std::string> words;
    {
        const std::string w = read_file(argv[2]);
        size_t a = 0;
        while (a < w.size()) {
            size_t b = w.find('\n', a);
            if (b == std::string::npos)
                b = w.size();
            if (b > a)
                words.push_back(w.substr(a, b - a));
            a = b + 1;
        }
    }
    if (argc > 3)
        check(ss_b200_set_sync_service(atoi(argv[3]), 0));
    printf("resident service kernel for synchronous calls: %s\n", (argc > 3 && atoi(argv[3]) == 0) ? "off" : "on");
    DeviceHaystack hay = DeviceHaystack::upload(i386);
    std::vector<DynamicB200Searcher> searchers;
    for (auto &w : words)
        searchers.push_back(DynamicB200Searcher::new_(w));
    unsigned long long sum = 0;
    for (int it = 0; it < 4; it++) {
        sum = 0;
        auto t0 = clk::now();
        for (auto &s : searchers)
            sum += *s.find_in(hay);
        double ms = std::chrono::duration<double, std::milli>(clk::now() - t0).count();
        printf("long sweep (one find_in per word, device haystack): %.3f ms/iteration, %.2f us/call, sum %llu\n", ms,
               ms * 1e3 / searchers.size(), sum);
    }
    // the same loop stream-ordered: one ss_b200_find_in_device_async per word, results in device memory,
    // one synchronisation at the end (what an async-aware caller of the drop-in would do)
    {
        uint64_t *d_res = nullptr;
        void *d_ws = nullptr;
        cudaMalloc(&d_res, searchers.size() * sizeof(uint64_t));
        cudaMalloc(&d_ws, 32);
        cudaMemset(d_ws, 0, 32);
        cudaStream_t st;
        cudaStreamCreate(&st);
        std::vector<uint64_t> res(searchers.size());
        for (int it = 0; it < 4; it++) {
            auto t0 = clk::now();
            for (size_t w = 0; w < searchers.size(); w++)
                searchers[w].find_in_device_async(hay.device_ptr(), hay.len(), 0, SS_B200_NPOS, d_ws, d_res + w, st);
            const double enq_ms = std::chrono::duration<double, std::milli>(clk::now() - t0).count();
            cudaMemcpyAsync(res.data(), d_res, res.size() * sizeof(uint64_t), cudaMemcpyDeviceToHost, st);
            cudaStreamSynchronize(st);
            double ms = std::chrono::duration<double, std::milli>(clk::now() - t0).count();
            printf("  (host enqueue alone: %.3f ms = %.2f us/call)\n", enq_ms, enq_ms * 1e3 / searchers.size());
            sum = 0;
            for (uint64_t v : res)
                sum += v;
            printf("long sweep (one find_in_device_async per word, one sync): %.3f ms/iteration, %.2f us/call, sum %llu\n",
                   ms, ms * 1e3 / searchers.size(), sum);
        }
        // the same number of stream-ordered calls with ONE searcher (one kernel variant, absent needle):
        // separates the cost of the call itself from the cost of switching kernels between launches
        auto same = DynamicB200Searcher::new_("ipsum");
        for (int it = 0; it < 3; it++) {
            auto t0 = clk::now();
            for (size_t w = 0; w < searchers.size(); w++)
                same.find_in_device_async(hay.device_ptr(), hay.len(), 0, SS_B200_NPOS, d_ws, d_res + w, st);
            const double enq_ms = std::chrono::duration<double, std::milli>(clk::now() - t0).count();
            cudaStreamSynchronize(st);
            double ms = std::chrono::duration<double, std::milli>(clk::now() - t0).count();
            printf("same searcher, %zu stream-ordered calls: %.2f us/call (host enqueue alone %.2f us)\n", searchers.size(),
                   ms * 1e3 / searchers.size(), enq_ms * 1e3 / searchers.size());
        }
        cudaFree(d_res);
        cudaFree(d_ws);
        cudaStreamDestroy(st);
    }
    auto absent = DynamicB200Searcher::new_("ipsum");
    for (int it = 0; it < 3; it++) {
        auto t0 = clk::now();
        for (int i = 0; i < 2000; i++)
            absent.search_in(hay);
        double us = std::chrono::duration<double, std::micro>(clk::now() - t0).count() / 2000;
        printf("absent needle over 857 KB device haystack: %.2f us/call = %.1f GB/s\n", us, i386.size() / us / 1e3);
    }
    for (int it = 0; it < 3; it++) {
        auto t0 = clk::now();
        for (int i = 0; i < 2000; i++)
            absent.search_in(Bytes("Lorem ipsum dolor sit amet, consectetur adipiscing elit"));
        double us = std::chrono::duration<double, std::micro>(clk::now() - t0).count() / 2000;
        printf("host slice of 55 bytes through search_in_host: %.2f us/call\n", us);
    }
    return 0;
}
