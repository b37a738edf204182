// bench_latency.cpp -- per-call latency of the synchronous C-ABI entry points (development tool).
//   bench_latency <i386.txt> <words.txt>
#include "sliceslice_b200.hpp"

#include <chrono>
#include <cstdio>
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
