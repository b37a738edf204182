// test_ctx.cpp -- the multi-GPU context driven from C++ through the host mirror, on however many GPUs
// the box has (ss_b200_ctx_create(0)): what a compiled host following the reference's own FFI pattern
// (bench/sse4-strstr/src/lib.rs:4-15) gets.  Shapes follow the reference's tests: a naive windows() search
// is the expectation (src/lib.rs:365-381, tests/i386.rs:6-10).
//
//   test_ctx <i386.txt>
#include "sliceslice_b200.hpp"

#include <cuda_runtime_api.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iterator>
#include <string>
#include <vector>

using namespace sliceslice_b200;

static int g_checks = 0;
#define CHECK(cond)                                                                                                  \
    do {                                                                                                             \
        g_checks++;                                                                                                  \
        if (!(cond)) {                                                                                               \
            fprintf(stderr, "%s:%d: CHECK failed: %s\n", __FILE__, __LINE__, #cond);                                 \
            exit(1);                                                                                                 \
        }                                                                                                            \
    } while (0)

static std::optional<size_t> naive(const std::string &h, const std::string &n)
{
    auto it = std::search(h.begin(), h.end(), n.begin(), n.end());
    if (it == h.end() && !n.empty())
        return std::nullopt;
    return (size_t)(it - h.begin());
}

int main(int argc, char **argv)
{
    if (argc < 2)
        return 2;
    std::ifstream f(argv[1], std::ios::binary);
    const std::string i386((std::istreambuf_iterator<char>(f)), {});
    Context ctx(0);
    const int ndev = ctx.device_count();
    CHECK(ndev >= 1);

    // one haystack, 24 copies of the text with a marker needle that does not occur in it
    std::string hay;
    for (int i = 0; i < 24; i++)
        hay += i386;
    const std::string marker = "\x01zq~sliceslice~qz\x02";
    CHECK(!naive(hay, marker).has_value());
    auto searcher = DynamicB200Searcher::new_(marker);
    const int exchanges[3] = {SS_B200_EXCHANGE_HOST, SS_B200_EXCHANGE_PEER, SS_B200_EXCHANGE_NCCL};
    const int n_ex = ndev > 1 ? 3 : 1;
    {
        auto sh = ctx.upload_sharded(hay, 256);
        CHECK(sh.len() == hay.size());
        for (int e = 0; e < n_ex; e++) {
            ctx.set_exchange(exchanges[e]);
            CHECK(!ctx.search_in(searcher, sh));
            CHECK(!ctx.find_in(searcher, sh).has_value());
        }
    }
    // plants, descending: the last k bytes; straddling every shard boundary; inside shard 0
    const size_t per = (((hay.size() + ndev - 1) / ndev) + 15) & ~(size_t)15, k = marker.size();
    std::vector<size_t> spots = {hay.size() - k};
    for (int d = ndev - 1; d >= 1; d--)
        spots.push_back(d * per - k / 2);
    spots.push_back(4242);
    for (size_t spot : spots) {
        hay.replace(spot, k, marker);
        auto sh = ctx.upload_sharded(hay, 256);
        for (int e = 0; e < n_ex; e++) {
            ctx.set_exchange(exchanges[e]);
            CHECK(ctx.search_in(searcher, sh));
            CHECK(ctx.find_in(searcher, sh) == naive(hay, marker));
            CHECK(ctx.find_in(searcher, sh) == std::optional<size_t>(spot));
        }
        // the same haystack as ONE host slice striped over all devices (pageable here; pinned below)
        CHECK(ctx.find_in(searcher, hay) == std::optional<size_t>(spot));
        CHECK(ctx.search_in(searcher, hay));
    }
    // every needle of the reference's long sweep that starts a line of the text: found at the same offset
    for (const char *w : {"segmentation", "the", "80386", "descriptor", "x", ""}) {
        auto s = DynamicB200Searcher::new_(w);
        auto sh = ctx.upload_sharded(i386, 64);
        CHECK(ctx.find_in(s, sh) == naive(i386, w));
        CHECK(ctx.find_in(s, i386) == naive(i386, w));
    }
    // pinned host slice: the DMA ring reads it directly
    {
        void *pinned = nullptr;
        CHECK(cudaHostAlloc(&pinned, hay.size(), cudaHostAllocDefault) == cudaSuccess);
        memcpy(pinned, hay.data(), hay.size());
        CHECK(ctx.find_in(searcher, Bytes(pinned, hay.size())) == naive(hay, marker));
        uint64_t h2d = 0, chunks = 0, chunk_bytes = 0;
        int mode = 0;
        check(ss_b200_ctx_last_host_stats(ctx.raw(), &h2d, &chunks, &chunk_bytes, &mode));
        CHECK(mode == 1 || mode == 2);
        CHECK(chunks >= 1);
        cudaFreeHost(pinned);
    }
    // many-haystack mode: lines of the text as the set
    {
        std::vector<uint64_t> off = {0};
        size_t pos = 0, taken = 0;
        while (pos < i386.size() && taken < 5000) {
            size_t nl = i386.find('\n', pos);
            if (nl == std::string::npos)
                nl = i386.size() - 1;
            off.push_back(nl + 1);
            pos = nl + 1;
            taken++;
        }
        auto set = ctx.upload_haystack_set((const uint8_t *)i386.data(), off.data(), off.size() - 1);
        for (const char *w : {"segment", "the", "ipsum", ""}) {
            auto s = DynamicB200Searcher::new_(w);
            auto flags = ctx.search_in(s, set);
            CHECK(flags.size() == off.size() - 1);
            for (size_t h = 0; h + 1 < off.size(); h++) {
                const std::string line = i386.substr(off[h], off[h + 1] - off[h]);
                CHECK((flags[h] != 0) == naive(line, w).has_value());
            }
        }
    }
    int v = 0;
    if (ndev > 1) {
        check(ss_b200_ctx_nccl_version(&v));
        CHECK(v >= 20000);
    }
    printf("ok: %d checks on %d device(s), nccl %d\n", g_checks, ndev, v);
    return 0;
}
