// test_host_mirror.cpp -- the reference's own test shapes, driven through the C++ host mirror
// (include/sliceslice_b200.hpp -> C ABI -> sm_100a kernels).
//
//   search()           src/lib.rs:365-381   every `position`, compared with naive windows().any()
//   KAT groups         src/lib.rs:422-544   (table generated from tests/golden/kats.json)
//   memchr KATs        src/lib.rs:299-331
//   panics             src/x86.rs:533-565
//   search_long_haystack / search_short_haystack   tests/i386.rs:46-70
//
// usage: test_host_mirror ctor            (no device needed: constructor contract only)
//        test_host_mirror all <i386.txt> <words.txt>
#include "sliceslice_b200.hpp"

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

struct Kat {
    const char *group, *haystack, *needle;
    bool found;
    long offset;
};
static const Kat KATS[] = {
#include "kats_table.inc"
};
struct MemchrKat {
    const char *haystack, *needle;
    bool found;
};
static const MemchrKat MEMCHR_KATS[] = {
#include "memchr_table.inc"
};

// tests/i386.rs:6-10
static std::optional<size_t> find_subsequence(const std::string &haystack, const std::string &needle)
{
    auto it = std::search(haystack.begin(), haystack.end(), needle.begin(), needle.end());
    if (it == haystack.end() && !needle.empty())
        return std::nullopt;
    return (size_t)(it - haystack.begin());
}

// src/lib.rs:365-381: the result must not depend on `position`, for both searcher flavours
static bool search(const std::string &haystack, const std::string &needle)
{
    const bool result = find_subsequence(haystack, needle).has_value();
    DeviceHaystack dev = DeviceHaystack::upload(haystack);
    for (size_t position = 0; position < needle.size(); position++) {
        auto dynamic = DynamicB200Searcher::with_position(needle, position);
        CHECK(dynamic.search_in(haystack) == result);
        CHECK(dynamic.inlined_search_in(dev) == result);
        CHECK(dynamic.find_in(dev) == find_subsequence(haystack, needle));
        auto strict = B200Searcher::with_position(needle, position);
        CHECK(strict.search_in(dev) == result);
    }
    return result;
}

template <typename F>
static bool panics(F f)
{
    try {
        f();
    } catch (const SearcherPanic &) {
        return true;
    }
    return false;
}

static void test_ctor_contract()
{
    // src/x86.rs:533-565
    CHECK(panics([] { B200Searcher::with_position("foo", 3); }));        // avx2_invalid_position
    CHECK(panics([] { DynamicB200Searcher::with_position("foo", 3); })); // dynamic_avx2_invalid_position
    CHECK(panics([] { B200Searcher::new_(""); }));                       // avx2_empty_needle
    CHECK(!panics([] { DynamicB200Searcher::new_(""); }));               // N0 is valid, src/x86.rs:470
    CHECK(panics([] { DynamicB200Searcher::with_position("a", 1); }));   // assert_eq!(position, 0), :473
    CHECK(!panics([] { DynamicB200Searcher::with_position("", 7); }));   // position ignored for N0
    CHECK(DynamicB200Searcher::new_("ipsum").position() == 4);
    CHECK(DynamicB200Searcher::with_position("ipsum", 2).position() == 2);
    // second-anchor choice is host arithmetic: built-in table and a caller histogram
    CHECK(DynamicB200Searcher::with_rarest_position("the quiz").position() == 7);
    CHECK(DynamicB200Searcher::with_rarest_position("x").position() == 0);
    CHECK(panics([] { B200Searcher::with_rarest_position(""); })); // Avx2Searcher::new(empty), src/x86.rs:285
    {
        std::vector<uint64_t> hist(256, 100);
        hist['p'] = 1;
        CHECK(DynamicB200Searcher::with_rarest_position("ipsum", hist.data()).position() == 1);
        hist['p'] = 100; // all equal: the reference's default, the last byte (src/x86.rs:457)
        CHECK(DynamicB200Searcher::with_rarest_position("ipsum", hist.data()).position() == 4);
    }
    // decided on the host, no device involved
    CHECK(DynamicB200Searcher::new_("").search_in(Bytes("")));
    CHECK(!DynamicB200Searcher::new_("abcd").search_in(Bytes("abc")));
    CHECK(!DynamicB200Searcher::new_("a").search_in(Bytes("")));
}

// src/lib.rs:35-104, :333-363: every needle carrier the reference accepts ([u8; N], [u8], Box/Rc/Arc, &N,
// Vec<u8>) reaches the searcher as the same bytes; here: literal, std::string, string_view, vector, pointer+len
static void test_needle_carriers()
{
    const std::string hay = "Lorem ipsum dolor sit amet, consectetur adipiscing elit";
    const std::string s = "ipsum";
    const std::vector<uint8_t> v(s.begin(), s.end());
    const std::string_view sv(s);
    const uint8_t raw[5] = {'i', 'p', 's', 'u', 'm'};
    for (const Bytes &needle : {Bytes("ipsum"), Bytes(s), Bytes(sv), Bytes(v), Bytes(raw, sizeof raw)}) {
        CHECK(needle.len == 5);
        auto searcher = DynamicB200Searcher::new_(needle);
        CHECK(searcher.needle() == v);
        CHECK(searcher.find_in(Bytes(hay)) == std::optional<size_t>(6));
    }
    // needles with embedded NUL and high bytes keep their full length (no C-string truncation)
    const uint8_t odd[4] = {0x00, 0xFF, 0x00, 0x80};
    const uint8_t hay2[9] = {1, 2, 0x00, 0xFF, 0x00, 0x80, 3, 0, 0};
    CHECK(DynamicB200Searcher::new_(Bytes(odd, 4)).find_in(Bytes(hay2, 9)) == std::optional<size_t>(2));
}

static std::string read_file(const char *path)
{
    std::ifstream f(path, std::ios::binary);
    if (!f) {
        fprintf(stderr, "cannot open %s\n", path);
        exit(2);
    }
    return std::string(std::istreambuf_iterator<char>(f), {});
}

int main(int argc, char **argv)
{
    const std::string mode = argc > 1 ? argv[1] : "ctor";
    test_ctor_contract();
    if (mode == "ctor") {
        printf("ok: %d checks (constructor contract, no device)\n", g_checks);
        return 0;
    }
    if (argc < 4) {
        fprintf(stderr, "usage: %s all <i386.txt> <words.txt>\n", argv[0]);
        return 2;
    }
    test_needle_carriers();
    // KAT groups with the literal expected results of src/lib.rs:422-544
    for (const Kat &k : KATS) {
        CHECK(search(k.haystack, k.needle) == k.found);
        auto off = DynamicB200Searcher::new_(k.needle).find_in(Bytes(k.haystack));
        CHECK(off.has_value() == k.found && (!k.found || (long)*off == k.offset));
    }
    for (const MemchrKat &k : MEMCHR_KATS) // src/lib.rs:299-331
        CHECK(DynamicB200Searcher::new_(k.needle).search_in(Bytes(k.haystack)) == k.found);
    // doctest, src/x86.rs:6-14
    CHECK(DynamicB200Searcher::new_("ipsum").search_in(Bytes("Lorem ipsum dolor sit amet, consectetur adipiscing elit")));
    CHECK(!DynamicB200Searcher::new_("ipsum").search_in(Bytes("foo bar baz qux quux quuz corge grault garply waldo fred")));

    const std::string i386 = read_file(argv[2]);
    std::vector<std::string> words;
    {
        const std::string w = read_file(argv[3]);
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
    CHECK(words.size() == 4585);

    // tests/i386.rs:58-70 search_long_haystack (device-resident haystack, one search_in per word)
    {
        DeviceHaystack hay = DeviceHaystack::upload(i386);
        unsigned long long sum = 0;
        for (const std::string &w : words) {
            auto s = DynamicB200Searcher::new_(w);
            auto got = s.find_in(hay);
            CHECK(got == find_subsequence(i386, w));
            CHECK(s.search_in(hay));
            sum += *got;
        }
        CHECK(sum == 809985317ull);
        CHECK(!DynamicB200Searcher::new_("ipsum").search_in(hay));
        // second anchor from the haystack's own histogram: same answers (src/lib.rs:375-378)
        const std::vector<uint64_t> hist = hay.byte_histogram();
        unsigned long long total = 0;
        for (uint64_t c : hist)
            total += c;
        CHECK(total == i386.size());
        for (size_t w = 0; w < words.size(); w += 97) {
            auto s = DynamicB200Searcher::with_rarest_position(words[w], hist.data());
            CHECK(s.find_in(hay) == find_subsequence(i386, words[w]));
        }
    }
    // tests/i386.rs:46-56 search_short_haystack: every word in every not-shorter word; a fixed stride
    // subsample keeps the one-call-per-pair form (the full sweep runs batched in tests/test_gpu_parity.py)
    {
        std::stable_sort(words.begin(), words.end(),
                         [](const std::string &a, const std::string &b) { return a.size() < b.size(); });
        size_t p = 0, tested = 0, matches = 0;
        for (size_t i = 0; i < words.size(); i++) {
            std::optional<DynamicB200Searcher> s;
            for (size_t j = i; j < words.size(); j++, p++) {
                if (p % 211 != 0)
                    continue;
                if (!s)
                    s.emplace(DynamicB200Searcher::new_(words[i]));
                const bool exp = find_subsequence(words[j], words[i]).has_value();
                CHECK(s->search_in(Bytes(words[j])) == exp);
                tested++;
                matches += exp;
            }
        }
        CHECK(p == 10513405ull);
        printf("short sweep subsample: %zu pairs, %zu matches\n", tested, matches);
    }
    printf("ok: %d checks\n", g_checks);
    return 0;
}
