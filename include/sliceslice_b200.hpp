// sliceslice_b200.hpp -- C++17 host-side mirror of the reference's searcher interface, above the C ABI.
//
// The reference's host language is Rust; no Rust toolchain exists in the build image, so the host
// side that a Rust caller would get from rust/sliceslice-b200 is provided here in C++ with the same
// names, argument meaning and failure behaviour:
//
//   reference (src/x86.rs)                               here
//   DynamicAvx2Searcher::new(needle)           :454  ->  DynamicB200Searcher::new_(needle)
//   DynamicAvx2Searcher::with_position(n, p)   :468  ->  DynamicB200Searcher::with_position(n, p)
//   .search_in(haystack) -> bool               :523  ->  .search_in(haystack) -> bool
//   .inlined_search_in(haystack)               :498  ->  .inlined_search_in(haystack)
//   Avx2Searcher::{new, with_position}    :282, :297 ->  B200Searcher::{new_, with_position}
//   panic!(..) at construction      :300, :304, :473 ->  throws SearcherPanic
//
// plus find_in() = the index at which the reference's scan returns true (leftmost occurrence), the
// contract of the reference's own FFI precedent avx2_strstr_v2 (bench/sse4-strstr/src/lib.rs:4-15).
// Header-only; link with -lsliceslice_b200.  No CPU fallback: device errors throw B200Error.
#pragma once
#include "sliceslice_b200.h"

#include <cstdint>
#include <cstring>
#include <optional>
#include <stdexcept>
#include <string>
#include <string_view>
#include <utility>
#include <vector>

namespace sliceslice_b200 {

// What the reference does with panic!: assert!(position < needle.size()) src/x86.rs:300,
// assert_eq!(position, 0) :473, Avx2Searcher::new(empty) :285.
struct SearcherPanic : std::logic_error {
    using std::logic_error::logic_error;
};
// CUDA / argument failures (the reference has no counterpart: its search_in is infallible).
struct B200Error : std::runtime_error {
    using std::runtime_error::runtime_error;
};

inline void check(int rc)
{
    if (rc == SS_B200_OK)
        return;
    if (rc == SS_B200_E_POSITION || rc == SS_B200_E_EMPTY_NEEDLE)
        throw SearcherPanic(ss_b200_strerror(rc));
    std::string msg = ss_b200_strerror(rc);
    if (rc == SS_B200_E_CUDA || rc == SS_B200_E_NOMEM || rc == SS_B200_E_NCCL || rc == SS_B200_E_ARG)
        msg += std::string(": ") + ss_b200_last_error();
    throw B200Error(msg);
}

// Borrowed bytes, the analogue of &[u8] / the Needle trait's as_bytes() (src/lib.rs:35-41).
struct Bytes {
    const uint8_t *ptr = nullptr;
    size_t len = 0;
    Bytes() = default;
    Bytes(const uint8_t *p, size_t n) : ptr(p), len(n) {}
    Bytes(const void *p, size_t n) : ptr(static_cast<const uint8_t *>(p)), len(n) {}
    Bytes(std::string_view s) : ptr(reinterpret_cast<const uint8_t *>(s.data())), len(s.size()) {}
    Bytes(const std::string &s) : Bytes(std::string_view(s)) {}
    Bytes(const char *s) : Bytes(std::string_view(s)) {}
    Bytes(const std::vector<uint8_t> &v) : ptr(v.data()), len(v.size()) {}
};

// A haystack resident in HBM: uploaded (owned) or borrowed device memory.
class DeviceHaystack {
public:
    static DeviceHaystack upload(Bytes host)
    {
        ss_b200_haystack *h = nullptr;
        check(ss_b200_haystack_upload(host.ptr, host.len, &h));
        return DeviceHaystack(h);
    }
    static DeviceHaystack from_device(const void *dptr, size_t len)
    {
        ss_b200_haystack *h = nullptr;
        check(ss_b200_haystack_from_device(dptr, len, &h));
        return DeviceHaystack(h);
    }
    DeviceHaystack(DeviceHaystack &&o) noexcept : h_(std::exchange(o.h_, nullptr)) {}
    DeviceHaystack &operator=(DeviceHaystack &&o) noexcept
    {
        if (this != &o) {
            ss_b200_haystack_free(h_);
            h_ = std::exchange(o.h_, nullptr);
        }
        return *this;
    }
    DeviceHaystack(const DeviceHaystack &) = delete;
    DeviceHaystack &operator=(const DeviceHaystack &) = delete;
    ~DeviceHaystack() { ss_b200_haystack_free(h_); }
    size_t len() const { return ss_b200_haystack_len(h_); }
    // 256 byte counts (sample_bytes = 0: a 16 MiB sample, exact for shorter haystacks), the input of
    // with_rarest_position
    std::vector<uint64_t> byte_histogram(size_t sample_bytes = 0) const
    {
        std::vector<uint64_t> hist(256);
        check(ss_b200_haystack_byte_histogram(h_, sample_bytes, hist.data()));
        return hist;
    }
    const void *device_ptr() const { return ss_b200_haystack_device_ptr(h_); }
    const ss_b200_haystack *raw() const { return h_; }

private:
    explicit DeviceHaystack(ss_b200_haystack *h) : h_(h) {}
    ss_b200_haystack *h_ = nullptr;
};

namespace detail {

template <bool STRICT>
class SearcherImpl {
public:
    // ::new(needle): second anchor = last byte (position = len.wrapping_sub(1)), src/x86.rs:282-287, :454-459.
    // (`new` is a C++ keyword, hence the trailing underscore.)
    static SearcherImpl new_(Bytes needle)
    {
        ss_b200_searcher *s = nullptr;
        check(STRICT ? ss_b200_searcher_new_strict(needle.ptr, needle.len, &s)
                     : ss_b200_searcher_new(needle.ptr, needle.len, &s));
        return SearcherImpl(s, needle);
    }
    // ::with_position(needle, position), src/x86.rs:297-316, :468-493.
    static SearcherImpl with_position(Bytes needle, size_t position)
    {
        ss_b200_searcher *s = nullptr;
        check(STRICT ? ss_b200_searcher_with_position_strict(needle.ptr, needle.len, position, &s)
                     : ss_b200_searcher_with_position(needle.ptr, needle.len, position, &s));
        return SearcherImpl(s, needle);
    }
    // with_position(needle, p), p = the index whose byte is rarest under `hist` (256 counts; nullptr =
    // built-in background table) -- SURVEY 8f-3; results do not depend on it (src/lib.rs:375-378).
    // The strict flavour panics on the empty needle like Avx2Searcher::new (src/x86.rs:285, :300).
    static SearcherImpl with_rarest_position(Bytes needle, const uint64_t *hist = nullptr)
    {
        size_t position = 0;
        check(ss_b200_rarest_position(needle.ptr, needle.len, hist, &position));
        return with_position(needle, position);
    }

    SearcherImpl(SearcherImpl &&o) noexcept : s_(std::exchange(o.s_, nullptr)), needle_(std::move(o.needle_)) {}
    SearcherImpl &operator=(SearcherImpl &&o) noexcept
    {
        if (this != &o) {
            ss_b200_searcher_free(s_);
            s_ = std::exchange(o.s_, nullptr);
            needle_ = std::move(o.needle_);
        }
        return *this;
    }
    SearcherImpl(const SearcherImpl &) = delete;
    SearcherImpl &operator=(const SearcherImpl &) = delete;
    ~SearcherImpl() { ss_b200_searcher_free(s_); }

    // private trait Searcher::{needle, position}, src/lib.rs:289-293
    const std::vector<uint8_t> &needle() const { return needle_; }
    size_t position() const { return ss_b200_searcher_position(s_); }

    // search_in(&self, haystack: &[u8]) -> bool, src/x86.rs:380, :523 -- host slice
    bool search_in(Bytes haystack) const
    {
        uint8_t found = 0;
        check(ss_b200_search_in_host(s_, haystack.ptr, haystack.len, &found));
        return found != 0;
    }
    // the same call on a device-resident haystack (the path the roofline is measured on)
    bool search_in(const DeviceHaystack &haystack) const
    {
        uint8_t found = 0;
        check(ss_b200_search_in(s_, haystack.raw(), &found));
        return found != 0;
    }
    template <typename H>
    bool inlined_search_in(const H &haystack) const
    {
        return search_in(haystack); // src/x86.rs:356, :498: #[inline] is a codegen attribute only
    }

    std::optional<size_t> find_in(Bytes haystack) const
    {
        size_t off = SS_B200_NPOS;
        check(ss_b200_find_in_host(s_, haystack.ptr, haystack.len, &off));
        return off == SS_B200_NPOS ? std::nullopt : std::optional<size_t>(off);
    }
    std::optional<size_t> find_in(const DeviceHaystack &haystack) const
    {
        size_t off = SS_B200_NPOS;
        check(ss_b200_find_in(s_, haystack.raw(), &off));
        return off == SS_B200_NPOS ? std::nullopt : std::optional<size_t>(off);
    }
    // stream-ordered device entry (ss_b200_find_in_device_async); see the C header for the arguments
    void find_in_device_async(const void *dptr, size_t len, uint64_t base_offset, size_t start_limit, void *workspace,
                              uint64_t *d_result, void *stream) const
    {
        check(ss_b200_find_in_device_async(s_, dptr, len, base_offset, start_limit, workspace, d_result, stream));
    }
    const ss_b200_searcher *raw() const { return s_; }

private:
    SearcherImpl(ss_b200_searcher *s, Bytes needle) : s_(s), needle_(needle.ptr, needle.ptr + needle.len) {}
    ss_b200_searcher *s_ = nullptr;
    std::vector<uint8_t> needle_;
};

} // namespace detail

// Drop-in for sliceslice::x86::DynamicAvx2Searcher (src/x86.rs:405-526): empty needle always matches.
using DynamicB200Searcher = detail::SearcherImpl<false>;
// Drop-in for sliceslice::x86::Avx2Searcher (src/x86.rs:266-383): empty needle panics.
using B200Searcher = detail::SearcherImpl<true>;

// ---------------------------------------------------------------------------------------------
// Every GPU of the box from one process (ss_b200_ctx).  The searcher surface stays the reference's;
// the context adds WHERE the haystack lives: sharded over the devices, or a host slice striped over
// all of them for the duration of one search_in.

class Context;

// One haystack as contiguous shards of start positions, shard d on device d of the context.
class ShardedHaystack {
public:
    ShardedHaystack(ShardedHaystack &&o) noexcept : h_(std::exchange(o.h_, nullptr)) {}
    ShardedHaystack(const ShardedHaystack &) = delete;
    ShardedHaystack &operator=(const ShardedHaystack &) = delete;
    ~ShardedHaystack() { ss_b200_sharded_free(h_); }
    size_t len() const { return ss_b200_sharded_len(h_); }
    const ss_b200_sharded *raw() const { return h_; }

private:
    friend class Context;
    explicit ShardedHaystack(ss_b200_sharded *h) : h_(h) {}
    ss_b200_sharded *h_ = nullptr;
};

// A set of haystacks partitioned over the devices of a context (many-haystack mode).
class ContextHaystackSet {
public:
    ContextHaystackSet(ContextHaystackSet &&o) noexcept : h_(std::exchange(o.h_, nullptr)) {}
    ContextHaystackSet(const ContextHaystackSet &) = delete;
    ContextHaystackSet &operator=(const ContextHaystackSet &) = delete;
    ~ContextHaystackSet() { ss_b200_ctx_hayset_free(h_); }
    size_t len() const { return ss_b200_ctx_hayset_len(h_); }
    const ss_b200_ctx_hayset *raw() const { return h_; }

private:
    friend class Context;
    explicit ContextHaystackSet(ss_b200_ctx_hayset *h) : h_(h) {}
    ss_b200_ctx_hayset *h_ = nullptr;
};

class Context {
public:
    // ndev <= 0: every visible device
    explicit Context(int ndev = 0, int exchange = SS_B200_EXCHANGE_HOST)
    {
        check(ss_b200_ctx_create(ndev, nullptr, &c_));
        if (exchange != SS_B200_EXCHANGE_HOST)
            set_exchange(exchange);
    }
    Context(const Context &) = delete;
    Context &operator=(const Context &) = delete;
    ~Context() { ss_b200_ctx_free(c_); }
    int device_count() const { return ss_b200_ctx_device_count(c_); }
    void set_exchange(int kind) { check(ss_b200_ctx_set_exchange(c_, kind)); }

    ShardedHaystack upload_sharded(Bytes host, size_t halo = 4096) const
    {
        ss_b200_sharded *h = nullptr;
        check(ss_b200_sharded_upload(c_, host.ptr, host.len, halo, &h));
        return ShardedHaystack(h);
    }
    ShardedHaystack sharded_from_device(const void *const *dptrs, const size_t *owned, const size_t *spans) const
    {
        ss_b200_sharded *h = nullptr;
        check(ss_b200_sharded_from_device(c_, dptrs, owned, spans, &h));
        return ShardedHaystack(h);
    }
    // searcher.search_in(haystack) with the haystack sharded over the devices (src/x86.rs:523)
    template <bool STRICT>
    bool search_in(const detail::SearcherImpl<STRICT> &s, const ShardedHaystack &h)
    {
        uint8_t found = 0;
        check(ss_b200_search_sharded(c_, s.raw(), h.raw(), &found, nullptr));
        return found != 0;
    }
    template <bool STRICT>
    std::optional<size_t> find_in(const detail::SearcherImpl<STRICT> &s, const ShardedHaystack &h)
    {
        size_t off = SS_B200_NPOS;
        check(ss_b200_find_sharded(c_, s.raw(), h.raw(), &off));
        return off == SS_B200_NPOS ? std::nullopt : std::optional<size_t>(off);
    }
    // searcher.search_in(&[u8]) with ONE host slice striped over all devices / PCIe links
    template <bool STRICT>
    bool search_in(const detail::SearcherImpl<STRICT> &s, Bytes haystack)
    {
        uint8_t found = 0;
        check(ss_b200_search_in_host_multi(c_, s.raw(), haystack.ptr, haystack.len, &found));
        return found != 0;
    }
    template <bool STRICT>
    std::optional<size_t> find_in(const detail::SearcherImpl<STRICT> &s, Bytes haystack)
    {
        size_t off = SS_B200_NPOS;
        check(ss_b200_find_in_host_multi(c_, s.raw(), haystack.ptr, haystack.len, &off));
        return off == SS_B200_NPOS ? std::nullopt : std::optional<size_t>(off);
    }
    // many-haystack mode: flags[h] = searcher.search_in(haystack h)
    ContextHaystackSet upload_haystack_set(const uint8_t *blob, const uint64_t *offsets, size_t n) const
    {
        ss_b200_ctx_hayset *h = nullptr;
        check(ss_b200_ctx_hayset_upload(c_, blob, offsets, n, &h));
        return ContextHaystackSet(h);
    }
    template <bool STRICT>
    std::vector<uint8_t> search_in(const detail::SearcherImpl<STRICT> &s, const ContextHaystackSet &set)
    {
        std::vector<uint8_t> flags(set.len());
        check(ss_b200_ctx_hayset_search(c_, s.raw(), set.raw(), flags.data()));
        return flags;
    }
    ss_b200_ctx *raw() const { return c_; }

private:
    ss_b200_ctx *c_ = nullptr;
};

} // namespace sliceslice_b200
