/*
 * sliceslice_b200.h -- C ABI of the B200-native single-pattern substring search.
 *
 * Drop-in boundary for ONE path of cloudflare/sliceslice-rs: the inner loop of
 * sliceslice::x86::DynamicAvx2Searcher::search_in (Mula's two-anchor-byte filter
 * + memcmp verify).  Every entry point cites the reference interface it replaces
 * (file:line relative to the reference checkout).  Plain C types only: pointer +
 * length pairs, SIZE_MAX for "not found" -- the same convention the reference's
 * own FFI precedent uses (bench/sse4-strstr/src/wrapper.h:7, src/lib.rs:4-15).
 *
 * There is no CPU fallback: every search entry point runs hand-written sm_100a
 * CUDA kernels and returns SS_B200_E_CUDA when no usable device is present.
 *
 * Error convention (the reference panics at construction only, src/x86.rs:300,
 * :304, :473; search_in is infallible): no unwinding across the boundary; every
 * function returns an int status and the Rust/C++/Python shims turn
 * SS_B200_E_POSITION / SS_B200_E_EMPTY_NEEDLE into the reference's panics.
 *
 * Threading (reference: searchers are immutable plain data, Send + Sync,
 * src/x86.rs:266-271): a const searcher / haystack handle may be used from any
 * number of host threads concurrently; the synchronous calls keep their stream,
 * workspace and result slot in thread-local storage.
 */
#ifndef SLICESLICE_B200_H
#define SLICESLICE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SS_B200_ABI_VERSION 2

/* "not found" for host-side offsets: usize::MAX / std::string::npos
 * (bench/sse4-strstr/src/lib.rs:12-14). */
#define SS_B200_NPOS ((size_t)-1)
/* "not found" in a DEVICE result slot written by the *_async calls.  INT64_MAX so
 * that a min-reduction over ranks works on both signed and unsigned views (NCCL /
 * gloo have no bitwise OR; min over first offsets gives found AND the offset). */
#define SS_B200_DEVICE_NONE 0x7FFFFFFFFFFFFFFFull

enum ss_b200_status {
    SS_B200_OK = 0,
    SS_B200_E_POSITION = 1,     /* assert!(position < needle.size()) src/x86.rs:300; assert_eq!(position, 0) :473 */
    SS_B200_E_EMPTY_NEEDLE = 2, /* Avx2Searcher::new(empty) panics, src/x86.rs:285,300 (strict ctors only) */
    SS_B200_E_ARG = 3,          /* NULL handle / pointer */
    SS_B200_E_CUDA = 4,         /* CUDA runtime error or no device; see ss_b200_last_error() */
    SS_B200_E_NOMEM = 5,
    SS_B200_E_NCCL = 6          /* libnccl not loadable, or an NCCL call failed; see ss_b200_last_error() */
};

typedef struct ss_b200_searcher ss_b200_searcher; /* opaque, immutable after creation */
typedef struct ss_b200_haystack ss_b200_haystack; /* opaque device-resident haystack   */
typedef struct ss_b200_batch ss_b200_batch;       /* opaque device-resident word sets  */
typedef struct ss_b200_ctx ss_b200_ctx;           /* opaque multi-GPU context: every GPU of the box, one process */
typedef struct ss_b200_sharded ss_b200_sharded;   /* opaque haystack sharded over the devices of a context */
typedef struct ss_b200_ctx_hayset ss_b200_ctx_hayset; /* opaque set of haystacks partitioned over a context */

const char *ss_b200_strerror(int status);
/* Thread-local detail string of the last SS_B200_E_CUDA on this thread. */
const char *ss_b200_last_error(void);
int ss_b200_abi_version(void);

/* ------------------------------------------------------------------------- */
/* Searcher construction.                                                     */

/* DynamicAvx2Searcher::new(needle) -- src/x86.rs:454-459.
 * position = len - 1 (wrapping; ignored for the empty needle). Never fails on
 * the arguments; the empty needle is valid (variant N0, src/x86.rs:470). */
int ss_b200_searcher_new(const uint8_t *needle, size_t len, ss_b200_searcher **out);

/* DynamicAvx2Searcher::with_position(needle, position) -- src/x86.rs:468-493.
 * len == 0: position ignored. len == 1: position must be 0 (:473).
 * len >= 2: position < len (:300) else SS_B200_E_POSITION. */
int ss_b200_searcher_with_position(const uint8_t *needle, size_t len, size_t position, ss_b200_searcher **out);

/* Avx2Searcher::new / ::with_position -- src/x86.rs:282-287, :297-316.
 * Same as above except that the empty needle is SS_B200_E_EMPTY_NEEDLE and a
 * one-byte needle takes the generic two-anchor path with position 0. */
int ss_b200_searcher_new_strict(const uint8_t *needle, size_t len, ss_b200_searcher **out);
int ss_b200_searcher_with_position_strict(const uint8_t *needle, size_t len, size_t position,
                                          ss_b200_searcher **out);

/* Second-anchor selection.  The reference leaves `position` to the caller
 * (with_position, src/x86.rs:289-297, :461-468; rationale :252-255) and its result never depends on
 * it (src/lib.rs:375-378): only the candidate rate does.  ss_b200_rarest_position picks the index
 * p in [1, min(len - 1, 2032)] whose byte needle[p] is rarest under `hist` (256 counts, e.g. from
 * ss_b200_haystack_byte_histogram; NULL = a built-in text/binary background table); a position of 16 or
 * more is charged 17/16 (the staged scan reads two more vectors per chunk then); equal costs go to the
 * larger index, so a needle of equally frequent bytes keeps the reference's default len - 1.
 * len < 2 gives 0.  ss_b200_searcher_new_rarest = with_position(needle, that index). */
int ss_b200_rarest_position(const uint8_t *needle, size_t len, const uint64_t *hist, size_t *position);
int ss_b200_searcher_new_rarest(const uint8_t *needle, size_t len, const uint64_t *hist, ss_b200_searcher **out);

void ss_b200_searcher_free(ss_b200_searcher *s); /* Drop */
size_t ss_b200_searcher_needle_len(const ss_b200_searcher *s);
size_t ss_b200_searcher_position(const ss_b200_searcher *s); /* private Searcher::position, src/lib.rs:289-293 */

/* ------------------------------------------------------------------------- */
/* Haystack residency (new concern: the reference borrows a host &[u8]).      */

/* Copy `len` host bytes into HBM on the current device (owned by the handle). */
int ss_b200_haystack_upload(const uint8_t *host, size_t len, ss_b200_haystack **out);
/* Borrow `len` bytes already in device memory (any byte alignment).  The synchronous search calls
 * run on a stream owned by the library: the bytes must be complete (no write still pending on another
 * stream) when ss_b200_search_in / ss_b200_find_in is called.  ss_b200_find_in_device_async is the
 * stream-ordered entry.
 *
 * READ PRECONDITION (differs from the reference).  The reference never touches a byte outside the
 * borrowed slice: its last block is re-aligned to end exactly at haystack[len) (src/lib.rs:276-284,
 * rationale src/x86.rs:257-261).  The kernels here load whole 16-byte aligned words, i.e. they read
 * [align_down(dptr, 16), align_up(dptr + len, 16)): up to 15 bytes before and up to 15 bytes after the
 * slice.  Those bytes never influence a result (positions outside [0, len - k] are masked before use)
 * and the loads cannot fault: an aligned 16-byte word never crosses a page, so it lies in a page that
 * also holds a byte of the slice, and device (and pinned host) mappings are page-granular.  What the
 * caller must accept is that memory-checking tools (compute-sanitizer initcheck) may report those bytes
 * as read while uninitialised, and that a slice carved out of a larger buffer has its neighbours' edge
 * bytes loaded (never used).  This applies to every entry point that takes device memory. */
int ss_b200_haystack_from_device(const void *dptr, size_t len, ss_b200_haystack **out);
void ss_b200_haystack_free(ss_b200_haystack *h);
size_t ss_b200_haystack_len(const ss_b200_haystack *h);
const void *ss_b200_haystack_device_ptr(const ss_b200_haystack *h);

/* 256-bin byte histogram of device memory, for ss_b200_rarest_position.  It is a SAMPLE by default:
 * sample_bytes == 0 means 16 MiB; about sample_bytes are counted in evenly spaced 4 KiB granules
 * (granule g starts at byte g * floor(G / ceil(sample_bytes / 4096)) * 4096, G = ceil(len / 4096)), which
 * ranks the needle bytes of any haystack in ~20 us.  sample_bytes >= len counts every byte (exact; every
 * haystack of up to 16 MiB is counted exactly by default) -- for a long haystack that is a full pass at
 * ~1.8 TB/s (one shared-memory atomic per byte), a quarter of the scan's own rate, and buys the position
 * choice nothing.  The async form writes 256 uint64 to device memory in stream order. */
int ss_b200_byte_histogram_device_async(const void *dptr, size_t len, size_t sample_bytes, uint64_t *d_hist,
                                        void *stream);
int ss_b200_haystack_byte_histogram(const ss_b200_haystack *h, size_t sample_bytes, uint64_t hist[256]);

/* ------------------------------------------------------------------------- */
/* The hot call.                                                              */

/* DynamicAvx2Searcher::search_in(&self, haystack) -> bool -- src/x86.rs:523-525
 * (body :498-519 -> :356-376 -> src/lib.rs:253-287 -> :199-251).
 * *found = 1/0. Semantics (k = needle len, n = haystack len): k==0 -> true;
 * k==1 -> n>0 && byte present (src/lib.rs:130-136); n<k -> false;
 * n==k -> haystack==needle (src/x86.rs:357-359); else any i in [0,n-k]. */
int ss_b200_search_in(const ss_b200_searcher *s, const ss_b200_haystack *h, uint8_t *found);

/* Same scan, returning the index at which the reference's loop returns true --
 * always the leftmost occurrence (src/lib.rs:263-274 ascending chunks, :221
 * ascending bits, :242-244 first verified candidate) -- or SS_B200_NPOS.
 * Mirrors avx2_strstr_v2's contract (bench/sse4-strstr/src/wrapper.cpp:18-28). */
int ss_b200_find_in(const ss_b200_searcher *s, const ss_b200_haystack *h, size_t *offset);

/* search_in(&[u8]) with a HOST slice: the literal analogue of src/x86.rs:523.
 * A slice of up to 32 KiB (the reference's short-haystack regime) is copied into the calling thread's
 * mapped pinned buffer and scanned in place over PCIe: no DMA, no events -- and, for needles of up to 17
 * bytes, no launch either (the thread's resident kernel reads it, see ss_b200_set_sync_service).  Longer slices
 * stream to the device in chunks (a ring of three device buffers sized from the slice: an eighth of
 * it, between 4 and 64 MiB each; copy/scan overlap; the host feeds at most three chunks ahead of the
 * results it has seen, so a match stops the feeding) and are scanned there; PCIe-bound by
 * construction.  Pinned (cudaHostAlloc / cudaHostRegister) memory is copied directly, or -- short
 * pinned slices, and always with ss_b200_set_host_path(2, ..) -- read in place by the scan kernel.  A
 * pageable slice of 8 MiB or more is staged through a pinned ring that a pool of memcpy worker threads
 * fills in parallel.  See ss_b200_set_host_path. */
int ss_b200_search_in_host(const ss_b200_searcher *s, const uint8_t *host, size_t len, uint8_t *found);
int ss_b200_find_in_host(const ss_b200_searcher *s, const uint8_t *host, size_t len, size_t *offset);

/* The same call with ONE host slice striped over every device of a context: chunk i of the slice goes
 * to device i % ndev, every device runs its own copy/scan ring, so all PCIe links of the box carry the
 * slice at once; the first offsets are MIN-reduced in the library.  Same semantics and result as
 * ss_b200_find_in_host. */
int ss_b200_find_in_host_multi(ss_b200_ctx *ctx, const ss_b200_searcher *s, const uint8_t *host, size_t len,
                               size_t *offset);
int ss_b200_search_in_host_multi(ss_b200_ctx *ctx, const ss_b200_searcher *s, const uint8_t *host, size_t len,
                                 uint8_t *found);
/* What the last ss_b200_find_in_host_multi of this context did: bytes handed to cudaMemcpyAsync, chunks
 * submitted, chunk size, and data path (1 DMA ring, 2 in place, 3 short slice; +10 = pageable input
 * staged through the pinned ring).  Any pointer may be NULL. */
int ss_b200_ctx_last_host_stats(const ss_b200_ctx *ctx, uint64_t *h2d_bytes, uint64_t *chunks, uint64_t *chunk_bytes,
                                int *mode);

/* Stream-ordered scan of device memory, no host synchronisation.
 *   dptr, len     haystack bytes in device memory (any alignment)
 *   base_offset   global coordinate of dptr[0]; added to the reported offset
 *                 (shard r of a sharded haystack passes r * shard_len)
 *   start_limit   number of start positions to test, counted from dptr[0]; pass
 *                 SIZE_MAX for "all" (len - k + 1). A shard that carries a k-1
 *                 byte right halo passes its own shard_len here so that
 *                 positions owned by the next shard are not reported twice.
 *   workspace     16 bytes of device memory, ZERO when the call is enqueued; the
 *                 kernel leaves it zero again (self-resetting), so one
 *                 cudaMemset at allocation time is enough for any number of
 *                 searches issued back to back on one stream.
 *   d_result      device (or mapped pinned host) uint64 slot: receives
 *                 base_offset + first offset, or SS_B200_DEVICE_NONE.
 *   stream        cudaStream_t (as void*); NULL = legacy default stream.
 * This is the entry the roofline is measured on and the one the multi-GPU
 * sharded mode calls before its min-allreduce. */
int ss_b200_find_in_device_async(const ss_b200_searcher *s, const void *dptr, size_t len, uint64_t base_offset,
                                 size_t start_limit, void *workspace, uint64_t *d_result, void *stream);

/* Count mode, stream-ordered (SURVEY 8f "find-all / count"): *d_count receives the number of start
 * positions i < start_limit with hay[i..i+k) == needle -- every occurrence, overlapping ones included
 * (the reference stops at the first, src/lib.rs:242-244; this is the same scan without the early
 * return).  workspace: 32 bytes of device memory, zero when enqueued.  The empty needle is
 * SS_B200_E_ARG.  Shards add their counts (ncclSum). */
int ss_b200_count_in_device_async(const ss_b200_searcher *s, const void *dptr, size_t len, size_t start_limit,
                                  void *workspace, uint64_t *d_count, void *stream);

/* ------------------------------------------------------------------------- */
/* Sharded search with the exchange fused into the scan (multi-GPU, one process per GPU).
 * Alternative to "ss_b200_find_in_device_async + ncclAllReduce(min)": the scan's last CTA stores this
 * rank's result straight into every rank's mailbox (peer HBM over NVLink, CUDA IPC mappings), and a
 * one-warp kernel enqueued behind the scan takes the minimum once all `world` results have landed.
 *   ss_b200_mailbox_create   4 * world uint64 slots in device memory (cudaMalloc, IPC-exportable)
 *   ss_b200_ipc_export/open  64-byte CUDA IPC handle out / mapped peer pointer in
 *   mailboxes[world]         host array of device pointers: [rank] = own mailbox, others = opened handles
 *   seq                      search counter, identical on all ranks, incremented per search
 *   workspace                32 bytes, zero when enqueued
 *   d_result                 receives min over ranks of (base_offset + first offset) or SS_B200_DEVICE_NONE
 * All ranks must issue the same sequence of searches. */
int ss_b200_mailbox_create(int world, void **d_mailbox);
int ss_b200_mailbox_free(void *d_mailbox);
int ss_b200_ipc_export(const void *dptr, uint8_t handle_out[64]);
int ss_b200_ipc_open(const uint8_t handle[64], void **dptr_out);
int ss_b200_ipc_close(void *dptr);
int ss_b200_find_in_device_exchange_async(const ss_b200_searcher *s, const void *dptr, size_t len,
                                          uint64_t base_offset, size_t start_limit, void *workspace,
                                          void *const *mailboxes, int world, int rank, uint64_t seq,
                                          uint64_t *d_result, void *stream);

/* ------------------------------------------------------------------------- */
/* Multi-GPU from ONE process.  The reference binds plain C functions from a single
 * process (bench/sse4-strstr/build.rs:8-23, bench/sse4-strstr/src/lib.rs:4-15, wrapper.h:7); a host that
 * follows that pattern reaches every GPU of the box through a context.  A context is used by one thread
 * at a time (calls on one context are serialised). */

/* devices == NULL: devices 0 .. ndev-1; ndev <= 0: every visible device.  Creates, per device, the
 * streams, workspace and mapped result slot of the synchronous calls, and enables peer access between
 * all pairs. */
int ss_b200_ctx_create(int ndev, const int *devices, ss_b200_ctx **out);
void ss_b200_ctx_free(ss_b200_ctx *ctx);
int ss_b200_ctx_device_count(const ss_b200_ctx *ctx);
int ss_b200_ctx_device(const ss_b200_ctx *ctx, int i); /* CUDA ordinal of lane i, -1 if out of range */

/* How ss_b200_search_sharded MIN-reduces the per-shard first offsets (NCCL has no bitwise OR; the
 * minimum over first offsets gives found AND the leftmost offset):
 *   HOST  every scan publishes into its own mapped host word; the calling thread takes the minimum
 *         (default: nothing but the scans on the critical path)
 *   PEER  the scan's last CTA stores its result into every device's mailbox (peer HBM over NVLink),
 *         a one-warp kernel per device takes the minimum: the result is also complete on every GPU
 *   NCCL  ncclAllReduce(ncclMin, ncclUint64, count 1) over communicators made with ncclCommInitAll;
 *         libnccl.so.2 is loaded with dlopen on first use -- SS_B200_E_NCCL if that fails */
enum ss_b200_exchange { SS_B200_EXCHANGE_HOST = 0, SS_B200_EXCHANGE_PEER = 1, SS_B200_EXCHANGE_NCCL = 2 };
int ss_b200_ctx_set_exchange(ss_b200_ctx *ctx, int kind);
int ss_b200_ctx_nccl_version(int *version); /* e.g. 22809; SS_B200_E_NCCL when libnccl cannot be loaded */

/* One haystack as contiguous shards of start positions, shard d on device d of the context: shard d
 * owns positions [d*per, (d+1)*per), per = ceil(len / ndev) rounded up to 16, and holds `halo` more
 * bytes (right halo only: start positions are independent, src/lib.rs:263-274).  A needle of up to
 * halo + 1 bytes can be searched; longer ones are SS_B200_E_ARG. */
int ss_b200_sharded_upload(const ss_b200_ctx *ctx, const uint8_t *host, size_t len, size_t halo,
                           ss_b200_sharded **out);
/* Borrow shards already in device memory: dptrs[d] (on device d of the context) holds spans[d] bytes
 * starting at global byte owned[0] + .. + owned[d-1] and owns the first owned[d] start positions. */
int ss_b200_sharded_from_device(const ss_b200_ctx *ctx, const void *const *dptrs, const size_t *owned,
                                const size_t *spans, ss_b200_sharded **out);
void ss_b200_sharded_free(ss_b200_sharded *sh);
size_t ss_b200_sharded_len(const ss_b200_sharded *sh);
int ss_b200_sharded_shard(const ss_b200_sharded *sh, int i, const void **dptr, size_t *start, size_t *owned,
                          size_t *span);

/* DynamicAvx2Searcher::search_in over the sharded haystack (src/x86.rs:523-525): every device scans
 * its shard concurrently, the first offsets are MIN-reduced by the context's exchange.  *found = 1/0;
 * *global_offset (nullable) = leftmost occurrence in the whole haystack or SS_B200_NPOS. */
int ss_b200_search_sharded(ss_b200_ctx *ctx, const ss_b200_searcher *s, const ss_b200_sharded *sh, uint8_t *found,
                           size_t *global_offset);
int ss_b200_find_sharded(ss_b200_ctx *ctx, const ss_b200_searcher *s, const ss_b200_sharded *sh, size_t *offset);

/* Many-haystack mode over the devices of a context: the set (CSR: blob + n+1 uint64 offsets, host
 * memory) is partitioned into contiguous index ranges of balanced bytes, one per device; every
 * haystack lives on exactly one device, so the OR over devices is a gather of disjoint flag slices.
 * flags[h] (host, n bytes) = search_in(haystack h). */
int ss_b200_ctx_hayset_upload(const ss_b200_ctx *ctx, const uint8_t *blob, const uint64_t *offsets, size_t n,
                              ss_b200_ctx_hayset **out);
void ss_b200_ctx_hayset_free(ss_b200_ctx_hayset *hs);
size_t ss_b200_ctx_hayset_len(const ss_b200_ctx_hayset *hs);
int ss_b200_ctx_hayset_part(const ss_b200_ctx_hayset *hs, int i, size_t *lo, size_t *hi);
int ss_b200_ctx_hayset_search(ss_b200_ctx *ctx, const ss_b200_searcher *s, const ss_b200_ctx_hayset *hs,
                              uint8_t *flags);

/* Many-haystack mode, stream-ordered: ONE needle against a device-resident SET of haystacks in a
 * single pass at the long-scan rate (the set is scanned as one blob; a match counts for haystack h
 * only if it lies wholly inside it).  Per haystack the result is search_in() of src/x86.rs:523.
 *   d_blob, blob_len  concatenated haystack bytes in device memory
 *   d_offsets         n_haystacks+1 uint64 in device memory, d_offsets[0] == 0, d_offsets[n] == blob_len
 *   d_flags           n_haystacks uint8 in device memory: set to 1/0 per haystack
 *   workspace         32 bytes of device memory, zero when enqueued (left zero again)
 * Multi-GPU, one process per GPU: each rank holds a subset of the haystacks, packs its flags into its
 * bit range of one global bitmap (ss_b200_pack_flags_async) and the bitmaps are OR-ed with
 * ncclAllReduce(ncclSum, ncclUint32) -- disjoint bits, and NCCL has no bitwise OR.  One process, all GPUs:
 * ss_b200_ctx_hayset_search gathers the disjoint flag slices. */
int ss_b200_search_many_async(const ss_b200_searcher *s, const void *d_blob, const uint64_t *d_offsets,
                              size_t n_haystacks, size_t blob_len, uint8_t *d_flags, void *workspace, void *stream);

/* The same mode over a PREPARED set ("construct once, search many times", the reference's searcher
 * pattern applied to the haystack side).  ss_b200_hayset_create borrows d_blob / d_offsets (they must
 * outlive the set and stay unchanged) and builds, in stream order, one lookup hint per 4 KiB of blob --
 * the index of the haystack holding that byte -- so that the hit path finds the haystack of a match with
 * one or two probes instead of a binary search over the whole offset table.  Results are identical to
 * ss_b200_search_many_async; only needles that occur in many haystacks run faster.  The search must be
 * ordered after the create call's stream work (same stream, or an event). */
typedef struct ss_b200_hayset ss_b200_hayset;
int ss_b200_hayset_create(const void *d_blob, const uint64_t *d_offsets, size_t n_haystacks, size_t blob_len,
                          void *stream, ss_b200_hayset **out);
void ss_b200_hayset_free(ss_b200_hayset *hs);
size_t ss_b200_hayset_len(const ss_b200_hayset *hs);
int ss_b200_hayset_search_async(const ss_b200_searcher *s, const ss_b200_hayset *hs, uint8_t *d_flags,
                                void *workspace, void *stream);

/* Bit-packed flags for the cross-GPU OR of the many-haystack mode.  Rank r holds haystacks
 * [first_bit, first_bit + n) of a global set of total_bits haystacks: the call zeroes the whole bitmap
 * d_words (ceil(total_bits / 32) uint32) and sets bit h of it for every local flag that is non-zero.
 * The ranks' bit ranges are disjoint, so ncclAllReduce(ncclSum, ncclUint32) over the bitmaps is the
 * bitwise OR NCCL lacks (/usr/include/nccl.h: sum, prod, max, min, avg) at 1 bit per haystack. */
int ss_b200_pack_flags_async(const uint8_t *d_flags, size_t n, size_t first_bit, uint32_t *d_words, size_t total_bits,
                             void *stream);

/* ------------------------------------------------------------------------- */
/* Batched modes (north-star "batched many-haystack mode"; workloads:
 * bench/benches/i386.rs:118-131 short sweep, :246-257 all needles over one
 * haystack).  Word sets are CSR: blob + (count+1) uint64 offsets, host memory. */

int ss_b200_batch_create(const uint8_t *needle_blob, const uint64_t *needle_off, size_t n_needles,
                         const uint8_t *hay_blob, const uint64_t *hay_off, size_t n_haystacks,
                         ss_b200_batch **out);
void ss_b200_batch_free(ss_b200_batch *b);

/* Explicit pair list: pair p searches needle pair_needle[p] in haystack
 * pair_hay[p] with DynamicAvx2Searcher::new semantics.  Outputs (host, either
 * may be NULL): bitmap bit p (LSB-first in uint32 words) = search_in();
 * offsets[p] = first offset or UINT64_MAX. */
int ss_b200_batch_search_pairs(const ss_b200_batch *b, const uint32_t *pair_needle, const uint32_t *pair_hay,
                               size_t n_pairs, uint32_t *bitmap, uint64_t *offsets);

/* Triangular rule of the reference's short-haystack bench (needle i against
 * every haystack j >= i, both sets being the same length-sorted word list):
 * pair index p = i*W - i*(i-1)/2 + (j-i).  bitmap must hold ceil(W(W+1)/2 / 32)
 * words.  *matches (nullable) receives the popcount. */
int ss_b200_batch_search_triangular(const ss_b200_batch *b, uint32_t *bitmap, uint64_t *matches);

/* Every needle of the batch over ONE long device-resident haystack in a single
 * launch (each haystack tile is staged once and tested against the whole needle
 * table).  offsets[w] = first offset or UINT64_MAX (host array, n_needles). */
int ss_b200_batch_find_all_in(const ss_b200_batch *b, const ss_b200_haystack *h, uint64_t *offsets);

/* Stream-ordered forms of the three batched searches: inputs and outputs in DEVICE memory, enqueued on
 * `stream` (cudaStream_t as void*, NULL = legacy default stream), no host synchronisation and no state
 * shared between calls -- any number of streams may use one batch handle concurrently.  (The
 * synchronous forms above run on the calling thread's own stream and copy the results back.)
 *   pairs       d_pair_needle / d_pair_hay: n_pairs uint32 indices, which must be in range;
 *               d_bitmap (nullable) ceil(n_pairs / 32) words, d_offsets (nullable) n_pairs uint64
 *   triangular  d_bitmap ceil(W(W+1)/2 / 32) words, d_matches one uint64
 *   find_all    dptr/len: the haystack; d_offsets n_needles uint64 (first offset or UINT64_MAX) */
int ss_b200_batch_search_pairs_async(const ss_b200_batch *b, const uint32_t *d_pair_needle,
                                     const uint32_t *d_pair_hay, size_t n_pairs, uint32_t *d_bitmap,
                                     uint64_t *d_offsets, void *stream);
int ss_b200_batch_search_triangular_async(const ss_b200_batch *b, uint32_t *d_bitmap, uint64_t *d_matches,
                                          void *stream);
int ss_b200_batch_find_all_in_device_async(const ss_b200_batch *b, const void *dptr, size_t len,
                                           uint64_t *d_offsets, void *stream);

/* ------------------------------------------------------------------------- */
/* Synthetic inputs of BASELINE configs 2'/4/5, generated in HBM so that multi-
 * GiB haystacks never cross PCIe.  Bit-identical CPU copies live in oracle/.  */

/* dst[t] = byte(global_start + t), byte(i) = (splitmix64(seed ^ (i>>3)) >> 8(i&7)) & 0xFF, 0xFF -> 0x00 */
int ss_b200_fill_random(void *d_dst, size_t len, uint64_t global_start, uint64_t seed, void *stream);
/* dst[t] = src[(global_start + t) % src_len]; src is device memory. */
int ss_b200_fill_tiled(void *d_dst, size_t len, uint64_t global_start, const void *d_src, size_t src_len,
                       void *stream);

/* ------------------------------------------------------------------------- */
/* Tuning / introspection (bench and tests; not part of the reference surface). */

/* Kernel variant for the long scan: 0 = auto, 1 = direct 16-byte LDG,
 * 2 = TMA bulk-staged shared-memory ring.  Process-wide. */
int ss_b200_set_scan_variant(int variant);
/* Overrides: ctas_per_sm (0 = auto), ... see DESIGN.md. */
int ss_b200_set_scan_tuning(int ctas_per_sm, int unroll, int tile_kib, int stages);
/* Extra anchors the long scan may fold into the two-anchor filter when it sees many candidates
 * (word-aligned needle offsets 4 and 8, or one of 1..3 for short needles): 0 = never, 1 or -1 =
 * adaptive per warp (default).  Changes the candidate rate only, never a result. */
int ss_b200_set_extra_anchors(int n);
/* The short-scan variant is launched with programmatic stream serialisation (back-to-back searches on
 * one stream overlap each launch with the previous scan): 1 = on (default), 0 = plain launches. */
int ss_b200_set_launch_pdl(int on);
/* Synchronous searches (ss_b200_find_in / ss_b200_search_in) over device-resident haystacks of up to
 * 4 MiB with needles of up to 64 bytes (and ss_b200_find_in_host / _search_in_host over host slices of up to
 * 32 KiB with needles of up to 17 bytes) do not launch a kernel per call: a resident grid per calling thread
 * (at most two per device) polls a request word in mapped pinned memory and answers into another, so a
 * call costs one PCIe round trip plus the scan -- the regime of the reference's per-needle loops
 * (bench/benches/i386.rs:252-256).  The grid retires by itself idle_us after the last call (default 100;
 * implicit synchronisations such as cudaFree wait at most that long) and at ss_b200_thread_release / thread
 * exit.  on = 0: always launch (idle_us == 0 keeps the current value). */
int ss_b200_set_sync_service(int on, int idle_us);
/* Host-slice path (ss_b200_find_in_host / _multi):
 *   mode          0 auto (pinned slices of up to 2 MiB per device in place, else the DMA ring), 1 always the DMA ring,
 *                 2 pinned input read in place by the direct-load kernel, 3 in place by the TMA kernel
 *   chunk_mib     0 auto (an eighth of a device's share of the slice, 4..64 MiB), else MiB per chunk
 *   copy_threads  memcpy workers that stage pageable input: -1 auto (min(15, cores-1)), 0 = hand pageable
 *                 memory to the driver.  Staged input goes through ONE device even in a multi-device
 *                 context: it is bound by the staging copy, not by a PCIe link */
int ss_b200_set_host_path(int mode, int chunk_mib, int copy_threads);
/* Measured host->device copy bandwidth of the current device: `bytes` of pinned memory, best of
 * `reps` cudaMemcpyAsync (the PCIe ceiling the host-slice path is reported against). */
int ss_b200_measure_h2d(size_t bytes, int reps, double *gb_per_s);
/* The synchronous calls keep a per-thread, per-device lane (two streams, a 64-byte workspace, a mapped
 * result word and -- after a host-slice search -- a staging ring sized from the slices searched).  It is
 * released when the thread exits; ss_b200_thread_release() releases it now (the next call rebuilds what
 * it needs).  ss_b200_thread_footprint reports what the calling thread currently holds. */
int ss_b200_thread_release(void);
int ss_b200_thread_footprint(size_t *device_bytes, size_t *pinned_bytes);
/* Number of kernel launches issued by this library in this process so far. */
uint64_t ss_b200_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* SLICESLICE_B200_H */
