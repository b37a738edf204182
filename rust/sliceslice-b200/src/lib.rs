//! B200 backend for sliceslice: same surface as `sliceslice::x86::DynamicAvx2Searcher`
//! (`new`, `with_position`, `search_in`, `inlined_search_in`), body replaced by the C ABI of
//! `include/sliceslice_b200.h`.  NOT COMPILED IN THIS ENVIRONMENT (no Rust toolchain).
//!
//! ```ignore
//! use sliceslice_b200::DynamicB200Searcher;
//! let searcher = DynamicB200Searcher::new(b"ipsum".to_owned());
//! assert!(searcher.search_in(b"Lorem ipsum dolor sit amet"));
//! ```
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)]
pub struct RawSearcher {
    _p: [u8; 0],
}
#[repr(C)]
pub struct RawHaystack {
    _p: [u8; 0],
}
#[repr(C)]
pub struct RawCtx {
    _p: [u8; 0],
}
#[repr(C)]
pub struct RawSharded {
    _p: [u8; 0],
}
#[repr(C)]
pub struct RawCtxHayset {
    _p: [u8; 0],
}
#[repr(C)]
pub struct RawBatch {
    _p: [u8; 0],
}
#[repr(C)]
pub struct RawHayset {
    _p: [u8; 0],
}

pub const SS_B200_OK: c_int = 0;
pub const SS_B200_E_POSITION: c_int = 1;
pub const SS_B200_E_EMPTY_NEEDLE: c_int = 2;
/// How `Context::search_in` MIN-reduces the per-shard first offsets (`enum ss_b200_exchange`).
#[repr(i32)]
#[derive(Clone, Copy, Debug, PartialEq, Eq)]
pub enum Exchange {
    Host = 0,
    Peer = 1,
    Nccl = 2,
}

use sys::*;

/// Every entry point of `include/sliceslice_b200.h`, as declared there (kept in step by
/// `tests/test_capi_cpu.py::test_rust_shim_declares_every_header_entry`).
pub mod sys {
    use super::{RawBatch, RawCtx, RawCtxHayset, RawHayset, RawHaystack, RawSearcher, RawSharded};
    use std::os::raw::{c_char, c_int, c_void};
    extern "C" {
        pub fn ss_b200_strerror(status: c_int) -> *const c_char;
        pub fn ss_b200_last_error() -> *const c_char;
        pub fn ss_b200_abi_version() -> c_int;
        pub fn ss_b200_searcher_new(needle: *const u8, len: usize, out: *mut *mut RawSearcher) -> c_int;
        pub fn ss_b200_searcher_with_position(needle: *const u8, len: usize, position: usize, out: *mut *mut RawSearcher) -> c_int;
        pub fn ss_b200_searcher_new_strict(needle: *const u8, len: usize, out: *mut *mut RawSearcher) -> c_int;
        pub fn ss_b200_searcher_with_position_strict(needle: *const u8, len: usize, position: usize, out: *mut *mut RawSearcher) -> c_int;
        pub fn ss_b200_rarest_position(needle: *const u8, len: usize, hist: *const u64, position: *mut usize) -> c_int;
        pub fn ss_b200_searcher_new_rarest(needle: *const u8, len: usize, hist: *const u64, out: *mut *mut RawSearcher) -> c_int;
        pub fn ss_b200_searcher_free(s: *mut RawSearcher);
        pub fn ss_b200_searcher_needle_len(s: *const RawSearcher) -> usize;
        pub fn ss_b200_searcher_position(s: *const RawSearcher) -> usize;
        pub fn ss_b200_haystack_upload(host: *const u8, len: usize, out: *mut *mut RawHaystack) -> c_int;
        pub fn ss_b200_haystack_from_device(dptr: *const c_void, len: usize, out: *mut *mut RawHaystack) -> c_int;
        pub fn ss_b200_haystack_free(h: *mut RawHaystack);
        pub fn ss_b200_haystack_len(h: *const RawHaystack) -> usize;
        pub fn ss_b200_haystack_device_ptr(h: *const RawHaystack) -> *const c_void;
        pub fn ss_b200_byte_histogram_device_async(dptr: *const c_void, len: usize, sample_bytes: usize, d_hist: *mut u64, stream: *mut c_void) -> c_int;
        pub fn ss_b200_haystack_byte_histogram(h: *const RawHaystack, sample_bytes: usize, hist: *mut u64) -> c_int;
        pub fn ss_b200_search_in(s: *const RawSearcher, h: *const RawHaystack, found: *mut u8) -> c_int;
        pub fn ss_b200_find_in(s: *const RawSearcher, h: *const RawHaystack, offset: *mut usize) -> c_int;
        pub fn ss_b200_search_in_host(s: *const RawSearcher, host: *const u8, len: usize, found: *mut u8) -> c_int;
        pub fn ss_b200_find_in_host(s: *const RawSearcher, host: *const u8, len: usize, offset: *mut usize) -> c_int;
        pub fn ss_b200_find_in_host_multi(ctx: *mut RawCtx, s: *const RawSearcher, host: *const u8, len: usize, offset: *mut usize) -> c_int;
        pub fn ss_b200_search_in_host_multi(ctx: *mut RawCtx, s: *const RawSearcher, host: *const u8, len: usize, found: *mut u8) -> c_int;
        pub fn ss_b200_ctx_last_host_stats(ctx: *const RawCtx, h2d_bytes: *mut u64, chunks: *mut u64, chunk_bytes: *mut u64, mode: *mut c_int) -> c_int;
        pub fn ss_b200_find_in_device_async(s: *const RawSearcher, dptr: *const c_void, len: usize, base_offset: u64, start_limit: usize, workspace: *mut c_void, d_result: *mut u64, stream: *mut c_void) -> c_int;
        pub fn ss_b200_count_in_device_async(s: *const RawSearcher, dptr: *const c_void, len: usize, start_limit: usize, workspace: *mut c_void, d_count: *mut u64, stream: *mut c_void) -> c_int;
        pub fn ss_b200_mailbox_create(world: c_int, d_mailbox: *mut *mut c_void) -> c_int;
        pub fn ss_b200_mailbox_free(d_mailbox: *mut c_void) -> c_int;
        pub fn ss_b200_ipc_export(dptr: *const c_void, handle_out: *mut u8) -> c_int;
        pub fn ss_b200_ipc_open(handle: *const u8, dptr_out: *mut *mut c_void) -> c_int;
        pub fn ss_b200_ipc_close(dptr: *mut c_void) -> c_int;
        pub fn ss_b200_find_in_device_exchange_async(s: *const RawSearcher, dptr: *const c_void, len: usize, base_offset: u64, start_limit: usize, workspace: *mut c_void, mailboxes: *const *mut c_void, world: c_int, rank: c_int, seq: u64, d_result: *mut u64, stream: *mut c_void) -> c_int;
        pub fn ss_b200_ctx_create(ndev: c_int, devices: *const c_int, out: *mut *mut RawCtx) -> c_int;
        pub fn ss_b200_ctx_free(ctx: *mut RawCtx);
        pub fn ss_b200_ctx_device_count(ctx: *const RawCtx) -> c_int;
        pub fn ss_b200_ctx_device(ctx: *const RawCtx, i: c_int) -> c_int;
        pub fn ss_b200_ctx_set_exchange(ctx: *mut RawCtx, kind: c_int) -> c_int;
        pub fn ss_b200_ctx_nccl_version(version: *mut c_int) -> c_int;
        pub fn ss_b200_sharded_upload(ctx: *const RawCtx, host: *const u8, len: usize, halo: usize, out: *mut *mut RawSharded) -> c_int;
        pub fn ss_b200_sharded_from_device(ctx: *const RawCtx, dptrs: *const *const c_void, owned: *const usize, spans: *const usize, out: *mut *mut RawSharded) -> c_int;
        pub fn ss_b200_sharded_free(sh: *mut RawSharded);
        pub fn ss_b200_sharded_len(sh: *const RawSharded) -> usize;
        pub fn ss_b200_sharded_shard(sh: *const RawSharded, i: c_int, dptr: *mut *const c_void, start: *mut usize, owned: *mut usize, span: *mut usize) -> c_int;
        pub fn ss_b200_search_sharded(ctx: *mut RawCtx, s: *const RawSearcher, sh: *const RawSharded, found: *mut u8, global_offset: *mut usize) -> c_int;
        pub fn ss_b200_find_sharded(ctx: *mut RawCtx, s: *const RawSearcher, sh: *const RawSharded, offset: *mut usize) -> c_int;
        pub fn ss_b200_ctx_hayset_upload(ctx: *const RawCtx, blob: *const u8, offsets: *const u64, n: usize, out: *mut *mut RawCtxHayset) -> c_int;
        pub fn ss_b200_ctx_hayset_free(hs: *mut RawCtxHayset);
        pub fn ss_b200_ctx_hayset_len(hs: *const RawCtxHayset) -> usize;
        pub fn ss_b200_ctx_hayset_part(hs: *const RawCtxHayset, i: c_int, lo: *mut usize, hi: *mut usize) -> c_int;
        pub fn ss_b200_ctx_hayset_search(ctx: *mut RawCtx, s: *const RawSearcher, hs: *const RawCtxHayset, flags: *mut u8) -> c_int;
        pub fn ss_b200_search_many_async(s: *const RawSearcher, d_blob: *const c_void, d_offsets: *const u64, n_haystacks: usize, blob_len: usize, d_flags: *mut u8, workspace: *mut c_void, stream: *mut c_void) -> c_int;
        pub fn ss_b200_hayset_create(d_blob: *const c_void, d_offsets: *const u64, n_haystacks: usize, blob_len: usize, stream: *mut c_void, out: *mut *mut RawHayset) -> c_int;
        pub fn ss_b200_hayset_free(hs: *mut RawHayset);
        pub fn ss_b200_hayset_len(hs: *const RawHayset) -> usize;
        pub fn ss_b200_hayset_search_async(s: *const RawSearcher, hs: *const RawHayset, d_flags: *mut u8, workspace: *mut c_void, stream: *mut c_void) -> c_int;
        pub fn ss_b200_pack_flags_async(d_flags: *const u8, n: usize, first_bit: usize, d_words: *mut u32, total_bits: usize, stream: *mut c_void) -> c_int;
        pub fn ss_b200_batch_create(needle_blob: *const u8, needle_off: *const u64, n_needles: usize, hay_blob: *const u8, hay_off: *const u64, n_haystacks: usize, out: *mut *mut RawBatch) -> c_int;
        pub fn ss_b200_batch_free(b: *mut RawBatch);
        pub fn ss_b200_batch_search_pairs(b: *const RawBatch, pair_needle: *const u32, pair_hay: *const u32, n_pairs: usize, bitmap: *mut u32, offsets: *mut u64) -> c_int;
        pub fn ss_b200_batch_search_triangular(b: *const RawBatch, bitmap: *mut u32, matches: *mut u64) -> c_int;
        pub fn ss_b200_batch_find_all_in(b: *const RawBatch, h: *const RawHaystack, offsets: *mut u64) -> c_int;
        pub fn ss_b200_batch_search_pairs_async(b: *const RawBatch, d_pair_needle: *const u32, d_pair_hay: *const u32, n_pairs: usize, d_bitmap: *mut u32, d_offsets: *mut u64, stream: *mut c_void) -> c_int;
        pub fn ss_b200_batch_search_triangular_async(b: *const RawBatch, d_bitmap: *mut u32, d_matches: *mut u64, stream: *mut c_void) -> c_int;
        pub fn ss_b200_batch_find_all_in_device_async(b: *const RawBatch, dptr: *const c_void, len: usize, d_offsets: *mut u64, stream: *mut c_void) -> c_int;
        pub fn ss_b200_fill_random(d_dst: *mut c_void, len: usize, global_start: u64, seed: u64, stream: *mut c_void) -> c_int;
        pub fn ss_b200_fill_tiled(d_dst: *mut c_void, len: usize, global_start: u64, d_src: *const c_void, src_len: usize, stream: *mut c_void) -> c_int;
        pub fn ss_b200_set_scan_variant(variant: c_int) -> c_int;
        pub fn ss_b200_set_scan_tuning(ctas_per_sm: c_int, unroll: c_int, tile_kib: c_int, stages: c_int) -> c_int;
        pub fn ss_b200_set_extra_anchors(n: c_int) -> c_int;
        pub fn ss_b200_set_launch_pdl(on: c_int) -> c_int;
        pub fn ss_b200_set_sync_service(on: c_int, idle_us: c_int) -> c_int;
        pub fn ss_b200_set_host_path(mode: c_int, chunk_mib: c_int, copy_threads: c_int) -> c_int;
        pub fn ss_b200_measure_h2d(bytes: usize, reps: c_int, gb_per_s: *mut f64) -> c_int;
        pub fn ss_b200_thread_release() -> c_int;
        pub fn ss_b200_thread_footprint(device_bytes: *mut usize, pinned_bytes: *mut usize) -> c_int;
        pub fn ss_b200_launch_count() -> u64;
    }
}

/// Synchronous searches over short device-resident haystacks go through a resident kernel (no launch per
/// call) by default; `false` launches a kernel per call.  `idle_us == 0` keeps the current idle time.
pub fn set_sync_service(on: bool, idle_us: u32) {
    check(unsafe { ss_b200_set_sync_service(on as c_int, idle_us as c_int) });
}

/// Frees the calling thread's streams, result slot and staging ring (they are also freed at thread exit).
pub fn thread_release() {
    check(unsafe { ss_b200_thread_release() });
}

fn check(rc: c_int) {
    if rc == SS_B200_OK {
        return;
    }
    let msg = unsafe { std::ffi::CStr::from_ptr(ss_b200_strerror(rc)) }.to_string_lossy().into_owned();
    match rc {
        // the same panics the reference raises: assert!(position < needle.size()) src/x86.rs:300,
        // assert_eq!(position, 0) :473, empty needle for Avx2Searcher :285
        SS_B200_E_POSITION | SS_B200_E_EMPTY_NEEDLE => panic!("{}", msg),
        _ => {
            let detail = unsafe { std::ffi::CStr::from_ptr(ss_b200_last_error()) }.to_string_lossy().into_owned();
            panic!("sliceslice-b200: {} {}", msg, detail)
        }
    }
}

/// A haystack resident in B200 HBM.
pub struct DeviceHaystack(*mut RawHaystack);
unsafe impl Send for DeviceHaystack {}
unsafe impl Sync for DeviceHaystack {}
impl DeviceHaystack {
    pub fn upload(bytes: &[u8]) -> Self {
        let mut h = std::ptr::null_mut();
        check(unsafe { ss_b200_haystack_upload(bytes.as_ptr(), bytes.len(), &mut h) });
        DeviceHaystack(h)
    }
    /// # Safety: `dptr` must point to `len` bytes of device memory that outlive the handle.
    pub unsafe fn from_device(dptr: *const c_void, len: usize) -> Self {
        let mut h = std::ptr::null_mut();
        check(ss_b200_haystack_from_device(dptr, len, &mut h));
        DeviceHaystack(h)
    }
    /// 256 byte counts of the haystack (`sample_bytes == 0`: a 16 MiB sample, exact for shorter haystacks;
    /// `sample_bytes >= len`: every byte), for `with_rarest_position`.
    pub fn byte_histogram(&self, sample_bytes: usize) -> [u64; 256] {
        let mut hist = [0u64; 256];
        check(unsafe { ss_b200_haystack_byte_histogram(self.0, sample_bytes, hist.as_mut_ptr()) });
        hist
    }
}
impl Drop for DeviceHaystack {
    fn drop(&mut self) {
        unsafe { ss_b200_haystack_free(self.0) }
    }
}

/// Implemented by both searcher flavours so that `Context` takes either.
pub trait RawSearcherRef {
    fn raw(&self) -> *const RawSearcher;
}

macro_rules! searcher {
    ($name:ident, $new:ident, $with:ident, $doc:expr) => {
        #[doc = $doc]
        pub struct $name<N: AsRef<[u8]>> {
            raw: *mut RawSearcher,
            needle: N,
        }
        // immutable after construction, like the reference's searchers (src/x86.rs:266-271)
        unsafe impl<N: AsRef<[u8]> + Send> Send for $name<N> {}
        unsafe impl<N: AsRef<[u8]> + Sync> Sync for $name<N> {}
        impl<N: AsRef<[u8]>> $name<N> {
            pub fn new(needle: N) -> Self {
                let mut raw = std::ptr::null_mut();
                let b = needle.as_ref();
                check(unsafe { $new(b.as_ptr(), b.len(), &mut raw) });
                Self { raw, needle }
            }
            pub fn with_position(needle: N, position: usize) -> Self {
                let mut raw = std::ptr::null_mut();
                let b = needle.as_ref();
                check(unsafe { $with(b.as_ptr(), b.len(), position, &mut raw) });
                Self { raw, needle }
            }
            /// `with_position` with the index whose byte is rarest under `hist` (`None`: built-in
            /// background table).  Never changes a result (src/lib.rs:375-378), only the candidate rate.
            pub fn with_rarest_position(needle: N, hist: Option<&[u64; 256]>) -> Self {
                let mut position = 0usize;
                let b = needle.as_ref();
                let hp = hist.map_or(std::ptr::null(), |h| h.as_ptr());
                check(unsafe { ss_b200_rarest_position(b.as_ptr(), b.len(), hp, &mut position) });
                Self::with_position(needle, position)
            }
            pub fn needle(&self) -> &[u8] {
                self.needle.as_ref()
            }
            #[inline]
            pub fn inlined_search_in(&self, haystack: &[u8]) -> bool {
                let mut found = 0u8;
                check(unsafe { ss_b200_search_in_host(self.raw, haystack.as_ptr(), haystack.len(), &mut found) });
                found != 0
            }
            pub fn search_in(&self, haystack: &[u8]) -> bool {
                self.inlined_search_in(haystack)
            }
            pub fn find_in(&self, haystack: &[u8]) -> Option<usize> {
                let mut off = usize::MAX;
                check(unsafe { ss_b200_find_in_host(self.raw, haystack.as_ptr(), haystack.len(), &mut off) });
                if off == usize::MAX { None } else { Some(off) }
            }
            pub fn search_in_device(&self, haystack: &DeviceHaystack) -> bool {
                let mut found = 0u8;
                check(unsafe { ss_b200_search_in(self.raw, haystack.0, &mut found) });
                found != 0
            }
            pub fn find_in_device(&self, haystack: &DeviceHaystack) -> Option<usize> {
                let mut off = usize::MAX;
                check(unsafe { ss_b200_find_in(self.raw, haystack.0, &mut off) });
                if off == usize::MAX { None } else { Some(off) }
            }
        }
        impl<N: AsRef<[u8]>> RawSearcherRef for $name<N> {
            fn raw(&self) -> *const RawSearcher {
                self.raw
            }
        }
        impl<N: AsRef<[u8]>> Drop for $name<N> {
            fn drop(&mut self) {
                unsafe { ss_b200_searcher_free(self.raw) }
            }
        }
    };
}

searcher!(DynamicB200Searcher, ss_b200_searcher_new, ss_b200_searcher_with_position,
          "Drop-in for `sliceslice::x86::DynamicAvx2Searcher` (src/x86.rs:405-526).");
searcher!(B200Searcher, ss_b200_searcher_new_strict, ss_b200_searcher_with_position_strict,
          "Drop-in for `sliceslice::x86::Avx2Searcher` (src/x86.rs:266-383): empty needle panics.");

/// One haystack as contiguous shards of start positions, shard `d` on device `d` of a `Context`.
pub struct ShardedHaystack(*mut RawSharded);
impl ShardedHaystack {
    pub fn len(&self) -> usize {
        unsafe { ss_b200_sharded_len(self.0) }
    }
}
impl Drop for ShardedHaystack {
    fn drop(&mut self) {
        unsafe { ss_b200_sharded_free(self.0) }
    }
}

/// A set of haystacks partitioned over the devices of a `Context` (many-haystack mode).
pub struct ContextHaystackSet(*mut RawCtxHayset);
impl ContextHaystackSet {
    pub fn len(&self) -> usize {
        unsafe { ss_b200_ctx_hayset_len(self.0) }
    }
}
impl Drop for ContextHaystackSet {
    fn drop(&mut self) {
        unsafe { ss_b200_ctx_hayset_free(self.0) }
    }
}

/// Every GPU of the box from one process.  `&mut self` on the search calls: a context is used by one
/// thread at a time (the searchers themselves stay `Send + Sync`).
pub struct Context(*mut RawCtx);
unsafe impl Send for Context {}
impl Context {
    /// `ndev == 0`: every visible device.
    pub fn new(ndev: usize) -> Self {
        let mut c = std::ptr::null_mut();
        check(unsafe { ss_b200_ctx_create(ndev as c_int, std::ptr::null(), &mut c) });
        Context(c)
    }
    pub fn device_count(&self) -> usize {
        unsafe { ss_b200_ctx_device_count(self.0) as usize }
    }
    pub fn set_exchange(&mut self, kind: Exchange) {
        check(unsafe { ss_b200_ctx_set_exchange(self.0, kind as c_int) });
    }
    /// Shards of `len / ndev` start positions each, every shard with `halo` more bytes: needles of up to
    /// `halo + 1` bytes can be searched.
    pub fn upload_sharded(&self, bytes: &[u8], halo: usize) -> ShardedHaystack {
        let mut h = std::ptr::null_mut();
        check(unsafe { ss_b200_sharded_upload(self.0, bytes.as_ptr(), bytes.len(), halo, &mut h) });
        ShardedHaystack(h)
    }
    /// # Safety: `dptrs[d]` must point to `spans[d]` bytes on device `d` of the context, outliving the handle.
    pub unsafe fn sharded_from_device(&self, dptrs: &[*const c_void], owned: &[usize], spans: &[usize]) -> ShardedHaystack {
        assert!(dptrs.len() == self.device_count() && owned.len() == dptrs.len() && spans.len() == dptrs.len());
        let mut h = std::ptr::null_mut();
        check(ss_b200_sharded_from_device(self.0, dptrs.as_ptr(), owned.as_ptr(), spans.as_ptr(), &mut h));
        ShardedHaystack(h)
    }
    /// `searcher.search_in(haystack)` with the haystack sharded over the devices.
    pub fn search_in_sharded<S: RawSearcherRef>(&mut self, searcher: &S, haystack: &ShardedHaystack) -> bool {
        self.find_in_sharded(searcher, haystack).is_some()
    }
    pub fn find_in_sharded<S: RawSearcherRef>(&mut self, searcher: &S, haystack: &ShardedHaystack) -> Option<usize> {
        let (mut found, mut off) = (0u8, usize::MAX);
        check(unsafe { ss_b200_search_sharded(self.0, searcher.raw(), haystack.0, &mut found, &mut off) });
        if found != 0 { Some(off) } else { None }
    }
    /// `searcher.search_in(&[u8])` with ONE host slice striped over every device / PCIe link of the box.
    pub fn search_in<S: RawSearcherRef>(&mut self, searcher: &S, haystack: &[u8]) -> bool {
        let mut found = 0u8;
        check(unsafe { ss_b200_search_in_host_multi(self.0, searcher.raw(), haystack.as_ptr(), haystack.len(), &mut found) });
        found != 0
    }
    pub fn find_in<S: RawSearcherRef>(&mut self, searcher: &S, haystack: &[u8]) -> Option<usize> {
        let mut off = usize::MAX;
        check(unsafe { ss_b200_find_in_host_multi(self.0, searcher.raw(), haystack.as_ptr(), haystack.len(), &mut off) });
        if off == usize::MAX { None } else { Some(off) }
    }
    /// Many-haystack mode: `blob` = concatenated haystacks, `offsets` = n + 1 ascending byte offsets.
    pub fn upload_haystack_set(&self, blob: &[u8], offsets: &[u64]) -> ContextHaystackSet {
        assert!(!offsets.is_empty() && offsets[0] == 0 && *offsets.last().unwrap() as usize == blob.len());
        let mut h = std::ptr::null_mut();
        check(unsafe { ss_b200_ctx_hayset_upload(self.0, blob.as_ptr(), offsets.as_ptr(), offsets.len() - 1, &mut h) });
        ContextHaystackSet(h)
    }
    /// `flags[h] = searcher.search_in(haystack h)` for every haystack of the set.
    pub fn search_in_set<S: RawSearcherRef>(&mut self, searcher: &S, set: &ContextHaystackSet) -> Vec<bool> {
        let mut flags = vec![0u8; set.len()];
        check(unsafe { ss_b200_ctx_hayset_search(self.0, searcher.raw(), set.0, flags.as_mut_ptr()) });
        flags.into_iter().map(|f| f != 0).collect()
    }
}
impl Drop for Context {
    fn drop(&mut self) {
        unsafe { ss_b200_ctx_free(self.0) }
    }
}

/// Needle and haystack sets for the batched modes (the reference's two bench loops,
/// `bench/benches/i386.rs:118-131` and `:246-257`).  Sets are given as slices of byte strings.
pub struct Batch {
    raw: *mut RawBatch,
    n_needles: usize,
}
unsafe impl Send for Batch {}
fn csr(items: &[&[u8]]) -> (Vec<u8>, Vec<u64>) {
    let mut blob = Vec::new();
    let mut off = vec![0u64];
    for it in items {
        blob.extend_from_slice(it);
        off.push(blob.len() as u64);
    }
    (blob, off)
}
impl Batch {
    pub fn new(needles: &[&[u8]], haystacks: &[&[u8]]) -> Self {
        let (nb, no) = csr(needles);
        let (hb, ho) = csr(haystacks);
        let mut raw = std::ptr::null_mut();
        check(unsafe {
            ss_b200_batch_create(nb.as_ptr(), no.as_ptr(), needles.len(), hb.as_ptr(), ho.as_ptr(), haystacks.len(), &mut raw)
        });
        Batch { raw, n_needles: needles.len() }
    }
    /// Every needle of the batch over one device-resident haystack in a single launch: first offsets.
    pub fn find_all_in(&self, haystack: &DeviceHaystack) -> Vec<Option<usize>> {
        let mut out = vec![u64::MAX; self.n_needles];
        check(unsafe { ss_b200_batch_find_all_in(self.raw, haystack.0, out.as_mut_ptr()) });
        out.into_iter().map(|v| if v == u64::MAX { None } else { Some(v as usize) }).collect()
    }
    /// `needle i` in every `haystack j >= i` of one length-sorted list: (bitmap, number of matches).
    pub fn search_triangular(&self) -> (Vec<u32>, u64) {
        let w = self.n_needles as u64;
        let mut bitmap = vec![0u32; ((w * (w + 1) / 2 + 31) / 32) as usize];
        let mut matches = 0u64;
        check(unsafe { ss_b200_batch_search_triangular(self.raw, bitmap.as_mut_ptr(), &mut matches) });
        (bitmap, matches)
    }
}
impl Drop for Batch {
    fn drop(&mut self) {
        unsafe { ss_b200_batch_free(self.raw) }
    }
}

/// Kernel launches issued by the library in this process so far.
pub fn launch_count() -> u64 {
    unsafe { ss_b200_launch_count() }
}
