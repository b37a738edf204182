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

pub const SS_B200_OK: c_int = 0;
pub const SS_B200_E_POSITION: c_int = 1;
pub const SS_B200_E_EMPTY_NEEDLE: c_int = 2;

extern "C" {
    fn ss_b200_strerror(status: c_int) -> *const c_char;
    fn ss_b200_last_error() -> *const c_char;
    fn ss_b200_searcher_new(needle: *const u8, len: usize, out: *mut *mut RawSearcher) -> c_int;
    fn ss_b200_searcher_with_position(needle: *const u8, len: usize, position: usize, out: *mut *mut RawSearcher) -> c_int;
    fn ss_b200_searcher_new_strict(needle: *const u8, len: usize, out: *mut *mut RawSearcher) -> c_int;
    fn ss_b200_searcher_with_position_strict(needle: *const u8, len: usize, position: usize, out: *mut *mut RawSearcher) -> c_int;
    fn ss_b200_rarest_position(needle: *const u8, len: usize, hist: *const u64, position: *mut usize) -> c_int;
    fn ss_b200_haystack_byte_histogram(h: *const RawHaystack, sample_bytes: usize, hist: *mut u64) -> c_int;
    fn ss_b200_searcher_free(s: *mut RawSearcher);
    fn ss_b200_haystack_upload(host: *const u8, len: usize, out: *mut *mut RawHaystack) -> c_int;
    fn ss_b200_haystack_from_device(dptr: *const c_void, len: usize, out: *mut *mut RawHaystack) -> c_int;
    fn ss_b200_haystack_free(h: *mut RawHaystack);
    fn ss_b200_search_in(s: *const RawSearcher, h: *const RawHaystack, found: *mut u8) -> c_int;
    fn ss_b200_find_in(s: *const RawSearcher, h: *const RawHaystack, offset: *mut usize) -> c_int;
    fn ss_b200_search_in_host(s: *const RawSearcher, host: *const u8, len: usize, found: *mut u8) -> c_int;
    fn ss_b200_find_in_host(s: *const RawSearcher, host: *const u8, len: usize, offset: *mut usize) -> c_int;
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
    /// 256 byte counts of the haystack (`sample_bytes == 0`: every byte), for `with_rarest_position`.
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
