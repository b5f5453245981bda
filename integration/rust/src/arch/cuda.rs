//! B200 backend for `mpc-iris-code` -- sits beside `generic.rs` and `sve.rs` in `src/arch/`.
//!
//! UNBUILT ARTEFACT: the image this repository is developed in has no Rust toolchain, so this file has
//! never been compiled.  It is the literal drop-in a maintainer adds to the reference crate
//! (`src/arch/mod.rs:1-5` gains `#[cfg(feature = "cuda")] mod cuda;`); every `extern "C"` item below is
//! declared in `include/iris_b200.h` and exercised from Python by this repository's GPU tests.
#![cfg(feature = "cuda")]
#![allow(unused)]
use crate::{bits::LIMBS, Bits, EncodedBits, Template, BITS};
use std::{
    ffi::CStr,
    ops::Range,
    os::raw::{c_char, c_int, c_void},
    ptr,
};

#[repr(C)]
pub struct IrisDb {
    _private: [u8; 0],
}
#[repr(C)]
pub struct IrisDistanceEngine {
    _private: [u8; 0],
}
#[repr(C)]
pub struct IrisMasksEngine {
    _private: [u8; 0],
}

pub const IRIS_DB_SHARES: u32 = 1;
pub const IRIS_DB_MASKS: u32 = 2;

#[link(name = "iris_b200")]
extern "C" {
    fn iris_last_error() -> *const c_char;
    fn iris_dot_u16(device: c_int, a: *const u16, b: *const u16, out: *mut u16) -> c_int;
    fn iris_dot_bool(device: c_int, a: *const u64, b: *const u64, out: *mut u16) -> c_int;
    fn iris_db_create(device: c_int, capacity_rows: u64, flags: u32, out: *mut *mut IrisDb) -> c_int;
    fn iris_db_destroy(db: *mut IrisDb) -> c_int;
    fn iris_db_append_shares(db: *mut IrisDb, rows: *const u16, n: u64) -> c_int;
    fn iris_db_append_masks(db: *mut IrisDb, rows: *const u64, n: u64) -> c_int;
    fn iris_db_load_shares_file(db: *mut IrisDb, path: *const c_char, first_row: u64, n_rows: u64) -> c_int;
    fn iris_db_load_masks_file(db: *mut IrisDb, path: *const c_char, first_row: u64, n_rows: u64) -> c_int;
    fn iris_distance_engine_new(device: c_int, query: *const u16, out: *mut *mut IrisDistanceEngine) -> c_int;
    fn iris_distance_engine_new_from_template(
        device: c_int,
        pattern: *const u64,
        mask: *const u64,
        out: *mut *mut IrisDistanceEngine,
    ) -> c_int;
    fn iris_distance_engine_free(e: *mut IrisDistanceEngine) -> c_int;
    fn iris_distance_engine_batch_process(
        e: *mut IrisDistanceEngine,
        out: *mut u16,
        out_len: u64,
        db: *const u16,
        db_len: u64,
    ) -> c_int;
    fn iris_distance_engine_batch_process_resident(
        e: *mut IrisDistanceEngine,
        out: *mut u16,
        out_len: u64,
        db: *mut IrisDb,
        row_begin: u64,
        row_end: u64,
    ) -> c_int;
    fn iris_masks_engine_new(device: c_int, query: *const u64, out: *mut *mut IrisMasksEngine) -> c_int;
    fn iris_masks_engine_free(e: *mut IrisMasksEngine) -> c_int;
    fn iris_masks_engine_batch_process(
        e: *mut IrisMasksEngine,
        out: *mut u16,
        out_len: u64,
        db: *const u64,
        db_len: u64,
    ) -> c_int;
    fn iris_masks_engine_batch_process_resident(
        e: *mut IrisMasksEngine,
        out: *mut u16,
        out_len: u64,
        db: *mut IrisDb,
        row_begin: u64,
        row_end: u64,
    ) -> c_int;
    fn iris_combine_min(
        device: c_int,
        distance_shares: *const *const u16,
        parties: u32,
        denominators: *const u16,
        n: u64,
        index_base: u64,
        distances_out: *mut f64,
        min_distance: *mut f64,
        min_index: *mut u64,
    ) -> c_int;
    // Page-locked result buffers (instead of a fresh Vec per chunk, main.rs:429, 514) and device buffers.
    #[allow(dead_code)]
    fn iris_host_alloc(bytes: u64, out: *mut *mut c_void) -> c_int;
    #[allow(dead_code)]
    fn iris_host_free(ptr: *mut c_void) -> c_int;
    #[allow(dead_code)]
    fn iris_device_alloc(device: c_int, bytes: u64, out: *mut *mut c_void) -> c_int;
    #[allow(dead_code)]
    fn iris_device_free(device: c_int, ptr: *mut c_void) -> c_int;
}

fn check(rc: c_int) {
    if rc != 0 {
        // The reference's only failure mode on this path is a panic (`assert_eq!`, lib.rs:43,70).
        let msg = unsafe { CStr::from_ptr(iris_last_error()) }.to_string_lossy();
        panic!("iris_b200: {msg}");
    }
}

/// Same signature as `generic::dot_u16` (generic.rs:11) so `pub use cuda::{dot_bool, dot_u16}` type-checks.
pub fn dot_u16(a: &[u16; BITS], b: &[u16; BITS]) -> u16 {
    let mut out = 0u16;
    check(unsafe { iris_dot_u16(0, a.as_ptr(), b.as_ptr(), &mut out) });
    out
}

/// Same signature as `generic::dot_bool` (generic.rs:4).
pub fn dot_bool(a: &[u64; LIMBS], b: &[u64; LIMBS]) -> u16 {
    let mut out = 0u16;
    check(unsafe { iris_dot_bool(0, a.as_ptr(), b.as_ptr(), &mut out) });
    out
}

/// HBM-resident database: replaces the `Arc<Mmap>` + `cast_slice` of main.rs:389-391 / 458-461.
pub struct Database(*mut IrisDb);
unsafe impl Send for Database {}

impl Database {
    pub fn from_shares(device: i32, rows: &[EncodedBits]) -> Self {
        let mut h = ptr::null_mut();
        check(unsafe { iris_db_create(device, rows.len().max(1) as u64, IRIS_DB_SHARES, &mut h) });
        check(unsafe { iris_db_append_shares(h, rows.as_ptr().cast(), rows.len() as u64) });
        Self(h)
    }

    pub fn from_masks(device: i32, rows: &[Bits]) -> Self {
        let mut h = ptr::null_mut();
        check(unsafe { iris_db_create(device, rows.len().max(1) as u64, IRIS_DB_MASKS, &mut h) });
        check(unsafe { iris_db_append_masks(h, rows.as_ptr().cast(), rows.len() as u64) });
        Self(h)
    }
}

impl Drop for Database {
    fn drop(&mut self) {
        unsafe { iris_db_destroy(self.0) };
    }
}

/// Drop-in for `DistanceEngine` (lib.rs:28-52).
pub struct DistanceEngine(*mut IrisDistanceEngine);
unsafe impl Send for DistanceEngine {}

impl DistanceEngine {
    pub fn new(query: &EncodedBits) -> Self {
        let mut h = ptr::null_mut();
        check(unsafe { iris_distance_engine_new(0, query.0.as_ptr(), &mut h) });
        Self(h)
    }

    /// `DistanceEngine::new(&encode(&template))` (main.rs:427) with `encode` done on the device.
    pub fn from_template(template: &Template) -> Self {
        let mut h = ptr::null_mut();
        check(unsafe {
            iris_distance_engine_new_from_template(0, template.pattern.0.as_ptr(), template.mask.0.as_ptr(), &mut h)
        });
        Self(h)
    }

    /// Literal signature of lib.rs:42.
    pub fn batch_process(&self, out: &mut [[u16; 31]], db: &[EncodedBits]) {
        check(unsafe {
            iris_distance_engine_batch_process(
                self.0,
                out.as_mut_ptr().cast(),
                out.len() as u64,
                db.as_ptr().cast(),
                db.len() as u64,
            )
        });
    }

    /// The participant loop of main.rs:428-431 with the database resident in HBM.
    pub fn batch_process_resident(&self, out: &mut [[u16; 31]], db: &Database, rows: Range<u64>) {
        check(unsafe {
            iris_distance_engine_batch_process_resident(
                self.0,
                out.as_mut_ptr().cast(),
                out.len() as u64,
                db.0,
                rows.start,
                rows.end,
            )
        });
    }
}

impl Drop for DistanceEngine {
    fn drop(&mut self) {
        unsafe { iris_distance_engine_free(self.0) };
    }
}

/// Drop-in for `MasksEngine` (lib.rs:55-79).
pub struct MasksEngine(*mut IrisMasksEngine);
unsafe impl Send for MasksEngine {}

impl MasksEngine {
    pub fn new(query: &Bits) -> Self {
        let mut h = ptr::null_mut();
        check(unsafe { iris_masks_engine_new(0, query.0.as_ptr(), &mut h) });
        Self(h)
    }

    /// Literal signature of lib.rs:69.
    pub fn batch_process(&self, out: &mut [[u16; 31]], db: &[Bits]) {
        check(unsafe {
            iris_masks_engine_batch_process(
                self.0,
                out.as_mut_ptr().cast(),
                out.len() as u64,
                db.as_ptr().cast(),
                db.len() as u64,
            )
        });
    }

    pub fn batch_process_resident(&self, out: &mut [[u16; 31]], db: &Database, rows: Range<u64>) {
        check(unsafe {
            iris_masks_engine_batch_process_resident(
                self.0,
                out.as_mut_ptr().cast(),
                out.len() as u64,
                db.0,
                rows.start,
                rows.end,
            )
        });
    }
}

impl Drop for MasksEngine {
    fn drop(&mut self) {
        unsafe { iris_masks_engine_free(self.0) };
    }
}

/// The coordinator's per-batch reduction (main.rs:597-621) on the device.
pub fn combine_min(shares: &[&[[u16; 31]]], denominators: &[[u16; 31]], index_base: u64) -> (f64, usize) {
    let ptrs: Vec<*const u16> = shares.iter().map(|s| s.as_ptr().cast()).collect();
    let (mut d, mut i) = (f64::INFINITY, u64::MAX);
    check(unsafe {
        iris_combine_min(
            0,
            ptrs.as_ptr(),
            ptrs.len() as u32,
            denominators.as_ptr().cast(),
            denominators.len() as u64,
            index_base,
            ptr::null_mut(),
            &mut d,
            &mut i,
        )
    });
    (d, i as usize)
}
