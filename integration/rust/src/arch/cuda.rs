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
#[repr(C)]
pub struct IrisCluster {
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
    // Arch-level grids: every pair of `a` and `b` vectors in one call (the criterion grid of arch/mod.rs:22-72).
    fn iris_dot_u16_batch(device: c_int, a: *const u16, n_a: u32, b: *const u16, n_b: u64, out: *mut u16) -> c_int;
    fn iris_dot_bool_batch(device: c_int, a: *const u64, n_a: u32, b: *const u64, n_b: u64, out: *mut u16) -> c_int;
    // One database row-sharded over the GPUs of the box.
    fn iris_cluster_create(
        devices: *const c_int,
        n_devices: u32,
        capacity_rows: u64,
        flags: u32,
        out: *mut *mut IrisCluster,
    ) -> c_int;
    fn iris_cluster_destroy(c: *mut IrisCluster) -> c_int;
    fn iris_cluster_len(c: *const IrisCluster, n_shards: *mut u32, n_shares: *mut u64, n_masks: *mut u64) -> c_int;
    fn iris_cluster_load_files(c: *mut IrisCluster, shares_path: *const c_char, masks_path: *const c_char) -> c_int;
    fn iris_cluster_load_rows(c: *mut IrisCluster, shares: *const u16, masks: *const u64, n: u64) -> c_int;
    fn iris_cluster_match(
        c: *mut IrisCluster,
        query: *const u16,
        query_mask: *const u64,
        distances_out: *mut u16,
        denominators_out: *mut u16,
    ) -> c_int;
    fn iris_cluster_match_template(
        c: *mut IrisCluster,
        pattern: *const u64,
        mask: *const u64,
        distances_out: *mut u16,
        denominators_out: *mut u16,
    ) -> c_int;
    fn iris_cluster_search(
        c: *mut IrisCluster,
        templates: *const u64,
        num_queries: u32,
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

/// The criterion grid of arch/mod.rs:46-72 in one call: `out[i][j] = dot_u16(a[j], b[i])`.
pub fn dot_u16_grid(a: &[[u16; BITS]], b: &[[u16; BITS]]) -> Vec<u16> {
    let mut out = vec![0u16; a.len() * b.len()];
    check(unsafe {
        iris_dot_u16_batch(0, a.as_ptr().cast(), a.len() as u32, b.as_ptr().cast(), b.len() as u64, out.as_mut_ptr())
    });
    out
}

/// The criterion grid of arch/mod.rs:22-44 for `dot_bool`.
pub fn dot_bool_grid(a: &[[u64; LIMBS]], b: &[[u64; LIMBS]]) -> Vec<u16> {
    let mut out = vec![0u16; a.len() * b.len()];
    check(unsafe {
        iris_dot_bool_batch(0, a.as_ptr().cast(), a.len() as u32, b.as_ptr().cast(), b.len() as u64, out.as_mut_ptr())
    });
    out
}

/// One database row-sharded over several GPUs (main.rs:386-400 mmaps the whole file; here every GPU holds a
/// contiguous block of rows in HBM and the library runs one host thread + stream per GPU).
pub struct Cluster(*mut IrisCluster);
unsafe impl Send for Cluster {}

impl Cluster {
    pub fn new(devices: &[i32], capacity_rows: u64, flags: u32) -> Self {
        let mut h = ptr::null_mut();
        check(unsafe { iris_cluster_create(devices.as_ptr(), devices.len() as u32, capacity_rows, flags, &mut h) });
        Self(h)
    }

    /// `mpc.share-i` / `mpc.masks` as written by `prepare` (main.rs:337-371); every GPU reads its own block.
    pub fn load_files(&self, shares: Option<&CStr>, masks: Option<&CStr>) {
        check(unsafe {
            iris_cluster_load_files(
                self.0,
                shares.map_or(ptr::null(), |p| p.as_ptr()),
                masks.map_or(ptr::null(), |p| p.as_ptr()),
            )
        });
    }

    pub fn len(&self) -> usize {
        let (mut s, mut m, mut n) = (0u64, 0u64, 0u32);
        check(unsafe { iris_cluster_len(self.0, &mut n, &mut s, &mut m) });
        s.max(m) as usize
    }

    /// The participant's request (main.rs:419-431): distances of the whole database for one template.
    pub fn distances(&self, template: &Template, out: &mut [[u16; 31]]) {
        check(unsafe {
            iris_cluster_match_template(
                self.0,
                template.pattern.0.as_ptr(),
                template.mask.0.as_ptr(),
                out.as_mut_ptr().cast(),
                ptr::null_mut(),
            )
        });
    }

    /// The coordinator's side (main.rs:510-516): denominators of the whole database for one query mask.
    pub fn denominators(&self, mask: &Bits, out: &mut [[u16; 31]]) {
        check(unsafe { iris_cluster_match(self.0, ptr::null(), mask.0.as_ptr(), ptr::null_mut(), out.as_mut_ptr().cast()) });
    }

    /// Whole search on the GPUs for a cluster holding plain encodings: (min distance, row) per template.
    pub fn search(&self, templates: &[Template]) -> Vec<(f64, usize)> {
        let mut d = vec![f64::INFINITY; templates.len()];
        let mut i = vec![u64::MAX; templates.len()];
        check(unsafe {
            iris_cluster_search(self.0, templates.as_ptr().cast(), templates.len() as u32, d.as_mut_ptr(), i.as_mut_ptr())
        });
        d.into_iter().zip(i.into_iter().map(|x| x as usize)).collect()
    }
}

impl Drop for Cluster {
    fn drop(&mut self) {
        unsafe { iris_cluster_destroy(self.0) };
    }
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
