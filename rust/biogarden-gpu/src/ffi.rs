//! Raw `extern "C"` declarations of include/bgalign.h (the part the drop-in uses).
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)] pub struct bg_ctx { _private: [u8; 0] }

pub const BG_GLOBAL: i32 = 0;
pub const BG_LOCAL: i32 = 1;
pub const BG_SEMIGLOBAL: i32 = 2;
pub const BG_FITTING: i32 = 3;
pub const BG_OVERLAP: i32 = 4;

pub const BG_OK: c_int = 0;
pub const BG_EINVAL_RANGE: c_int = 1;
pub const BG_EINVAL_SIZE: c_int = 2;
pub const BG_ST_OK: u8 = 0;

#[repr(C)]
pub struct bg_batch {
    pub n_pairs: u64, pub residues: *const u8, pub seq_off: *const u64,
    /// 0 = one byte per residue, 2 / 5 = packed (bgalign.h BG_PACK_*); `alphabet` maps code -> residue byte
    pub packing: u32, pub reserved_: u32, pub alphabet: *const u8,
}

#[repr(C)]
pub struct bg_params {
    pub mode: i32, pub gap_open: i32, pub gap_extend: i32, pub flags: u32,
    pub table: *const i32, pub n_rows: i32, pub n_cols: i32,
    pub row_code: *const u8, pub col_code: *const u8,
}

#[repr(C)]
pub struct bg_result {
    pub n_pairs: u64, pub score: *mut i32, pub status: *mut u8, pub arena: *mut u8, pub off: *mut u64,
    pub owner_: *mut c_void,
}

/// Compact results (bgalign.h): one 2-bit op per alignment column instead of the aligned strings.
#[repr(C)]
pub struct bg_ops_result {
    pub n_pairs: u64, pub score: *mut i32, pub status: *mut u8, pub len: *mut u32, pub first: *mut u32,
    pub ops: *mut u32, pub ops_off: *mut u64, pub owner_: *mut c_void,
}

extern "C" {
    pub fn bg_align_batch_ops(ctx: *mut bg_ctx, input: *const bg_batch, p: *const bg_params, out: *mut bg_ops_result) -> c_int;
    pub fn bg_ops_result_free(r: *mut bg_ops_result);
    /// ops -> the two aligned strings (`len` bytes each); AVX-512 VBMI2 when the host has it.
    pub fn bg_expand_ops(seq1_from: *const u8, seq2_from: *const u8, ops: *const u32, len: u64, a_out: *mut u8, b_out: *mut u8) -> c_int;
    pub fn bg_create(devices: *const c_int, n_dev: c_int, out: *mut *mut bg_ctx) -> c_int;
    pub fn bg_destroy(ctx: *mut bg_ctx);
    pub fn bg_strerror(err: c_int) -> *const c_char;
    pub fn bg_align_batch(ctx: *mut bg_ctx, input: *const bg_batch, p: *const bg_params, out: *mut bg_result) -> c_int;
    pub fn bg_result_free(r: *mut bg_result);
    pub fn bg_edit_distance_batch(ctx: *mut bg_ctx, input: *const bg_batch, out: *mut u64) -> c_int;
    /// Trace memory per launch of the long-pair path (0 = automatic); pairs that need more are aligned with
    /// bounded-memory traceback (row checkpoints + block-wise re-fill), same results.
    pub fn bg_set_long_trace_budget(ctx: *mut bg_ctx, bytes: u64) -> c_int;
    /// Page-lock the arena a Tile was flattened into (once), so that the H2D copy is an asynchronous DMA.
    pub fn bg_pin_host(ptr: *const c_void, bytes: u64) -> c_int;
    pub fn bg_unpin_host(ptr: *const c_void) -> c_int;
}
