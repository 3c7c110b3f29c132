//! Raw bindings of `include/b2f.h` (C ABI of libb2f.so).  Field order and widths mirror the header exactly.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

pub const B2F_OK: c_int = 0;
pub const B2F_ERR_INVALID_DATA: c_int = -1;
pub const B2F_ERR_UNEXPECTED_EOF: c_int = -2;
pub const B2F_ERR_OUTPUT_TOO_SMALL: c_int = -3;
pub const B2F_ERR_NOMEM: c_int = -4;
pub const B2F_ERR_CUDA: c_int = -5;
pub const B2F_ERR_INVALID_ARG: c_int = -6;

pub const B2F_FMT_DEFLATE: c_int = 0;
pub const B2F_FMT_ZLIB: c_int = 1;
pub const B2F_FMT_GZIP: c_int = 2;
pub const B2F_FMT_GZIP_MULTI: c_int = 3;

pub const B2F_MODE_DYNAMIC: i32 = 0;
pub const B2F_MODE_FIXED: i32 = 1;
pub const B2F_MODE_STORED: i32 = 2;
pub const B2F_SCHED_FLUSH: i64 = -1;

#[repr(C)]
pub struct b2f_ctx {
    _p: [u8; 0],
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct b2f_encode_opts {
    pub block_size: u64,
    pub window_size: u32,
    pub max_length: u32,
    pub mode: i32,
    pub zlib_flush_sync: i32,
    pub gzip_mtime: u32,
    pub gzip_os: u8,
    pub gzip_is_text: u8,
    pub gzip_is_verified: u8,
    pub gzip_has_extra: u8,
    pub gzip_extra: *const u8,
    pub gzip_extra_len: u32,
    pub gzip_filename: *const c_char,
    pub gzip_comment: *const c_char,
}

extern "C" {
    pub fn b2f_encode_opts_default(o: *mut b2f_encode_opts);
    pub fn b2f_ctx_create(device: c_int, out: *mut *mut b2f_ctx) -> c_int;
    pub fn b2f_ctx_destroy(ctx: *mut b2f_ctx);
    pub fn b2f_last_error(ctx: *const b2f_ctx) -> *const c_char;
    pub fn b2f_host_alloc(bytes: usize, out: *mut *mut c_void) -> c_int;
    pub fn b2f_host_free(p: *mut c_void);
    pub fn b2f_lz77_default(
        ctx: *mut b2f_ctx, buf: *const u8, len: usize, window_size: u32, max_length: u32, codes: *mut u32, n_codes: *mut usize,
    ) -> c_int;
    pub fn b2f_encode_batch(
        ctx: *mut b2f_ctx, fmt: c_int, opts: *const b2f_encode_opts, n_streams: usize, input: *const *const u8, in_len: *const usize,
        sched: *const *const i64, n_sched: *const usize, out: *const *mut u8, out_cap: *const usize, out_len: *mut usize, status: *mut c_int,
    ) -> c_int;
    pub fn b2f_encode_bound(in_len: usize, n_sched: usize, opts: *const b2f_encode_opts) -> usize;
    pub fn b2f_decode_batch(
        ctx: *mut b2f_ctx, fmt: c_int, n_streams: usize, input: *const *const u8, in_len: *const usize, out: *const *mut u8,
        out_cap: *const usize, out_len: *mut usize, in_consumed: *mut usize, status: *mut c_int,
    ) -> c_int;
    pub fn b2f_adler32_batch(ctx: *mut b2f_ctx, n: usize, buf: *const *const u8, len: *const usize, init: *const u32, out: *mut u32) -> c_int;
    pub fn b2f_crc32_batch(ctx: *mut b2f_ctx, n: usize, buf: *const *const u8, len: *const usize, init: *const u32, out: *mut u32) -> c_int;
}
