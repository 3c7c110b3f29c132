//! Owning wrapper of a `b2f_ctx` (one CUDA device, its streams and scratch memory) and the status -> io::Error mapping.
use crate::ffi;
use std::ffi::CStr;
use std::io;
use std::rc::Rc;

/// One context per host thread (`b2f.h`: not thread-safe) -- hence `Rc`, not `Arc`: the codecs built on it are `!Send`.
pub struct RawCtx(pub(crate) *mut ffi::b2f_ctx);
impl Drop for RawCtx {
    fn drop(&mut self) {
        unsafe { ffi::b2f_ctx_destroy(self.0) }
    }
}
#[derive(Clone)]
pub struct Ctx(pub(crate) Rc<RawCtx>);
impl Ctx {
    /// `device`: CUDA ordinal.  Fails with `io::ErrorKind::Other` when no device is usable (there is no CPU fallback).
    pub fn new(device: i32) -> io::Result<Self> {
        let mut p = std::ptr::null_mut();
        let rc = unsafe { ffi::b2f_ctx_create(device, &mut p) };
        if rc != ffi::B2F_OK {
            let msg = unsafe { CStr::from_ptr(ffi::b2f_last_error(std::ptr::null())) }.to_string_lossy().into_owned();
            return Err(io::Error::new(io::ErrorKind::Other, format!("b2f_ctx_create: {msg}")));
        }
        Ok(Ctx(Rc::new(RawCtx(p))))
    }
    pub(crate) fn raw(&self) -> *mut ffi::b2f_ctx {
        (self.0).0
    }
    pub(crate) fn last_error(&self) -> String {
        unsafe { CStr::from_ptr(ffi::b2f_last_error(self.raw())) }.to_string_lossy().into_owned()
    }
    /// Maps a call / stream status to the `io::Error` the reference raises at the same point
    /// (`src/lib.rs:10-29` InvalidData, `src/bit.rs:137` UnexpectedEof).
    pub(crate) fn check(&self, rc: i32) -> io::Result<()> {
        match rc {
            ffi::B2F_OK => Ok(()),
            ffi::B2F_ERR_INVALID_DATA => Err(io::Error::new(io::ErrorKind::InvalidData, "invalid DEFLATE/ZLIB/GZIP data")),
            ffi::B2F_ERR_UNEXPECTED_EOF => Err(io::Error::new(io::ErrorKind::UnexpectedEof, "unexpected end of stream")),
            ffi::B2F_ERR_NOMEM => Err(io::Error::new(io::ErrorKind::OutOfMemory, "libb2f: out of memory")),
            ffi::B2F_ERR_INVALID_ARG => Err(io::Error::new(io::ErrorKind::InvalidInput, self.last_error())),
            _ => Err(io::Error::new(io::ErrorKind::Other, self.last_error())),
        }
    }
}

/// Page-locked byte buffer (`b2f_host_alloc`): the DMA engines read/write it in place, ordinary `Vec<u8>` memory is staged
/// by the library instead (both work).
pub struct PinnedBuf {
    p: *mut u8,
    len: usize,
}
impl PinnedBuf {
    pub fn new(len: usize) -> io::Result<Self> {
        let mut p = std::ptr::null_mut();
        if unsafe { ffi::b2f_host_alloc(len, &mut p) } != ffi::B2F_OK {
            return Err(io::Error::new(io::ErrorKind::OutOfMemory, "b2f_host_alloc"));
        }
        Ok(PinnedBuf { p: p as *mut u8, len })
    }
}
impl std::ops::Deref for PinnedBuf {
    type Target = [u8];
    fn deref(&self) -> &[u8] {
        unsafe { std::slice::from_raw_parts(self.p, self.len) }
    }
}
impl std::ops::DerefMut for PinnedBuf {
    fn deref_mut(&mut self) -> &mut [u8] {
        unsafe { std::slice::from_raw_parts_mut(self.p, self.len) }
    }
}
impl Drop for PinnedBuf {
    fn drop(&mut self) {
        unsafe { ffi::b2f_host_free(self.p as *mut _) }
    }
}
