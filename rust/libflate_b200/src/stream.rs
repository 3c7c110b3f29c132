//! The Read/Write surface: `Encoder<W>` / `Decoder<R>` with the method names and semantics of
//! src/deflate/encode.rs:132-258 and src/deflate/decode.rs:8-165 (zlib.rs / gzip.rs wrap the same machinery).
//!
//! Encoder: `write` always consumes everything and records the write size -- libflate's output depends on the write
//! schedule (chunk = bytes buffered when 8 windows are reached, block = writes until block_size; encode.rs:277-286,
//! default.rs:60-68) and the library reproduces it from that record; `flush` records a block boundary; `finish` runs ONE
//! batch call and writes the stream to `W`.  Difference from the reference: bytes reach `W` at `finish()`, not after every block.
//! Decoder: reads the compressed stream from `R` to its end on the first `read`, decodes once, serves `read` calls.
use crate::ctx::Ctx;
use crate::ffi;
use libflate::finish::{Complete, Finish};
use std::ffi::CString;
use std::io::{self, Read, Write};

/// Encoder options = deflate::EncodeOptions + DefaultLz77EncoderBuilder + the header fields that influence output bytes.
#[derive(Clone)]
pub struct EncodeOptions {
    pub(crate) o: ffi::b2f_encode_opts,
    filename: Option<CString>,
    comment: Option<CString>,
    extra: Option<Vec<u8>>,
}
impl Default for EncodeOptions {
    fn default() -> Self {
        let mut o = std::mem::MaybeUninit::<ffi::b2f_encode_opts>::zeroed();
        unsafe { ffi::b2f_encode_opts_default(o.as_mut_ptr()) };
        EncodeOptions { o: unsafe { o.assume_init() }, filename: None, comment: None, extra: None }
    }
}
impl EncodeOptions {
    pub fn new() -> Self {
        Self::default()
    }
    /// `EncodeOptions::block_size` (encode.rs:84-87)
    pub fn block_size(mut self, size: usize) -> Self {
        self.o.block_size = size as u64;
        self
    }
    /// `EncodeOptions::no_compression` (encode.rs:66-72)
    pub fn no_compression(mut self) -> Self {
        self.o.mode = ffi::B2F_MODE_STORED;
        self
    }
    /// `EncodeOptions::fixed_huffman_codes` (encode.rs:103-106)
    pub fn fixed_huffman_codes(mut self) -> Self {
        self.o.mode = ffi::B2F_MODE_FIXED;
        self
    }
    pub fn window_size(mut self, size: u16) -> Self {
        self.o.window_size = size as u32;
        self
    }
    pub fn max_length(mut self, len: u16) -> Self {
        self.o.max_length = len as u32;
        self
    }
    /// zlib `FlushMode::Sync` (zlib.rs:150-157)
    pub fn zlib_flush_sync(mut self, on: bool) -> Self {
        self.o.zlib_flush_sync = on as i32;
        self
    }
    /// gzip `HeaderBuilder::modification_time` (gzip.rs:171-174).  The reference defaults to now(); here 0 unless set.
    pub fn gzip_mtime(mut self, t: u32) -> Self {
        self.o.gzip_mtime = t;
        self
    }
    pub fn gzip_filename(mut self, name: CString) -> Self {
        self.filename = Some(name);
        self
    }
    pub fn gzip_comment(mut self, c: CString) -> Self {
        self.comment = Some(c);
        self
    }
    pub fn gzip_extra(mut self, serialized_subfields: Vec<u8>) -> Self {
        self.extra = Some(serialized_subfields);
        self
    }
    /// pointers into `self` are only valid while `self` lives: resolved right before the FFI call
    fn resolved(&self) -> ffi::b2f_encode_opts {
        let mut o = self.o;
        o.gzip_filename = self.filename.as_ref().map_or(std::ptr::null(), |s| s.as_ptr());
        o.gzip_comment = self.comment.as_ref().map_or(std::ptr::null(), |s| s.as_ptr());
        if let Some(e) = &self.extra {
            o.gzip_has_extra = 1;
            o.gzip_extra = e.as_ptr();
            o.gzip_extra_len = e.len() as u32;
        }
        o
    }
}

pub struct Encoder<W: Write> {
    inner: W,
    ctx: Ctx,
    fmt: i32,
    opts: EncodeOptions,
    data: Vec<u8>,
    sched: Vec<i64>,
}
impl<W: Write> Encoder<W> {
    pub(crate) fn make(inner: W, ctx: Ctx, fmt: i32, opts: EncodeOptions) -> Self {
        Encoder { inner, ctx, fmt, opts, data: Vec::new(), sched: Vec::new() }
    }
    /// `Encoder::finish` (encode.rs:203-209): returns the writer AND the error, if any
    pub fn finish(mut self) -> Finish<W, io::Error> {
        let r = self.run();
        Finish::new(self.inner, r.err())
    }
    fn run(&mut self) -> io::Result<()> {
        let o = self.opts.resolved();
        let cap = unsafe { ffi::b2f_encode_bound(self.data.len(), self.sched.len(), &o) };
        let mut out = vec![0u8; cap];
        let (mut n, mut st) = (0usize, 0i32);
        let (ip, il) = (self.data.as_ptr(), self.data.len());
        // zero writes is an explicit empty schedule (not "one write_all"): pass a valid pointer with n_sched = 0
        let empty = [0i64];
        let (sp, sn) = if self.sched.is_empty() { (empty.as_ptr(), 0usize) } else { (self.sched.as_ptr(), self.sched.len()) };
        let op = out.as_mut_ptr();
        let rc = unsafe { ffi::b2f_encode_batch(self.ctx.raw(), self.fmt, &o, 1, &ip, &il, &sp, &sn, &op, &cap, &mut n, &mut st) };
        self.ctx.check(rc)?;
        self.ctx.check(st)?;
        self.inner.write_all(&out[..n])
    }
    pub fn as_inner_ref(&self) -> &W {
        &self.inner
    }
    pub fn as_inner_mut(&mut self) -> &mut W {
        &mut self.inner
    }
    pub fn into_inner(self) -> W {
        self.inner
    }
}
impl<W: Write> Write for Encoder<W> {
    /// always consumes all of `buf` and returns `Ok(buf.len())`, like `Encoder::write` (encode.rs:241-244)
    fn write(&mut self, buf: &[u8]) -> io::Result<usize> {
        self.data.extend_from_slice(buf);
        self.sched.push(buf.len() as i64);
        Ok(buf.len())
    }
    /// forces a (possibly empty) non-final block (encode.rs:245-248)
    fn flush(&mut self) -> io::Result<()> {
        self.sched.push(ffi::B2F_SCHED_FLUSH);
        Ok(())
    }
}
impl<W: Write> Complete for Encoder<W> {
    fn complete(self) -> io::Result<()> {
        self.finish().into_result().map(|_| ())
    }
}

pub struct Decoder<R: Read> {
    inner: R,
    ctx: Ctx,
    fmt: i32,
    out: Vec<u8>,
    pos: usize,
    status: i32,
    consumed: usize,
    decoded: bool,
}
impl<R: Read> Decoder<R> {
    pub(crate) fn make(inner: R, ctx: Ctx, fmt: i32) -> Self {
        Decoder { inner, ctx, fmt, out: Vec::new(), pos: 0, status: 0, consumed: 0, decoded: false }
    }
    fn run(&mut self) -> io::Result<()> {
        let mut input = Vec::new();
        self.inner.read_to_end(&mut input)?;
        let mut cap = (input.len() * 8 + 4096).max(1 << 16);
        loop {
            self.out.resize(cap, 0);
            let (ip, il, op) = (input.as_ptr(), input.len(), self.out.as_mut_ptr());
            let (mut ol, mut ic, mut st) = (0usize, 0usize, 0i32);
            let rc = unsafe { ffi::b2f_decode_batch(self.ctx.raw(), self.fmt, 1, &ip, &il, &op, &cap, &mut ol, &mut ic, &mut st) };
            self.ctx.check(rc)?;
            if st == ffi::B2F_ERR_OUTPUT_TOO_SMALL {
                cap = (cap * 2).max(ol + 64);     // out_len is always the size needed
                continue;
            }
            self.out.truncate(ol.min(cap));
            self.status = st;
            self.consumed = ic;
            self.decoded = true;
            return Ok(());
        }
    }
    pub fn as_inner_ref(&self) -> &R {
        &self.inner
    }
    pub fn as_inner_mut(&mut self) -> &mut R {
        &mut self.inner
    }
    pub fn into_inner(self) -> R {
        self.inner
    }
    /// `Decoder::unread_decoded_data` (decode.rs:71-79): on error, the bytes decoded before it
    pub fn unread_decoded_data(&self) -> &[u8] {
        &self.out[self.pos..]
    }
    /// bytes of the underlying stream that belong to the decoded container (the rest was read but is not part of it)
    pub fn consumed(&self) -> usize {
        self.consumed
    }
}
impl<R: Read> Read for Decoder<R> {
    fn read(&mut self, buf: &mut [u8]) -> io::Result<usize> {
        if !self.decoded {
            self.run()?;
        }
        self.ctx.check(self.status)?;
        let k = buf.len().min(self.out.len() - self.pos);
        buf[..k].copy_from_slice(&self.out[self.pos..self.pos + k]);
        self.pos += k;
        Ok(k)
    }
}
