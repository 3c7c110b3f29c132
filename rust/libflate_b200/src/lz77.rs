//! The trait path: a `libflate_lz77::Lz77Encode` whose match finder runs on the GPU.
//! Replaces `DefaultLz77Encoder` (libflate_lz77/src/default.rs:59-113); plugs into the UNCHANGED libflate encoders through
//! `EncodeOptions::with_lz77` (src/deflate/encode.rs:59-65).  Huffman coding stays on the CPU on this path; the whole-stream
//! path in `crate::stream` keeps everything on the device.
use crate::ctx::Ctx;
use crate::ffi;
use libflate_lz77::{Code, Lz77Encode, Sink, MAX_LENGTH, MAX_WINDOW_SIZE};

pub struct B200Lz77Encoder {
    ctx: Ctx,
    window_size: u16,
    max_length: u16,
    buf: Vec<u8>,
    codes: Vec<u32>,
}
impl B200Lz77Encoder {
    pub fn new(ctx: Ctx) -> Self {
        Self::with_window_size(ctx, MAX_WINDOW_SIZE)
    }
    /// `DefaultLz77Encoder::with_window_size` (default.rs:48-57)
    pub fn with_window_size(ctx: Ctx, size: u16) -> Self {
        B200Lz77Encoder { ctx, window_size: size.min(MAX_WINDOW_SIZE), max_length: MAX_LENGTH, buf: Vec::new(), codes: Vec::new() }
    }
    /// `DefaultLz77EncoderBuilder::max_length` (default.rs:234-239)
    pub fn max_length(mut self, len: u16) -> Self {
        self.max_length = len.min(MAX_LENGTH);
        self
    }
}
impl Lz77Encode for B200Lz77Encoder {
    fn encode<S: Sink>(&mut self, buf: &[u8], sink: S) {
        self.buf.extend_from_slice(buf);
        if self.buf.len() >= self.window_size as usize * 8 {
            // same threshold as default.rs:64: a chunk is whatever has been buffered when 8 windows are reached
            self.flush(sink);
        }
    }
    fn flush<S: Sink>(&mut self, mut sink: S) {
        if self.buf.is_empty() {
            return;
        }
        self.codes.resize(self.buf.len(), 0);
        let mut n = 0usize;
        let rc = unsafe {
            ffi::b2f_lz77_default(self.ctx.raw(), self.buf.as_ptr(), self.buf.len(), self.window_size as u32, self.max_length as u32,
                                  self.codes.as_mut_ptr(), &mut n)
        };
        // the trait is infallible (libflate_lz77/src/lib.rs:83-107): a device failure is a panic, like an allocation failure
        assert_eq!(rc, ffi::B2F_OK, "b2f_lz77_default: {}", self.ctx.last_error());
        for &c in &self.codes[..n] {
            sink.consume(if c & 0x8000_0000 != 0 {
                Code::Pointer { length: ((c >> 16) & 0x1FF) as u16, backward_distance: (c & 0xFFFF) as u16 }
            } else {
                Code::Literal(c as u8)
            });
        }
        self.buf.clear();
    }
    fn window_size(&self) -> u16 {
        self.window_size
    }
}
