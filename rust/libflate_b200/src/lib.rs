//! `libflate_b200`: the B200 (sm_100a) backend of libflate's hot path, behind libflate's own names.
//!
//! ```no_run
//! use std::io::Write;
//! let ctx = libflate_b200::Ctx::new(0)?;
//! // (a) the trait path: GPU match finder inside the unchanged libflate encoder
//! let lz77 = libflate_b200::B200Lz77Encoder::new(ctx.clone());
//! let mut e = libflate::gzip::Encoder::with_options(Vec::new(), libflate::gzip::EncodeOptions::with_lz77(lz77))?;
//! e.write_all(b"Hello World!")?;
//! let _gz = e.finish().into_result()?;
//! // (b) the whole-stream path: LZ77, Huffman, bit packing and CRC-32 on the device
//! let mut e = libflate_b200::gzip::encoder(Vec::new(), ctx.clone())?;
//! e.write_all(b"Hello World!")?;
//! let _gz = e.finish().into_result()?;
//! # Ok::<(), std::io::Error>(())
//! ```
mod ctx;
pub mod ffi;
mod lz77;
mod stream;

pub use ctx::{Ctx, PinnedBuf};
pub use lz77::B200Lz77Encoder;
pub use stream::EncodeOptions;

macro_rules! codec_module {
    ($name:ident, $fmt:expr, $doc:expr) => {
        #[doc = $doc]
        pub mod $name {
            use crate::{ctx::Ctx, stream};
            use std::io::{self, Read, Write};
            pub use crate::stream::EncodeOptions;
            pub type Encoder<W> = stream::Encoder<W>;
            pub type Decoder<R> = stream::Decoder<R>;
            /// `Encoder::new(inner)` of the reference, plus the device context
            pub fn encoder<W: Write>(inner: W, ctx: Ctx) -> io::Result<Encoder<W>> {
                encoder_with_options(inner, ctx, EncodeOptions::default())
            }
            /// `Encoder::with_options(inner, options)`
            pub fn encoder_with_options<W: Write>(inner: W, ctx: Ctx, options: EncodeOptions) -> io::Result<Encoder<W>> {
                Ok(stream::Encoder::make(inner, ctx, $fmt, options))
            }
            /// `Decoder::new(inner)`
            pub fn decoder<R: Read>(inner: R, ctx: Ctx) -> io::Result<Decoder<R>> {
                Ok(stream::Decoder::make(inner, ctx, $fmt))
            }
        }
    };
}
codec_module!(deflate, crate::ffi::B2F_FMT_DEFLATE, "raw DEFLATE: mirrors `libflate::deflate::{Encoder, Decoder}` (src/deflate/{encode,decode}.rs)");
codec_module!(zlib, crate::ffi::B2F_FMT_ZLIB, "ZLIB: mirrors `libflate::zlib::{Encoder, Decoder}` (src/zlib.rs:284-410, 522-681); Adler-32 on the device");
codec_module!(gzip, crate::ffi::B2F_FMT_GZIP, "GZIP: mirrors `libflate::gzip::{Encoder, Decoder}` (src/gzip.rs:754-1048); CRC-32 on the device");

/// `libflate::gzip::MultiDecoder` (src/gzip.rs:1052-1167): decodes every member of a concatenated gzip stream
pub fn gzip_multi_decoder<R: std::io::Read>(inner: R, ctx: Ctx) -> std::io::Result<stream::Decoder<R>> {
    Ok(stream::Decoder::make(inner, ctx, ffi::B2F_FMT_GZIP_MULTI))
}

/// `checksum::Crc32` / `checksum::Adler32` (src/checksum.rs:4-33) for callers that only want the reductions; `init` chains calls.
pub fn crc32(ctx: &Ctx, data: &[u8], init: u32) -> std::io::Result<u32> {
    let (p, n, mut out) = (data.as_ptr(), data.len(), 0u32);
    ctx.check(unsafe { ffi::b2f_crc32_batch(ctx.raw(), 1, &p, &n, &init, &mut out) })?;
    Ok(out)
}
pub fn adler32(ctx: &Ctx, data: &[u8], init: u32) -> std::io::Result<u32> {
    let (p, n, mut out) = (data.as_ptr(), data.len(), 0u32);
    ctx.check(unsafe { ffi::b2f_adler32_batch(ctx.raw(), 1, &p, &n, &init, &mut out) })?;
    Ok(out)
}

#[cfg(test)]
mod tests {
    //! The reference's own golden vectors (src/deflate/encode.rs:152-154, src/zlib.rs:547-549, src/lz77.rs:16-31); need a B200.
    use super::*;
    use libflate_lz77::{Code, Lz77Encode};
    use std::io::{Read, Write};

    #[test]
    fn hello_world_bytes_equal_libflate() {
        let ctx = Ctx::new(0).unwrap();
        let mut e = deflate::encoder(Vec::new(), ctx.clone()).unwrap();
        e.write_all(b"Hello World!").unwrap();
        let got = e.finish().into_result().unwrap();
        assert_eq!(got, [5, 192, 49, 13, 0, 0, 8, 3, 65, 43, 224, 6, 7, 24, 128, 237, 147, 38, 245, 63, 244, 230, 65, 181, 50, 215, 1]);
        let mut want = libflate::deflate::Encoder::new(Vec::new());
        want.write_all(b"Hello World!").unwrap();
        assert_eq!(got, want.finish().into_result().unwrap());
        let mut d = deflate::decoder(&got[..], ctx).unwrap();
        let mut s = Vec::new();
        d.read_to_end(&mut s).unwrap();
        assert_eq!(s, b"Hello World!");
    }

    #[test]
    fn lz77_trait_path() {
        let ctx = Ctx::new(0).unwrap();
        let mut enc = B200Lz77Encoder::new(ctx);
        let mut codes: Vec<Code> = Vec::new();
        enc.encode(b"aaaaa", &mut codes);
        enc.flush(&mut codes);
        assert_eq!(codes, [Code::Literal(97), Code::Pointer { length: 4, backward_distance: 1 }]);
    }
}
