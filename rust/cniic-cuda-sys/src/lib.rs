//! cniic-cuda-sys: raw FFI declarations of include/cniic_b200.h (`ffi.rs`, GENERATED from the header by
//! tools/gen_rust_ffi.py -- every entry point, constant and repr(C) struct) plus a safe wrapper a `Codec` adapter can sit
//! on (rust/cniic-side/gpuc.rs is that adapter, the one file a cniic maintainer adds under src/codec/).
//! Source-only in this repository (no Rust toolchain in the image); tests/test_cpu_host.py checks ffi.rs against the header.
mod ffi;
pub use ffi::*;

use std::ffi::{c_char, CStr, CString};

thread_local! {
    // bench.rs:27 calls codecs from rayon workers: one context per worker thread (a ctx is not thread-safe).  The large
    // scratch arrays (histogram bins) are shared per DEVICE inside the library, not per ctx, so 32 workers do not hold 32 copies.
    static CTX: Ctx = Ctx::new().expect("no CUDA device: cniic-cuda has no CPU fallback");
}

pub struct Ctx(*mut cniic_ctx);
impl Ctx {
    pub fn new() -> Result<Self, i32> {
        let mut p = std::ptr::null_mut();
        let rc = unsafe { cniic_ctx_create(-1, &mut p) };
        if rc == CNIIC_OK { Ok(Ctx(p)) } else { Err(rc) }
    }
    pub fn raw(&self) -> *mut cniic_ctx { self.0 }
    fn err(&self) -> String { unsafe { CStr::from_ptr(cniic_last_error(self.0)) }.to_string_lossy().into_owned() }
}
impl Drop for Ctx { fn drop(&mut self) { unsafe { cniic_ctx_destroy(self.0) } } }

/// A GPU codec selected by the reference's own codec expression ("voronoi(2048)", "cluster-colors(256)", "delta", ...).
pub struct GpuCodec { expr: CString }
impl GpuCodec {
    pub fn new(expr: &str) -> Option<Self> {
        let c = CString::new(expr).ok()?;
        let mut buf = [0 as c_char; 64];
        (unsafe { cniic_codec_name(c.as_ptr(), buf.as_mut_ptr(), 64) } == CNIIC_OK).then(|| GpuCodec { expr: c })
    }
    /// Codec::encode (codec.rs:15): same bytes as the CPU codec with the deterministic tie rules of DESIGN.md.
    pub fn encode_rgb8(&self, rgb: &[u8], w: u32, h: u32, writer: &mut impl std::io::Write) -> std::io::Result<()> {
        CTX.with(|ctx| {
            let mut need = 0usize;
            let mut buf = vec![0u8; 1 << 16];
            let mut rc = unsafe { cniic_codec_encode(ctx.0, self.expr.as_ptr(), rgb.as_ptr(), w, h, buf.as_mut_ptr(), buf.len(), &mut need) };
            if rc == CNIIC_ERR_BUFFER_TOO_SMALL {
                // the finished stream waits in the ctx: fetch it, do not encode twice
                buf.resize(need, 0);
                rc = unsafe { cniic_codec_encode_fetch(ctx.0, buf.as_mut_ptr(), buf.len(), &mut need) };
            }
            match rc {
                CNIIC_OK => writer.write_all(&buf[..need]),
                // kmeans.rs:54-56, 67-68 are assert!s in the reference: keep the panic behaviour
                CNIIC_ERR_TOO_FEW_POINTS | CNIIC_ERR_TOO_FEW_ACTIVE => panic!("{}", ctx.err()),
                _ => Err(std::io::Error::new(std::io::ErrorKind::Other, ctx.err())),
            }
        })
    }
    /// Codec::decode (codec.rs:16): None on a malformed stream.
    pub fn decode_rgb8(&self, data: &[u8]) -> Option<(u32, u32, Vec<u8>)> {
        CTX.with(|ctx| {
            let (mut w, mut h) = (0u32, 0u32);
            let rc = unsafe { cniic_codec_decode(ctx.0, self.expr.as_ptr(), data.as_ptr(), data.len(), &mut w, &mut h, std::ptr::null_mut(), 0) };
            if rc != CNIIC_OK { return None; }
            let mut out = vec![0u8; w as usize * h as usize * 3];
            let rc = unsafe { cniic_codec_decode(ctx.0, self.expr.as_ptr(), data.as_ptr(), data.len(), &mut w, &mut h, out.as_mut_ptr(), out.len() / 3) };
            (rc == CNIIC_OK).then_some((w, h, out))
        })
    }
    pub fn name(&self) -> String {
        let mut buf = [0 as c_char; 64];
        unsafe { cniic_codec_name(self.expr.as_ptr(), buf.as_mut_ptr(), 64); CStr::from_ptr(buf.as_ptr()) }.to_string_lossy().into_owned()
    }
}

/// kmeans::cluster for ColorCount points (kmeans.rs:21, clusterc.rs:28): colours + counts in, k centroids + assignment out.
pub fn kmeans_rgb(colours: &[u8], counts: Option<&[u32]>, k: u32, max_iters: u32) -> Result<(Vec<u8>, Vec<u64>, Vec<u16>, cniic_kmeans_stats), String> {
    let n = colours.len() / 3;
    CTX.with(|ctx| {
        let (mut cen, mut wt, mut asg) = (vec![0u8; 3 * k as usize], vec![0u64; k as usize], vec![0u16; n]);
        let mut st: cniic_kmeans_stats = unsafe { std::mem::zeroed() };
        let rc = unsafe { cniic_kmeans_rgb(ctx.0, colours.as_ptr(), counts.map_or(std::ptr::null(), |c| c.as_ptr()), n, k, max_iters,
                                           CNIIC_TIE_KEEP_CURRENT, cen.as_mut_ptr(), wt.as_mut_ptr(), asg.as_mut_ptr(), &mut st) };
        match rc {
            CNIIC_OK => Ok((cen, wt, asg, st)),
            CNIIC_ERR_TOO_FEW_POINTS | CNIIC_ERR_TOO_FEW_ACTIVE => panic!("{}", ctx.err()),
            _ => Err(ctx.err()),
        }
    })
}
