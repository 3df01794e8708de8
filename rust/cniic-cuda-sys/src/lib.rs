//! Raw FFI declarations for include/cniic_b200.h plus a safe wrapper that implements cniic's `Codec` trait
//! (src/codec.rs:14-19).  Source-only in this repository (no Rust toolchain in the image).
#![allow(non_camel_case_types)]
use std::ffi::{c_char, c_int, c_void, CStr, CString};

#[repr(C)] pub struct cniic_ctx { _p: [u8; 0] }
#[repr(C)] pub struct cniic_kmeans { _p: [u8; 0] }

#[repr(C)]
#[derive(Default, Clone, Copy, Debug)]
pub struct cniic_kmeans_stats {
    pub iterations: u32, pub empty_events: u32, pub moved_last: u64, pub moved_total: u64,
    pub converged: u32, pub gpu_launches: u32, pub device_ms: f32, pub assign_ms_avg: f32, pub pairs_scored: u64,
}

pub const CNIIC_OK: c_int = 0;
pub const CNIIC_ERR_TOO_FEW_POINTS: c_int = 2;
pub const CNIIC_ERR_TOO_FEW_ACTIVE: c_int = 3;
pub const CNIIC_ERR_DECODE: c_int = 6;
pub const CNIIC_ERR_BUFFER_TOO_SMALL: c_int = 7;
pub const CNIIC_TIE_KEEP_CURRENT: c_int = 0;

extern "C" {
    pub fn cniic_ctx_create(device: c_int, out: *mut *mut cniic_ctx) -> c_int;
    pub fn cniic_ctx_destroy(ctx: *mut cniic_ctx);
    pub fn cniic_last_error(ctx: *const cniic_ctx) -> *const c_char;
    pub fn cniic_ctx_set_max_iters(ctx: *mut cniic_ctx, max_iters: u32) -> c_int;
    pub fn cniic_kmeans_rgb(ctx: *mut cniic_ctx, rgb: *const u8, counts: *const u32, n: usize, k: u32, max_iters: u32,
        tie_rule: c_int, out_centroids: *mut u8, out_weight: *mut u64, out_assign: *mut u16, stats: *mut cniic_kmeans_stats) -> c_int;
    pub fn cniic_kmeans_xyrgb(ctx: *mut cniic_ctx, rgb: *const u8, w: u32, h: u32, k: u32, max_iters: u32, tie_rule: c_int,
        out_xy: *mut u32, out_rgb: *mut u8, out_weight: *mut u64, out_assign: *mut u16, stats: *mut cniic_kmeans_stats) -> c_int;
    /// `count` independent images advanced in lock step, one launch per stage for the whole batch (bench.rs:15-34).
    pub fn cniic_kmeans_rgb_batch(ctx: *mut cniic_ctx, rgb: *const *const u8, n: *const usize, count: u32, k: u32, max_iters: u32,
                                  tie_rule: c_int, out_centroids: *mut u8, out_weight: *mut u64, out_assign: *const *mut u16,
                                  stats: *mut cniic_kmeans_stats) -> c_int;
    pub fn cniic_kmeans_xyrgb_batch(ctx: *mut cniic_ctx, rgb: *const *const u8, w: *const u32, h: *const u32, count: u32, k: u32, max_iters: u32,
                                    tie_rule: c_int, out_xy: *mut u32, out_rgb: *mut u8, out_weight: *mut u64, out_assign: *const *mut u16,
                                    stats: *mut cniic_kmeans_stats) -> c_int;
    pub fn cniic_delta_i16_range_device(ctx: *mut cniic_ctx, d_rgb: *const u8, w: u32, h: u32, i_begin: u64, i_end: u64, d_out: *mut i16) -> c_int;
    pub fn cniic_hist_delta_range_device(ctx: *mut cniic_ctx, d_rgb: *const u8, w: u32, h: u32, i_begin: u64, i_end: u64, out_keys: *mut u32,
                                         out_counts: *mut u64, cap: usize, out_n: *mut usize) -> c_int;
    pub fn cniic_voronoi_fill(ctx: *mut cniic_ctx, cxy: *const u32, crgb: *const u8, k: u32, w: u32, h: u32, out_rgb: *mut u8) -> c_int;
    pub fn cniic_hist_rgb(ctx: *mut cniic_ctx, rgb: *const u8, n: usize, out_keys: *mut u32, out_counts: *mut u64, cap: usize, out_n: *mut usize) -> c_int;
    pub fn cniic_hist_delta(ctx: *mut cniic_ctx, rgb: *const u8, w: u32, h: u32, out_keys: *mut u32, out_counts: *mut u64, cap: usize, out_n: *mut usize) -> c_int;
    pub fn cniic_delta_i16(ctx: *mut cniic_ctx, rgb: *const u8, w: u32, h: u32, out: *mut i16) -> c_int;
    pub fn cniic_hilbert_xy(ctx: *mut cniic_ctx, w: u32, h: u32, out_xy: *mut u32) -> c_int;
    pub fn cniic_sse_rgb(ctx: *mut cniic_ctx, a: *const u8, b: *const u8, n_pixels: usize, out_sse: *mut u64) -> c_int;
    pub fn cniic_codec_encode(ctx: *mut cniic_ctx, codec: *const c_char, rgb: *const u8, w: u32, h: u32, out: *mut u8, cap: usize, out_len: *mut usize) -> c_int;
    pub fn cniic_codec_decode(ctx: *mut cniic_ctx, codec: *const c_char, data: *const u8, len: usize, w: *mut u32, h: *mut u32, out_rgb: *mut u8, cap_pixels: usize) -> c_int;
    pub fn cniic_codec_name(codec: *const c_char, out: *mut c_char, cap: usize) -> c_int;
}

thread_local! {
    // bench.rs:27 calls codecs from rayon workers: one context per worker thread (a ctx is not thread-safe)
    static CTX: Ctx = Ctx::new().expect("no CUDA device: cniic-cuda has no CPU fallback");
}

pub struct Ctx(*mut cniic_ctx);
impl Ctx {
    pub fn new() -> Result<Self, c_int> {
        let mut p = std::ptr::null_mut();
        let rc = unsafe { cniic_ctx_create(-1, &mut p) };
        if rc == CNIIC_OK { Ok(Ctx(p)) } else { Err(rc) }
    }
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
            loop {
                let rc = unsafe { cniic_codec_encode(ctx.0, self.expr.as_ptr(), rgb.as_ptr(), w, h, buf.as_mut_ptr(), buf.len(), &mut need) };
                match rc {
                    CNIIC_OK => return writer.write_all(&buf[..need]),
                    CNIIC_ERR_BUFFER_TOO_SMALL => buf.resize(need, 0),
                    // kmeans.rs:54-56, 67-68 are assert!s in the reference: keep the panic behaviour
                    CNIIC_ERR_TOO_FEW_POINTS | CNIIC_ERR_TOO_FEW_ACTIVE => panic!("{}", ctx.err()),
                    _ => return Err(std::io::Error::new(std::io::ErrorKind::Other, ctx.err())),
                }
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
#[allow(dead_code)]
fn _unused(_: *mut c_void) {}
