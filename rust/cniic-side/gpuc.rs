//! src/codec/gpuc.rs -- the ONE new file on the cniic side: GPU codecs behind the unchanged `Codec` trait (codec.rs:14-19).
//!
//! Install: copy to cniic's `src/codec/gpuc.rs`, add `cniic-cuda-sys = { path = ".../rust/cniic-cuda-sys" }` to its
//! Cargo.toml, and add the line `gpuc, Gpu;` to `gen_all!` in `src/codec.rs:120-127`.  `bench.rs`, `main.rs`, the CSV
//! and the plot scripts stay untouched: `cniic --codec="gpu:voronoi(2048)" imgs/*` runs `measure_all` over the GPU codec.
//! Source-only in this repository (no Rust toolchain in the image).
use super::{Codec, Img};
use cniic_cuda_sys::GpuCodec;
use std::{io, str::FromStr};

/// Any of "cluster-colors(N)", "voronoi(N)", "delta", "hufman", "hilbert(rle)" executed on the GPU.
pub struct Gpu(GpuCodec);

impl Codec for Gpu {
    fn encode<W: io::Write>(&self, img: &Img, writer: &mut W) -> io::Result<()> {
        let rgb = img.to_rgb8(); // row-major packed RGB8 == the layout the C ABI takes (clusterc.rs:19, hufc.rs:13)
        self.0.encode_rgb8(rgb.as_raw(), rgb.width(), rgb.height(), writer)
    }

    fn decode<I: Iterator<Item = u8>>(&self, reader: &mut I) -> Option<Img> {
        let bytes: Vec<u8> = reader.collect();
        let (w, h, rgb) = self.0.decode_rgb8(&bytes)?;
        image::RgbImage::from_raw(w, h, rgb).map(Into::into)
    }

    /// "voronoi_2048", "cluster-colors_256", "delta", "Hufman", "hilbert-rle": the CSV names of the CPU codecs
    /// (clusterc.rs:59-61, 191-193; hilbertc.rs:81-87, 433-435; hufc.rs:42-44), so output/<name>.csv and the plots keep working.
    fn name(&self) -> String {
        self.0.name()
    }

    fn is_lossless(&self) -> bool {
        matches!(self.0.name().as_str(), "delta" | "Hufman" | "hilbert-rle")
    }
}

impl FromStr for Gpu {
    type Err = String;

    /// "gpu:voronoi(2048)" -- the prefix keeps the CPU codecs reachable under their own expressions.
    fn from_str(s: &str) -> Result<Self, String> {
        let expr = s.strip_prefix("gpu:").ok_or("not a gpu codec")?;
        GpuCodec::new(expr).map(Gpu).ok_or_else(|| format!("unknown gpu codec {expr}"))
    }
}
