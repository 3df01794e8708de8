// Pipe-throughput microbenchmarks for sm_100a (B200). Standalone binary; not part of the product path.
// Measures warp-instructions per clock per SM for the instruction mixes the Lloyd kernels are built from,
// so the kernel design (FFMA vs FFMA2, FMNMX vs FMNMX3, IDP.4A, shared atomics) rests on measured numbers.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench tools/ubench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

constexpr int ITERS = 4096;

__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long r;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ float fmin3(float a, float b, float c) {
    float r;
    asm volatile("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}
__device__ __forceinline__ float ffma(float a, float b, float c) {
    float r;
    asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}
__device__ __forceinline__ float fmin2(float a, float b) {
    float r;
    asm volatile("min.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}

// mode 0: FFMA x8 chains
__global__ void k_ffma(float* out, float s) {
    float a[8];
    for (int i = 0; i < 8; i++) a[i] = threadIdx.x + i;
    float b = s, c = s * 0.5f;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
            for (int i = 0; i < 8; i++) a[i] = ffma(a[i], b, c);
    }
    float r = 0; for (int i = 0; i < 8; i++) r += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
// mode 1: FFMA2 x8 chains (16 fp32 lanes of work per thread-instr pair)
__global__ void k_ffma2(float* out, float s) {
    unsigned long long a[8];
    for (int i = 0; i < 8; i++) { float2 v = make_float2(threadIdx.x + i, i); a[i] = *reinterpret_cast<unsigned long long*>(&v); }
    float2 bv = make_float2(s, s), cv = make_float2(s * 0.5f, s);
    unsigned long long b = *reinterpret_cast<unsigned long long*>(&bv), c = *reinterpret_cast<unsigned long long*>(&cv);
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
            for (int i = 0; i < 8; i++) a[i] = fma2(a[i], b, c);
    }
    float r = 0; for (int i = 0; i < 8; i++) { float2 v = *reinterpret_cast<float2*>(&a[i]); r += v.x + v.y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
// mode 2: Lloyd-like scalar: per (pixel p of 8, centroid): 3 FFMA + 1 FMNMX, centroid regs shared
__global__ void k_mix31(float* out, float s) {
    float xr[8], xg[8], xb[8], m[8];
    for (int i = 0; i < 8; i++) { xr[i] = threadIdx.x + i; xg[i] = i * s; xb[i] = i + s; m[i] = 1e30f; }
    float c0 = s, c1 = s * 2, c2 = s * 3, c3 = s * 4;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            float t = ffma(xr[i], c0, c3); t = ffma(xg[i], c1, t); t = ffma(xb[i], c2, t); m[i] = fmin2(m[i], t);
        }
        c3 += 1.0f;
    }
    float r = 0; for (int i = 0; i < 8; i++) r += m[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
// mode 3: scalar FFMA, two centroids + FMNMX3: 6 FFMA + 1 FMNMX3
__global__ void k_mix61(float* out, float s) {
    float xr[8], xg[8], xb[8], m[8];
    for (int i = 0; i < 8; i++) { xr[i] = threadIdx.x + i; xg[i] = i * s; xb[i] = i + s; m[i] = 1e30f; }
    float c0 = s, c1 = s * 2, c2 = s * 3, c3 = s * 4, d0 = s * 5, d1 = s * 6, d2 = s * 7, d3 = s * 8;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            float t = ffma(xr[i], c0, c3); t = ffma(xg[i], c1, t); t = ffma(xb[i], c2, t);
            float u = ffma(xr[i], d0, d3); u = ffma(xg[i], d1, u); u = ffma(xb[i], d2, u);
            m[i] = fmin3(m[i], t, u);
        }
        c3 += 1.0f; d3 += 1.0f;
    }
    float r = 0; for (int i = 0; i < 8; i++) r += m[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
// mode 4: FFMA2 across two centroids + FMNMX3: 3 FFMA2 + 1 FMNMX3 per (pixel, centroid pair)
__global__ void k_mix2(float* out, float s) {
    unsigned long long xr[8], xg[8], xb[8]; float m[8];
    for (int i = 0; i < 8; i++) {
        float2 v = make_float2(threadIdx.x + i, threadIdx.x + i); xr[i] = *reinterpret_cast<unsigned long long*>(&v);
        v = make_float2(i * s, i * s); xg[i] = *reinterpret_cast<unsigned long long*>(&v);
        v = make_float2(i + s, i + s); xb[i] = *reinterpret_cast<unsigned long long*>(&v);
        m[i] = 1e30f;
    }
    float2 v0 = make_float2(s, s * 5), v1 = make_float2(s * 2, s * 6), v2 = make_float2(s * 3, s * 7), v3 = make_float2(s * 4, s * 8);
    unsigned long long c0 = *reinterpret_cast<unsigned long long*>(&v0), c1 = *reinterpret_cast<unsigned long long*>(&v1),
                       c2 = *reinterpret_cast<unsigned long long*>(&v2), c3 = *reinterpret_cast<unsigned long long*>(&v3);
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            unsigned long long t = fma2(xr[i], c0, c3); t = fma2(xg[i], c1, t); t = fma2(xb[i], c2, t);
            float2 tv = *reinterpret_cast<float2*>(&t);
            m[i] = fmin3(m[i], tv.x, tv.y);
        }
        c3 += 1;  // integer poke keeps the loop from being hoisted
    }
    float r = 0; for (int i = 0; i < 8; i++) r += m[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
// mode 5: IDP.4A chains
__global__ void k_dp4a(float* out, int s) {
    int a[8];
    for (int i = 0; i < 8; i++) a[i] = threadIdx.x + i;
    int b = s;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
            for (int i = 0; i < 8; i++) a[i] = __dp4a(a[i], b, a[i]);
    }
    int r = 0; for (int i = 0; i < 8; i++) r += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
// mode 6: IMAD chains
__global__ void k_imad(float* out, int s) {
    int a[8];
    for (int i = 0; i < 8; i++) a[i] = threadIdx.x + i;
    int b = s, c = s + 3;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
            for (int i = 0; i < 8; i++) a[i] = a[i] * b + c;
    }
    int r = 0; for (int i = 0; i < 8; i++) r += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
// mode 7: IDP.4A + IMNMX (1:1)
__global__ void k_dp4a_min(float* out, int s) {
    int x[8], m[8];
    for (int i = 0; i < 8; i++) { x[i] = threadIdx.x * 77 + i; m[i] = 0x7fffffff; }
    int c = s, acc = s * 3;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
#pragma unroll
            for (int i = 0; i < 8; i++) { int t = __dp4a(x[i], c, acc); m[i] = min(m[i], t); }
            c += 0x01010101; acc -= 3;
        }
    }
    int r = 0; for (int i = 0; i < 8; i++) r += m[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
// mode 8: shared atomics, pseudo-random bins in a table of `bins` u32
__global__ void k_atoms(float* out, int bins) {
    extern __shared__ unsigned int tab[];
    for (int i = threadIdx.x; i < bins; i += blockDim.x) tab[i] = 0;
    __syncthreads();
    unsigned int h = threadIdx.x * 2654435761u + blockIdx.x * 40503u;
    for (int it = 0; it < ITERS / 4; it++) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
            h = h * 1664525u + 1013904223u;
            atomicAdd(&tab[(h >> 8) % bins], h & 255u);
        }
    }
    __syncthreads();
    unsigned int r = 0;
    for (int i = threadIdx.x; i < bins; i += blockDim.x) r += tab[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
// mode 9: shared atomics, warp-coherent bins (all lanes same cluster, 4 consecutive words: r,g,b,count)
__global__ void k_atoms_same(float* out, int bins) {
    extern __shared__ unsigned int tab[];
    for (int i = threadIdx.x; i < bins; i += blockDim.x) tab[i] = 0;
    __syncthreads();
    unsigned int h = (threadIdx.x >> 5) * 2654435761u + blockIdx.x * 40503u;
    for (int it = 0; it < ITERS / 4; it++) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
            h = h * 1664525u + 1013904223u;
            atomicAdd(&tab[(h >> 8) % bins], threadIdx.x);
        }
    }
    __syncthreads();
    unsigned int r = 0;
    for (int i = threadIdx.x; i < bins; i += blockDim.x) r += tab[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
// mode 10: match_any
__global__ void k_match(float* out, int s) {
    unsigned int h = threadIdx.x * 2654435761u + blockIdx.x * 40503u, acc = 0;
    for (int it = 0; it < ITERS / 4; it++) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
            h = h * 1664525u + 1013904223u;
            acc += __match_any_sync(0xffffffffu, (h >> 8) & s);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
// mode 11: LDS.128 broadcast + scalar FFMA/FMNMX3 (the real inner-loop shape, P=8, centroids from smem)
__global__ void k_lds_mix(float* out, float s, int k) {
    extern __shared__ float4 cs[];
    for (int i = threadIdx.x; i < k; i += blockDim.x) cs[i] = make_float4(s * i, s + i, s - i, i);
    __syncthreads();
    float xr[8], xg[8], xb[8], m[8];
    for (int i = 0; i < 8; i++) { xr[i] = threadIdx.x + i; xg[i] = i * s; xb[i] = i + s; m[i] = 1e30f; }
    for (int it = 0; it < ITERS / 64; it++) {
#pragma unroll 4
        for (int j = 0; j < k; j += 2) {
            float4 c = cs[j], d = cs[j + 1];
#pragma unroll
            for (int i = 0; i < 8; i++) {
                float t = ffma(xr[i], c.x, c.w); t = ffma(xg[i], c.y, t); t = ffma(xb[i], c.z, t);
                float u = ffma(xr[i], d.x, d.w); u = ffma(xg[i], d.y, u); u = ffma(xb[i], d.z, u);
                m[i] = fmin3(m[i], t, u);
            }
        }
        xr[0] += 1.0f;
    }
    float r = 0; for (int i = 0; i < 8; i++) r += m[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
// mode 12: LDS.128 (two centroids interleaved) + FFMA2/FMNMX3
__global__ void k_lds_mix2(float* out, float s, int k) {
    extern __shared__ float4 cs[];   // pair j: cs[2j] = {c0.x,c1.x,c0.y,c1.y}, cs[2j+1] = {c0.z,c1.z,c0.w,c1.w}
    for (int i = threadIdx.x; i < k; i += blockDim.x) cs[i] = make_float4(s * i, s + i, s - i, i);
    __syncthreads();
    unsigned long long xr[8], xg[8], xb[8]; float m[8];
    for (int i = 0; i < 8; i++) {
        float2 v = make_float2(threadIdx.x + i, threadIdx.x + i); xr[i] = *reinterpret_cast<unsigned long long*>(&v);
        v = make_float2(i * s, i * s); xg[i] = *reinterpret_cast<unsigned long long*>(&v);
        v = make_float2(i + s, i + s); xb[i] = *reinterpret_cast<unsigned long long*>(&v);
        m[i] = 1e30f;
    }
    const ulonglong2* cp = reinterpret_cast<const ulonglong2*>(cs);
    for (int it = 0; it < ITERS / 64; it++) {
#pragma unroll 4
        for (int j = 0; j < k; j += 2) {
            ulonglong2 a = cp[j], b = cp[j + 1];
#pragma unroll
            for (int i = 0; i < 8; i++) {
                unsigned long long t = fma2(xr[i], a.x, b.y); t = fma2(xg[i], a.y, t); t = fma2(xb[i], b.x, t);
                float2 tv = *reinterpret_cast<float2*>(&t);
                m[i] = fmin3(m[i], tv.x, tv.y);
            }
        }
        xr[0] += 1;
    }
    float r = 0; for (int i = 0; i < 8; i++) r += m[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <typename F>
static float timeit(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); f();
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int i = 0; i < 5; i++) f();
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms / 5;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    printf("device %s SMs %d maxclk %d kHz\n", p.name, sms, clk_khz);
    float* out; CK(cudaMalloc(&out, sizeof(float) * 148 * 16 * 1024));
    const int TPB = 256;
    for (int bps : {1, 2, 4}) {
        int grid = sms * bps;
        double warps = (double)grid * TPB / 32;
        auto rep = [&](const char* name, float ms, double instr_per_thread, double work_per_instr) {
            double wi = warps * instr_per_thread;              // warp-instructions
            double per_clk_sm = wi / (ms * 1e-3 * clk_khz * 1e3) / sms;
            printf("  bps=%d %-34s %8.3f ms  %6.3f warp-instr/clk/SM (at max clk)  %7.2f Gop-lanes/s x%g\n", bps, name, ms,
                   per_clk_sm, wi * 32 / (ms * 1e-3) * 1e-9, work_per_instr);
        };
        rep("FFMA", timeit([&] { k_ffma<<<grid, TPB>>>(out, 1.0001f); }), ITERS * 32.0, 1);
        rep("FFMA2", timeit([&] { k_ffma2<<<grid, TPB>>>(out, 1.0001f); }), ITERS * 32.0, 2);
        rep("3FFMA+FMNMX (32 instr/it)", timeit([&] { k_mix31<<<grid, TPB>>>(out, 1.0001f); }), ITERS * 32.0, 1);
        rep("6FFMA+FMNMX3 (56 instr/it)", timeit([&] { k_mix61<<<grid, TPB>>>(out, 1.0001f); }), ITERS * 56.0, 1);
        rep("3FFMA2+FMNMX3 (32 instr/it)", timeit([&] { k_mix2<<<grid, TPB>>>(out, 1.0001f); }), ITERS * 32.0, 1);
        rep("IDP.4A", timeit([&] { k_dp4a<<<grid, TPB>>>(out, 3); }), ITERS * 32.0, 1);
        rep("IMAD", timeit([&] { k_imad<<<grid, TPB>>>(out, 3); }), ITERS * 32.0, 1);
        rep("IDP.4A+IMNMX (64 instr/it)", timeit([&] { k_dp4a_min<<<grid, TPB>>>(out, 3); }), ITERS * 64.0, 1);
        rep("ATOMS random 1024 bins", timeit([&] { k_atoms<<<grid, TPB, 1024 * 4>>>(out, 1024); }), ITERS, 1);
        rep("ATOMS random 12288 bins", timeit([&] { k_atoms<<<grid, TPB, 12288 * 4>>>(out, 12288); }), ITERS, 1);
        rep("ATOMS warp-same 1024 bins", timeit([&] { k_atoms_same<<<grid, TPB, 1024 * 4>>>(out, 1024); }), ITERS, 1);
        rep("MATCH.ANY 8 distinct", timeit([&] { k_match<<<grid, TPB>>>(out, 7); }), ITERS, 1);
        rep("MATCH.ANY 256 distinct", timeit([&] { k_match<<<grid, TPB>>>(out, 255); }), ITERS, 1);
        rep("LDS.128+6FFMA+FMNMX3 k=256 (58/2c)", timeit([&] { k_lds_mix<<<grid, TPB, 256 * 16>>>(out, 1.0001f, 256); }),
            (ITERS / 64) * 128.0 * 58.0, 1);
        rep("LDS.128+3FFMA2+FMNMX3 k=256 (34/2c)", timeit([&] { k_lds_mix2<<<grid, TPB, 256 * 16>>>(out, 1.0001f, 256); }),
            (ITERS / 64) * 128.0 * 34.0, 1);
    }
    CK(cudaDeviceSynchronize());
    printf("done\n");
    return 0;
}
