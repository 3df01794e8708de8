// selftest.cu -- TEST INFRASTRUCTURE: tiny kernels that exercise the emulation itself (tests/test_emu_kernels.py runs the
// built program once per case and expects the named diagnosis).  Written in CUDA syntax and passed through build_emu.py
// like the product sources.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <vector>

__global__ void __launch_bounds__(256) k_reduce(const int *in, int n, int *out, unsigned *ballots) {
    __shared__ int s_part[8];
    extern __shared__ int s_dyn[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int v = 0;
    for (int i = blockIdx.x * blockDim.x + tid; i < n; i += gridDim.x * blockDim.x) v += in[i];
    s_dyn[tid] = v;
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) s_part[warp] = v;
    __syncthreads();
    if (tid == 0) {
        int t = 0;
        for (int i = 0; i < 8; i++) t += s_part[i];
        int t2 = 0;
        for (int i = 0; i < 256; i++) t2 += s_dyn[i];
        atomicAdd(out, t);
        atomicAdd(out + 1, t2);
    }
    // partial-mask collectives: the even lanes vote among themselves, the odd lanes have left the function
    const unsigned act = __ballot_sync(0xffffffffu, (lane & 1) == 0);
    if (lane & 1) return;
    const unsigned peers = __match_any_sync(act, lane % 3 == 0);
    if (blockIdx.x == 0 && warp == 0) ballots[lane] = peers;
}

__global__ void k_oob(int *p, int n) { p[n + threadIdx.x] = 1; }

__global__ void k_deadlock(int *p) {
    if (threadIdx.x == 0) {
        __syncthreads();
    } else {
        p[threadIdx.x] = __shfl_sync(0xffffffffu, (int)threadIdx.x, 0);
        __syncthreads();
    }
}

__global__ void k_divergent(int *p) {
    if (threadIdx.x < 16) p[threadIdx.x] = (int)__ballot_sync(0xffffffffu, 1);
    else p[threadIdx.x] = __shfl_xor_sync(0xffffffffu, (int)threadIdx.x, 1);
}

// a missing barrier that a forward sweep over the threads hides: every thread reads its LEFT neighbour's slot
__global__ void k_race(int *out) {
    __shared__ int s_v[64];
    s_v[threadIdx.x] = (int)threadIdx.x * 3;
    /* __syncthreads() is missing here */
    out[threadIdx.x] = threadIdx.x ? s_v[threadIdx.x - 1] : 0;
}

// a 128-bit load from an address that is only 4-byte aligned: the GPU raises "misaligned address"
__global__ void k_misaligned(const int *p, int *out) {
    const uint4 v = *reinterpret_cast<const uint4 *>(p + 1);
    out[0] = (int)(v.x + v.y + v.z + v.w);
}

int main(int argc, char **argv) {
    const char *mode = argc > 1 ? argv[1] : "ok";
    int *d = nullptr;
    cudaMalloc(&d, 64 * sizeof(int));
    cudaMemset(d, 0, 64 * sizeof(int));
    if (!strcmp(mode, "oob")) {
        k_oob<<<1, 32>>>(d, 64);
        cudaFree(d);  // the canary check fires here
        return 0;
    }
    if (!strcmp(mode, "deadlock")) { k_deadlock<<<1, 32>>>(d); return 0; }
    if (!strcmp(mode, "divergent")) { k_divergent<<<1, 32>>>(d); return 0; }
    if (!strcmp(mode, "misaligned")) { k_misaligned<<<1, 1>>>(d, d + 32); return 0; }
    if (!strcmp(mode, "race")) {  // exit code 0 = the hazard stayed hidden, 3 = it produced a wrong value
        k_race<<<1, 64>>>(d);
        int h[64];
        cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        for (int i = 1; i < 64; i++)
            if (h[i] != (i - 1) * 3) { fprintf(stderr, "race exposed at thread %d\n", i); return 3; }
        printf("race hidden\n");
        return 0;
    }
    const int n = 100000;
    std::vector<int> h(n);
    long long want = 0;
    for (int i = 0; i < n; i++) { h[i] = (i * 7919) % 101 - 50; want += h[i]; }
    int *d_in = nullptr, *d_out = nullptr;
    unsigned *d_b = nullptr;
    cudaMalloc(&d_in, n * sizeof(int));
    cudaMalloc(&d_out, 2 * sizeof(int));
    cudaMalloc(&d_b, 32 * sizeof(unsigned));
    cudaMemcpy(d_in, h.data(), n * sizeof(int), cudaMemcpyHostToDevice);
    cudaMemset(d_out, 0, 2 * sizeof(int));
    cudaMemset(d_b, 0, 32 * sizeof(unsigned));
    k_reduce<<<5, 256, 256 * sizeof(int)>>>(d_in, n, d_out, d_b);
    int got[2];
    unsigned b[32];
    cudaMemcpy(got, d_out, sizeof(got), cudaMemcpyDeviceToHost);
    cudaMemcpy(b, d_b, sizeof(b), cudaMemcpyDeviceToHost);
    unsigned m0 = 0, m1 = 0;  // even lanes with lane % 3 == 0, and the other even lanes
    for (int l = 0; l < 32; l += 2) (l % 3 == 0 ? m0 : m1) |= 1u << l;
    bool ok = got[0] == want && got[1] == want;
    for (int l = 0; l < 32; l += 2) ok = ok && b[l] == (l % 3 == 0 ? m0 : m1);
    cudaFree(d); cudaFree(d_in); cudaFree(d_out); cudaFree(d_b);
    if (!ok) { fprintf(stderr, "selftest FAILED: %d %d want %lld\n", got[0], got[1], want); return 1; }
    printf("selftest ok\n");
    return 0;
}
