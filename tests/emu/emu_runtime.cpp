// emu_runtime.cpp -- TEST INFRASTRUCTURE (see shim/cuda_runtime.h): fiber scheduler + CUDA runtime stand-ins.
//
// Execution model: a launch runs its CTAs one after another; inside a CTA every CUDA thread is a fiber, scheduled
// round-robin and switched only at __syncthreads() and at warp collectives.  That is one legal interleaving of the CUDA
// program (weakly ordered hardware has many more), so the emulation proves nothing about races -- it checks the logic
// that does not depend on timing, deterministically, and it catches:
//   * barriers / collectives that not all required threads reach (reported as a deadlock with the waiting threads),
//   * lanes of one rendezvous calling different collectives (divergent-collective bug),
//   * writes outside cudaMalloc'ed blocks (canaries checked at cudaFree and at exit), reads of never-written device
//     memory showing up as 0xCB.. garbage, dynamic shared memory poisoned per CTA.
#include <cuda_runtime.h>
#include <sys/mman.h>
#if !defined(__x86_64__)
#include <ucontext.h>
#endif

#include <chrono>
#include <map>
#include <mutex>
#include <vector>

namespace emu {

ThreadCtx *g_cur = nullptr;

namespace {

constexpr size_t STACK_BYTES = 256 << 10;
constexpr size_t CANARY = 64;

struct Rendezvous {
    bool used = false, draining = false;
    unsigned mask = 0, arrived = 0, readers = 0;
    int op = 0;
    unsigned gen = 0;
    unsigned long long slot[32];
    WarpVals out;
};

struct WarpState {
    unsigned live = 0;  // lanes that have not returned from the kernel
    Rendezvous rv[8];
};

// Context switch.  glibc's swapcontext makes a signal-mask system call per switch; on x86-64 a dozen instructions that swap
// the callee-saved registers and the stack pointer do the same job ~20x faster.  Other hosts fall back to ucontext.
#if defined(__x86_64__)
struct Context { void *sp = nullptr; };
extern "C" void emu_switch(void **save_sp, void *load_sp);
asm(R"(
    .text
    .globl emu_switch
    .type emu_switch,@function
emu_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
    .size emu_switch,.-emu_switch
)");
void context_make(Context *c, void *stack, size_t bytes, void (*entry)()) {
    uintptr_t top = (reinterpret_cast<uintptr_t>(stack) + bytes) & ~uintptr_t(15);
    void **sp = reinterpret_cast<void **>(top);
    *--sp = nullptr;                               // fake return address of `entry` (it never returns)
    *--sp = reinterpret_cast<void *>(entry);       // popped by `ret`
    for (int i = 0; i < 6; i++) *--sp = nullptr;   // rbp rbx r12 r13 r14 r15
    c->sp = sp;
}
void context_switch(Context *from, Context *to) { emu_switch(&from->sp, to->sp); }
#else
struct Context { ucontext_t uc; };
void context_make(Context *c, void *stack, size_t bytes, void (*entry)()) {
    getcontext(&c->uc);
    c->uc.uc_stack.ss_sp = stack;
    c->uc.uc_stack.ss_size = bytes;
    c->uc.uc_link = nullptr;
    makecontext(&c->uc, entry, 0);
}
void context_switch(Context *from, Context *to) { swapcontext(&from->uc, &to->uc); }
#endif

struct Fiber {
    Context uc;
    ThreadCtx tc;
    void *stack = nullptr;
    bool done = false;
    // wait condition: runnable when *wait_ptr != wait_val (nullptr = runnable)
    const volatile unsigned *wait_ptr = nullptr;
    unsigned wait_val = 0;
    const char *wait_what = "";
};

struct Machine {
    std::vector<Fiber> fibers;
    std::vector<WarpState> warps;
    std::vector<void *> stacks;
    Context sched;
    unsigned n_threads = 0, n_live = 0;
    unsigned bar_count = 0;
    volatile unsigned bar_gen = 0;
    // __syncthreads_or/and/count: accumulators of the barrier in flight, indexed by the parity of its generation
    int red_or[2] = {0, 0}, red_nand[2] = {0, 0}, red_cnt[2] = {0, 0};
    std::vector<unsigned char> smem;
    const std::function<void()> *body = nullptr;
    const char *kernel = "";
    Fiber *cur = nullptr;
};

Machine M;
std::recursive_mutex g_mu;

void yield_to_scheduler() { context_switch(&M.cur->uc, &M.sched); }

void wait_on(const volatile unsigned *p, unsigned val, const char *what) {
    Fiber *f = M.cur;
    f->wait_ptr = p;
    f->wait_val = val;
    f->wait_what = what;
    while (*p == val) yield_to_scheduler();
    f->wait_ptr = nullptr;
}

void barrier_completed() {
    M.bar_count = 0;
    M.bar_gen = M.bar_gen + 1;
    const unsigned nxt = M.bar_gen & 1;  // the accumulators of the NEXT barrier (its results were read two barriers ago)
    M.red_or[nxt] = M.red_nand[nxt] = M.red_cnt[nxt] = 0;
}

void release_barrier_if_complete() {
    if (M.bar_count && M.bar_count == M.n_live) barrier_completed();
}

void complete_rendezvous(WarpState &w, Rendezvous &r) {
    for (int i = 0; i < 32; i++) r.out.v[i] = r.slot[i];
    r.out.present = r.arrived;
    r.readers = (unsigned)__builtin_popcount(r.arrived);
    r.draining = true;
    r.gen++;
    (void)w;
}

void check_rendezvous_after_exit(WarpState &w) {
    for (Rendezvous &r : w.rv)
        if (r.used && !r.draining && r.arrived && (r.arrived & w.live) == (r.mask & w.live)) complete_rendezvous(w, r);
}

void trampoline() {
    Fiber *f = M.cur;
    (*M.body)();
    f->done = true;
    M.n_live--;
    WarpState &w = M.warps[f->tc.warp];
    w.live &= ~(1u << f->tc.lane);
    check_rendezvous_after_exit(w);
    release_barrier_if_complete();
    yield_to_scheduler();
    abort();  // never resumed
}

void *get_stack(size_t i) {
    while (M.stacks.size() <= i) {
        void *p = mmap(nullptr, STACK_BYTES, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (p == MAP_FAILED) { perror("emu: mmap stack"); abort(); }
        M.stacks.push_back(p);
    }
    return M.stacks[i];
}

[[noreturn]] void deadlock() {
    fprintf(stderr, "emu: DEADLOCK in kernel %s, block (%u,%u,%u): no thread can make progress\n", M.kernel,
            M.fibers[0].tc.bid.x, M.fibers[0].tc.bid.y, M.fibers[0].tc.bid.z);
    int shown = 0;
    for (Fiber &f : M.fibers)
        if (!f.done && shown++ < 16) fprintf(stderr, "  thread %u (warp %u lane %u) waits at %s\n", f.tc.linear, f.tc.warp, f.tc.lane, f.wait_what);
    abort();
}

void run_cta(const LaunchCfg &cfg, Idx bid) {
    const unsigned T = cfg.block.x * cfg.block.y * cfg.block.z;
    M.n_threads = M.n_live = T;
    M.bar_count = 0;
    M.red_or[0] = M.red_or[1] = M.red_nand[0] = M.red_nand[1] = M.red_cnt[0] = M.red_cnt[1] = 0;
    M.fibers.assign(T, Fiber());
    M.warps.assign((T + 31) / 32, WarpState());
    memset(M.smem.data(), 0xCD, M.smem.size());
    for (unsigned t = 0; t < T; t++) {
        Fiber &f = M.fibers[t];
        f.tc.tid = Idx{t % cfg.block.x, (t / cfg.block.x) % cfg.block.y, t / (cfg.block.x * cfg.block.y)};
        f.tc.bid = bid;
        f.tc.bdim = Idx{cfg.block.x, cfg.block.y, cfg.block.z};
        f.tc.gdim = Idx{cfg.grid.x, cfg.grid.y, cfg.grid.z};
        f.tc.linear = t;
        f.tc.lane = t & 31;
        f.tc.warp = t >> 5;
        M.warps[t >> 5].live |= 1u << (t & 31);
        f.stack = get_stack(t);
        context_make(&f.uc, f.stack, STACK_BYTES, trampoline);
    }
    // Scheduling order between yield points.  The CUDA model allows any; running the fibers of a CTA in reverse or in a
    // random order (EMU_SCHED=reverse | random[:seed]) makes a missing __syncthreads() show up as a wrong result with high
    // probability, because a thread then runs ahead of (or behind) the neighbours a forward sweep would have kept it next to.
    static int sched_mode = -1;
    static unsigned long long rng = 0x9E3779B97F4A7C15ull;
    if (sched_mode < 0) {
        const char *e = getenv("EMU_SCHED");
        sched_mode = !e ? 0 : (!strncmp(e, "reverse", 7) ? 1 : (!strncmp(e, "random", 6) ? 2 : 0));
        if (e && sched_mode == 2 && strchr(e, ':')) rng ^= strtoull(strchr(e, ':') + 1, nullptr, 10) * 0xD1342543DE82EF95ull;
    }
    std::vector<unsigned> order(T);
    for (unsigned t = 0; t < T; t++) order[t] = sched_mode == 1 ? T - 1 - t : t;
    while (M.n_live) {
        bool progress = false;
        if (sched_mode == 2)  // a fresh permutation for every sweep
            for (unsigned t = T - 1; t > 0; t--) {
                rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17;
                std::swap(order[t], order[rng % (t + 1)]);
            }
        for (unsigned oi = 0; oi < T; oi++) {
            const unsigned t = order[oi];
            Fiber &f = M.fibers[t];
            if (f.done) continue;
            if (f.wait_ptr && *f.wait_ptr == f.wait_val) continue;
            M.cur = &f;
            g_cur = &f.tc;
            context_switch(&M.sched, &f.uc);
            progress = true;
        }
        if (!progress) deadlock();
    }
    M.cur = nullptr;
    g_cur = nullptr;
}

// ---- device memory with canaries ----
struct Alloc { size_t bytes; bool host; };
std::map<void *, Alloc> g_allocs;
size_t g_alloc_bytes = 0;

void check_canary(void *user, const Alloc &a) {
    const unsigned char *b = static_cast<unsigned char *>(user);
    for (size_t i = 0; i < CANARY; i++)
        if (b[-(ptrdiff_t)CANARY + (ptrdiff_t)i] != 0xA5 || b[a.bytes + i] != 0xA5) {
            fprintf(stderr, "emu: OUT-OF-BOUNDS WRITE around a %zu-byte device block (%s side)\n", a.bytes,
                    b[-(ptrdiff_t)CANARY + (ptrdiff_t)i] != 0xA5 ? "low" : "high");
            abort();
        }
}

#ifdef EMU_ASAN
// AddressSanitizer build (tests/emu/run_asan.sh): blocks carry no padding of ours, so ASan's own red zones sit right behind the
// last byte and catch out-of-bounds READS as well (the GPU may fault on them, or silently read a neighbour's data)
cudaError_t alloc_common(void **p, size_t bytes, bool host) {
    void *raw = nullptr;
    if (posix_memalign(&raw, 256, bytes ? bytes : 1)) return cudaErrorMemoryAllocation;
    memset(raw, host ? 0x00 : 0xCB, bytes);
    *p = raw;
    g_allocs[*p] = Alloc{bytes, host};
    return cudaSuccess;
}
cudaError_t free_common(void *p) {
    if (!p) return cudaSuccess;
    g_allocs.erase(p);
    free(p);
    return cudaSuccess;
}
#define EMU_NO_CANARY 1
#endif
#ifndef EMU_NO_CANARY
cudaError_t alloc_common(void **p, size_t bytes, bool host) {
    unsigned char *raw = nullptr;
    if (posix_memalign(reinterpret_cast<void **>(&raw), 256, bytes + 256 + CANARY)) return cudaErrorMemoryAllocation;
    memset(raw, 0xA5, 256);
    memset(raw + 256, host ? 0x00 : 0xCB, bytes);
    memset(raw + 256 + bytes, 0xA5, CANARY);
    *p = raw + 256;
    g_allocs[*p] = Alloc{bytes, host};
    g_alloc_bytes += bytes;
    return cudaSuccess;
}

cudaError_t free_common(void *p) {
    if (!p) return cudaSuccess;
    auto it = g_allocs.find(p);
    if (it == g_allocs.end()) { fprintf(stderr, "emu: cudaFree of an unknown pointer %p\n", p); abort(); }
    check_canary(p, it->second);
    g_alloc_bytes -= it->second.bytes;
    g_allocs.erase(it);
    free(static_cast<unsigned char *>(p) - 256);
    return cudaSuccess;
}

struct AtExit {
    ~AtExit() {
        for (auto &kv : g_allocs) check_canary(kv.first, kv.second);
    }
} g_at_exit;
#endif

}  // namespace

// a fiber waits until *p != val, yielding to the other threads of its CTA (used by the shim's mbarrier stand-in)
void spin_while_equal(const volatile unsigned *p, unsigned val, const char *what) { wait_on(p, val, what); }

// 128-byte aligned (kernels may declare `extern __shared__ __align__(128)`: TMA destinations)
void *dyn_smem() { return reinterpret_cast<void *>((reinterpret_cast<uintptr_t>(M.smem.data()) + 127) & ~uintptr_t(127)); }

void cta_barrier() {
    M.bar_count++;
    const unsigned gen = M.bar_gen;
    if (M.bar_count == M.n_live) {
        barrier_completed();
        return;
    }
    wait_on(&M.bar_gen, gen, "__syncthreads()");
}

int cta_barrier_reduce(int pred, int op) {
    const unsigned slot = M.bar_gen & 1;
    M.red_or[slot] |= pred != 0;
    M.red_nand[slot] |= pred == 0;
    M.red_cnt[slot] += pred != 0;
    cta_barrier();
    return op == 0 ? M.red_or[slot] : (op == 1 ? !M.red_nand[slot] : M.red_cnt[slot]);
}

void warp_exchange(unsigned mask, unsigned long long v, int op, WarpVals *out) {
    Fiber *f = M.cur;
    WarpState &w = M.warps[f->tc.warp];
    const unsigned bit = 1u << f->tc.lane;
    if (!(mask & bit)) { fprintf(stderr, "emu: kernel %s: lane %u calls a warp collective with mask %08x that does not name it\n", M.kernel, f->tc.lane, mask); abort(); }
    Rendezvous *r = nullptr;
    for (Rendezvous &c : w.rv)
        if (c.used && !c.draining && c.mask == mask && !(c.arrived & bit)) { r = &c; break; }
    if (!r) {
        for (Rendezvous &c : w.rv)
            if (!c.used) { r = &c; break; }
        if (!r) { fprintf(stderr, "emu: kernel %s: too many concurrent warp rendezvous\n", M.kernel); abort(); }
        r->used = true; r->draining = false; r->mask = mask; r->arrived = 0; r->op = op;
    }
    if (r->op != op) {
        fprintf(stderr, "emu: kernel %s, warp %u: lanes of one rendezvous (mask %08x) call DIFFERENT collectives (%d vs %d)\n", M.kernel,
                f->tc.warp, mask, r->op, op);
        abort();
    }
    r->slot[f->tc.lane] = v;
    r->arrived |= bit;
    if ((r->arrived & w.live) == (mask & w.live)) {
        complete_rendezvous(w, *r);
    } else {
        const unsigned gen = r->gen;
        wait_on(&r->gen, gen, "a warp collective (__shfl/__ballot/... _sync)");
    }
    *out = r->out;
    if (--r->readers == 0) { r->used = false; r->draining = false; }
}

void launch(const LaunchCfg &cfg, const std::function<void()> &body, const char *name) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    const unsigned T = cfg.block.x * cfg.block.y * cfg.block.z;
    if (T == 0 || T > 1024 || cfg.grid.x == 0 || cfg.grid.y == 0 || cfg.grid.z == 0 || cfg.smem > (size_t)emu_max_dyn_smem()) {
        fprintf(stderr, "emu: invalid launch configuration for %s: grid (%u,%u,%u) block (%u,%u,%u) smem %zu\n", name, cfg.grid.x, cfg.grid.y,
                cfg.grid.z, cfg.block.x, cfg.block.y, cfg.block.z, cfg.smem);
        abort();
    }
    M.body = &body;
    M.kernel = name;
#ifdef EMU_ASAN
    M.smem = std::vector<unsigned char>(cfg.smem + 127);  // fresh block (+ alignment slack): ASan sees accesses past the dynamic shared memory
#else
    M.smem.resize(cfg.smem + 64 + 127);
#endif
    for (unsigned z = 0; z < cfg.grid.z; z++)
        for (unsigned y = 0; y < cfg.grid.y; y++)
            for (unsigned x = 0; x < cfg.grid.x; x++) run_cta(cfg, Idx{x, y, z});
    M.body = nullptr;
}

}  // namespace emu

// ---- runtime API ----------------------------------------------------------------------------------------------
struct emu_stream { int id; };
struct emu_event { double t_ms; };
static emu_stream g_default_stream{0};

int emu_max_dyn_smem() { return 227 * 1024; }
int emu_blocks_per_sm() { const char *e = getenv("EMU_BLOCKS_PER_SM"); return e ? atoi(e) : 2; }

cudaError_t cudaMalloc(void **p, size_t bytes) { return emu::alloc_common(p, bytes, false); }
cudaError_t cudaFree(void *p) { return emu::free_common(p); }
cudaError_t cudaMallocHost(void **p, size_t bytes) { return emu::alloc_common(p, bytes, true); }
cudaError_t cudaFreeHost(void *p) { return emu::free_common(p); }
cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { if (n) memmove(d, s, n); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t) { if (n) memmove(d, s, n); return cudaSuccess; }
cudaError_t cudaMemset(void *d, int v, size_t n) { if (n) memset(d, v, n); return cudaSuccess; }
cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t) { if (n) memset(d, v, n); return cudaSuccess; }
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = new emu_stream{1}; return cudaSuccess; }
cudaError_t cudaStreamDestroy(cudaStream_t s) { if (s != &g_default_stream) delete s; return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = new emu_event{0.0}; return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { return cudaEventCreate(e); }
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }  // kernels run synchronously here
cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) { e->t_ms = now_ms(); return cudaSuccess; }
cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b) { *ms = float(b->t_ms - a->t_ms); return cudaSuccess; }
cudaError_t cudaGetLastError() { return cudaSuccess; }
const char *cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : (e == cudaErrorMemoryAllocation ? "out of memory" : "emulated CUDA error"); }
cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return cudaSuccess; }
cudaError_t cudaGetDevice(int *d) { *d = 0; return cudaSuccess; }
cudaError_t cudaSetDevice(int d) { return d == 0 ? cudaSuccess : cudaErrorInvalidValue; }
cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int) {
    memset(p, 0, sizeof(*p));
    snprintf(p->name, sizeof(p->name), "cniic host emulation (tests/emu)");
    const char *e = getenv("EMU_SM_COUNT");
    p->multiProcessorCount = e ? atoi(e) : 3;
    p->major = 10;
    p->minor = 0;
    p->totalGlobalMem = size_t(8) << 30;
    p->sharedMemPerBlockOptin = emu_max_dyn_smem();
    return cudaSuccess;
}
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t *, void *) { return cudaErrorNotSupported; }
cudaError_t cudaIpcOpenMemHandle(void **, cudaIpcMemHandle_t, unsigned) { return cudaErrorNotSupported; }
cudaError_t cudaIpcCloseMemHandle(void *) { return cudaErrorNotSupported; }
