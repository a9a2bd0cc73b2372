// Host stand-in for <cuda_runtime.h>: TEST INFRASTRUCTURE, never part of the product.
//
// The device code of this repository (sdf_b200/csrc/sdf_ops.cuh, sdf_kernels.cuh) is written as
// templated `__device__` bodies.  tests/util_hostsim.py compiles those very files with g++ against
// this header, so that the CPU test suite (no GPU in the build container) executes the kernels' own
// source: CUDA threads become fibers of one host thread, a CTA is a set of fibers run round-robin
// (ascending, descending or randomly permuted: State::schedule), `__syncthreads` / warp collectives are rendezvous points between fibers, atomics are
// plain read-modify-writes (one host thread), `__shared__` is function-static storage that every CTA
// finds filled with garbage (CTAs run one after another), a TMA bulk copy is a memcpy that completes its mbarrier at once.
//
// What this does NOT model: memory ordering, bank conflicts, MUFU approximation error (the FAST math
// policy runs its own formulas -- Cody-Waite reduction, polynomial atan2, the rotation form of the
// circular array -- on correctly rounded host functions, optionally perturbed: State::fast_ulps),
// timing.  It checks the LOGIC of the kernels: index arithmetic, masks, compaction, list protocols,
// table look-ups, the marching-cubes arithmetic.
#pragma once
#include <algorithm>
#include <float.h>
#include <functional>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <ucontext.h>
#include <vector>

#define HOSTSIM 1
#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __constant__ static
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))
#define __launch_bounds__(...)
#define __grid_constant__

struct uint2 { unsigned int x, y; };
struct uint3 { unsigned int x, y, z; };
struct alignas(16) uint4 { unsigned int x, y, z, w; };
struct alignas(16) int4 { int x, y, z, w; };
struct float3 { float x, y, z; };
struct dim3 { unsigned int x, y, z; dim3(unsigned int a = 1, unsigned int b = 1, unsigned int c = 1) : x(a), y(b), z(c) {} };
static inline uint2 make_uint2(unsigned int x, unsigned int y) { uint2 r; r.x = x; r.y = y; return r; }

namespace hostsim {

struct Warp {
    uint32_t alive = 0, arrived = 0, gen = 0;
    uint64_t buf[2][32];
    uint32_t part[2] = {0, 0};          // lanes that took part in the collective of buffer 0 / 1
};

// Context switch between fibers: a dozen instructions on x86-64 (glibc's swapcontext makes a signal-mask system call
// per switch, a third of the suite's run time); ucontext elsewhere and under AddressSanitizer, which knows swapcontext.
#if defined(__x86_64__) && !defined(__SANITIZE_ADDRESS__) && !defined(HOSTSIM_UCONTEXT)
#define HOSTSIM_ASM_SWITCH 1
extern "C" void hostsim_switch(void** save_sp, void* load_sp);
__asm__(".text\n"
        ".weak hostsim_switch\n"
        ".type hostsim_switch,@function\n"
        "hostsim_switch:\n"
        "  pushq %rbp\n  pushq %rbx\n  pushq %r12\n  pushq %r13\n  pushq %r14\n  pushq %r15\n"
        "  movq %rsp, (%rdi)\n"
        "  movq %rsi, %rsp\n"
        "  popq %r15\n  popq %r14\n  popq %r13\n  popq %r12\n  popq %rbx\n  popq %rbp\n"
        "  ret\n"
        ".size hostsim_switch,.-hostsim_switch\n");
struct Context { void* sp = nullptr; };
#else
struct Context { ucontext_t uc; };
#endif

struct Fiber {
    Context ctx;
    bool done = false;
};

struct State {
    uint3 thread_idx{0, 0, 0}, block_idx{0, 0, 0}, block_dim{1, 1, 1}, grid_dim{1, 1, 1};
    int cur = -1;
    Context sched;
    std::vector<Fiber> fibers;
    std::vector<char*> stacks;
    unsigned alive = 0, bar_arrived = 0, bar_gen = 0;
    Warp warps[32];
    unsigned long progress = 0;
    unsigned char* dyn = nullptr;
    size_t dyn_cap = 0;
    std::function<void()> body;
    unsigned long launches = 0, switches = 0;
    // FAST-policy error injection: relative perturbation of approximate div / sqrt / sin / cos results, in units
    // of 2^-23 (0 = none); a deterministic hash of the operands picks the sign and magnitude
    double fast_ulps = 0.0;
    // thread scheduling order (see launch): 0 ascending, 1 descending, >= 2 random permutation per pass, seeded by it
    int schedule = 0;
    unsigned long long rng = 1;
    std::vector<unsigned> order;
};

inline State S;
static constexpr size_t kStackBytes = 256 * 1024;

inline void die(const char* what) {
    fprintf(stderr, "hostsim: %s (block %u, thread %u)\n", what, S.block_idx.x, S.thread_idx.x);
    abort();
}

inline void switch_to(Context& from, Context& to) {
#ifdef HOSTSIM_ASM_SWITCH
    hostsim_switch(&from.sp, to.sp);
#else
    swapcontext(&from.uc, &to.uc);
#endif
}

// AddressSanitizer tracks one stack per thread: every fiber switch is announced (sanitizer/common_interface_defs.h), or
// its stack bookkeeping -- and the reports that rest on it -- would be guesses
#ifdef __SANITIZE_ADDRESS__
extern "C" void __sanitizer_start_switch_fiber(void** fake_stack_save, const void* bottom, size_t size);
extern "C" void __sanitizer_finish_switch_fiber(void* fake_stack_save, const void** bottom_old, size_t* size_old);
inline const void* g_sched_stack = nullptr;
inline size_t g_sched_stack_size = 0;
inline void* g_sched_fake = nullptr;
inline std::vector<void*> g_fiber_fake;
#endif

// scheduler -> fiber t
inline void enter_fiber(unsigned t) {
#ifdef __SANITIZE_ADDRESS__
    __sanitizer_start_switch_fiber(&g_sched_fake, S.stacks[t], kStackBytes);
    switch_to(S.sched, S.fibers[t].ctx);
    __sanitizer_finish_switch_fiber(g_sched_fake, nullptr, nullptr);
#else
    switch_to(S.sched, S.fibers[t].ctx);
#endif
}

// fiber -> scheduler (`dying`: the fiber never runs again)
inline void leave_fiber(bool dying) {
    Fiber& f = S.fibers[(size_t)S.cur];
#ifdef __SANITIZE_ADDRESS__
    const size_t me = (size_t)S.cur;
    __sanitizer_start_switch_fiber(dying ? nullptr : &g_fiber_fake[me], g_sched_stack, g_sched_stack_size);
    switch_to(f.ctx, S.sched);
    __sanitizer_finish_switch_fiber(g_fiber_fake[me], &g_sched_stack, &g_sched_stack_size);
#else
    (void)dying;
    switch_to(f.ctx, S.sched);
#endif
}

inline void yield() {
    ++S.switches;
    leave_fiber(false);
}

inline void release_block_barrier() { S.bar_arrived = 0; ++S.bar_gen; ++S.progress; }
inline void release_warp(Warp& w) { w.part[w.gen & 1u] = w.alive; w.arrived = 0; ++w.gen; ++S.progress; }

inline void sync_block() {
    const unsigned g = S.bar_gen;
    ++S.bar_arrived; ++S.progress;
    if (S.bar_arrived == S.alive) { release_block_barrier(); return; }
    while (S.bar_gen == g) yield();
    ++S.progress;
}

// every live lane of the warp deposits a 64-bit word; returns the buffer all of them can read until the
// warp's next collective has completed (two buffers alternate)
inline const uint64_t* warp_gather(unsigned mask, uint64_t v, uint32_t* alive_out = nullptr) {
    if (mask != 0xffffffffu) die("warp collectives are modelled for the full mask only");
    const unsigned tid = S.thread_idx.x;
    Warp& w = S.warps[tid >> 5];
    const uint32_t g = w.gen;
    uint64_t* b = w.buf[g & 1u];
    b[tid & 31u] = v;
    ++w.arrived; ++S.progress;
    if (w.arrived == (uint32_t)__builtin_popcount(w.alive)) release_warp(w);
    else {
        while (w.gen == g) yield();
        ++S.progress;
    }
    if (alive_out) *alive_out = w.part[g & 1u];       // (not w.alive: lanes resumed earlier may have exited since)
    return b;
}

inline void fiber_exit() {
    const unsigned tid = S.thread_idx.x;
    Fiber& f = S.fibers[(size_t)S.cur];
    f.done = true;
    --S.alive; ++S.progress;
    Warp& w = S.warps[tid >> 5];
    w.alive &= ~(1u << (tid & 31u));
    // an exited thread no longer takes part: barriers the others wait at may now be complete
    if (w.arrived != 0 && w.arrived == (uint32_t)__builtin_popcount(w.alive)) release_warp(w);
    if (S.bar_arrived != 0 && S.bar_arrived == S.alive) release_block_barrier();
    leave_fiber(true);
    die("resumed a finished fiber");
}

inline void trampoline() {
#ifdef __SANITIZE_ADDRESS__
    __sanitizer_finish_switch_fiber(nullptr, &g_sched_stack, &g_sched_stack_size);        // first time on this stack
#endif
    S.body();
    fiber_exit();
}

inline unsigned char* dyn_smem() { return S.dyn; }

// Shared memory is neither zero nor the previous CTA's on arrival: the patched sources call note_shared() right after
// every `__shared__` declaration (tests/util_hostsim.py); the first thread of a CTA to get there fills the variable
// with garbage.  Dynamic shared memory is filled by launch().
struct SharedSeen { unsigned long launch; unsigned block; };
inline std::vector<std::pair<void*, SharedSeen>>& shared_seen() { static std::vector<std::pair<void*, SharedSeen>> v; return v; }
inline void note_shared(void* p, size_t bytes) {
    for (auto& e : shared_seen())
        if (e.first == p) {
            if (e.second.launch == S.launches && e.second.block == S.block_idx.x) return;
            e.second = SharedSeen{S.launches, S.block_idx.x};
            memset(p, 0xcd, bytes);
            return;
        }
    shared_seen().push_back({p, SharedSeen{S.launches, S.block_idx.x}});
    memset(p, 0xcd, bytes);
}

// Run `body` as a grid of `grid` CTAs of `block` threads (1-D), CTA after CTA.
inline void launch(unsigned grid, unsigned block, size_t dyn_bytes, std::function<void()> body) {
    if (block == 0 || block > 1024) die("bad block size");
    static const bool env_read = [] {              // $HOSTSIM_SCHEDULE: the schedule for a whole process (pytest --hostsim)
        if (const char* e = getenv("HOSTSIM_SCHEDULE")) { S.schedule = atoi(e); S.rng = 0x9e3779b97f4a7c15ull * (unsigned long long)(S.schedule + 1); }
        return true;
    }();
    (void)env_read;
    if (dyn_bytes > S.dyn_cap) {
        free(S.dyn);
        S.dyn = (unsigned char*)aligned_alloc(128, (dyn_bytes + 127) & ~(size_t)127);
        S.dyn_cap = dyn_bytes;
    }
    while (S.stacks.size() < block) {
        void* p = mmap(nullptr, kStackBytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_STACK, -1, 0);
        if (p == MAP_FAILED) die("fiber stack allocation failed");
        S.stacks.push_back((char*)p);
    }
    ++S.launches;
#ifdef __SANITIZE_ADDRESS__
    if (g_fiber_fake.size() < block) g_fiber_fake.resize(block, nullptr);
#endif
    S.body = std::move(body);
    S.grid_dim = uint3{grid, 1, 1};
    S.block_dim = uint3{block, 1, 1};
    S.fibers.resize(block);
    // CTAs run one after another, in the order the schedule names (which CTA claims which work item, which one finishes last)
    std::vector<unsigned> cta_order(grid);
    for (unsigned i = 0; i < grid; ++i) cta_order[i] = S.schedule == 1 ? grid - 1 - i : i;
    if (S.schedule >= 2)
        for (unsigned i = grid; i > 1; --i) {
            S.rng = S.rng * 6364136223846793005ull + 1442695040888963407ull;
            std::swap(cta_order[i - 1], cta_order[(unsigned)((S.rng >> 33) % i)]);
        }
    for (unsigned bi = 0; bi < grid; ++bi) {
        const unsigned b = cta_order[bi];
        S.block_idx = uint3{b, 0, 0};
        S.alive = block; S.bar_arrived = 0;
        if (S.dyn && dyn_bytes) memset(S.dyn, 0xcd, dyn_bytes);
        for (unsigned wi = 0; wi < 32; ++wi) {
            Warp& w = S.warps[wi];
            const unsigned lo = wi * 32;
            const unsigned n = lo >= block ? 0u : std::min(32u, block - lo);
            w.alive = n == 32 ? 0xffffffffu : ((1u << n) - 1u);
            w.arrived = 0;
        }
        for (unsigned t = 0; t < block; ++t) {
            Fiber& f = S.fibers[t];
            f.done = false;
#ifdef HOSTSIM_ASM_SWITCH
            // a fresh stack that hostsim_switch "returns" into trampoline() from: six callee-saved registers, then the
            // return address in a 16-byte aligned slot (so that the stack is aligned as after a call)
            void** slot = (void**)(S.stacks[t] + kStackBytes - 16);
            slot[0] = (void*)trampoline;
            for (int r = 1; r <= 6; ++r) slot[-r] = nullptr;
            f.ctx.sp = (void*)(slot - 6);
#else
            getcontext(&f.ctx.uc);
            f.ctx.uc.uc_stack.ss_sp = S.stacks[t];
            f.ctx.uc.uc_stack.ss_size = kStackBytes;
            f.ctx.uc.uc_link = nullptr;
            makecontext(&f.ctx.uc, (void (*)())trampoline, 0);
#endif
        }
        for (;;) {
            const unsigned long before = S.progress;
            bool any = false;
            for (unsigned slot = 0; slot < block; ++slot) {
                // the order in which ready threads get the host thread: ascending, descending, or a fresh pseudo-random
                // permutation every pass -- a result that depends on it is a missing barrier (tests: racecheck)
                unsigned t = slot;
                if (S.schedule == 1) t = block - 1 - slot;
                else if (S.schedule >= 2) {
                    if (slot == 0) {
                        S.order.resize(block);
                        for (unsigned i = 0; i < block; ++i) S.order[i] = i;
                        for (unsigned i = block - 1; i > 0; --i) {
                            S.rng = S.rng * 6364136223846793005ull + 1442695040888963407ull;
                            std::swap(S.order[i], S.order[(unsigned)((S.rng >> 33) % (i + 1))]);
                        }
                    }
                    t = S.order[slot];
                }
                Fiber& f = S.fibers[t];
                if (f.done) continue;
                any = true;
                S.cur = (int)t;
                S.thread_idx = uint3{t, 0, 0};
                enter_fiber(t);
            }
            if (!any) break;
            if (S.progress == before) die("deadlock: a full pass over the CTA's threads changed nothing");
        }
    }
    S.cur = -1;
    S.body = nullptr;
}

template <class T> inline uint64_t to_bits(T v) { uint64_t b = 0; static_assert(sizeof(T) <= 8, ""); memcpy(&b, &v, sizeof(T)); return b; }
template <class T> inline T from_bits(uint64_t b) { T v; memcpy(&v, &b, sizeof(T)); return v; }

// deterministic pseudo-random relative perturbation in [-ulps, +ulps] * 2^-23 (FAST-policy error injection)
inline float perturb(float r, float a, float b) {
    if (S.fast_ulps == 0.0 || !isfinite(r)) return r;
    uint32_t h = (uint32_t)to_bits(a) * 0x9e3779b1u ^ ((uint32_t)to_bits(b) + 0x7f4a7c15u) * 0x85ebca6bu;
    h ^= h >> 15; h *= 0x2c1b3c6du; h ^= h >> 12;
    const double u = ((double)(h & 0xffffffu) / (double)0x7fffff) - 1.0;           // [-1, 1]
    return (float)((double)r * (1.0 + u * S.fast_ulps * 1.1920928955078125e-07));
}
inline float perturb_abs(float r, float a, double abs_err) {
    if (S.fast_ulps == 0.0 || !isfinite(r)) return r;
    uint32_t h = (uint32_t)to_bits(a) * 0x9e3779b1u + 0x165667b1u;
    h ^= h >> 15; h *= 0x2c1b3c6du; h ^= h >> 12;
    const double u = ((double)(h & 0xffffffu) / (double)0x7fffff) - 1.0;
    return (float)((double)r + u * abs_err);
}
inline float sqrt_approx(float x) { return perturb(sqrtf(x), x, 0.5f); }

// mbarrier + TMA bulk copy: the 64-bit barrier word counts completed phases; a bulk copy is a memcpy that
// completes the phase at once
inline void mbar_init(uint64_t* bar, unsigned) { *bar = 0; }
inline void mbar_expect_tx(uint64_t*, unsigned) {}
inline void mbar_wait(uint64_t* bar, unsigned parity) {
    while ((unsigned)(*bar & 1u) == (parity & 1u)) yield();
    ++S.progress;
}
inline void tma_bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
    if ((bytes & 15u) != 0 || ((uintptr_t)dst & 15u) != 0 || ((uintptr_t)src & 15u) != 0) die("cp.async.bulk: 16-byte alignment / size");
    memcpy(dst, src, bytes);
    *bar += 1; ++S.progress;
}

}  // namespace hostsim

#define threadIdx (hostsim::S.thread_idx)
#define blockIdx (hostsim::S.block_idx)
#define blockDim (hostsim::S.block_dim)
#define gridDim (hostsim::S.grid_dim)

// ------------------------------------------------------------------------------- synchronisation, collectives
static inline void __syncthreads() { hostsim::sync_block(); }
static inline void __syncwarp(unsigned mask = 0xffffffffu) { hostsim::warp_gather(mask, 0); }
static inline void __threadfence() {}
static inline void __threadfence_system() {}
static inline void __threadfence_block() {}

template <class T> static inline T __shfl_sync(unsigned mask, T v, int src) {
    return hostsim::from_bits<T>(hostsim::warp_gather(mask, hostsim::to_bits(v))[src & 31]);
}
template <class T> static inline T __shfl_xor_sync(unsigned mask, T v, int lanemask) {
    const unsigned lane = threadIdx.x & 31u;
    return hostsim::from_bits<T>(hostsim::warp_gather(mask, hostsim::to_bits(v))[(lane ^ (unsigned)lanemask) & 31u]);
}
template <class T> static inline T __shfl_up_sync(unsigned mask, T v, unsigned delta) {
    const unsigned lane = threadIdx.x & 31u;
    const uint64_t* b = hostsim::warp_gather(mask, hostsim::to_bits(v));
    return lane >= delta ? hostsim::from_bits<T>(b[lane - delta]) : v;
}
static inline unsigned __ballot_sync(unsigned mask, bool pred) {
    uint32_t alive = 0;
    const uint64_t* b = hostsim::warp_gather(mask, pred ? 1u : 0u, &alive);
    unsigned r = 0;
    for (int i = 0; i < 32; ++i) if (((alive >> i) & 1u) && b[i]) r |= 1u << i;
    return r;
}
static inline bool __any_sync(unsigned mask, bool pred) { return __ballot_sync(mask, pred) != 0u; }
static inline bool __all_sync(unsigned mask, bool pred) { return __ballot_sync(mask, !pred) == 0u; }

// ------------------------------------------------------------------------------- atomics (one host thread)
template <class T> static inline T hostsim_rmw_add(T* p, T v) { const T o = *p; *p = (T)(o + v); return o; }
static inline unsigned int atomicAdd(unsigned int* p, unsigned int v) { return hostsim_rmw_add(p, v); }
static inline int atomicAdd(int* p, int v) { return hostsim_rmw_add(p, v); }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return hostsim_rmw_add(p, v); }
static inline unsigned long long atomicAdd_system(unsigned long long* p, unsigned long long v) { return hostsim_rmw_add(p, v); }
static inline int atomicMin(int* p, int v) { const int o = *p; if (v < o) *p = v; return o; }
static inline int atomicMax(int* p, int v) { const int o = *p; if (v > o) *p = v; return o; }
static inline unsigned int atomicMax(unsigned int* p, unsigned int v) { const unsigned int o = *p; if (v > o) *p = v; return o; }
static inline unsigned int atomicOr(unsigned int* p, unsigned int v) { const unsigned int o = *p; *p = o | v; return o; }

// ------------------------------------------------------------------------------- loads, bit tricks
template <class T> static inline T __ldg(const T* p) { return *p; }
template <class T> static inline T __ldcg(const T* p) { return *p; }
static inline size_t __cvta_generic_to_shared(const void* p) { return (size_t)p; }
static inline int __popc(unsigned int v) { return __builtin_popcount(v); }
static inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __ffsll(long long v) { return __builtin_ffsll(v); }
static inline unsigned int __umulhi(unsigned int a, unsigned int b) { return (unsigned int)(((unsigned long long)a * b) >> 32); }
static inline int __float_as_int(float f) { return hostsim::from_bits<int>(hostsim::to_bits(f)); }
static inline unsigned int __float_as_uint(float f) { return hostsim::from_bits<unsigned int>(hostsim::to_bits(f)); }
static inline float __int_as_float(int i) { return hostsim::from_bits<float>(hostsim::to_bits(i)); }
static inline double __longlong_as_double(long long v) { return hostsim::from_bits<double>(hostsim::to_bits(v)); }
static inline long long clock64() { return 0; }
static inline void __nanosleep(unsigned) { hostsim::yield(); }
static inline void __trap() { hostsim::die("__trap"); }

// ------------------------------------------------------------------------------- arithmetic with fixed rounding
// (the translation unit is compiled with -ffp-contract=off: a plain operation IS the _rn intrinsic)
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline double __drcp_rn(double a) { return 1.0 / a; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __fsqrt_rn(float a) { return sqrtf(a); }
// approximate units of the FAST math policy
static inline float __fdividef(float a, float b) { return hostsim::perturb(a / b, a, b); }
static inline void hostsim_fast_sincosf(float a, float* s, float* c) {
    float ss, cc;
    sincosf(a, &ss, &cc);
    *s = hostsim::perturb_abs(ss, a, 4.0e-7 * (hostsim::S.fast_ulps > 0 ? 1.0 : 0.0));
    *c = hostsim::perturb_abs(cc, -a, 4.0e-7 * (hostsim::S.fast_ulps > 0 ? 1.0 : 0.0));
}

#define __sincosf hostsim_fast_sincosf          // (glibc declares a __sincosf of its own)

// CUDA's global min / max overloads
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline unsigned int min(unsigned int a, unsigned int b) { return a < b ? a : b; }
static inline unsigned int max(unsigned int a, unsigned int b) { return a > b ? a : b; }
static inline unsigned int min(int a, unsigned int b) { return (unsigned int)a < b ? (unsigned int)a : b; }
static inline unsigned int min(unsigned int a, int b) { return a < (unsigned int)b ? a : (unsigned int)b; }
static inline long long min(long long a, long long b) { return a < b ? a : b; }
static inline long long max(long long a, long long b) { return a > b ? a : b; }
static inline unsigned long long min(unsigned long long a, unsigned long long b) { return a < b ? a : b; }
static inline unsigned long long max(unsigned long long a, unsigned long long b) { return a > b ? a : b; }

// =====================================================================================================
// Host API stand-in: what sdfb200.cu / sdfb200_weld.cu call.  Everything executes at call time (a stream is a
// label), device memory is host memory kept in a registry so that cudaPointerGetAttributes can tell it from
// the caller's buffers, one fake device with HOSTSIM_SMS "SMs" (default 3: small grids, fast emulation).
// A captured stream records closures that a graph launch replays; a real cubin cannot be loaded (the library then
// runs its interpreter kernels), the "cubin" of the stand-in NVRTC (fake_nvrtc.cpp) is a host shared object.
#include <dlfcn.h>
#include <unistd.h>
#include <chrono>
#include <map>
#include <tuple>
#include <utility>

typedef int cudaError_t;
enum {
    cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorMemoryAllocation = 2, cudaErrorNoDevice = 100,
    cudaErrorInvalidDevice = 101, cudaErrorNotReady = 600, cudaErrorNotSupported = 801
};
// A stream is a label; while it is being captured its operations are recorded as closures (kernel parameters copied at
// that moment, as CUDA does) and a graph launch replays them.
struct hostsim_stream_ { int id; bool capturing = false; std::vector<std::function<void()>> ops; };
struct hostsim_event_ { std::chrono::steady_clock::time_point t; };
typedef hostsim_stream_* cudaStream_t;
typedef hostsim_event_* cudaEvent_t;
struct hostsim_graph_ { std::vector<std::function<void()>> ops; };
struct hostsim_graph_exec_ { std::vector<std::function<void()>> ops; };
typedef hostsim_graph_* cudaGraph_t;
typedef hostsim_graph_exec_* cudaGraphExec_t;
typedef struct hostsim_kernel_* cudaKernel_t;
typedef struct hostsim_library_* cudaLibrary_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3, cudaMemcpyDefault = 4 };
enum cudaMemoryType { cudaMemoryTypeUnregistered = 0, cudaMemoryTypeHost = 1, cudaMemoryTypeDevice = 2, cudaMemoryTypeManaged = 3 };
struct cudaPointerAttributes { cudaMemoryType type; int device; void* devicePointer; void* hostPointer; };
struct cudaIpcMemHandle_t { char reserved[64]; };
enum { cudaStreamNonBlocking = 1, cudaHostAllocPortable = 1, cudaHostAllocMapped = 2, cudaHostRegisterPortable = 1,
       cudaEventRecordExternal = 1, cudaEventDisableTiming = 2, cudaIpcMemLazyEnablePeerAccess = 1,
       cudaStreamCaptureModeThreadLocal = 1, cudaDevAttrMultiProcessorCount = 16,
       cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };

namespace hostsim {
struct Region { size_t bytes; cudaMemoryType type; bool owned; };
inline std::map<uintptr_t, Region>& regions() { static std::map<uintptr_t, Region> r; return r; }
inline const Region* find_region(const void* p, uintptr_t* base = nullptr) {
    auto& r = regions();
    auto it = r.upper_bound((uintptr_t)p);
    if (it == r.begin()) return nullptr;
    --it;
    if ((uintptr_t)p >= it->first + it->second.bytes) return nullptr;
    if (base) *base = it->first;
    return &it->second;
}
inline cudaError_t alloc(void** out, size_t bytes, cudaMemoryType type) {
    void* p = aligned_alloc(256, (std::max<size_t>(bytes, 1) + 255) & ~(size_t)255);
    if (!p) return cudaErrorMemoryAllocation;
    memset(p, 0xcd, std::max<size_t>(bytes, 1));                // device memory is not zero on arrival
    regions()[(uintptr_t)p] = Region{std::max<size_t>(bytes, 1), type, true};
    *out = p;
    return cudaSuccess;
}
inline int device_count() { const char* e = getenv("HOSTSIM_DEVICES"); return e ? atoi(e) : 1; }
inline int sm_count() { const char* e = getenv("HOSTSIM_SMS"); return e && atoi(e) > 0 ? atoi(e) : 3; }

inline bool capturing(cudaStream_t s) { return s != nullptr && s->capturing; }

template <class... KA, class... A>
inline void launch_kernel(void (*k)(KA...), unsigned grid, unsigned block, size_t smem, cudaStream_t s, A&&... args) {
    std::tuple<std::decay_t<KA>...> held(std::forward<A>(args)...);       // kernel parameters are copied at launch
    if (capturing(s)) { s->ops.push_back([=] { launch(grid, block, smem, [&] { std::apply(k, held); }); }); return; }
    launch(grid, block, smem, [&] { std::apply(k, held); });
}
// cudaLaunchKernel's `void** args` unpacked with the parameter types of the kernel's own signature
template <class... KA, size_t... I>
inline void launch_args_impl(void (*k)(KA...), unsigned grid, unsigned block, size_t smem, void** args, cudaStream_t s,
                             std::index_sequence<I...>) {
    std::tuple<std::decay_t<KA>...> held(*(std::decay_t<KA>*)args[I]...);
    if (capturing(s)) { s->ops.push_back([=] { launch(grid, block, smem, [&] { std::apply(k, held); }); }); return; }
    launch(grid, block, smem, [&] { std::apply(k, held); });
}
template <class... KA>
inline void launch_args(void (*k)(KA...), unsigned grid, unsigned block, size_t smem, void** args, cudaStream_t s) {
    launch_args_impl(k, grid, block, smem, args, s, std::index_sequence_for<KA...>());
}
}  // namespace hostsim

static inline const char* cudaGetErrorString(cudaError_t e) {
    switch (e) {
        case cudaSuccess: return "no error";
        case cudaErrorMemoryAllocation: return "out of memory";
        case cudaErrorNotSupported: return "operation not supported (host execution)";
        case cudaErrorNotReady: return "not ready";
        default: return "error (host execution)";
    }
}
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = hostsim::device_count(); return *n > 0 ? cudaSuccess : cudaErrorNoDevice; }
static inline cudaError_t cudaSetDevice(int d) { return d >= 0 && d < hostsim::device_count() ? cudaSuccess : cudaErrorInvalidDevice; }
static inline cudaError_t cudaDeviceGetAttribute(int* v, int attr, int) {
    *v = attr == cudaDevAttrMultiProcessorCount ? hostsim::sm_count() : 0;
    return cudaSuccess;
}
template <class P> static inline cudaError_t cudaMalloc(P** p, size_t bytes) { return hostsim::alloc((void**)p, bytes, cudaMemoryTypeDevice); }
template <class P> static inline cudaError_t cudaMallocHost(P** p, size_t bytes) { return hostsim::alloc((void**)p, bytes, cudaMemoryTypeHost); }
static inline cudaError_t cudaHostAlloc(void** p, size_t bytes, unsigned) { return hostsim::alloc(p, bytes, cudaMemoryTypeHost); }
static inline cudaError_t cudaHostGetDevicePointer(void** d, void* h, unsigned) { *d = h; return cudaSuccess; }
static inline cudaError_t cudaFree(void* p) {
    if (!p) return cudaSuccess;
    auto it = hostsim::regions().find((uintptr_t)p);
    if (it == hostsim::regions().end() || !it->second.owned) return cudaErrorInvalidValue;
    hostsim::regions().erase(it);
    free(p);
    return cudaSuccess;
}
static inline cudaError_t cudaFreeHost(void* p) { return cudaFree(p); }
static inline cudaError_t cudaHostRegister(void* p, size_t bytes, unsigned) {
    hostsim::regions()[(uintptr_t)p] = hostsim::Region{bytes, cudaMemoryTypeHost, false};
    return cudaSuccess;
}
static inline cudaError_t cudaHostUnregister(void* p) { hostsim::regions().erase((uintptr_t)p); return cudaSuccess; }
static inline cudaError_t cudaPointerGetAttributes(cudaPointerAttributes* a, const void* p) {
    const hostsim::Region* r = hostsim::find_region(p);
    a->type = r ? r->type : cudaMemoryTypeUnregistered;
    a->device = 0; a->devicePointer = (void*)p; a->hostPointer = (void*)p;
    return cudaSuccess;
}
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { if (n) memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t st = nullptr) {
    if (hostsim::capturing(st)) { st->ops.push_back([=] { if (n) memmove(d, s, n); }); return cudaSuccess; }
    if (n) memmove(d, s, n);
    return cudaSuccess;
}
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t st = nullptr) {
    if (hostsim::capturing(st)) { st->ops.push_back([=] { if (n) memset(d, v, n); }); return cudaSuccess; }
    if (n) memset(d, v, n);
    return cudaSuccess;
}
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = new hostsim_stream_(); (*s)->id = 1; return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t s) { delete s; return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamQuery(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new hostsim_event_(); return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = new hostsim_event_(); return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t st = nullptr) {
    if (hostsim::capturing(st)) { st->ops.push_back([=] { e->t = std::chrono::steady_clock::now(); }); return cudaSuccess; }
    e->t = std::chrono::steady_clock::now();
    return cudaSuccess;
}
static inline cudaError_t cudaEventRecordWithFlags(cudaEvent_t e, cudaStream_t st, unsigned) { return cudaEventRecord(e, st); }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) {
    *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count();
    return cudaSuccess;
}
// stream capture -> graph: the recorded closures, replayed in order by every launch ($HOSTSIM_NO_CAPTURE refuses the
// capture, which sends the library down its plain-launch fallback)
static inline cudaError_t cudaStreamBeginCapture(cudaStream_t s, int) {
    if (!s || s->capturing || getenv("HOSTSIM_NO_CAPTURE")) return cudaErrorNotSupported;
    s->capturing = true; s->ops.clear();
    return cudaSuccess;
}
static inline cudaError_t cudaStreamEndCapture(cudaStream_t s, cudaGraph_t* g) {
    *g = nullptr;
    if (!s || !s->capturing) return cudaErrorInvalidValue;
    s->capturing = false;
    *g = new hostsim_graph_();
    (*g)->ops.swap(s->ops);
    return cudaSuccess;
}
static inline cudaError_t cudaGraphInstantiate(cudaGraphExec_t* e, cudaGraph_t g, unsigned long long) {
    if (!g) return cudaErrorInvalidValue;
    *e = new hostsim_graph_exec_();
    (*e)->ops = g->ops;
    return cudaSuccess;
}
static inline cudaError_t cudaGraphDestroy(cudaGraph_t g) { delete g; return cudaSuccess; }
static inline cudaError_t cudaGraphExecDestroy(cudaGraphExec_t e) { delete e; return cudaSuccess; }
static inline cudaError_t cudaGraphLaunch(cudaGraphExec_t e, cudaStream_t) {
    if (!e) return cudaErrorInvalidValue;
    for (auto& op : e->ops) op();
    return cudaSuccess;
}
// Run-time compiled kernels: a real cubin (NVRTC output) cannot be loaded -- the library then runs its interpreter
// kernels; the "cubin" of tests/hostsim/fake_nvrtc.cpp is a host shared object and is dlopen'ed.
static inline cudaError_t cudaLibraryLoadData(cudaLibrary_t* lib, const void* image, void*, void*, unsigned, void*, void*, unsigned) {
    const unsigned char* e = (const unsigned char*)image;
    if (memcmp(e, "\x7f" "ELF", 4) != 0 || e[4] != 2 || e[16] != 3) return cudaErrorNotSupported;      // not a 64-bit ET_DYN image
    if (*(const uint16_t*)(e + 18) != 62) return cudaErrorNotSupported;                                 // not x86-64: a real cubin
    uint64_t shoff; uint16_t shentsize, shnum;
    memcpy(&shoff, e + 40, 8); memcpy(&shentsize, e + 58, 2); memcpy(&shnum, e + 60, 2);
    const size_t bytes = (size_t)shoff + (size_t)shentsize * shnum;                                      // section headers close the file
    char path[] = "/tmp/hostsim_lib_XXXXXX";
    const int fd = mkstemp(path);
    if (fd < 0) return cudaErrorInvalidValue;
    const bool ok = write(fd, image, bytes) == (ssize_t)bytes;
    close(fd);
    void* h = ok ? dlopen(path, RTLD_NOW | RTLD_LOCAL) : nullptr;
    unlink(path);
    if (!h) { fprintf(stderr, "hostsim: dlopen of a host-compiled kernel failed: %s\n", ok ? dlerror() : "short write"); return cudaErrorInvalidValue; }
    *lib = (cudaLibrary_t)h;
    return cudaSuccess;
}
static inline cudaError_t cudaLibraryUnload(cudaLibrary_t lib) { if (lib) dlclose((void*)lib); return cudaSuccess; }
static inline cudaError_t cudaLibraryGetKernel(cudaKernel_t* k, cudaLibrary_t lib, const char*) {
    void* f = dlsym((void*)lib, "hostsim_jit_launch");
    if (!f) return cudaErrorInvalidValue;
    *k = (cudaKernel_t)f;
    return cudaSuccess;
}
static inline cudaError_t cudaLaunchKernel(const void* k, dim3 grid, dim3 block, void** args, size_t smem, cudaStream_t s) {
    if (!k) return cudaErrorInvalidValue;
    ((void (*)(unsigned, unsigned, size_t, void**, cudaStream_t))k)(grid.x, block.x, smem, args, s);
    return cudaSuccess;
}
static inline cudaError_t cudaKernelSetAttributeForDevice(cudaKernel_t, int, int, int) { return cudaSuccess; }
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, int, int) { return cudaSuccess; }
// one process, one address space: an IPC handle is the pointer itself (several "ranks" of a test share a process)
static inline cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t* h, void* p) { memset(h, 0, sizeof(*h)); memcpy(h->reserved, &p, sizeof(p)); return cudaSuccess; }
static inline cudaError_t cudaIpcOpenMemHandle(void** p, cudaIpcMemHandle_t h, unsigned) { memcpy(p, h.reserved, sizeof(*p)); return cudaSuccess; }
static inline cudaError_t cudaIpcCloseMemHandle(void*) { return cudaSuccess; }
