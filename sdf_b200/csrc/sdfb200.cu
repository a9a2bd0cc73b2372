// sdfb200.cu -- kernels and C-ABI of libsdfb200.so (sm_100a).
//
// Hot path of fogleman/sdf's core.generate (sdf/core.py:84-150), B200-native:
//
//   skip_prepass_kernel   core._skip (core.py:28-43): float64, 1 + 8 samples per batch (16 lanes per batch)
//   tile_sign_kernel      (A, precision `exact`) lattice samples generated in registers from the float64
//                         axes, whole SDF tree evaluated in fast float32, sign / uncertainty bit masks by
//                         warp ballot, cube classification + compaction by the CTA that completes a tile
//   tile_eval_kernel      (A, the other precisions) same with a shared-memory float32 value tile (core.py:51-52)
//   refine_points_kernel  (B) float64 re-evaluation of every sample that can reach the output
//   extract_kernel        (C) Lewiner marching cubes over the candidate cubes, tables in shared memory,
//                         CTA prefix-sum compaction (core.py:16-18), world mapping (core.py:58-60)
//   batch_stats_kernel    skipped / empty / nonempty (core.py:135-141)
//   eval_stream_*         SDF3.__call__ on an explicit point block (d3.py:24-25): float3
//                         points streamed HBM -> shared memory with cp.async.bulk (TMA
//                         bulk copy, mbarrier completion), distances streamed back
//
// No tensor cores: this is elementwise math.  No CPU fallback anywhere.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <float.h>
#include <stdarg.h>
#include <stdlib.h>
#include <algorithm>
#include <dlfcn.h>
#include <sys/stat.h>
#include <unistd.h>
#include <math.h>
#include <memory>
#include <new>
#include <mutex>
#include <future>
#include <atomic>
#include <chrono>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/sdfb200.h"
#include "sdf_kernels.cuh"

// ---------------------------------------------------------------------------------------
// error plumbing

static thread_local std::string g_last_error;

static int fail(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

// the same, for the other translation units of the library (sdfb200_weld.cu)
int sdfb200_fail_(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

#define CUDA_TRY(expr)                                                                      \
    do {                                                                                    \
        cudaError_t err__ = (expr);                                                         \
        if (err__ != cudaSuccess)                                                           \
            return fail(SDFB200_E_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(err__), \
                        __FILE__, __LINE__);                                                \
    } while (0)

// ---------------------------------------------------------------------------------------
// statically compiled kernels: the interpreter front end, program image in the parameter bank

__global__ void __launch_bounds__(128)
skip_prepass_kernel(const __grid_constant__ ProgramImage img, const GridDesc g, int sparse, unsigned char* flags,
                    unsigned int* active_list, unsigned int* batch_tris, Counters* ctr) {
    skip_prepass_body(InterpEval{&img}, g, sparse, flags, active_list, batch_tris, ctr);
}

#define INTERP_TILE_THREADS 512

template <int TS, int PREC>
__global__ void __launch_bounds__(INTERP_TILE_THREADS, 1)
tile_eval_kernel(const __grid_constant__ ProgramImage img, const GridDesc g, const float tau,
                 const unsigned int* active_list, const unsigned int work_begin, const unsigned int work_end,
                 const TileLists L, Counters* ctr) {
    tile_eval_body<TS, PREC, INTERP_TILE_THREADS>(InterpEval{&img}, g, tau, active_list, work_begin, work_end, L, ctr);
}

template <int TS>
__global__ void __launch_bounds__(SIGN_THREADS, SIGN_BLOCKS_PER_SM)
tile_sign_kernel(const __grid_constant__ ProgramImage img, const GridDesc g, const float tau,
                 const unsigned int* active_list, const unsigned int work_begin, const unsigned int work_end,
                 Counters* ctr, unsigned char* scratch) {
    tile_sign_body<TS, SIGN_THREADS>(InterpEval{&img}, g, tau, active_list, work_begin, work_end, ctr, scratch);
}

// SHARE = the share axis the sign kernel of the call ran with (fixes which lattice axis the mask bits run along)
template <int TS, int SHARE>
__global__ void __launch_bounds__(CLASSIFY_THREADS)
tile_classify_kernel(const GridDesc g, const unsigned int* active_list, const unsigned int work_begin,
                     const unsigned int work_end, const TileLists L, Counters* ctr, const unsigned char* scratch) {
    tile_classify_body<TS, SHARE>(g, active_list, work_begin, work_end, L, ctr, scratch);
}

template <int TS>
__global__ void __launch_bounds__(128)
refine_points_kernel(const __grid_constant__ ProgramImage img, const GridDesc g, const TileLists L, Counters* ctr) {
    refine_points_body<TS>(InterpEval{&img}, g, L, ctr);
}

template <int TS>
__global__ void __launch_bounds__(EXTRACT_THREADS, EXTRACT_BLOCKS_PER_SM)
extract_kernel(const GridDesc g, const TileLists L, Counters* ctr, unsigned int* batch_tris, float* tri_out,
               const unsigned long long tri_capacity, unsigned long long* ext_cursor, const McTablesSmem* tables,
               const unsigned char* stats_flags, RoundProgress* progress, unsigned long long* done_signal) {
    extract_body<TS>(g, L, ctr, batch_tris, tri_out, tri_capacity, ext_cursor, tables, stats_flags, progress, done_signal);
}

__global__ void shared_wait_ready_kernel(const unsigned long long* ready, unsigned long long* my_seq) { shared_wait_ready_body(ready, my_seq); }
__global__ void shared_wait_done_kernel(const unsigned long long* done, unsigned long long world) { shared_wait_done_body(done, world); }
__global__ void shared_release_kernel(unsigned long long* cursor, unsigned long long* done, unsigned long long* ready) {
    shared_release_body(cursor, done, ready);
}

// Candidate cubes of ONE dense user volume (sdfb200_mc_volume): the volume sits in tile slot 0 in its own C-order
// linearisation; a cube with mixed `v > 0` corners goes to the candidate list the extraction kernel walks.
__global__ void volume_candidates_kernel(const TileLists L, Counters* ctr, int px, int py, int pz) {
    const int n = (px - 1) * (py - 1) * (pz - 1);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int cz = i % (pz - 1), cy = (i / (pz - 1)) % (py - 1), cx = i / ((pz - 1) * (py - 1));
    const float* p = L.vol + (size_t)((cx * py + cy) * pz + cz);
    const int pyz = py * pz;
    const float f[8] = {p[0], p[1], p[pz + 1], p[pz], p[pyz], p[pyz + 1], p[pyz + pz + 1], p[pyz + pz]};
    int pos = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) pos += f[k] > 0.0f ? 1 : 0;
    if (pos == 0 || pos == 8) return;
    const unsigned int at = atomicAdd(&ctr->n_cubes, 1u);
    if (at < L.cube_cap) L.cubes[at] = make_uint2(0u, (unsigned int)((cx << 10) | (cy << 5) | cz));
}

// float32 copies of the lattice axes for the fast classification pass (same rounding as a host-side cast)
__global__ void axes_to_float_kernel(const double* a, float* f, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) f[i] = (float)a[i];
}

__global__ void batch_stats_kernel(const GridDesc g, const unsigned char* flags, const unsigned int* batch_tris,
                                   Counters* ctr) {
    batch_stats_body(g, flags, batch_tris, ctr);
}

__global__ void __launch_bounds__(STL_THREADS)
stl_records_kernel(const float* tris, long long n, unsigned char* out) { stl_records_body(tris, n, out); }

template <bool FAST>
__global__ void __launch_bounds__(STREAM_THREADS)
eval_stream_f32_kernel(const __grid_constant__ ProgramImage img, const float* xyz, long long n, float* out) {
    eval_stream_f32_body<FAST>(InterpEval{&img}, xyz, n, out);
}

__global__ void __launch_bounds__(256)
eval_stream_f64_kernel(const __grid_constant__ ProgramImage img, const double* xyz, long long n, double* out) {
    eval_stream_f64_body(InterpEval{&img}, xyz, n, out);
}

// ---------------------------------------------------------------------------------------
// host side


static constexpr int kMaxPipe = 16;      // most rounds a pipelined (host output) call is cut into

struct DeviceCtx {
    int device = -1;
    int sms = 0;
    // workspace, grown on demand
    double* axes = nullptr; size_t axes_cap = 0; std::vector<double> axes_host;
    unsigned char* flags = nullptr; unsigned int* active = nullptr; unsigned int* batch_tris = nullptr; size_t nb_cap = 0;
    Counters* ctr = nullptr;
    Counters* ctr_host = nullptr;       // pinned
    void* stage = nullptr; size_t stage_cap = 0;   // staging for host in/out buffers
    // tile pipeline work lists (TileLists)
    float* vol = nullptr; size_t slot_cap = 0; TileMeta* tile_meta = nullptr; size_t meta_cap = 0;
    unsigned char* scratch = nullptr; size_t scratch_cap = 0;       // tile_sign_kernel: per-tile mask blocks
    McTablesSmem* mc_tables = nullptr;                              // packed decision tables (one bulk copy per extraction CTA)
    RoundProgress* progress = nullptr;                              // page-locked + device mapped: one slot per pipelined round
    RoundProgress* progress_dev = nullptr;                          //   (the same memory as the device sees it)
    // the whole device-resident step as one CUDA graph, re-used while every launch parameter stays the same
    cudaStream_t cap_stream = nullptr;
    struct StepGraph { std::vector<unsigned char> key; cudaGraphExec_t exec = nullptr; unsigned long long used = 0; };
    std::vector<StepGraph> step_graphs;                             // a few recent ones (the multi-GPU cursors ping-pong: two keys alternate)
    unsigned long long step_clock = 0;
    uint2* cubes = nullptr; size_t cube_cap = 0;
    uint2* pts = nullptr; size_t pt_cap = 0;
    cudaEvent_t ev[9] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    // pipelined host output: the triangles of round k travel over PCIe while round k+1 computes
    cudaStream_t copy_stream = nullptr;
};

// Rounds a call is cut into when the triangles go to HOST memory, so that the triangles of round k cross PCIe
// while round k+1 computes.  SDFB200_PIPE fixes the number (1 = one round + one copy at the end; read on every
// call so tests can switch it); by default it follows the expected output size -- every round costs four more
// latency-bound kernels, which only pays while its copy (36 B per triangle at ~50 GB/s) is longer than that:
// about one round per 12 MB.  Measured on B200 (profiles/r2_e2e_variants.jsonl): gearlike 512^3, 32 MB: 1 round
// 0.85 ms, 3 rounds 0.77; blobby 768^3, 77 MB: 2.01 -> 1.73 (4); knurling 1024^3, 212 MB: 5.75 -> 4.43 (8).
static int pipe_rounds(int64_t expected_triangles) {
    int v;
    if (const char* e = getenv("SDFB200_PIPE")) v = atoi(e);
    else v = (int)((expected_triangles * 36 + (6ll << 20)) / (12ll << 20));
    const int lo = getenv("SDFB200_PIPE") ? 1 : 2;
    const int hi = getenv("SDFB200_PIPE") ? kMaxPipe : 8;
    return v < lo ? lo : (v > hi ? hi : v);
}

enum JitKernel { JK_SKIP = 0, JK_TILE0 = 1, JK_STREAM_F32 = 5, JK_STREAM_F32_FAST = 6, JK_STREAM_F64 = 7, JK_REFINE = 8, JK_COUNT = 9 };

// Outcome of compiling one kernel of one program: pure CPU work (NVRTC or the disk cache), produced on a
// background thread, loaded into the CUDA context by the thread that owns the program lock.
struct JitResult {
    bool ok = false;
    std::vector<char> cubin;
    int share = -1, n = 0;          // the tile kernels' sample grouping
    std::string log;                // NVRTC failure, if any
};

struct JitEntry {
    bool started = false;           // a compile has been launched (fut is valid until harvested)
    bool tried = false;             // harvested: kern is final (nullptr = unavailable)
    std::future<JitResult> fut;
    cudaLibrary_t lib = nullptr;
    cudaKernel_t kern = nullptr;
    bool smem_set[64] = {false};
};

struct sdfb200_prog {
    ProgramImage img;                   // interpreter image (valid only when the program fits it)
    bool img_valid = false;
    std::vector<uint32_t> code;         // the whole program (the specialised kernels are generated from it)
    std::vector<double> consts;
    int n_consts = 0;
    std::mutex lock;
    std::vector<DeviceCtx> ctxs;
    JitEntry jit[JK_COUNT];
    int plan_share[JK_COUNT] = {}, plan_n[JK_COUNT] = {};      // grouping the last compile of each kernel chose
    std::string jit_log;        // last NVRTC failure, if any
};

// ---------------------------------------------------------------------------------------
// run-time specialisation: the lowered tree as straight-line CUDA, compiled with NVRTC for sm_100a
//
// Mode (env SDFB200_JIT): "1"/unset = use a specialised kernel when NVRTC is available and fall back
// to the interpreter kernels otherwise; "0" = interpreter only.  NVRTC is loaded with dlopen so that a
// box without it still runs (on the interpreter).

struct NvrtcApi {
    void* h = nullptr;
    int (*CreateProgram)(void**, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
    int (*CompileProgram)(void*, int, const char* const*) = nullptr;
    int (*GetCUBINSize)(void*, size_t*) = nullptr;
    int (*GetCUBIN)(void*, char*) = nullptr;
    int (*GetProgramLogSize)(void*, size_t*) = nullptr;
    int (*GetProgramLog)(void*, char*) = nullptr;
    int (*DestroyProgram)(void**) = nullptr;
    bool ok = false;
};

static NvrtcApi& nvrtc_api() {
    static NvrtcApi api = [] {
        NvrtcApi a;
        const char* names[] = {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12"};
        for (const char* n : names) { a.h = dlopen(n, RTLD_NOW | RTLD_LOCAL); if (a.h) break; }
        if (!a.h) return a;
        a.CreateProgram = (decltype(a.CreateProgram))dlsym(a.h, "nvrtcCreateProgram");
        a.CompileProgram = (decltype(a.CompileProgram))dlsym(a.h, "nvrtcCompileProgram");
        a.GetCUBINSize = (decltype(a.GetCUBINSize))dlsym(a.h, "nvrtcGetCUBINSize");
        a.GetCUBIN = (decltype(a.GetCUBIN))dlsym(a.h, "nvrtcGetCUBIN");
        a.GetProgramLogSize = (decltype(a.GetProgramLogSize))dlsym(a.h, "nvrtcGetProgramLogSize");
        a.GetProgramLog = (decltype(a.GetProgramLog))dlsym(a.h, "nvrtcGetProgramLog");
        a.DestroyProgram = (decltype(a.DestroyProgram))dlsym(a.h, "nvrtcDestroyProgram");
        a.ok = a.CreateProgram && a.CompileProgram && a.GetCUBINSize && a.GetCUBIN && a.GetProgramLogSize &&
               a.GetProgramLog && a.DestroyProgram;
        return a;
    }();
    return api;
}

// 0 = interpreter only, 1 = specialise when the job is big enough to repay the compile (default),
// 2 = always specialise.  Read on every call so tests can switch it.
static int jit_mode() {
    const char* e = getenv("SDFB200_JIT");
    return e ? atoi(e) : 1;
}

static std::string csrc_dir() {
    Dl_info info;
    if (dladdr((const void*)&csrc_dir, &info) && info.dli_fname) {
        std::string p = info.dli_fname;
        const size_t k = p.rfind('/');
        return (k == std::string::npos ? std::string(".") : p.substr(0, k)) + "/csrc";
    }
    return "csrc";
}

// Which lattice axis the tile kernel's per-thread sample group should run along (tile_eval_body, "axis
// sharing").  A dependency mask over the lattice axes (bit 0 = x, 1 = y, 2 = z) is pushed through the
// program for every machine register; an operation whose inputs do not depend on axis A is computed
// once per group when the group runs along A.  Costs are rough instruction counts of the float32
// throughput path -- the choice only affects speed, never results.  Returns the best axis (z when no
// axis saves anything: z-columns store into the tile without bank conflicts) and, through
// `saved_fraction`, the estimated share of the work that does not depend on it.
static int share_axis(const sdfb200_prog* p, double* saved_fraction = nullptr, double* saved_cost = nullptr) {
    struct Masks {
        unsigned x = 1, y = 2, z = 4, d = 0, s0 = 0;
        unsigned r[SDF_MAX_PSLOTS][3] = {};
        unsigned vs[SDF_MAX_VSTACK + 1] = {};
    } m;
    double inv[3] = {0, 0, 0}, total = 0;
    auto piece = [&](double cost, unsigned mask) {
        total += cost;
        for (int ax = 0; ax < 3; ++ax) if (!((mask >> ax) & 1u)) inv[ax] += cost;
    };
    auto push = [&](int a) { if (a > 0) m.vs[a - 1] = m.s0; };
    auto pop = [&](int a) { if (a > 0) m.s0 = m.vs[a - 1]; };
    const double kZero = 3.7252902984619141e-09;        // 2^-28: coefficients the FAST path drops (sdf_ops.cuh: dot3z)
    for (size_t i = 0; i < p->code.size(); ++i) {
        const uint32_t w = p->code[i];
        const int op = (int)(w & 0xff), a = (int)((w >> 8) & 0xff);
        const double* c = p->consts.data() + (w >> 16);
        const unsigned xyz = m.x | m.y | m.z, xy = m.x | m.y;
        auto nz = [&](double v, double scale) { return fabs(v) > kZero * scale; };
        switch (op) {
            case OP_SPHERE: piece(8, xyz); m.d = xyz; break;
            case OP_PLANE: {
                const double sc = std::max(fabs(c[0]), std::max(fabs(c[1]), fabs(c[2])));
                const unsigned mk = (nz(c[0], sc) ? m.x : 0u) | (nz(c[1], sc) ? m.y : 0u) | (nz(c[2], sc) ? m.z : 0u);
                piece(3, mk); m.d = mk; break;
            }
            case OP_BOX: case OP_ROUNDED_BOX:
                piece(2, m.x); piece(2, m.y); piece(2, m.z); piece(12, xyz); m.d = xyz; break;
            case OP_WIREFRAME_BOX: piece(5, m.x); piece(5, m.y); piece(5, m.z); piece(40, xyz); m.d = xyz; break;
            case OP_TORUS: piece(5, xy); piece(5, xyz); m.d = xyz; break;
            case OP_CAPSULE: piece(22, xyz); m.d = xyz; break;
            case OP_CYLINDER: piece(5, xy); m.d = xy; break;
            case OP_CAPPED_CYLINDER: piece(35, xyz); m.d = xyz; break;
            case OP_ROUNDED_CYLINDER: piece(5, xy); piece(2, m.z); piece(12, xyz); m.d = xyz; break;
            case OP_CAPPED_CONE: piece(45, xyz); m.d = xyz; break;
            case OP_ROUNDED_CONE: piece(5, xy); piece(20, xyz); m.d = xyz; break;
            case OP_ELLIPSOID: piece(5, m.x); piece(5, m.y); piece(5, m.z); piece(14, xyz); m.d = xyz; break;
            case OP_PYRAMID: piece(3, xy); piece(40, xyz); m.d = xyz; break;
            case OP_TETRAHEDRON: piece(4, xy); piece(5, xyz); m.d = xyz; break;
            case OP_OCTAHEDRON: piece(1, m.x); piece(1, m.y); piece(1, m.z); piece(3, xyz); m.d = xyz; break;
            case OP_DODECAHEDRON: case OP_ICOSAHEDRON:
                piece(3, m.x); piece(3, m.y); piece(3, m.z); piece(14, xyz); m.d = xyz; break;
            case OP_TRANSLATE: piece(1, m.x); piece(1, m.y); piece(1, m.z); break;
            case OP_SCALE_P: piece(2, m.x); piece(2, m.y); piece(2, m.z); break;
            case OP_ROTATE: {
                double sc = 0;
                for (int k = 0; k < 9; ++k) sc = std::max(sc, fabs(c[k]));
                unsigned out[3];
                for (int col = 0; col < 3; ++col) {        // p @ M: out[col] = sum_row p[row] * M[row][col]
                    const unsigned in[3] = {m.x, m.y, m.z};
                    unsigned mk = 0; int terms = 0;
                    for (int row = 0; row < 3; ++row) if (nz(c[3 * row + col], sc)) { mk |= in[row]; ++terms; }
                    piece(terms, mk); out[col] = mk;
                }
                m.x = out[0]; m.y = out[1]; m.z = out[2]; break;
            }
            case OP_TWIST: piece(10, m.z); piece(4, xyz); m.x = m.y = xyz; break;
            case OP_BEND: piece(10, m.x); piece(4, xy); m.x = m.y = xy; break;
            case OP_ELONGATE: push(a); piece(3, m.x); piece(3, m.y); piece(3, m.z); piece(3, xyz); m.s0 = xyz; break;
            case OP_CIRCULAR_PRE:
                piece(55, xy);
                m.r[a][0] = xy; m.r[a][1] = xy; m.r[a][2] = m.z;
                m.x = m.y = xy; break;
            case OP_REPEAT_CELL: piece(5, m.x); piece(5, m.y); piece(5, m.z); break;
            case OP_MULC: case OP_NEGATE: case OP_ADDC: piece(1, m.d); break;
            case OP_SHELL: piece(2, m.d); break;
            case OP_ADD_POP: case OP_UNION: case OP_DIFFERENCE: case OP_INTERSECTION:
                piece(1, m.d | m.s0); m.d |= m.s0; pop(a); break;
            case OP_BLEND: piece(3, m.d | m.s0); m.d |= m.s0; pop(a); break;
            case OP_UNION_K: case OP_DIFFERENCE_K: case OP_INTERSECTION_K:
                piece(12, m.d | m.s0); m.d |= m.s0; pop(a); break;
            case OP_CIRCLE: piece(6, xy); m.d = xy; break;
            case OP_LINE2: piece(2, xy); m.d = xy; break;
            case OP_RECTANGLE: piece(2, m.x); piece(2, m.y); piece(8, xy); m.d = xy; break;
            case OP_ROUNDED_RECTANGLE: piece(16, xy); m.d = xy; break;
            case OP_EQUILATERAL_TRIANGLE: case OP_HEXAGON: piece(20, xy); m.d = xy; break;
            case OP_ROUNDED_X: piece(10, xy); m.d = xy; break;
            case OP_POLYGON: piece(14.0 * a + 6, xy); m.d = xy; break;
            case OP_VESICA: piece(14, xy); m.d = xy; break;
            case OP_REVOLVE: piece(5, xy); m.x = xy; m.y = m.z; break;
            case OP_EXTRUDE_PRE: push(a); piece(2, m.z); m.s0 = m.z; break;
            case OP_EXTRUDE_POST: piece(10, m.d | m.s0); m.d |= m.s0; pop(a); break;
            case OP_SLICE_PRE: m.z = 0; break;
            case OP_SLICE_SELECT: piece(2, m.d | m.s0); m.d |= m.s0; pop(a); break;
            case OP_TPARAM_LINEAR: piece(8, xyz); m.d = xyz; break;
            case OP_TPARAM_RADIAL: piece(8, xy); m.d = xy; break;
            case OP_TPARAM_Z: piece(4, m.z); m.d = m.z; break;
            case OP_EASE: if (a) piece(10, m.d); break;
            case OP_DISPLACE: piece(3, xyz | m.d); m.x |= m.d; m.y |= m.d; m.z |= m.d; break;
            case OP_LERP_POP2: case OP_MIX_POP2:
                piece(3, m.d | m.s0 | m.vs[a]); m.d |= m.s0 | m.vs[a]; pop(a); break;
            case OP_WRAP_PRE: piece(35, xy); m.d = xy; m.y = xy; break;
            case OP_WRAP_POST: piece(1, m.d); m.x = m.d; break;
            case OP_PUSH: push(a); m.s0 = m.d; break;
            case OP_SAVEP: m.r[a][0] = m.x; m.r[a][1] = m.y; m.r[a][2] = m.z; break;
            case OP_LOADP: m.x = m.r[a][0]; m.y = m.r[a][1]; m.z = m.r[a][2]; break;
            default: piece(10, 7u); break;
        }
    }
    int best = 2;                                       // ties go to z: conflict-free tile stores, the usual extrusion axis
    if (inv[1] > inv[best]) best = 1;
    if (inv[0] > inv[best]) best = 0;
    const double frac = total > 0 ? inv[best] / total : 0.0;
    if (saved_fraction) *saved_fraction = frac;
    if (saved_cost) *saved_cost = inv[best];
    return best;
}

// Column groups are used when they save at least this much estimated evaluation work per sample (in
// share_axis()'s instruction units); below it the flat 4-sample mapping is faster.  Measured on B200
// (profiles/r1_ab_grouping.jsonl): a column tile costs ~4 us more than a flat one outside the
// evaluation, a saved instruction per sample is worth ~0.2 us per tile -> gearlike (75 saved, tile
// kernel -27 %) and knurling (56, -4 %) use columns; csg (15: +6 % with columns), blobby, sphere stay flat.
static constexpr double kColumnMinSavedCost = 30.0;

static int env_int(const char* name, int dflt) {
    const char* e = getenv(name);
    return (e && *e) ? atoi(e) : dflt;
}

// The generated evaluator: one SdfOps call per program word, constants as literals.
// `share` / `kn`: the tile kernels' sample grouping (tile_eval_body); ignored by the other kernels.
static std::string jit_source(const sdfb200_prog* p, int kernel, int share, int kn) {
    std::ostringstream o;
    // experiments: SDFB200_JIT_DEFINES="NAME[=V] NAME2 ..." become #defines of the translation unit
    // (part of the source, hence of the cache key); unset in normal operation
    if (const char* defs = getenv("SDFB200_JIT_DEFINES")) {
        std::istringstream in(defs);
        std::string tok;
        while (in >> tok) {
            const size_t eq = tok.find('=');
            o << "#define " << (eq == std::string::npos ? tok : tok.substr(0, eq) + " " + tok.substr(eq + 1)) << "\n";
        }
    }
    o << "#define SDF_JIT_KERNEL " << kernel << "\n#include \"sdf_kernels.cuh\"\n";
    if (kernel >= 1 && kernel <= 4) {
        double frac = 0;
        const int ax = share_axis(p, &frac);
        char note[96];
        snprintf(note, sizeof(note), "// share_axis: %d (%.3f of the estimated work does not depend on it)\n", ax, frac);
        o << note;
    }
    o << "struct SpecEval {\n  static constexpr int kN = " << kn << ";\n  static constexpr int kShare = " << share
      << ";\n  template <typename T, int N, bool FAST>\n"
         "  __device__ __forceinline__ void run(T (&x)[N], T (&y)[N], T (&z)[N], T (&d)[N]) const {\n"
         "    typedef SdfOps<T, N, FAST> O;\n    typename O::State st(x, y, z, d);\n";
    o << "    const T C[" << (p->n_consts > 0 ? p->n_consts : 1) << "] = {";
    for (int i = 0; i < p->n_consts; ++i) {
        const double v = p->consts[i];
        char buf[80];
        if (isfinite(v)) snprintf(buf, sizeof(buf), "(T)%a", v);
        else { long long bits; memcpy(&bits, &v, 8); snprintf(buf, sizeof(buf), "(T)__longlong_as_double(0x%llxLL)", bits); }
        o << (i ? ", " : "") << buf;
    }
    if (p->n_consts == 0) o << "(T)0";
    o << "};\n";
    for (size_t i = 0; i < p->code.size(); ++i) {
        const uint32_t w = p->code[i];
        o << "    O::" << SDF_OP_NAME[w & 0xff] << "(st, C + " << (w >> 16) << ", " << ((w >> 8) & 0xff) << ");\n";
    }
    o << "  }\n};\n#include \"sdf_jit_entry.cuh\"\n";
    return o.str();
}

// Disk cache of the specialised kernels: <dir>/<128-bit key>.cubin, the key covering the generated
// source, every header it includes, the compile options and the NVRTC version.  Directory: env
// SDFB200_JIT_CACHE (empty string = no cache), default `.jitcache` next to libsdfb200.so; a directory
// that cannot be written is simply not used.  A script that meshes the same tree twice pays the
// NVRTC compile (seconds) once.
static uint64_t fnv1a(const void* data, size_t n, uint64_t h) {
    const unsigned char* b = (const unsigned char*)data;
    for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 0x100000001b3ull; }
    return h;
}

static const char* const kJitOptions[] = {"--gpu-architecture=sm_100a", "--std=c++17", "-lineinfo", "--extra-device-vectorization",
                                          "--ptxas-options=-v"};

static void jit_env_hash(uint64_t h[2]) {
    static uint64_t cached[2] = {0, 0};
    static std::once_flag once;
    std::call_once(once, [] {
        uint64_t a = 0xcbf29ce484222325ull, b = 0x84222325cbf29ce4ull;
        const char* files[] = {"sdf_kernels.cuh", "sdf_ops.cuh", "sdf_opcodes.h", "sdf_jit_entry.cuh", "mc_tables.inc"};
        for (const char* f : files) {
            const std::string path = csrc_dir() + "/" + f;
            if (FILE* fp = fopen(path.c_str(), "rb")) {
                char buf[65536];
                size_t n;
                while ((n = fread(buf, 1, sizeof(buf), fp)) > 0) { a = fnv1a(buf, n, a); b = fnv1a(buf, n, b); }
                fclose(fp);
            }
        }
        for (const char* o : kJitOptions) { a = fnv1a(o, strlen(o), a); b = fnv1a(o, strlen(o), b); }
        int ver[2] = {0, 0};
        if (nvrtc_api().h)
            if (auto fn = (int (*)(int*, int*))dlsym(nvrtc_api().h, "nvrtcVersion")) fn(&ver[0], &ver[1]);
        a = fnv1a(ver, sizeof(ver), a); b = fnv1a(ver, sizeof(ver), b);
        cached[0] = a; cached[1] = b;
    });
    h[0] = cached[0]; h[1] = cached[1];
}

static std::string jit_cache_path(const std::string& src) {
    std::string dir;
    if (const char* e = getenv("SDFB200_JIT_CACHE")) { if (!*e) return std::string(); dir = e; }
    else { dir = csrc_dir(); dir = dir.substr(0, dir.size() - 5) + "/.jitcache"; }      // strip "/csrc"
    uint64_t h[2];
    jit_env_hash(h);
    h[0] = fnv1a(src.data(), src.size(), h[0]);
    h[1] = fnv1a(src.data(), src.size(), h[1]);
    char name[64];
    snprintf(name, sizeof(name), "/%016llx%016llx.cubin", (unsigned long long)h[0], (unsigned long long)h[1]);
    return dir + name;
}

static bool jit_cache_load(const std::string& path, std::vector<char>* out) {
    if (path.empty()) return false;
    FILE* fp = fopen(path.c_str(), "rb");
    if (!fp) return false;
    fseek(fp, 0, SEEK_END);
    const long sz = ftell(fp);
    fseek(fp, 0, SEEK_SET);
    bool ok = sz > 64;
    if (ok) { out->resize((size_t)sz); ok = fread(out->data(), 1, (size_t)sz, fp) == (size_t)sz; }
    fclose(fp);
    ok = ok && memcmp(out->data(), "\x7f" "ELF", 4) == 0;          // a cubin is an ELF image
    if (!ok) out->clear();
    return ok;
}

static void jit_cache_store(const std::string& path, const std::vector<char>& cubin) {
    if (path.empty()) return;
    const size_t k = path.rfind('/');
    if (k != std::string::npos) mkdir(path.substr(0, k).c_str(), 0777);          // fails when it exists: fine
    const std::string tmp = path + "." + std::to_string((long long)getpid()) + ".tmp";
    FILE* fp = fopen(tmp.c_str(), "wb");
    if (!fp) return;
    const bool ok = fwrite(cubin.data(), 1, cubin.size(), fp) == cubin.size();
    fclose(fp);
    if (!ok || rename(tmp.c_str(), path.c_str()) != 0) remove(tmp.c_str());
}

// NVRTC: source -> cubin for sm_100a (or the cached cubin of the same source).  Needs no GPU; thread safe.
// *spill = bytes of register spill stores ptxas reported for the kernel.
static bool jit_compile_source(int kernel, const std::string& src, std::vector<char>* cubin_out, long* spill, std::string* err) {
    const std::string cpath = jit_cache_path(src);
    *spill = 0;
    if (jit_cache_load(cpath, cubin_out)) {
        if (FILE* f = fopen((cpath + ".spill").c_str(), "r")) { if (fscanf(f, "%ld", spill) != 1) *spill = 0; fclose(f); }
        return true;
    }
    NvrtcApi& rt = nvrtc_api();
    if (!rt.ok) { *err = "NVRTC (libnvrtc.so.12) not available"; return false; }
    void* prog = nullptr;
    if (rt.CreateProgram(&prog, src.c_str(), "sdfb200_jit.cu", 0, nullptr, nullptr) != 0) { *err = "nvrtcCreateProgram failed"; return false; }
    const std::string inc = "-I" + csrc_dir();
    const char* opts[] = {kJitOptions[0], kJitOptions[1], kJitOptions[2], kJitOptions[3], kJitOptions[4], inc.c_str()};
    const int rc = rt.CompileProgram(prog, 6, opts);
    size_t n = 0;
    rt.GetProgramLogSize(prog, &n);
    std::string log(n, '\0');
    if (n) rt.GetProgramLog(prog, &log[0]);
    if (rc != 0) {
        *err = "NVRTC compile failed (kernel " + std::to_string(kernel) + "): " + log;
        if (getenv("SDFB200_JIT_VERBOSE")) fprintf(stderr, "%s\n", err->c_str());
        rt.DestroyProgram(&prog);
        return false;
    }
    // "ptxas info : ... N bytes spill stores, M bytes spill loads" (one line per entry point; this unit has one)
    for (size_t at = log.find("bytes spill stores"); at != std::string::npos; at = log.find("bytes spill stores", at + 1)) {
        size_t b = at;
        while (b > 0 && log[b - 1] == ' ') --b;
        size_t a = b;
        while (a > 0 && isdigit((unsigned char)log[a - 1])) --a;
        if (a < b) *spill = std::max(*spill, atol(log.substr(a, b - a).c_str()));
    }
    if (getenv("SDFB200_JIT_VERBOSE")) fprintf(stderr, "[sdfb200 jit] kernel %d: %s", kernel, log.c_str());
    size_t sz = 0;
    rt.GetCUBINSize(prog, &sz);
    cubin_out->resize(sz);
    rt.GetCUBIN(prog, cubin_out->data());
    rt.DestroyProgram(&prog);
    jit_cache_store(cpath, *cubin_out);
    if (*spill > 0 && !cpath.empty())
        if (FILE* f = fopen((cpath + ".spill").c_str(), "w")) { fprintf(f, "%ld\n", *spill); fclose(f); }
    return true;
}

// Source generation + compilation of one kernel -- a pure function of the program (code, constants) and the
// environment, safe to run on a background thread.  The tile kernels (1..4) choose their sample grouping here:
// thread-runs of 11 samples along the axis share_axis() names (three per 33-sample line), backing off to 6
// and 3 (4 and the flat 4-sample mapping for the value-tile kernels) when ptxas had to spill registers; the
// candidates compile CONCURRENTLY (a cold first call waits for the slowest one, not for their sum).
// SDFB200_JIT_SHARE (-1 = flat, 0..2) / SDFB200_JIT_N force a grouping (experiments, A/B timing).
static JitResult jit_compile(const sdfb200_prog* p, int kernel) {
    JitResult res;
    const bool tile = kernel >= JK_TILE0 && kernel < JK_TILE0 + 4;
    const bool sign = kernel == JK_TILE0 + SDFB200_PREC_EXACT;          // tile_sign_body: always thread-runs along one axis
    int cand[4][2];
    int n_cand = 0;
    const int env_share = env_int("SDFB200_JIT_SHARE", -2), env_n = env_int("SDFB200_JIT_N", 0);
    int share = -1;
    if (tile) {
        if (env_share >= -1 && env_share <= 2) share = env_share;
        else {
            double saved = 0;
            share = share_axis(p, nullptr, &saved);
            if (!sign && saved < kColumnMinSavedCost) share = -1;
        }
        if (sign && share < 0) share = 2;
    }
    if (env_n >= 1 && env_n <= 16) { cand[0][0] = share; cand[0][1] = env_n; n_cand = 1; }
    else if (sign) {
        // 33 samples per line: 3 runs of 11, 6 of 6 (3 idle slots), 11 of 3
        const int ns[3] = {11, 6, 3};
        for (int k = 0; k < 3; ++k) { cand[k][0] = share; cand[k][1] = ns[k]; }
        n_cand = 3;
    }
    else if (share < 0) { cand[0][0] = -1; cand[0][1] = 4; n_cand = 1; }
    else {
        const int ns[3] = {11, 6, 4};
        for (int k = 0; k < 3; ++k) { cand[k][0] = share; cand[k][1] = ns[k]; }
        cand[3][0] = -1; cand[3][1] = 4;
        n_cand = 4;
    }
    // a few spilled bytes cost less than a shorter run (SDFB200_JIT_SPILL_OK: bytes of spill stores tolerated)
    const long spill_ok = env_int("SDFB200_JIT_SPILL_OK", 0);
    struct One { std::string src; std::vector<char> cubin; long spill = 0; bool ok = false; std::string err; };
    std::vector<One> ones((size_t)n_cand);
    for (int k = 0; k < n_cand; ++k) ones[(size_t)k].src = jit_source(p, kernel, cand[k][0], cand[k][1]);
    auto build = [&](int k) { One& o = ones[(size_t)k]; o.ok = jit_compile_source(kernel, o.src, &o.cubin, &o.spill, &o.err); };
    // the preferred candidate first (normally it does not spill, or sits in the disk cache); only when it spills
    // are the shorter runs built, side by side
    build(0);
    int best = 0;
    if (ones[0].ok && ones[0].spill > spill_ok && n_cand > 1) {
        std::vector<std::future<void>> jobs;
        for (int k = 1; k < n_cand; ++k) jobs.push_back(std::async(std::launch::async, build, k));
        for (auto& j : jobs) j.get();
        for (int k = 1; k < n_cand; ++k) {
            if (!ones[(size_t)k].ok) continue;
            if (ones[(size_t)best].spill > spill_ok && ones[(size_t)k].spill < ones[(size_t)best].spill) best = k;
        }
    }
    One& o = ones[(size_t)best];
    if (!o.ok) { res.log = o.err; return res; }
    res.ok = true;
    res.cubin.swap(o.cubin);
    res.share = cand[best][0]; res.n = cand[best][1];
    if (const char* dump = getenv("SDFB200_JIT_DUMP")) {       // debugging aid: keep source + cubin
        const std::string base = std::string(dump) + "/jit_k" + std::to_string(kernel);
        if (FILE* f = fopen((base + ".cu").c_str(), "w")) { fputs(o.src.c_str(), f); fclose(f); }
        if (FILE* f = fopen((base + ".cubin").c_str(), "wb")) { fwrite(res.cubin.data(), 1, res.cubin.size(), f); fclose(f); }
    }
    return res;
}

// Launch the compile of `kernel` on a background thread (once).  Called for every kernel a generate call will
// need before the first of them is used, so that a cold call compiles them side by side.
static void jit_start(sdfb200_prog* p, int kernel) {
    JitEntry& e = p->jit[kernel];
    if (e.started || e.tried) return;
    e.started = true;
    const sdfb200_prog* cp = p;
    e.fut = std::async(std::launch::async, [cp, kernel] { return jit_compile(cp, kernel); });
}

// The specialised kernel, or nullptr when the interpreter kernels should run.
//   mode 0: never specialise.  mode 2: always, waiting for the compiler.  mode 1 (default): for jobs big enough
//   to repay it the compile runs in the BACKGROUND -- until the cubin is ready the call in hand (a one-shot
//   `f.save()` script's only call) runs on the interpreter kernels instead of waiting seconds for NVRTC; later
//   calls pick the specialised kernel up.  Programs too large for the interpreter image always wait.
static cudaKernel_t jit_get(sdfb200_prog* p, int kernel, bool big_enough = true) {
    const int mode = jit_mode();
    if (p->img_valid && (mode == 0 || (mode == 1 && !big_enough))) return nullptr;      // (no image: only this path exists)
    JitEntry& e = p->jit[kernel];
    if (e.tried) return e.kern;
    jit_start(p, kernel);
    const bool may_defer = mode == 1 && p->img_valid;
    if (may_defer && e.fut.wait_for(std::chrono::seconds(0)) != std::future_status::ready) return nullptr;
    JitResult r = e.fut.get();
    e.tried = true;
    p->plan_share[kernel] = r.share; p->plan_n[kernel] = r.n;
    if (!r.ok) { p->jit_log = r.log; return nullptr; }
    cudaLibrary_t lib = nullptr;
    if (cudaLibraryLoadData(&lib, r.cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0) != cudaSuccess) {
        p->jit_log = std::string("cudaLibraryLoadData failed: ") + cudaGetErrorString(cudaGetLastError());
        return nullptr;
    }
    cudaKernel_t k = nullptr;
    if (cudaLibraryGetKernel(&k, lib, "jit_kernel") != cudaSuccess) {
        p->jit_log = std::string("cudaLibraryGetKernel failed: ") + cudaGetErrorString(cudaGetLastError());
        cudaLibraryUnload(lib);
        return nullptr;
    }
    e.lib = lib;
    e.kern = k;
    return k;
}

#define REQUIRE_IMAGE(prog)                                                                                   \
    do {                                                                                                      \
        if (!(prog)->img_valid)                                                                               \
            return fail(SDFB200_E_INVALID, "program (%zu words) exceeds the interpreter image and its run-time " \
                        "specialisation is unavailable: %s", (prog)->code.size(), (prog)->jit_log.c_str());  \
    } while (0)

static cudaError_t jit_launch(cudaKernel_t k, dim3 grid, dim3 block, size_t smem, cudaStream_t stream, void** args) {
    return cudaLaunchKernel((const void*)k, grid, block, args, smem, stream);
}

static int get_ctx(sdfb200_prog* prog, int device, DeviceCtx** out) {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(SDFB200_E_NOGPU, "no CUDA device available (%s); libsdfb200 has no CPU fallback",
                    e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    if (device < 0 || device >= count) return fail(SDFB200_E_INVALID, "device %d out of range (0..%d)", device, count - 1);
    for (auto& c : prog->ctxs)
        if (c.device == device) { *out = &c; CUDA_TRY(cudaSetDevice(device)); return 0; }
    CUDA_TRY(cudaSetDevice(device));
    DeviceCtx c;
    c.device = device;
    CUDA_TRY(cudaDeviceGetAttribute(&c.sms, cudaDevAttrMultiProcessorCount, device));
    CUDA_TRY(cudaMalloc(&c.ctr, sizeof(Counters)));
    CUDA_TRY(cudaMallocHost(&c.ctr_host, sizeof(Counters)));
    for (int i = 0; i < 9; ++i) CUDA_TRY(cudaEventCreate(&c.ev[i]));
    CUDA_TRY(cudaStreamCreateWithFlags(&c.copy_stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&c.cap_stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaHostAlloc((void**)&c.progress, sizeof(RoundProgress) * kMaxPipe, cudaHostAllocMapped));
    memset(c.progress, 0, sizeof(RoundProgress) * kMaxPipe);
    CUDA_TRY(cudaHostGetDevicePointer((void**)&c.progress_dev, c.progress, 0));
    {   // the marching-cubes decision tables, packed exactly as the extraction kernel keeps them in shared memory
        std::unique_ptr<McTablesSmem> t(new McTablesSmem());
        memset(t.get(), 0, sizeof(McTablesSmem));
        memcpy(t->sub_base, MC_SUB_BASE, sizeof(t->sub_base));
        memcpy(t->tile_off, MC_TILE_OFF, sizeof(t->tile_off));
        memcpy(t->sub_true, MC_SUB_TRUE, sizeof(t->sub_true));
        memcpy(t->sub_false, MC_SUB_FALSE, sizeof(t->sub_false));
        memcpy(t->sub_interior, MC_SUB_INTERIOR, sizeof(t->sub_interior));
        memcpy(t->nface, MC_NFACE, sizeof(t->nface));
        memcpy(t->faces, MC_FACES, sizeof(t->faces));
        memcpy(t->tile_ntri, MC_TILE_NTRI, sizeof(t->tile_ntri));
        memcpy(t->tile_data, MC_TILE_DATA, sizeof(t->tile_data));
        CUDA_TRY(cudaMalloc(&c.mc_tables, sizeof(McTablesSmem)));
        CUDA_TRY(cudaMemcpy(c.mc_tables, t.get(), sizeof(McTablesSmem), cudaMemcpyHostToDevice));
    }
    prog->ctxs.reserve(16);
    prog->ctxs.push_back(c);
    *out = &prog->ctxs.back();
    return 0;
}

static bool is_device_ptr(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// 0 = pageable host, 1 = device, 2 = pinned (page-locked, device-mapped) host
static int mem_kind(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return 0; }
    if (a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged) return 1;
    if (a.type == cudaMemoryTypeHost) return 2;
    return 0;
}

static int ensure_stage(DeviceCtx* c, size_t bytes) {
    if (bytes <= c->stage_cap) return 0;
    if (c->stage) cudaFree(c->stage);
    c->stage = nullptr; c->stage_cap = 0;
    CUDA_TRY(cudaMalloc(&c->stage, bytes));
    c->stage_cap = bytes;
    return 0;
}

extern "C" int sdfb200_abi_version(void) { return SDFB200_ABI_VERSION; }

extern "C" const char* sdfb200_last_error(void) { return g_last_error.c_str(); }

extern "C" int sdfb200_device_count(int* count) {
    if (!count) return fail(SDFB200_E_INVALID, "count is NULL");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { cudaGetLastError(); n = 0; }
    *count = n;
    return 0;
}

extern "C" int sdfb200_host_alloc(size_t bytes, void** out) {
    if (!out) return fail(SDFB200_E_INVALID, "host_alloc: out is NULL");
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
        cudaGetLastError();
        return fail(SDFB200_E_NOGPU, "no CUDA device available; libsdfb200 has no CPU fallback");
    }
    CUDA_TRY(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocPortable | cudaHostAllocMapped));
    return 0;
}

extern "C" void sdfb200_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

// Page-lock memory the caller owns (e.g. a POSIX shared-memory segment several rank processes write their slabs
// into, each over its own PCIe link: sdf_b200/multi.py HostGather) so that device -> host copies into it run at
// DMA speed; sdfb200_copy_to_host is that copy (asynchronous on `stream`, then synchronised).
extern "C" int sdfb200_host_register(void* p, size_t bytes) {
    if (!p || !bytes) return fail(SDFB200_E_INVALID, "host_register: bad argument");
    CUDA_TRY(cudaHostRegister(p, bytes, cudaHostRegisterPortable));
    return 0;
}

extern "C" void sdfb200_host_unregister(void* p) {
    if (p) { cudaHostUnregister(p); cudaGetLastError(); }
}

extern "C" int sdfb200_copy_to_host(void* dst_host, const void* src_device, size_t bytes, int device, void* stream) {
    if (bytes == 0) return 0;
    if (!dst_host || !src_device) return fail(SDFB200_E_INVALID, "copy_to_host: NULL argument");
    CUDA_TRY(cudaSetDevice(device));
    CUDA_TRY(cudaMemcpyAsync(dst_host, src_device, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
    return 0;
}

extern "C" int sdfb200_program_create(const uint32_t* code, int n_code, const double* consts, int n_consts,
                                      sdfb200_prog** out) {
    if (!code || !out || n_code <= 0 || n_consts < 0 || (n_consts > 0 && !consts))
        return fail(SDFB200_E_INVALID, "program_create: NULL/empty argument");
    if (n_code > SDF_PROG_MAX_CODE || n_consts > SDF_PROG_MAX_CONSTS)
        return fail(SDFB200_E_INVALID, "program too large: %d words / %d constants (limits %d / %d)", n_code, n_consts,
                    SDF_PROG_MAX_CODE, SDF_PROG_MAX_CONSTS);
    // validate: opcodes known, constant windows inside the table, stack / slot operands in range
    for (int i = 0; i < n_code; ++i) {
        const uint32_t w = code[i];
        const int op = (int)(w & 0xff), a = (int)((w >> 8) & 0xff), ci = (int)(w >> 16);
        if (op <= 0 || op >= OP__COUNT) return fail(SDFB200_E_INVALID, "instruction %d: unknown opcode %d", i, op);
        const int nc = SDF_OP_NCONST[op] >= 0 ? SDF_OP_NCONST[op] : 2 * a;        // POLYGON: a points, 2a constants
        if (ci + nc > n_consts)
            return fail(SDFB200_E_INVALID, "instruction %d: constants [%d,%d) outside table of %d", i, ci, ci + nc, n_consts);
        const bool slot_op = (op == OP_SAVEP || op == OP_LOADP || op == OP_CIRCULAR_PRE);
        const bool free_op = (op == OP_POLYGON || op == OP_EASE);                 // operand is a count / an id
        if (slot_op && a >= SDF_MAX_PSLOTS) return fail(SDFB200_E_INVALID, "instruction %d: point slot %d too deep", i, a);
        if (!slot_op && !free_op && a > SDF_MAX_VSTACK) return fail(SDFB200_E_INVALID, "instruction %d: value stack %d too deep", i, a);
        if (op == OP_EASE && a > 33) return fail(SDFB200_E_INVALID, "instruction %d: unknown ease id %d", i, a);
    }
    sdfb200_prog* p = new (std::nothrow) sdfb200_prog();
    if (!p) return fail(SDFB200_E_INVALID, "out of host memory");
    memset(&p->img, 0, sizeof(ProgramImage));
    p->code.assign(code, code + n_code);
    p->consts.assign(consts, consts + n_consts);
    p->n_consts = n_consts;
    p->img_valid = (n_code <= SDF_MAX_CODE && n_consts <= SDF_MAX_CONSTS);
    if (p->img_valid) {
        p->img.n_code = (uint32_t)n_code;
        memcpy(p->img.code, code, sizeof(uint32_t) * n_code);
        for (int i = 0; i < n_consts; ++i) { p->img.cd[i] = consts[i]; p->img.cf[i] = (float)consts[i]; }
    }
    *out = p;
    return 0;
}

extern "C" void sdfb200_program_free(sdfb200_prog* prog) {
    if (!prog) return;
    for (auto& e : prog->jit) if (e.started && !e.tried && e.fut.valid()) e.fut.wait();      // background compiles read the program
    for (auto& c : prog->ctxs) {
        cudaSetDevice(c.device);
        cudaFree(c.axes); cudaFree(c.flags); cudaFree(c.active); cudaFree(c.batch_tris); cudaFree(c.ctr);
        cudaFreeHost(c.ctr_host); cudaFree(c.stage);
        cudaFree(c.vol); cudaFree(c.tile_meta); cudaFree(c.cubes); cudaFree(c.pts);
        cudaFree(c.scratch); cudaFree(c.mc_tables);
        if (c.progress) cudaFreeHost(c.progress);
        for (auto& sg : c.step_graphs) if (sg.exec) cudaGraphExecDestroy(sg.exec);
        if (c.cap_stream) cudaStreamDestroy(c.cap_stream);
        for (int i = 0; i < 9; ++i) if (c.ev[i]) cudaEventDestroy(c.ev[i]);
        if (c.copy_stream) cudaStreamDestroy(c.copy_stream);
    }
    for (auto& e : prog->jit) if (e.lib) cudaLibraryUnload(e.lib);
    delete prog;
}

static int eval_f32_impl(sdfb200_prog* prog, const float* xyz, int64_t n, float* d_out, int device, void* stream_, bool fast) {
    if (!prog || n < 0 || (n > 0 && (!xyz || !d_out))) return fail(SDFB200_E_INVALID, "eval_f32: bad argument");
    if (n == 0) return 0;
    std::lock_guard<std::mutex> guard(prog->lock);
    DeviceCtx* c; int rc = get_ctx(prog, device, &c); if (rc) return rc;
    cudaStream_t stream = (cudaStream_t)stream_;
    const bool in_dev = is_device_ptr(xyz), out_dev = is_device_ptr(d_out);
    const float* dx = xyz; float* dd = d_out;
    if (!in_dev || !out_dev) {
        const size_t in_b = in_dev ? 0 : (size_t)n * 12, out_b = out_dev ? 0 : (size_t)n * 4;
        const size_t in_pad = (in_b + 255) & ~(size_t)255;
        rc = ensure_stage(c, in_pad + out_b); if (rc) return rc;
        if (!in_dev) { CUDA_TRY(cudaMemcpyAsync(c->stage, xyz, in_b, cudaMemcpyHostToDevice, stream)); dx = (const float*)c->stage; }
        if (!out_dev) dd = (float*)((char*)c->stage + in_pad);
    }
    const long long tiles = (n + STREAM_TILE - 1) / STREAM_TILE;
    const int grid = (int)std::min<long long>(tiles, (long long)c->sms * 4);
    if (cudaKernel_t jk = jit_get(prog, fast ? JK_STREAM_F32_FAST : JK_STREAM_F32, n >= (1 << 20))) {
        long long nn = (long long)n;
        void* args[] = {(void*)&dx, (void*)&nn, (void*)&dd};
        CUDA_TRY(jit_launch(jk, dim3(grid), dim3(STREAM_THREADS), 0, stream, args));
    } else {
        REQUIRE_IMAGE(prog);
        if (fast) eval_stream_f32_kernel<true><<<grid, STREAM_THREADS, 0, stream>>>(prog->img, dx, (long long)n, dd);
        else eval_stream_f32_kernel<false><<<grid, STREAM_THREADS, 0, stream>>>(prog->img, dx, (long long)n, dd);
    }
    CUDA_TRY(cudaGetLastError());
    if (!out_dev) CUDA_TRY(cudaMemcpyAsync(d_out, dd, (size_t)n * 4, cudaMemcpyDeviceToHost, stream));
    if (!in_dev || !out_dev) CUDA_TRY(cudaStreamSynchronize(stream));
    return 0;
}

extern "C" int sdfb200_eval_f32(sdfb200_prog* prog, const float* xyz, int64_t n, float* d_out, int device, void* stream) {
    return eval_f32_impl(prog, xyz, n, d_out, device, stream, false);
}

extern "C" int sdfb200_eval_f32_fast(sdfb200_prog* prog, const float* xyz, int64_t n, float* d_out, int device, void* stream) {
    return eval_f32_impl(prog, xyz, n, d_out, device, stream, true);
}

extern "C" int sdfb200_eval_f64(sdfb200_prog* prog, const double* xyz, int64_t n, double* d_out, int device, void* stream_) {
    if (!prog || n < 0 || (n > 0 && (!xyz || !d_out))) return fail(SDFB200_E_INVALID, "eval_f64: bad argument");
    if (n == 0) return 0;
    std::lock_guard<std::mutex> guard(prog->lock);
    DeviceCtx* c; int rc = get_ctx(prog, device, &c); if (rc) return rc;
    cudaStream_t stream = (cudaStream_t)stream_;
    const bool in_dev = is_device_ptr(xyz), out_dev = is_device_ptr(d_out);
    const double* dx = xyz; double* dd = d_out;
    if (!in_dev || !out_dev) {
        const size_t in_b = in_dev ? 0 : (size_t)n * 24, out_b = out_dev ? 0 : (size_t)n * 8;
        const size_t in_pad = (in_b + 255) & ~(size_t)255;
        rc = ensure_stage(c, in_pad + out_b); if (rc) return rc;
        if (!in_dev) { CUDA_TRY(cudaMemcpyAsync(c->stage, xyz, in_b, cudaMemcpyHostToDevice, stream)); dx = (const double*)c->stage; }
        if (!out_dev) dd = (double*)((char*)c->stage + in_pad);
    }
    const int grid = (int)std::min<long long>((n + 255) / 256, (long long)c->sms * 8);
    if (cudaKernel_t jk = jit_get(prog, JK_STREAM_F64, n >= (1 << 20))) {     // small probes: not worth a compile
        long long nn = (long long)n;
        void* args[] = {(void*)&dx, (void*)&nn, (void*)&dd};
        CUDA_TRY(jit_launch(jk, dim3(grid), dim3(256), 0, stream, args));
    } else {
        REQUIRE_IMAGE(prog);
        eval_stream_f64_kernel<<<grid, 256, 0, stream>>>(prog->img, dx, (long long)n, dd);
    }
    CUDA_TRY(cudaGetLastError());
    if (!out_dev) CUDA_TRY(cudaMemcpyAsync(d_out, dd, (size_t)n * 8, cudaMemcpyDeviceToHost, stream));
    if (!in_dev || !out_dev) CUDA_TRY(cudaStreamSynchronize(stream));
    return 0;
}

static constexpr int kTS = 32;

// The batch range of a call: a Z-slab of whole batch layers [zb_begin, zb_end) (zb_end < 0: to the end) when
// ub_begin < 0, else the window [ub_begin, ub_end) of the Z-major batch order (multi-GPU: slabs cut inside a layer).
static int setup_grid(DeviceCtx* c, cudaStream_t stream, const double* xs, int64_t nx, const double* ys, int64_t ny,
                      const double* zs, int64_t nz, int batch_size, int zb_begin, int zb_end, int64_t ub_begin, int64_t ub_end,
                      GridDesc* g) {
    if (!xs || !ys || !zs || nx < 1 || ny < 1 || nz < 1 || batch_size < 1)
        return fail(SDFB200_E_INVALID, "generate: bad lattice (%lld x %lld x %lld, batch %d)", (long long)nx, (long long)ny,
                    (long long)nz, batch_size);
    if (nx > (1 << 20) || ny > (1 << 20) || nz > (1 << 20)) return fail(SDFB200_E_INVALID, "lattice axis too long");
    const size_t need = (size_t)(nx + ny + nz);
    if (need > c->axes_cap) {
        if (c->axes) cudaFree(c->axes);
        c->axes = nullptr; c->axes_cap = 0;
        CUDA_TRY(cudaMalloc(&c->axes, need * (sizeof(double) + sizeof(float))));      // float64 axes, then their float32 copies
        c->axes_cap = need;
        c->axes_host.clear();
    }
    float* const axes_f = reinterpret_cast<float*>(c->axes + c->axes_cap);
    // upload the axes unless they are the ones already resident (host copy kept for the comparison)
    const bool host_axes = !is_device_ptr(xs) && !is_device_ptr(ys) && !is_device_ptr(zs);
    bool same = host_axes && c->axes_host.size() == need;
    if (same) same = memcmp(c->axes_host.data(), xs, nx * 8) == 0 && memcmp(c->axes_host.data() + nx, ys, ny * 8) == 0 &&
                     memcmp(c->axes_host.data() + nx + ny, zs, nz * 8) == 0;
    if (!same) {
        if (host_axes) {
            c->axes_host.resize(need);
            memcpy(c->axes_host.data(), xs, nx * 8);
            memcpy(c->axes_host.data() + nx, ys, ny * 8);
            memcpy(c->axes_host.data() + nx + ny, zs, nz * 8);
            CUDA_TRY(cudaMemcpyAsync(c->axes, c->axes_host.data(), need * sizeof(double), cudaMemcpyHostToDevice, stream));
        } else {
            c->axes_host.clear();
            const cudaMemcpyKind kind = cudaMemcpyDefault;
            CUDA_TRY(cudaMemcpyAsync(c->axes, xs, nx * sizeof(double), kind, stream));
            CUDA_TRY(cudaMemcpyAsync(c->axes + nx, ys, ny * sizeof(double), kind, stream));
            CUDA_TRY(cudaMemcpyAsync(c->axes + nx + ny, zs, nz * sizeof(double), kind, stream));
        }
        axes_to_float_kernel<<<(unsigned)((need + 255) / 256), 256, 0, stream>>>(c->axes, axes_f, (int)need);
        CUDA_TRY(cudaGetLastError());
    }
    g->xs = c->axes; g->ys = c->axes + nx; g->zs = c->axes + nx + ny;
    g->xf = axes_f; g->yf = axes_f + nx; g->zf = axes_f + nx + ny;
    g->nx = (int)nx; g->ny = (int)ny; g->nz = (int)nz;
    g->bs = batch_size;
    g->nbx = (int)((nx + batch_size - 1) / batch_size);
    g->nby = (int)((ny + batch_size - 1) / batch_size);
    g->nbz = (int)((nz + batch_size - 1) / batch_size);
    if (zb_end < 0) zb_end = g->nbz;
    if (zb_begin < 0 || zb_begin > zb_end || zb_end > g->nbz)
        return fail(SDFB200_E_INVALID, "z batch range [%d,%d) outside [0,%d)", zb_begin, zb_end, g->nbz);
    g->sub = (batch_size + kTS - 1) / kTS;
    const size_t nb = (size_t)g->nbx * g->nby * g->nbz;
    if (nb >= (size_t)1 << 31) return fail(SDFB200_E_INVALID, "too many batches");
    const int64_t nxy = (int64_t)g->nbx * g->nby;
    if (ub_begin < 0) { ub_begin = (int64_t)zb_begin * nxy; ub_end = (int64_t)zb_end * nxy; }
    if (ub_begin > ub_end || ub_end > (int64_t)nb)
        return fail(SDFB200_E_INVALID, "batch window [%lld,%lld) outside [0,%lld)", (long long)ub_begin, (long long)ub_end, (long long)nb);
    g->ub0 = (int)ub_begin; g->ub1 = (int)ub_end;
    if (nb > c->nb_cap) {
        cudaFree(c->flags); cudaFree(c->active); cudaFree(c->batch_tris);
        c->flags = nullptr; c->active = nullptr; c->batch_tris = nullptr; c->nb_cap = 0;
        CUDA_TRY(cudaMalloc(&c->flags, nb));
        CUDA_TRY(cudaMalloc(&c->active, nb * sizeof(unsigned int)));
        CUDA_TRY(cudaMalloc(&c->batch_tris, nb * sizeof(unsigned int)));
        c->nb_cap = nb;
    }
    CUDA_TRY(cudaMemsetAsync(c->ctr, 0, sizeof(Counters), stream));     // flags / batch_tris: written by the pre-pass
    return 0;
}

template <typename T>
static int grow(T** p, size_t* cap, size_t need) {
    if (need <= *cap) return 0;
    if (*p) cudaFree(*p);
    *p = nullptr; *cap = 0;
    CUDA_TRY(cudaMalloc(p, need * sizeof(T)));
    *cap = need;
    return 0;
}

template <int PREC>
static cudaError_t launch_tile_eval(int grid, cudaStream_t stream, const ProgramImage& img, const GridDesc& g, float tau,
                                    const unsigned int* active, unsigned int wb, unsigned int we, const TileLists& L,
                                    Counters* ctr) {
    const size_t smem = sizeof(TileSmem<kTS, PREC == SDFB200_PREC_F64>);
    tile_eval_kernel<kTS, PREC><<<grid, INTERP_TILE_THREADS, smem, stream>>>(img, g, tau, active, wb, we, L, ctr);
    return cudaGetLastError();
}

// dynamic shared-memory opt-in of the value-tile kernels (once per device; never inside a stream capture)
static cudaError_t prepare_tile_eval(int precision) {
    cudaError_t e = cudaSuccess;
    if (precision == SDFB200_PREC_F32)
        e = cudaFuncSetAttribute(tile_eval_kernel<kTS, SDFB200_PREC_F32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TileSmem<kTS, false>));
    else if (precision == SDFB200_PREC_MIXED)
        e = cudaFuncSetAttribute(tile_eval_kernel<kTS, SDFB200_PREC_MIXED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TileSmem<kTS, false>));
    else if (precision == SDFB200_PREC_F64)
        e = cudaFuncSetAttribute(tile_eval_kernel<kTS, SDFB200_PREC_F64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TileSmem<kTS, true>));
    return e;
}

static const size_t kVolBudgetBytes = (size_t)24 << 30;      // tile volumes kept per round

// the shared (root GPU) words a step of the fused multi-GPU gather works with
struct SharedStep {
    unsigned long long* cursor = nullptr;      // output ranges are reserved here (system-scope atomics)
    unsigned long long* done = nullptr;        // += 1 when all of this rank's triangles are written
    const unsigned long long* ready = nullptr; // == this rank's sequence number when the buffer may be used
    unsigned long long* my_seq = nullptr;      // this rank's sequence number for the cursor (device memory of THIS GPU)
};

static bool graph_enabled() {
    const char* e = getenv("SDFB200_GRAPH");
    return e ? atoi(e) != 0 : true;
}

// One generate call.  Everything is queued on `stream` without waiting for the device; the host
// synchronises ONCE, at the end, when the output stays on the device (or is only counted), and only
// where it needs a number to go on when the triangles travel to host memory (the count of active
// batches to cut the work list into copy-overlapped rounds, the triangle count for the last copy).
// In the first case the whole step -- counters reset, pre-pass, A, B, C (+ batch statistics), counter
// read-back -- is ONE CUDA graph, captured the first time and re-launched for as long as every launch
// parameter stays the same (SDFB200_GRAPH=0 disables it): four API calls per step instead of twenty,
// and no launch gaps between the kernels.
static int generate_impl(sdfb200_prog* prog, const double* xs, int64_t nx, const double* ys, int64_t ny,
                         const double* zs, int64_t nz, int batch_size, int sparse, int precision, double tau_,
                         int zb_begin, int zb_end, float* tri_out, int64_t tri_capacity, int64_t* n_tri,
                         int64_t stats[6], double timings_ms[8], int device, void* stream_, unsigned char* stl_out,
                         const SharedStep* shared = nullptr, int64_t ub_begin = -1, int64_t ub_end = -1) {
    if (!prog || !n_tri) return fail(SDFB200_E_INVALID, "generate: NULL argument");
    unsigned long long* const ext_cursor = shared ? shared->cursor : nullptr;
    if (precision < 0 || precision > 3) return fail(SDFB200_E_INVALID, "generate: precision %d", precision);
    if (tri_capacity < 0) return fail(SDFB200_E_INVALID, "generate: negative capacity");
    std::lock_guard<std::mutex> guard(prog->lock);
    DeviceCtx* c; int rc = get_ctx(prog, device, &c); if (rc) return rc;
    cudaStream_t stream = (cudaStream_t)stream_;
    GridDesc g;
    CUDA_TRY(cudaEventRecord(c->ev[0], stream));
    rc = setup_grid(c, stream, xs, nx, ys, ny, zs, nz, batch_size, zb_begin, zb_end, ub_begin, ub_end, &g); if (rc) return rc;
    // output placement: device pointer -> written in place; pinned host -> either written in place by
    // the kernel over PCIe (zero copy, overlaps the transfer with the computation) or staged + one async
    // copy; pageable host -> staged + copy
    // STL mode: triangles stay in the device staging area, 50-byte records follow them there
    unsigned char* drec = nullptr;
    if (stl_out) {
        rc = ensure_stage(c, (size_t)tri_capacity * 36 + 256 + (size_t)tri_capacity * 50 + 256); if (rc) return rc;
        tri_out = (float*)c->stage;
        drec = (unsigned char*)c->stage + (((size_t)tri_capacity * 36 + 255) & ~(size_t)255);
    }
    const int out_kind = tri_out ? mem_kind(tri_out) : 1;
    static const bool zero_copy = [] { const char* e = getenv("SDFB200_ZEROCOPY"); return e ? atoi(e) != 0 : false; }();
    bool out_dev = (out_kind == 1) || (out_kind == 2 && zero_copy);
    float* dtri = tri_out;
    if (tri_out && out_kind == 2 && zero_copy) {
        void* dp = nullptr;
        CUDA_TRY(cudaHostGetDevicePointer(&dp, tri_out, 0));
        dtri = (float*)dp;
    } else if (tri_out && !out_dev) {
        rc = ensure_stage(c, (size_t)tri_capacity * 36 + 256); if (rc) return rc;
        dtri = (float*)c->stage;
    }
    const long long total_b = (long long)g.ub1 - g.ub0;
    const unsigned long long cap = tri_out ? (unsigned long long)tri_capacity : 0ull;
    const float tau = (float)tau_;
    const long long lattice = (long long)nx * ny * nz;
    const bool big = lattice >= (1 << 21);
    const bool sign_path = precision == SDFB200_PREC_EXACT;           // ballot masks, no value tile
    const bool refine = precision == SDFB200_PREC_MIXED || precision == SDFB200_PREC_EXACT;
    Counters& h = *c->ctr_host;
    const bool tail = stl_out != nullptr || (tri_out != nullptr && !out_dev);   // something still follows the counters
    bool ran_pipeline = false, have_counters = false;
    unsigned long long copied = 0;      // triangles already on their way to tri_out (pipelined host output)
    if (total_b > 0) {
        const int pb = (int)((total_b * PREPASS_LANES + 127) / 128);
        if (jit_mode() >= 2 || (jit_mode() == 1 && big) || !prog->img_valid) {      // the kernels of this call compile side by side
            jit_start(prog, JK_SKIP);
            jit_start(prog, JK_TILE0 + precision);
            if (refine) jit_start(prog, JK_REFINE);
        }
        // Work-list sizing.  When one tile volume per batch of the range fits the budget, the kernels
        // read the number of active batches from the pre-pass's counter themselves and the host never
        // synchronises in the middle; otherwise the host reads it and runs rounds of `max_round` tiles.
        const unsigned long long max_round = std::max<unsigned long long>(1ull, kVolBudgetBytes / ((size_t)VOL_STRIDE * 4));
        const unsigned long long upper = (unsigned long long)total_b * g.sub * g.sub * g.sub;
        // Host output: cut the work list into a few rounds so that the triangles of round k cross PCIe
        // while round k+1 is computed (the copy is the longer leg: 36 B per triangle at PCIe speed).
        int pipe = (tri_out && !out_dev && !stl_out && !ext_cursor && big) ? pipe_rounds(tri_capacity) : 1;
        if (upper < 2ull * c->sms) pipe = 1;                      // less than two tiles per SM: nothing to overlap
        // Rounds are cut ON THE DEVICE (round k of `pipe` equal parts of the active tiles, work_range()): the host
        // queues the whole call without knowing how many batches survive the pre-pass.
        const unsigned long long round_tiles = (upper + pipe - 1) / pipe;
        const bool sync_free = round_tiles <= max_round;
        if (!sync_free) pipe = 1;                                 // (a lattice beyond the volume budget: host-cut rounds below)

        // every kernel of the step resolved up front (nothing may compile or load inside a stream capture)
        cudaKernel_t k_skip = jit_get(prog, JK_SKIP, big);
        cudaKernel_t k_tile = jit_get(prog, JK_TILE0 + precision, big);
        cudaKernel_t k_refine = refine ? jit_get(prog, JK_REFINE, big) : nullptr;
        if ((!k_skip || !k_tile || (refine && !k_refine)) && !prog->img_valid) REQUIRE_IMAGE(prog);
        if (k_tile && !sign_path) {
            const bool f64 = precision == SDFB200_PREC_F64;
            const size_t smem = f64 ? sizeof(TileSmem<kTS, true>) : sizeof(TileSmem<kTS, false>);
            JitEntry& je = prog->jit[JK_TILE0 + precision];
            if (!je.smem_set[device]) {
                CUDA_TRY(cudaKernelSetAttributeForDevice(k_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem, device));
                je.smem_set[device] = true;
            }
        }

        if (!k_tile && !sign_path) CUDA_TRY(prepare_tile_eval(precision));

        auto rec = [&](cudaEvent_t ev, cudaStream_t s, bool captured) -> cudaError_t {
            return captured ? cudaEventRecordWithFlags(ev, s, cudaEventRecordExternal) : cudaEventRecord(ev, s);
        };
        auto launch_prepass = [&](cudaStream_t s) -> int {
            if (k_skip) {
                unsigned char* fl = c->flags; unsigned int* act = c->active; unsigned int* bt = c->batch_tris; Counters* ctrp = c->ctr;
                int sp = sparse;
                void* args[] = {(void*)&g, (void*)&sp, (void*)&fl, (void*)&act, (void*)&bt, (void*)&ctrp};
                CUDA_TRY(jit_launch(k_skip, dim3(pb), dim3(128), 0, s, args));
            } else {
                skip_prepass_kernel<<<pb, 128, 0, s>>>(prog->img, g, sparse, c->flags, c->active, c->batch_tris, c->ctr);
                CUDA_TRY(cudaGetLastError());
            }
            return 0;
        };
        // one round = (A) evaluate + classify, (B) float64 refinement, (C) extraction [+ batch statistics when last]
        auto launch_round = [&](cudaStream_t s, bool captured, const TileLists& L, unsigned int wbu, unsigned int we,
                                unsigned long long tiles, bool first, bool last, RoundProgress* progress) -> int {
            cudaError_t e = cudaSuccess;
            if (first) CUDA_TRY(rec(c->ev[4], s, captured));
            if (sign_path) {
                // (A'1) autonomous warps pull 1/13-tile batches from a cursor: fill the machine, surplus warps leave at once
                const int grid = (int)std::min<unsigned long long>(tiles * 16ull, (unsigned long long)c->sms * SIGN_BLOCKS_PER_SM);
                int share = 2;                                        // the interpreter kernel: runs along z
                if (k_tile) {
                    const unsigned int* act = c->active; Counters* ctrp = c->ctr; float tauv = tau;
                    unsigned char* scr = c->scratch;
                    void* args[] = {(void*)&g, (void*)&tauv, (void*)&act, (void*)&wbu, (void*)&we, (void*)&ctrp, (void*)&scr};
                    e = jit_launch(k_tile, dim3(grid), dim3(SIGN_THREADS), 0, s, args);
                    share = prog->plan_share[JK_TILE0 + precision] < 0 ? 2 : prog->plan_share[JK_TILE0 + precision];
                } else {
                    tile_sign_kernel<kTS><<<grid, SIGN_THREADS, 0, s>>>(prog->img, g, tau, c->active, wbu, we, c->ctr, c->scratch);
                    e = cudaGetLastError();
                }
                if (e == cudaSuccess) {
                    // (A'2) every tile of the round classified side by side, one CTA each
                    if (first) CUDA_TRY(rec(c->ev[8], s, captured));
                    const int cgrid = (int)std::min<unsigned long long>(tiles, (unsigned long long)c->sms * 8);
                    if (share == 0) tile_classify_kernel<kTS, 0><<<cgrid, CLASSIFY_THREADS, 0, s>>>(g, c->active, wbu, we, L, c->ctr, c->scratch);
                    else if (share == 1) tile_classify_kernel<kTS, 1><<<cgrid, CLASSIFY_THREADS, 0, s>>>(g, c->active, wbu, we, L, c->ctr, c->scratch);
                    else tile_classify_kernel<kTS, 2><<<cgrid, CLASSIFY_THREADS, 0, s>>>(g, c->active, wbu, we, L, c->ctr, c->scratch);
                    e = cudaGetLastError();
                }
            } else {
                const int grid = (int)std::min<unsigned long long>(tiles, (unsigned long long)c->sms);
                if (k_tile) {
                    const bool f64 = precision == SDFB200_PREC_F64;
                    const size_t smem = f64 ? sizeof(TileSmem<kTS, true>) : sizeof(TileSmem<kTS, false>);
                    const unsigned int* act = c->active; Counters* ctrp = c->ctr; float tauv = tau;
                    void* args[] = {(void*)&g, (void*)&tauv, (void*)&act, (void*)&wbu, (void*)&we, (void*)&L, (void*)&ctrp};
                    e = jit_launch(k_tile, dim3(grid), dim3(f64 ? 512 : 1024), smem, s, args);
                } else if (precision == SDFB200_PREC_F32)
                    e = launch_tile_eval<SDFB200_PREC_F32>(grid, s, prog->img, g, tau, c->active, wbu, we, L, c->ctr);
                else if (precision == SDFB200_PREC_MIXED)
                    e = launch_tile_eval<SDFB200_PREC_MIXED>(grid, s, prog->img, g, tau, c->active, wbu, we, L, c->ctr);
                else
                    e = launch_tile_eval<SDFB200_PREC_F64>(grid, s, prog->img, g, tau, c->active, wbu, we, L, c->ctr);
            }
            if (e != cudaSuccess) return fail(SDFB200_E_CUDA, "tile kernel launch failed: %s", cudaGetErrorString(e));
            if (first) CUDA_TRY(rec(c->ev[5], s, captured));
            if (refine) {
                const int rgrid = c->sms * env_int("SDFB200_REFINE_BLOCKS", 16);      // (A/B hook; 128-thread CTAs per SM)
                if (k_refine) {
                    Counters* ctrp = c->ctr;
                    void* args[] = {(void*)&g, (void*)&L, (void*)&ctrp};
                    CUDA_TRY(jit_launch(k_refine, dim3(rgrid), dim3(128), 0, s, args));
                } else {
                    refine_points_kernel<kTS><<<rgrid, 128, 0, s>>>(prog->img, g, L, c->ctr);
                    CUDA_TRY(cudaGetLastError());
                }
            }
            if (first) CUDA_TRY(rec(c->ev[6], s, captured));
            if (shared && first) {            // the shared buffer of this step's parity is free (the root has consumed its last use)
                shared_wait_ready_kernel<<<1, 32, 0, s>>>(shared->ready, shared->my_seq);
                CUDA_TRY(cudaGetLastError());
            }
            extract_kernel<kTS><<<c->sms * EXTRACT_BLOCKS_PER_SM, EXTRACT_THREADS, 0, s>>>(
                g, L, c->ctr, c->batch_tris, dtri, cap, ext_cursor, c->mc_tables, last ? c->flags : nullptr, progress,
                (shared && last) ? shared->done : nullptr);
            CUDA_TRY(cudaGetLastError());
            if (first) CUDA_TRY(rec(c->ev[7], s, captured));
            return 0;
        };

        unsigned long long n_work = upper;
        if (!sync_free) {
            rc = launch_prepass(stream); if (rc) return rc;
            CUDA_TRY(cudaEventRecord(c->ev[1], stream));
            CUDA_TRY(cudaMemcpyAsync(c->ctr_host, c->ctr, sizeof(Counters), cudaMemcpyDeviceToHost, stream));
            CUDA_TRY(cudaStreamSynchronize(stream));
            n_work = (unsigned long long)h.n_active * g.sub * g.sub * g.sub;
        }
        const unsigned int round = (unsigned int)std::min<unsigned long long>(sync_free ? round_tiles : n_work, max_round);
        if (n_work > 0) {
            ran_pipeline = true;
            rc = grow(&c->vol, &c->slot_cap, (size_t)round * VOL_STRIDE); if (rc) return rc;
            rc = grow(&c->tile_meta, &c->meta_cap, (size_t)round); if (rc) return rc;
            if (sign_path) {
                rc = grow(&c->scratch, &c->scratch_cap, (size_t)round * SIGN_STRIDE); if (rc) return rc;
            }
            const size_t est_tiles = sync_free ? std::max<size_t>(256, round / 4) : round;     // most batches are skipped / empty
            size_t want_cubes = std::max<size_t>(c->cube_cap, std::min<size_t>((size_t)round * 32768, std::max<size_t>((size_t)1 << 20, est_tiles * 3072)));
            size_t want_pts = std::max<size_t>(c->pt_cap, precision == SDFB200_PREC_EXACT ? want_cubes * 3 : want_cubes);
            for (int attempt = 0; attempt < 3; ++attempt) {
                rc = grow(&c->cubes, &c->cube_cap, want_cubes); if (rc) return rc;
                rc = grow(&c->pts, &c->pt_cap, want_pts); if (rc) return rc;
                TileLists L;
                memset(&L, 0, sizeof(L));
                L.vol = c->vol; L.tile_meta = c->tile_meta; L.cubes = c->cubes; L.pts = c->pts;
                L.slot_cap = round; L.cube_cap = (unsigned int)std::min<size_t>(c->cube_cap, 0xffffffffu);
                L.pt_cap = (unsigned int)std::min<size_t>(c->pt_cap, 0xffffffffu);
                if (attempt > 0 && shared)
                    return fail(SDFB200_E_CUDA, "work lists overflowed in a shared (multi-GPU) call: size them with a count-only pass of the same window first");
                if (attempt > 0) {      // redo from a clean slate (the pre-pass results stay valid)
                    CUDA_TRY(cudaStreamSynchronize(c->copy_stream));
                    copied = 0;
                    const size_t nb = (size_t)g.nbx * g.nby * g.nbz;
                    CUDA_TRY(cudaMemsetAsync(c->batch_tris, 0, nb * sizeof(unsigned int), stream));
                    CUDA_TRY(cudaMemsetAsync(&c->ctr->tri_cursor, 0, 4 * sizeof(unsigned long long), stream));
                    CUDA_TRY(cudaMemsetAsync(&c->ctr->work_cursor, 0, sizeof(unsigned int), stream));
                    CUDA_TRY(cudaMemsetAsync(&c->ctr->empty, 0, 2 * sizeof(unsigned int), stream));
                    CUDA_TRY(cudaMemsetAsync(&c->ctr->n_slots, 0, 6 * sizeof(unsigned int), stream));      // n_slots .. extract_done
                }
                if (sync_free) {
                    // ---- the whole step in one go: pre-pass, A, B, C + statistics, counters to the host
                    auto enqueue = [&](cudaStream_t s, bool captured) -> int {
                        int r = 0;
                        if (attempt == 0) { r = launch_prepass(s); if (r) return r; }       // (a retry keeps the pre-pass results)
                        CUDA_TRY(rec(c->ev[1], s, captured));
                        for (int k = 0; k < pipe; ++k) {
                            if (k > 0) {         // (the first round starts from the zeroed counters)
                                CUDA_TRY(cudaMemsetAsync(&c->ctr->work_cursor, 0, sizeof(unsigned int), s));
                                CUDA_TRY(cudaMemsetAsync(&c->ctr->n_slots, 0, 4 * sizeof(unsigned int), s));     // n_slots .. extract_done
                            }
                            const unsigned int wbu = pipe > 1 ? (unsigned int)k : 0u;
                            const unsigned int we = pipe > 1 ? (0xfffe0000u | (unsigned int)pipe) : 0xffffffffu;
                            r = launch_round(s, captured, L, wbu, we, round_tiles, k == 0, k == pipe - 1,
                                             pipe > 1 ? c->progress_dev + k : nullptr);
                            if (r) return r;
                        }
                        CUDA_TRY(rec(c->ev[2], s, captured));
                        CUDA_TRY(cudaMemcpyAsync(c->ctr_host, c->ctr, sizeof(Counters), cudaMemcpyDeviceToHost, s));
                        if (!tail) CUDA_TRY(rec(c->ev[3], s, captured));
                        return 0;
                    };
                    for (int k = 0; k < pipe; ++k) ((volatile RoundProgress*)c->progress)[k].done = 0u;
                    bool launched = false;
                    if (graph_enabled() && attempt == 0) {
                        // key: every value the launches above depend on
                        struct { GridDesc g; TileLists L; float tau; int sparse, precision, tail; float* dtri; unsigned long long cap, n_work;
                                 unsigned long long* ext; const void *sh1, *sh2, *sh3; void *k0, *k1, *k2; void *flags, *active, *bt, *ctr, *ctrh, *scr; } key;
                        memset(&key, 0, sizeof(key));
                        key.g = g; key.L = L; key.tau = tau; key.sparse = sparse; key.precision = precision; key.tail = tail ? 1 : 0;
                        key.sh1 = shared ? (const void*)shared->done : nullptr; key.sh2 = shared ? (const void*)shared->ready : nullptr;
                        key.sh3 = shared ? (const void*)shared->my_seq : nullptr;
                        key.dtri = dtri; key.cap = cap; key.n_work = n_work * 64ull + (unsigned long long)pipe; key.ext = ext_cursor;
                        key.k0 = (void*)k_skip; key.k1 = (void*)k_tile; key.k2 = (void*)k_refine;
                        key.flags = c->flags; key.active = c->active; key.bt = c->batch_tris; key.ctr = c->ctr; key.ctrh = c->ctr_host;
                        key.scr = c->scratch;
                        cudaGraphExec_t exec = nullptr;
                        for (auto& sg : c->step_graphs)
                            if (sg.exec && sg.key.size() == sizeof(key) && memcmp(sg.key.data(), &key, sizeof(key)) == 0) {
                                exec = sg.exec; sg.used = ++c->step_clock; break;
                            }
                        if (!exec) {
                            cudaGraph_t graph = nullptr;
                            if (cudaStreamBeginCapture(c->cap_stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
                                const int r = enqueue(c->cap_stream, true);
                                const cudaError_t ee = cudaStreamEndCapture(c->cap_stream, &graph);
                                if (r != 0 || ee != cudaSuccess || !graph || cudaGraphInstantiate(&exec, graph, 0) != cudaSuccess) exec = nullptr;
                                if (graph) cudaGraphDestroy(graph);
                            }
                            cudaGetLastError();          // a failed capture falls back to plain launches below
                            if (exec) {
                                constexpr size_t kMaxGraphs = 4;
                                if (c->step_graphs.size() >= kMaxGraphs) {          // evict the least recently used
                                    size_t lru = 0;
                                    for (size_t i = 1; i < c->step_graphs.size(); ++i)
                                        if (c->step_graphs[i].used < c->step_graphs[lru].used) lru = i;
                                    cudaGraphExecDestroy(c->step_graphs[lru].exec);
                                    c->step_graphs.erase(c->step_graphs.begin() + (long)lru);
                                }
                                DeviceCtx::StepGraph sg;
                                sg.key.assign((const unsigned char*)&key, (const unsigned char*)&key + sizeof(key));
                                sg.exec = exec; sg.used = ++c->step_clock;
                                c->step_graphs.push_back(std::move(sg));
                            }
                        }
                        if (exec) {
                            CUDA_TRY(cudaGraphLaunch(exec, stream));
                            launched = true;
                        }
                    }
                    if (!launched) { rc = enqueue(stream, false); if (rc) return rc; }
                    // pipelined host output: every round is queued; follow the rounds' progress words (mapped host
                    // memory, written by each round's extraction kernel) and send a round's triangles over PCIe as
                    // soon as it is complete, while the next rounds compute
                    for (int k = 0; k < pipe && pipe > 1; ++k) {
                        volatile RoundProgress* pr = c->progress + k;
                        while (pr->done == 0u) {
                            if (cudaStreamQuery(stream) != cudaErrorNotReady) break;      // finished or failed: the counters decide
                        }
                        std::atomic_thread_fence(std::memory_order_acquire);
                        if (pr->done == 0u) break;
                        if (pr->max_cubes > L.cube_cap || pr->max_pts > L.pt_cap) break;       // lists overflowed: retried below
                        const unsigned long long produced = pr->tri_cursor;
                        const unsigned long long upto = std::min<unsigned long long>(produced, cap);
                        if (upto > copied) {
                            CUDA_TRY(cudaMemcpyAsync(tri_out + copied * 9, dtri + copied * 9, (size_t)(upto - copied) * 36,
                                                     cudaMemcpyDeviceToHost, c->copy_stream));
                            copied = upto;
                        }
                    }
                } else {
                    for (unsigned long long wb = 0; wb < n_work; wb += round) {
                        const unsigned int we = (unsigned int)std::min<unsigned long long>(n_work, wb + round);
                        if (wb > 0) {        // (the first round starts from the zeroed counters)
                            CUDA_TRY(cudaMemsetAsync(&c->ctr->work_cursor, 0, sizeof(unsigned int), stream));
                            CUDA_TRY(cudaMemsetAsync(&c->ctr->n_slots, 0, 4 * sizeof(unsigned int), stream));     // n_slots .. extract_done
                        }
                        const unsigned long long tiles = std::min<unsigned long long>(n_work, wb + round) - wb;
                        rc = launch_round(stream, false, L, (unsigned int)wb, we, tiles, wb == 0, wb + round >= n_work, nullptr); if (rc) return rc;
                    }
                    CUDA_TRY(cudaEventRecord(c->ev[2], stream));
                    CUDA_TRY(cudaMemcpyAsync(c->ctr_host, c->ctr, sizeof(Counters), cudaMemcpyDeviceToHost, stream));
                    if (!tail) CUDA_TRY(cudaEventRecord(c->ev[3], stream));     // nothing follows on the device
                }
                CUDA_TRY(cudaStreamSynchronize(stream));
                have_counters = true;
                if (const char* dd = getenv("SDFB200_DEBUG_DUMP")) {
                    auto dump = [&](const char* name, const void* dptr, size_t bytes) {
                        std::vector<char> hb(bytes);
                        cudaMemcpy(hb.data(), dptr, bytes, cudaMemcpyDeviceToHost);
                        std::string fn = std::string(dd) + "/" + name;
                        if (FILE* f = fopen(fn.c_str(), "wb")) { fwrite(hb.data(), 1, bytes, f); fclose(f); }
                    };
                    dump("vol.bin", c->vol, (size_t)h.n_slots * VOL_STRIDE * 4);
                    dump("meta.bin", c->tile_meta, (size_t)h.n_slots * sizeof(TileMeta));
                    dump("cubes.bin", c->cubes, (size_t)std::min<unsigned>(h.n_cubes, L.cube_cap) * 8);
                    dump("pts.bin", c->pts, (size_t)std::min<unsigned>(h.n_pts, L.pt_cap) * 8);
                }
                if (getenv("SDFB200_DEBUG"))
                    fprintf(stderr, "[sdfb200] n_active=%u n_slots=%u n_cubes=%u n_pts=%u max=%u/%u caps=%u/%u tri=%llu f64=%llu graph=%d\n",
                            h.n_active, h.n_slots, h.n_cubes, h.n_pts, h.max_cubes, h.max_pts, L.cube_cap, L.pt_cap,
                            h.tri_cursor, h.pts_f64, (int)c->step_graphs.size());
                if (h.max_cubes <= L.cube_cap && h.max_pts <= L.pt_cap) break;
                want_cubes = std::max<size_t>(want_cubes, (size_t)h.max_cubes + (h.max_cubes >> 3));
                want_pts = std::max<size_t>(want_pts, (size_t)h.max_pts + (h.max_pts >> 3));
                if (attempt == 2) return fail(SDFB200_E_CUDA, "work lists kept overflowing (%u cubes, %u samples)", h.max_cubes, h.max_pts);
            }
        }
    } else {
        CUDA_TRY(cudaEventRecord(c->ev[1], stream));
    }
    if (!have_counters) {           // nothing ran behind the pre-pass (no batch at all, or no active batch)
        CUDA_TRY(cudaEventRecord(c->ev[2], stream));
        if (total_b > 0) {
            batch_stats_kernel<<<(int)((total_b + 127) / 128), 128, 0, stream>>>(g, c->flags, c->batch_tris, c->ctr);
            CUDA_TRY(cudaGetLastError());
        }
        CUDA_TRY(cudaMemcpyAsync(c->ctr_host, c->ctr, sizeof(Counters), cudaMemcpyDeviceToHost, stream));
        if (!tail) CUDA_TRY(cudaEventRecord(c->ev[3], stream));
        CUDA_TRY(cudaStreamSynchronize(stream));
    }
    *n_tri = (int64_t)h.tri_cursor;
    if (stats) {
        stats[0] = h.skipped; stats[1] = h.empty; stats[2] = h.nonempty;
        stats[3] = (int64_t)h.pts_f32; stats[4] = (int64_t)h.pts_f64; stats[5] = (int64_t)h.active_cubes;
    }
    const bool overflow = tri_out && !ext_cursor && h.tri_cursor > (unsigned long long)tri_capacity;
    if (tail) {            // the triangles (or their STL records) still have to reach the host
        if (stl_out && !overflow && h.tri_cursor > 0) {
            const long long nt = (long long)h.tri_cursor;
            stl_records_kernel<<<(unsigned)((nt + STL_THREADS - 1) / STL_THREADS), STL_THREADS, 0, stream>>>(dtri, nt, drec);
            CUDA_TRY(cudaGetLastError());
            CUDA_TRY(cudaMemcpyAsync(stl_out, drec, (size_t)nt * 50, cudaMemcpyDefault, stream));
        } else if (!stl_out && !overflow && h.tri_cursor > copied) {
            CUDA_TRY(cudaMemcpyAsync(tri_out + copied * 9, dtri + copied * 9, (size_t)(h.tri_cursor - copied) * 36,
                                     cudaMemcpyDeviceToHost, stream));
        }
        if (copied > 0) CUDA_TRY(cudaStreamSynchronize(c->copy_stream));
        CUDA_TRY(cudaEventRecord(c->ev[3], stream));
        CUDA_TRY(cudaStreamSynchronize(stream));
    }
    if (timings_ms) {
        float a = 0, b = 0, d = 0, ta = 0, tb = 0, tc = 0, tk = 0;
        cudaEventElapsedTime(&a, c->ev[0], c->ev[1]);
        cudaEventElapsedTime(&b, c->ev[1], c->ev[2]);
        cudaEventElapsedTime(&d, c->ev[2], c->ev[3]);
        if (ran_pipeline) {
            if (sign_path) { cudaEventElapsedTime(&ta, c->ev[4], c->ev[8]); cudaEventElapsedTime(&tk, c->ev[8], c->ev[5]); }
            else cudaEventElapsedTime(&ta, c->ev[4], c->ev[5]);
            cudaEventElapsedTime(&tb, c->ev[5], c->ev[6]);
            cudaEventElapsedTime(&tc, c->ev[6], c->ev[7]);
        }
        timings_ms[0] = a; timings_ms[1] = b; timings_ms[2] = a + b; timings_ms[3] = d;
        timings_ms[4] = ta; timings_ms[5] = tb; timings_ms[6] = tc; timings_ms[7] = tk;
    }
    if (overflow)
        return fail(SDFB200_E_CAPACITY, "triangle buffer too small: %lld needed, %lld given", (long long)h.tri_cursor,
                    (long long)tri_capacity);
    return 0;
}

extern "C" int sdfb200_generate(sdfb200_prog* prog, const double* xs, int64_t nx, const double* ys, int64_t ny,
                                const double* zs, int64_t nz, int batch_size, int sparse, int precision, double tau,
                                int zb_begin, int zb_end, float* tri_out, int64_t tri_capacity, int64_t* n_tri,
                                int64_t stats[6], double timings_ms[8], int device, void* stream) {
    return generate_impl(prog, xs, nx, ys, ny, zs, nz, batch_size, sparse, precision, tau, zb_begin, zb_end, tri_out,
                         tri_capacity, n_tri, stats, timings_ms, device, stream, nullptr);
}

extern "C" int sdfb200_generate_stl(sdfb200_prog* prog, const double* xs, int64_t nx, const double* ys, int64_t ny,
                                    const double* zs, int64_t nz, int batch_size, int sparse, int precision, double tau,
                                    void* records_out, int64_t tri_capacity, int64_t* n_tri, int64_t stats[6],
                                    double timings_ms[8], int device, void* stream) {
    if (!records_out) return fail(SDFB200_E_INVALID, "generate_stl: records_out is NULL");
    return generate_impl(prog, xs, nx, ys, ny, zs, nz, batch_size, sparse, precision, tau, 0, -1, nullptr, tri_capacity,
                         n_tri, stats, timings_ms, device, stream, (unsigned char*)records_out);
}

// ---------------------------------------------------------------------------------------
// shared output: one triangle buffer + cursors in the ROOT GPU's memory, mapped into the other ranks'
// address spaces with CUDA IPC, written by their extraction kernels over NVLink (sdfb200_generate_shared)

struct sdfb200_shared {
    unsigned char* base = nullptr;      // [0,256): header words; then TWO triangle buffers of `capacity` (step parity)
    unsigned long long* my_seq = nullptr;      // this process' sequence numbers for the two parities (its own GPU's memory)
    int64_t capacity = 0;
    int device = -1;
    bool owner = false;
};

// header: cursor[2] @ 0, done[2] @ 16, ready[2] @ 32 (unsigned long long each)
static unsigned long long* shared_word(sdfb200_shared* s, int field, int which) { return (unsigned long long*)(s->base + 16 * field + 8 * which); }
static float* shared_data(sdfb200_shared* s, int which) { return (float*)(s->base + 256 + (size_t)which * (size_t)s->capacity * 36); }

static int shared_init_seq(sdfb200_shared* s) {
    CUDA_TRY(cudaMalloc(&s->my_seq, 2 * sizeof(unsigned long long)));
    const unsigned long long init[2] = {0ull, 1ull};
    CUDA_TRY(cudaMemcpy(s->my_seq, init, sizeof(init), cudaMemcpyHostToDevice));
    return 0;
}

extern "C" int sdfb200_shared_create(int64_t tri_capacity, int device, sdfb200_shared** out, unsigned char handle_out[64]) {
    if (!out || !handle_out || tri_capacity <= 0) return fail(SDFB200_E_INVALID, "shared_create: bad argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    CUDA_TRY(cudaSetDevice(device));
    std::unique_ptr<sdfb200_shared> s(new (std::nothrow) sdfb200_shared());
    if (!s) return fail(SDFB200_E_INVALID, "out of host memory");
    CUDA_TRY(cudaMalloc(&s->base, 256 + 2 * (size_t)tri_capacity * 36));
    cudaIpcMemHandle_t h;
    unsigned long long header[32] = {0};
    header[4] = 0ull; header[5] = 1ull;                  // ready[0] = 0, ready[1] = 1: steps 0 and 1 may start
    cudaError_t e = cudaMemcpy(s->base, header, 256, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, s->base);
    if (e != cudaSuccess) {
        cudaFree(s->base);
        return fail(SDFB200_E_CUDA, "shared_create: %s", cudaGetErrorString(e));
    }
    memcpy(handle_out, &h, 64);
    s->capacity = tri_capacity; s->device = device; s->owner = true;
    int rc = shared_init_seq(s.get()); if (rc) { cudaFree(s->base); return rc; }
    *out = s.release();
    return 0;
}

extern "C" int sdfb200_shared_open(const unsigned char handle[64], int64_t tri_capacity, int device, sdfb200_shared** out) {
    if (!out || !handle || tri_capacity <= 0) return fail(SDFB200_E_INVALID, "shared_open: bad argument");
    CUDA_TRY(cudaSetDevice(device));
    std::unique_ptr<sdfb200_shared> s(new (std::nothrow) sdfb200_shared());
    if (!s) return fail(SDFB200_E_INVALID, "out of host memory");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    void* p = nullptr;
    CUDA_TRY(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    s->base = (unsigned char*)p; s->capacity = tri_capacity; s->device = device; s->owner = false;
    int rc = shared_init_seq(s.get()); if (rc) { cudaIpcCloseMemHandle(p); return rc; }
    *out = s.release();
    return 0;
}

extern "C" void sdfb200_shared_free(sdfb200_shared* s) {
    if (!s) return;
    cudaSetDevice(s->device);
    cudaFree(s->my_seq);
    if (s->owner) cudaFree(s->base); else cudaIpcCloseMemHandle(s->base);
    delete s;
}

extern "C" void* sdfb200_shared_buffer(sdfb200_shared* s, int which) {
    return (s && which >= 0 && which <= 1) ? (void*)shared_data(s, which) : nullptr;
}

// owner only: wait (on the device) until all `world` ranks have delivered the triangles of the step that uses parity
// `which`, then read their number (synchronises `stream`)
extern "C" int sdfb200_shared_wait(sdfb200_shared* s, int which, int world, int64_t* n, void* stream) {
    if (!s || !s->owner || !n || which < 0 || which > 1 || world < 1) return fail(SDFB200_E_INVALID, "shared_wait: bad argument");
    CUDA_TRY(cudaSetDevice(s->device));
    shared_wait_done_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(shared_word(s, 1, which), (unsigned long long)world);
    CUDA_TRY(cudaGetLastError());
    unsigned long long v = 0;
    CUDA_TRY(cudaMemcpyAsync(&v, shared_word(s, 0, which), 8, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
    *n = (int64_t)v;
    return 0;
}

// owner only: the buffer of parity `which` has been consumed -- reset its cursor and hand it to the step after next
extern "C" int sdfb200_shared_release(sdfb200_shared* s, int which, void* stream) {
    if (!s || !s->owner || which < 0 || which > 1) return fail(SDFB200_E_INVALID, "shared_release: bad argument");
    CUDA_TRY(cudaSetDevice(s->device));
    shared_release_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(shared_word(s, 0, which), shared_word(s, 1, which), shared_word(s, 2, which));
    CUDA_TRY(cudaGetLastError());
    return 0;
}

static SharedStep shared_step(sdfb200_shared* dst, int which) {
    SharedStep st;
    st.cursor = shared_word(dst, 0, which); st.done = shared_word(dst, 1, which); st.ready = shared_word(dst, 2, which);
    st.my_seq = dst->my_seq + which;
    return st;
}

extern "C" int sdfb200_generate_shared(sdfb200_prog* prog, const double* xs, int64_t nx, const double* ys, int64_t ny,
                                       const double* zs, int64_t nz, int batch_size, int sparse, int precision, double tau,
                                       int zb_begin, int zb_end, sdfb200_shared* dst, int which, int64_t* n_tri_local,
                                       int64_t stats[6], double timings_ms[8], int device, void* stream) {
    if (!dst || which < 0 || which > 1) return fail(SDFB200_E_INVALID, "generate_shared: bad argument");
    const SharedStep st = shared_step(dst, which);
    return generate_impl(prog, xs, nx, ys, ny, zs, nz, batch_size, sparse, precision, tau, zb_begin, zb_end,
                         shared_data(dst, which), dst->capacity, n_tri_local, stats, timings_ms, device, stream, nullptr, &st);
}

extern "C" int sdfb200_generate_window(sdfb200_prog* prog, const double* xs, int64_t nx, const double* ys, int64_t ny,
                                       const double* zs, int64_t nz, int batch_size, int sparse, int precision, double tau,
                                       int64_t ub_begin, int64_t ub_end, float* tri_out, int64_t tri_capacity,
                                       sdfb200_shared* dst, int which, int64_t* n_tri, int64_t stats[6],
                                       double timings_ms[8], int device, void* stream) {
    if (ub_begin < 0 || ub_end < ub_begin) return fail(SDFB200_E_INVALID, "generate_window: bad window");
    if (dst) {
        if (which < 0 || which > 1) return fail(SDFB200_E_INVALID, "generate_window: bad cursor");
        const SharedStep st = shared_step(dst, which);
        return generate_impl(prog, xs, nx, ys, ny, zs, nz, batch_size, sparse, precision, tau, 0, -1,
                             shared_data(dst, which), dst->capacity, n_tri, stats, timings_ms, device, stream, nullptr, &st,
                             ub_begin, ub_end);
    }
    return generate_impl(prog, xs, nx, ys, ny, zs, nz, batch_size, sparse, precision, tau, 0, -1, tri_out, tri_capacity,
                         n_tri, stats, timings_ms, device, stream, nullptr, nullptr, ub_begin, ub_end);
}

// skimage.measure.marching_cubes(volume, 0) + world mapping for ONE batch volume supplied by the caller (the values
// of an SDF the evaluator cannot lower: a user's own numpy closure) -- see include/sdfb200.h.
extern "C" int sdfb200_mc_volume(sdfb200_prog* prog, const float* volume, int64_t nx, int64_t ny, int64_t nz,
                                 const double* xs, const double* ys, const double* zs, float* tri_out, int64_t tri_capacity,
                                 int64_t* n_tri, int device, void* stream_) {
    if (!prog || !volume || !n_tri || !xs || !ys || !zs) return fail(SDFB200_E_INVALID, "mc_volume: NULL argument");
    if (nx < 1 || ny < 1 || nz < 1 || nx > kTS + 1 || ny > kTS + 1 || nz > kTS + 1)
        return fail(SDFB200_E_INVALID, "mc_volume: a batch volume has 1..%d samples per axis (got %lld x %lld x %lld)", kTS + 1,
                    (long long)nx, (long long)ny, (long long)nz);
    if (tri_capacity < 0) return fail(SDFB200_E_INVALID, "mc_volume: negative capacity");
    std::lock_guard<std::mutex> guard(prog->lock);
    DeviceCtx* c; int rc = get_ctx(prog, device, &c); if (rc) return rc;
    cudaStream_t stream = (cudaStream_t)stream_;
    *n_tri = 0;
    if (nx < 2 || ny < 2 || nz < 2) return 0;                    // skimage raises, the reference maps it to [] (core.py:53-56)
    GridDesc g;
    rc = setup_grid(c, stream, xs, nx, ys, ny, zs, nz, kTS, 0, -1, -1, -1, &g); if (rc) return rc;
    rc = grow(&c->vol, &c->slot_cap, (size_t)VOL_STRIDE); if (rc) return rc;
    rc = grow(&c->tile_meta, &c->meta_cap, (size_t)1); if (rc) return rc;
    rc = grow(&c->cubes, &c->cube_cap, (size_t)kTS * kTS * kTS); if (rc) return rc;
    TileLists L;
    memset(&L, 0, sizeof(L));
    L.vol = c->vol; L.tile_meta = c->tile_meta; L.cubes = c->cubes; L.pts = c->pts;
    L.slot_cap = 1; L.cube_cap = (unsigned int)std::min<size_t>(c->cube_cap, 0xffffffffu); L.pt_cap = 0;
    const size_t nvox = (size_t)nx * ny * nz;
    CUDA_TRY(cudaMemcpyAsync(c->vol, volume, nvox * sizeof(float), cudaMemcpyDefault, stream));
    TileMeta m;
    memset(&m, 0, sizeof(m));
    m.ext = (unsigned int)nx | ((unsigned int)ny << 8) | ((unsigned int)nz << 16);
    CUDA_TRY(cudaMemcpyAsync(c->tile_meta, &m, sizeof(m), cudaMemcpyHostToDevice, stream));
    const size_t nb = (size_t)g.nbx * g.nby * g.nbz;
    CUDA_TRY(cudaMemsetAsync(c->batch_tris, 0, nb * sizeof(unsigned int), stream));
    const int ncubes = (int)((nx - 1) * (ny - 1) * (nz - 1));
    volume_candidates_kernel<<<(ncubes + 255) / 256, 256, 0, stream>>>(L, c->ctr, (int)nx, (int)ny, (int)nz);
    CUDA_TRY(cudaGetLastError());
    const int out_kind = tri_out ? mem_kind(tri_out) : 1;
    float* dtri = tri_out;
    if (tri_out && out_kind != 1) {
        rc = ensure_stage(c, (size_t)tri_capacity * 36 + 256); if (rc) return rc;
        dtri = (float*)c->stage;
    }
    extract_kernel<kTS><<<std::min(c->sms, (ncubes + EXTRACT_THREADS - 1) / EXTRACT_THREADS), EXTRACT_THREADS, 0, stream>>>(
        g, L, c->ctr, c->batch_tris, dtri, tri_out ? (unsigned long long)tri_capacity : 0ull, nullptr, c->mc_tables, nullptr, nullptr, nullptr);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(c->ctr_host, c->ctr, sizeof(Counters), cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    const Counters& h = *c->ctr_host;
    *n_tri = (int64_t)h.tri_cursor;
    if (tri_out && h.tri_cursor > (unsigned long long)tri_capacity)
        return fail(SDFB200_E_CAPACITY, "triangle buffer too small: %lld needed, %lld given", (long long)h.tri_cursor, (long long)tri_capacity);
    if (tri_out && out_kind != 1 && h.tri_cursor > 0) {
        CUDA_TRY(cudaMemcpyAsync(tri_out, dtri, (size_t)h.tri_cursor * 36, cudaMemcpyDeviceToHost, stream));
        CUDA_TRY(cudaStreamSynchronize(stream));
    }
    return 0;
}

extern "C" int sdfb200_jit_probe(sdfb200_prog* prog, int kernel, int64_t* cubin_bytes) {
    if (!prog || kernel < 0 || kernel >= JK_COUNT) return fail(SDFB200_E_INVALID, "jit_probe: bad argument");
    std::lock_guard<std::mutex> guard(prog->lock);
    const JitResult r = jit_compile(prog, kernel);
    if (!r.ok) return fail(SDFB200_E_INVALID, "%s", r.log.c_str());
    if (cubin_bytes) *cubin_bytes = (int64_t)r.cubin.size();
    return 0;
}

extern "C" int sdfb200_jit_plan(sdfb200_prog* prog, int precision, int* share_axis_out, int* samples_per_thread,
                                double* invariant_fraction) {
    if (!prog || precision < 0 || precision > 3) return fail(SDFB200_E_INVALID, "jit_plan: bad argument");
    std::lock_guard<std::mutex> guard(prog->lock);
    const int kernel = JK_TILE0 + precision;
    const JitResult r = jit_compile(prog, kernel);
    if (!r.ok) return fail(SDFB200_E_INVALID, "%s", r.log.c_str());
    prog->plan_share[kernel] = r.share; prog->plan_n[kernel] = r.n;
    if (share_axis_out) *share_axis_out = r.share;
    if (samples_per_thread) *samples_per_thread = r.n;
    if (invariant_fraction) share_axis(prog, invariant_fraction);
    return 0;
}

extern "C" int sdfb200_skip_flags(sdfb200_prog* prog, const double* xs, int64_t nx, const double* ys, int64_t ny,
                                  const double* zs, int64_t nz, int batch_size, uint8_t* flags, int device, void* stream_) {
    if (!prog || !flags) return fail(SDFB200_E_INVALID, "skip_flags: NULL argument");
    std::lock_guard<std::mutex> guard(prog->lock);
    DeviceCtx* c; int rc = get_ctx(prog, device, &c); if (rc) return rc;
    cudaStream_t stream = (cudaStream_t)stream_;
    GridDesc g;
    rc = setup_grid(c, stream, xs, nx, ys, ny, zs, nz, batch_size, 0, -1, -1, -1, &g); if (rc) return rc;
    const long long total_b = (long long)g.nbx * g.nby * g.nbz;
    const int pb = (int)((total_b * PREPASS_LANES + 127) / 128);
    if (cudaKernel_t jk = jit_get(prog, JK_SKIP, total_b >= 4096)) {
        int sparse = 1;
        unsigned char* fl = c->flags; unsigned int* act = c->active; unsigned int* bt = c->batch_tris; Counters* ctrp = c->ctr;
        void* args[] = {(void*)&g, (void*)&sparse, (void*)&fl, (void*)&act, (void*)&bt, (void*)&ctrp};
        CUDA_TRY(jit_launch(jk, dim3(pb), dim3(128), 0, stream, args));
    } else {
        REQUIRE_IMAGE(prog);
        skip_prepass_kernel<<<pb, 128, 0, stream>>>(prog->img, g, 1, c->flags, c->active, c->batch_tris, c->ctr);
    }
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(flags, c->flags, (size_t)total_b, cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    return 0;
}
