// sdf_kernels.cuh -- device side of the hot path (bodies of every kernel), templated on the SDF
// evaluator EV: InterpEval (interpreter over a ProgramImage, sdf_ops.cuh) in the statically compiled
// library, SpecEval (generated straight-line code) in the run-time compiled one.  Compiles under nvcc
// and NVRTC.
//
//   skip_prepass_body     core._skip (core.py:28-43): float64, 1 + 8 samples per batch
//   fused_tile_body       per batch tile: lattice samples generated in registers from the float64
//                         axes, whole SDF tree evaluated into a shared-memory float32 tile
//                         (core.py:51-52), float64 re-evaluation of everything that can reach the
//                         output, warp-shuffle min/max early-out (skimage's "level within data range"
//                         check), Lewiner marching cubes on the staged tile with the tables in shared
//                         memory and warp/CTA prefix-sum compaction (core.py:16-18), world mapping
//                         (core.py:58-60)
//   batch_stats_body      skipped / empty / nonempty (core.py:135-141)
//   eval_stream_*_body    SDF3.__call__ on an explicit point block (d3.py:24-25): float3 points
//                         streamed HBM -> shared memory with cp.async.bulk (TMA bulk copy, mbarrier
//                         completion), distances streamed back
#pragma once
#include "sdf_ops.cuh"

#ifdef __CUDACC_RTC__
#define FLT_EPSILON 1.192092896e-07F
#define FLT_MAX 3.402823466e+38F
typedef unsigned long long uintptr_t;
#else
#include <float.h>
#endif

#ifndef SDFB200_PREC_F32
#define SDFB200_PREC_F32 0
#define SDFB200_PREC_MIXED 1
#define SDFB200_PREC_F64 2
#define SDFB200_PREC_EXACT 3
#endif

#define MC_TABLE_QUAL static const          // host-side data: packed into one device image by sdfb200.cu (get_ctx)
#include "mc_tables.inc"

// ---------------------------------------------------------------------------------------
// device-side job description

struct GridDesc {
    const double* xs;
    const double* ys;
    const double* zs;
    const float* xf;       // the same axes rounded to float32 (the fast classification pass reads these)
    const float* yf;
    const float* zf;
    int nx, ny, nz;        // lattice points per axis
    int bs;                // cubes per batch side (core.py BATCH_SIZE)
    int nbx, nby, nbz;     // batches per axis = ceil(n / bs)   (core.py:115-117)
    int ub0, ub1;          // batch window of this call, in Z-MAJOR order u = bz * (nbx * nby) + bx * nby + by
                           // (a Z-slab of whole batch layers [zb0, zb1) is the window [zb0, zb1) * nbx * nby)
    int sub;               // sub-tiles per batch side = ceil(bs / TS)
};

struct Counters {
    unsigned long long tri_cursor;      // triangles allocated so far
    unsigned long long pts_f32;         // samples evaluated in float32 (or f64 in PREC_F64)
    unsigned long long pts_f64;         // samples re-evaluated in float64
    unsigned long long active_cubes;
    unsigned int n_active;              // non-skipped batches
    unsigned int work_cursor;
    unsigned int skipped, empty, nonempty;
    unsigned int n_slots, n_cubes, n_pts;       // per round: stored tiles, candidate cubes, queued samples
    unsigned int extract_done;                  // per round: CTAs of the extraction kernel that have finished (its tail work)
    unsigned int max_cubes, max_pts;            // high-water marks over rounds (capacity check)
    unsigned int pad_[1];
};

// What the host needs to know when a round of a pipelined call is done, written by the round's extraction kernel
// into page-locked, device-mapped host memory: the host polls `done` and starts the copy of the round's triangles
// without a stream synchronisation.
struct RoundProgress {
    unsigned long long tri_cursor;              // triangles produced so far (all rounds up to this one)
    unsigned int max_cubes, max_pts;            // list high-water marks (overflow -> the host retries the call)
    unsigned int done;                          // 0 -> 1, written last (after a system-scope fence)
    unsigned int pad_[3];
};

// The tiles [wb, we) of the active list a kernel works on:
//   work_end_arg == 0xffffffff              every active tile (the count is read from the pre-pass's counter)
//   work_end_arg == 0xfffe0000 | R          round `work_begin` of R equal rounds of the active tiles (decided on the
//                                           device: the host queues all rounds without knowing the count)
//   otherwise                               [work_begin, work_end_arg) as given by the host
__device__ __forceinline__ void work_range(const Counters* ctr, unsigned int subs3, unsigned int work_begin,
                                           unsigned int work_end_arg, unsigned int* wb, unsigned int* we) {
    if (work_end_arg == 0xffffffffu) { *wb = work_begin; *we = ctr->n_active * subs3; return; }
    if ((work_end_arg & 0xffff0000u) == 0xfffe0000u) {
        const unsigned int total = ctr->n_active * subs3, rounds = work_end_arg & 0xffffu;
        const unsigned int per = (total + rounds - 1u) / rounds;
        *wb = min(work_begin * per, total);
        *we = min(*wb + per, total);
        return;
    }
    *wb = work_begin; *we = work_end_arg;
}

// t-th batch of the call's window -> batch coordinates; the batch id keeps core.py:119's order (X-major)
__device__ __forceinline__ unsigned int window_batch(const GridDesc& g, int t, int* bx, int* by, int* bz) {
    const int u = g.ub0 + t, nxy = g.nbx * g.nby;
    *bz = u / nxy;
    const int r = u - *bz * nxy;
    *bx = r / g.nby;
    *by = r - *bx * g.nby;
    return (unsigned int)((*bx * g.nby + *by) * g.nbz + *bz);
}

// ---------------------------------------------------------------------------------------
// float64 helpers that the compiler may not contract into FMAs: the marching-cubes
// arithmetic must round exactly like the C restatement (and skimage's C code).

__device__ __forceinline__ double xmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double xadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double xsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double xdiv(double a, double b) { return __ddiv_rn(a, b); }

// Lewiner corner k = (x,y,z): 0(0,0,0) 1(1,0,0) 2(1,1,0) 3(0,1,0) 4..7 same at z=1; edges 0-3 bottom ring,
// 4-7 top ring, 8-11 verticals.  Computed arithmetically: per-thread edge ids diverge, and divergent
// __constant__ reads serialise.
__device__ __forceinline__ int corner_x(int k) { return (k ^ (k >> 1)) & 1; }
__device__ __forceinline__ int corner_y(int k) { return (k >> 1) & 1; }
__device__ __forceinline__ int corner_z(int k) { return (k >> 2) & 1; }
__device__ __forceinline__ int edge_corner(int e, int s) {
    if (e < 8) return (e & 4) | ((e + s) & 3);
    return e - 8 + 4 * s;
}
__constant__ int c_face_corners[7][4] = {{0, 0, 0, 0}, {0, 4, 5, 1}, {1, 5, 6, 2}, {2, 6, 7, 3}, {3, 7, 4, 0}, {0, 3, 2, 1}, {4, 7, 6, 5}};
__constant__ int c_interior_edge[12][8] = {
    {0, 1, 3, 2, 7, 6, 4, 5}, {1, 2, 0, 3, 4, 7, 5, 6}, {2, 3, 1, 0, 5, 4, 6, 7}, {3, 0, 2, 1, 6, 5, 7, 4},
    {4, 5, 7, 6, 3, 2, 0, 1}, {5, 6, 4, 7, 0, 3, 1, 2}, {6, 7, 5, 4, 1, 0, 2, 3}, {7, 4, 6, 5, 2, 1, 3, 0},
    {0, 4, 3, 7, 2, 6, 1, 5}, {1, 5, 0, 4, 3, 7, 2, 6}, {2, 6, 1, 5, 0, 4, 3, 7}, {3, 7, 2, 6, 1, 5, 0, 4}};

// skimage test_internal; mode 12 = cases 4/10, 0..11 = reference edge
__device__ __forceinline__ bool mc_test_interior(const double* v, int mode, int s) {
    double t, At = 0, Bt = 0, Ct = 0, Dt = 0;
    const double eps = (double)FLT_EPSILON;
    if (mode == 12) {
        const double a = xsub(xmul(xsub(v[4], v[0]), xsub(v[6], v[2])), xmul(xsub(v[7], v[3]), xsub(v[5], v[1])));
        const double b = xsub(xsub(xadd(xmul(v[2], xsub(v[4], v[0])), xmul(v[0], xsub(v[6], v[2]))),
                                   xmul(v[1], xsub(v[7], v[3]))),
                              xmul(v[3], xsub(v[5], v[1])));
        t = xdiv(-b, xadd(xmul(2.0, a), eps));
        if (t < 0 || t > 1) return s > 0;
        At = xadd(v[0], xmul(xsub(v[4], v[0]), t));
        Bt = xadd(v[3], xmul(xsub(v[7], v[3]), t));
        Ct = xadd(v[2], xmul(xsub(v[6], v[2]), t));
        Dt = xadd(v[1], xmul(xsub(v[5], v[1]), t));
    } else {
        const int* e = c_interior_edge[mode];
        t = xdiv(v[e[0]], xadd(xsub(v[e[0]], v[e[1]]), eps));
        At = 0;
        Bt = xadd(v[e[2]], xmul(xsub(v[e[3]], v[e[2]]), t));
        Ct = xadd(v[e[4]], xmul(xsub(v[e[5]], v[e[4]]), t));
        Dt = xadd(v[e[6]], xmul(xsub(v[e[7]], v[e[6]]), t));
    }
    int test = 0;
    if (At >= 0) test += 1;
    if (Bt >= 0) test += 2;
    if (Ct >= 0) test += 4;
    if (Dt >= 0) test += 8;
    switch (test) {
        case 0: case 1: case 2: case 3: case 4: case 6: case 8: case 9: case 12:
            return s > 0;
        case 5:
            if (xsub(xmul(At, Ct), xmul(Bt, Dt)) < eps) return s > 0;
            break;
        case 10:
            if (xsub(xmul(At, Ct), xmul(Bt, Dt)) >= eps) return s > 0;
            break;
        default:
            return s < 0;
    }
    return s < 0;
}

// ---------------------------------------------------------------------------------------
// mbarrier + TMA bulk copy (cp.async.bulk: UBLKCP in SASS) helpers

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    }
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ---------------------------------------------------------------------------------------
// shared-memory copies of the decision tables ("edge table in shared memory")

struct __align__(16) McTablesSmem {
    unsigned short sub_base[256];
    unsigned short tile_off[MC_TILE_OFF_LEN];
    short sub_true[MC_SUB_TRUE_LEN];
    short sub_false[MC_SUB_FALSE_LEN];
    signed char sub_interior[MC_SUB_INTERIOR_LEN];
    unsigned char nface[256];
    unsigned char faces[256 * 6];
    unsigned char tile_ntri[MC_TILE_NTRI_LEN];
    unsigned char tile_data[MC_TILE_DATA_LEN];
};

// The host packs the tables once per device into one McTablesSmem image in global memory (sdfb200.cu:
// mc_tables_device); a CTA stages it with ONE cp.async.bulk (TMA bulk copy, mbarrier completion) that is in
// flight while the CTA's first cubes are gathered.
static_assert(sizeof(McTablesSmem) % 16 == 0, "bulk copy size");

__device__ __forceinline__ void load_mc_tables_async(McTablesSmem* dst, const McTablesSmem* __restrict__ src, uint64_t* bar, int tid) {
    if (tid == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect_tx(bar, (unsigned)sizeof(McTablesSmem));
        tma_bulk_g2s(dst, src, (unsigned)sizeof(McTablesSmem), bar);
    }
}

// ---------------------------------------------------------------------------------------
// skip pre-pass (core.py:28-43), one thread per batch, float64

template <class EV>
__device__ __forceinline__ double eval1_f64(const EV& ev, double px, double py, double pz) {
    double x[1] = {px}, y[1] = {py}, z[1] = {pz}, d[1];
    ev.template run<double, 1, false>(x, y, z, d);
    return d[0];
}

// One batch = 16 consecutive lanes (two batches per warp): lane 0 evaluates the centre, lanes 1..8 the
// corners in itertools.product((x0,x1),(y0,y1),(z0,z1)) order, lanes 9..15 idle.  The reference evaluates
// the corners only when |centre| > half diagonal and stops at the first sign mismatch -- the flag is the
// same when all nine are evaluated side by side (the field is a pure function), and the pre-pass then
// costs ONE float64 evaluation of latency instead of nine.
#define PREPASS_LANES 16
template <class EV>
__device__ __forceinline__ void skip_prepass_body(const EV& img, const GridDesc& g, int sparse,
                                                  unsigned char* __restrict__ flags,
                                                  unsigned int* __restrict__ active_list,
                                                  unsigned int* __restrict__ batch_tris,
                                                  Counters* __restrict__ ctr) {
    const int total = g.ub1 - g.ub0;
    const int gt = blockIdx.x * blockDim.x + threadIdx.x;
    const int tb = gt / PREPASS_LANES, k = gt % PREPASS_LANES;
    const bool in = tb < total;
    const int t = in ? tb : 0;
    int bx, by, bz;
    const unsigned int bid = window_batch(g, t, &bx, &by, &bz);
    bool skipped = false;
    if (sparse) {                                               // (uniform: every lane reaches the shuffles)
        const int ix0 = bx * g.bs, iy0 = by * g.bs, iz0 = bz * g.bs;
        const int ix1 = min(ix0 + g.bs, g.nx - 1), iy1 = min(iy0 + g.bs, g.ny - 1), iz1 = min(iz0 + g.bs, g.nz - 1);
        const double x0 = g.xs[ix0], x1 = g.xs[ix1], y0 = g.ys[iy0], y1 = g.ys[iy1], z0 = g.zs[iz0], z1 = g.zs[iz1];
        const double x = __ddiv_rn(__dadd_rn(x0, x1), 2.0), y = __ddiv_rn(__dadd_rn(y0, y1), 2.0),
                     z = __ddiv_rn(__dadd_rn(z0, z1), 2.0);
        double v = 0.0;
        if (in && k < 9) {
            const int c = k - 1;                                // corner id: bit 2 -> x, bit 1 -> y, bit 0 -> z
            const double px = k == 0 ? x : ((c & 4) ? x1 : x0), py = k == 0 ? y : ((c & 2) ? y1 : y0),
                         pz = k == 0 ? z : ((c & 1) ? z1 : z0);
            v = eval1_f64(img, px, py, pz);
        }
        const int base = (int)(threadIdx.x & 31u) & ~(PREPASS_LANES - 1);      // lane 0 of this batch inside the warp
        const double vc = __shfl_sync(0xffffffffu, v, base);
        const double v0 = __shfl_sync(0xffffffffu, v, base + 1);
        const bool pos = v0 > 0;                                // values[0] decides the sign (core.py:41-42)
        const bool mine = (k >= 1 && k <= 8) ? (pos ? (v > 0) : (v < 0)) : true;
        const unsigned int bal = __ballot_sync(0xffffffffu, mine);
        const bool same = ((bal >> base) & 0xffffu) == 0xffffu;
        const double r = fabs(vc);
        const double ex = __dsub_rn(x, x0), ey = __dsub_rn(y, y0), ez = __dsub_rn(z, z0);
        const double dd = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(ex, ex), __dmul_rn(ey, ey)), __dmul_rn(ez, ez)));
        skipped = !(r <= dd) && same;
    }
    if (!in || k != 0) return;
    flags[bid] = skipped ? 1 : 0;
    batch_tris[bid] = 0u;
    if (!skipped) {
        const unsigned int slot = atomicAdd(&ctr->n_active, 1u);
        active_list[slot] = bid;
    } else {
        atomicAdd(&ctr->skipped, 1u);
    }
}

// ---------------------------------------------------------------------------------------
// tile pipeline: (A) fused evaluate + classify per tile, (B) float64 refinement of the samples that
// can reach the output, (C) marching-cubes extraction -- B and C are flat, full-occupancy kernels
// over the compacted lists that A builds, so the sparse, latency-bound work never runs inside one
// lonely CTA.

#define VOL_STRIDE 35968            // floats per stored tile: 33^3 = 35937 rounded up to a multiple of 32

// Geometry of a stored tile, written once by the kernel that allots the slot, so that the flat kernels
// over the compacted lists (B, C) do not redo tile_geom()'s integer divisions per sample / per cube.
struct TileMeta {
    int bx0, by0, bz0;              // lattice index of the batch origin (world mapping of core.py:58-59)
    unsigned int bid;               // batch id (batch_tris flag)
    int ox, oy, oz;                 // sub-tile offset inside the batch (samples)
    unsigned int ext;               // px | py << 8 | pz << 16: sub-tile extent (samples, <= 33)
};

struct TileLists {
    float* vol;                     // [slot][VOL_STRIDE] float32 tile volumes (tile-local linearisation)
    TileMeta* tile_meta;            // [slot]
    uint2* cubes;                   // candidate cubes {slot, tile-local cube id}
    uint2* pts;                     // samples to re-evaluate in float64 {slot, tile-local sample id}
    unsigned int slot_cap, cube_cap, pt_cap;
};

struct TileGeom {
    int bx0, by0, bz0;              // lattice index of the batch origin
    int ox, oy, oz;                 // sub-tile offset inside the batch (samples)
    int px, py, pz;                 // sub-tile extent (samples)
};

__device__ __forceinline__ TileMeta make_tile_meta(const TileGeom& t, unsigned int bid) {
    TileMeta m;
    m.bx0 = t.bx0; m.by0 = t.by0; m.bz0 = t.bz0; m.bid = bid;
    m.ox = t.ox; m.oy = t.oy; m.oz = t.oz;
    m.ext = (unsigned int)t.px | ((unsigned int)t.py << 8) | ((unsigned int)t.pz << 16);
    return m;
}
__device__ __forceinline__ TileMeta load_tile_meta(const TileMeta* p) {
    const int4 a = __ldg(reinterpret_cast<const int4*>(p)), b = __ldg(reinterpret_cast<const int4*>(p) + 1);
    TileMeta m;
    m.bx0 = a.x; m.by0 = a.y; m.bz0 = a.z; m.bid = (unsigned int)a.w;
    m.ox = b.x; m.oy = b.y; m.oz = b.z; m.ext = (unsigned int)b.w;
    return m;
}

// batch id + sub-tile index -> extents (core.py:115-117: slices [i, i+s+1) clipped to the axis)
template <int TS>
__device__ __forceinline__ bool tile_geom(const GridDesc& g, unsigned int bid, int sidx, TileGeom& t) {
    const int sz = sidx % g.sub, sy = (sidx / g.sub) % g.sub, sx = sidx / (g.sub * g.sub);
    const int bz = (int)(bid % (unsigned)g.nbz), by = (int)((bid / (unsigned)g.nbz) % (unsigned)g.nby),
              bx = (int)(bid / (unsigned)(g.nbz * g.nby));
    t.bx0 = bx * g.bs; t.by0 = by * g.bs; t.bz0 = bz * g.bs;
    const int blx = min(g.bs + 1, g.nx - t.bx0), bly = min(g.bs + 1, g.ny - t.by0), blz = min(g.bs + 1, g.nz - t.bz0);
    t.ox = sx * TS; t.oy = sy * TS; t.oz = sz * TS;
    t.px = min(TS + 1, blx - t.ox); t.py = min(TS + 1, bly - t.oy); t.pz = min(TS + 1, blz - t.oz);
    if (t.px <= 0 || t.py <= 0 || t.pz <= 0) return false;
    if (sidx != 0 && (t.px < 2 || t.py < 2 || t.pz < 2)) return false;   // a 1-thick remainder holds no cube
    return true;
}

// n / d for n < 2^16 * d, d >= 1, through a precomputed reciprocal (the tile loops divide a lot)
struct FastDiv {
    unsigned int m; unsigned int d;
    __device__ __forceinline__ void init(unsigned int d_) { d = d_; m = (unsigned int)((0xffffffffull / d_) + 1ull); }
    __device__ __forceinline__ unsigned int div(unsigned int n) const {
        unsigned int q = __umulhi(n, m);
        return (d == 1) ? n : q;
    }
};

template <int TS, bool F64>
struct TileSmem {
    static constexpr int TP = TS + 1;
    static constexpr int NPTS = TP * TP * TP;
    float tile[(NPTS + 3) & ~3];
    unsigned long long posm[TP * TP];               // per z-line: bit j <=> sample j of the line is > 0
    unsigned long long uncm[TP * TP];               //             bit j <=> |sample j| < tau (sign not trusted)
    unsigned long long refm[TP * TP];               //             bit j <=> sample j is queued for float64
    unsigned int cand[TS * TS];                     // per (cx, cy): bit cz <=> cube (cx, cy, cz) is a candidate
    float axf[3][TP];
    double axd[F64 ? 3 : 1][F64 ? TP : 1];
    int minmax[2];
    unsigned int work_item, slot, any_cand;
    unsigned int warp_sums[32];
    unsigned int cube_gbase, pt_gbase;
};

__device__ __forceinline__ int float_to_ordered(float f) {
    const int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7fffffff;
}

// (A) for the precisions that keep float32 values: one CTA per tile evaluates the lattice samples into shared
// memory, classifies every cube, queues the candidate cubes and the samples float64 must confirm, stores
// the tile volume.
//   PREC_F32   accurate float32; candidates = cubes with mixed `v > 0` corners
//   PREC_MIXED accurate float32; samples with |d| < tau are uncertain -> queued for float64
//   PREC_F64   float64 evaluation of every sample
//   (PREC_EXACT -- the default -- needs no values at all: tile_sign_body below)
template <int TS, int PREC, int THREADS, class EV>
__device__ __forceinline__ void tile_eval_body(const EV& ev, const GridDesc& g, const float tau,
                                               const unsigned int* __restrict__ active_list,
                                               const unsigned int work_begin_arg, const unsigned int work_end_arg,
                                               const TileLists L, Counters* __restrict__ ctr) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr bool F64 = (PREC == SDFB200_PREC_F64);
    constexpr bool REFINE = (PREC == SDFB200_PREC_MIXED || PREC == SDFB200_PREC_EXACT);
    static_assert(PREC != SDFB200_PREC_EXACT, "precision `exact` runs tile_sign_body (no value tile)");
    constexpr bool EXACT = false;
    using S = TileSmem<TS, F64>;
    S& sm = *reinterpret_cast<S*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31;
    const unsigned int subs3 = (unsigned int)(g.sub * g.sub * g.sub);
    // work_end == ~0u: "every active batch" -- the count is read from the pre-pass's counter, so the
    // host does not have to synchronise between the two kernels
    unsigned int work_begin, work_end;
    work_range(ctr, subs3, work_begin_arg, work_end_arg, &work_begin, &work_end);
    unsigned long long my_pts = 0;

    // Every barrier below is executed unconditionally by every thread in every iteration: the
    // per-tile decisions (degenerate tile, early out, no candidates) only predicate the work
    // between the barriers.
    if (tid == 0) {
        sm.work_item = work_begin + atomicAdd(&ctr->work_cursor, 1u);
        sm.minmax[0] = 0x7fffffff;
        sm.minmax[1] = (int)0x80000000;
        sm.any_cand = 0;
    }
    __syncthreads();
    for (;;) {
        const unsigned int w = sm.work_item;
        if (w >= work_end) break;
        const unsigned int bid = active_list[w / subs3];
        const int sidx = (int)(w % subs3);
        TileGeom t;
        const bool valid = tile_geom<TS>(g, bid, sidx, t);
        if (!valid) { t.px = t.py = t.pz = 1; t.bx0 = t.by0 = t.bz0 = t.ox = t.oy = t.oz = 0; }
        const int px = t.px, py = t.py, pz = t.pz;
        const int npts = px * py * pz;
        const unsigned int pyz = (unsigned int)(py * pz);
        FastDiv dpyz, dpz;
        dpyz.init(pyz); dpz.init((unsigned int)pz);

        for (int i = tid; i < px; i += THREADS) { const double v = g.xs[t.bx0 + t.ox + i]; sm.axf[0][i] = (float)v; if (F64) sm.axd[0][F64 ? i : 0] = v; }
        for (int i = tid; i < py; i += THREADS) { const double v = g.ys[t.by0 + t.oy + i]; sm.axf[1][i] = (float)v; if (F64) sm.axd[F64 ? 1 : 0][F64 ? i : 0] = v; }
        for (int i = tid; i < pz; i += THREADS) { const double v = g.zs[t.bz0 + t.oz + i]; sm.axf[2][i] = (float)v; if (F64) sm.axd[F64 ? 2 : 0][F64 ? i : 0] = v; }
        __syncthreads();

        // ---------------------------------------------------------------- evaluation
        // A tile that is one sample thick along some axis holds no cube (the reference still evaluates
        // it, skimage raises, the batch counts as empty: core.py:53-56): nothing to evaluate.
        const bool thin = !valid || px < 2 || py < 2 || pz < 2;
        float vmin = FLT_MAX, vmax = -FLT_MAX;
        if (thin) {
        } else if (F64) {
            for (int idx = tid; idx < npts; idx += THREADS) {
                const unsigned int ix = dpyz.div(idx), rem = idx - ix * pyz, iy = dpz.div(rem), iz = rem - iy * pz;
                double x[1] = {sm.axd[0][F64 ? ix : 0]}, y[1] = {sm.axd[F64 ? 1 : 0][F64 ? iy : 0]},
                       z[1] = {sm.axd[F64 ? 2 : 0][F64 ? iz : 0]}, d[1];
                ev.template run<double, 1, false>(x, y, z, d);
                const float f = (float)d[0];
                sm.tile[idx] = f;
                vmin = fminf(vmin, f); vmax = fmaxf(vmax, f);
            }
        } else {
            constexpr int N = EV::kN;
            constexpr int SH = EV::kShare;
            if (SH < 0) {
                for (int base = 0; base < npts; base += THREADS * N) {
                    float x[N], y[N], z[N], d[N];
#pragma unroll
                    for (int k = 0; k < N; ++k) {
                        int idx = base + k * THREADS + tid;
                        if (idx >= npts) idx = npts - 1;
                        const unsigned int ix = dpyz.div(idx), rem = idx - ix * pyz, iy = dpz.div(rem), iz = rem - iy * pz;
                        x[k] = sm.axf[0][ix]; y[k] = sm.axf[1][iy]; z[k] = sm.axf[2][iz];
                    }
                    ev.template run<float, N, EXACT>(x, y, z, d);
#pragma unroll
                    for (int k = 0; k < N; ++k) {
                        const int idx = base + k * THREADS + tid;
                        if (idx < npts) {
                            sm.tile[idx] = d[k];
                            vmin = fminf(vmin, d[k]); vmax = fmaxf(vmax, d[k]);
                        }
                    }
                }
            } else {
                // Axis sharing: a thread's N samples are consecutive along lattice axis SH and share the
                // other two coordinates -- literally the same registers -- so that in the generated
                // straight-line evaluator every sub-expression that does not depend on axis SH (a
                // cylinder along it, an extruded 2-D profile, a circular array around it ...) is
                // computed ONCE for the N samples by the compiler's common-subexpression elimination.
                // The host picks SH from a dependency analysis of the program (sdfb200.cu: share_axis).
                const int pa = SH == 0 ? px : (SH == 1 ? py : pz);
                const int nrest = SH == 0 ? (int)pyz : (SH == 1 ? px * pz : px * py);
                const int ngroups = ((pa + N - 1) / N) * nrest;
                const int astride = SH == 0 ? (int)pyz : (SH == 1 ? pz : 1);
                FastDiv drest, dfast;
                drest.init((unsigned int)nrest);
                dfast.init((unsigned int)(SH == 2 ? py : pz));
                for (int gb = 0; gb < ngroups; gb += THREADS) {
                    int gi = gb + tid;
                    const bool on = gi < ngroups;
                    if (!on) gi = ngroups - 1;
                    const unsigned int ga = drest.div(gi), rest = gi - ga * nrest;
                    const unsigned int hi = dfast.div(rest), lo = rest - hi * dfast.d;
                    const int ia0 = (int)ga * N;
                    float x[N], y[N], z[N], d[N];
                    int idx0;
                    if (SH == 2) {                    // rest = ix * py + iy
                        const float xv = sm.axf[0][hi], yv = sm.axf[1][lo];
                        idx0 = (int)rest * pz + ia0;
#pragma unroll
                        for (int k = 0; k < N; ++k) { x[k] = xv; y[k] = yv; z[k] = sm.axf[2][min(ia0 + k, pz - 1)]; }
                    } else if (SH == 1) {             // rest = ix * pz + iz
                        const float xv = sm.axf[0][hi], zv = sm.axf[2][lo];
                        idx0 = (int)(hi * pyz) + ia0 * pz + (int)lo;
#pragma unroll
                        for (int k = 0; k < N; ++k) { x[k] = xv; z[k] = zv; y[k] = sm.axf[1][min(ia0 + k, py - 1)]; }
                    } else {                          // rest = iy * pz + iz
                        const float yv = sm.axf[1][hi], zv = sm.axf[2][lo];
                        idx0 = ia0 * (int)pyz + (int)rest;
#pragma unroll
                        for (int k = 0; k < N; ++k) { y[k] = yv; z[k] = zv; x[k] = sm.axf[0][min(ia0 + k, px - 1)]; }
                    }
                    ev.template run<float, N, EXACT>(x, y, z, d);
#pragma unroll
                    for (int k = 0; k < N; ++k) {
                        if (on && ia0 + k < pa) {
                            sm.tile[idx0 + k * astride] = d[k];
                            vmin = fminf(vmin, d[k]); vmax = fmaxf(vmax, d[k]);
                        }
                    }
                }
            }
        }
        if (tid == 0 && !thin) my_pts += (unsigned long long)npts;
        // warp-shuffle min/max reduction, one shared atomic per warp
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            vmin = fminf(vmin, __shfl_xor_sync(0xffffffffu, vmin, o));
            vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
        }
        if (lane == 0) {
            atomicMin(&sm.minmax[0], float_to_ordered(vmin));
            atomicMax(&sm.minmax[1], float_to_ordered(vmax));
        }
        __syncthreads();
        // early out: no sample can be > 0, or none can be <= 0: every cube index is 0 or 255 (skimage
        // raises ValueError / RuntimeError there; the reference maps both to an empty batch)
        bool live;
        {
            const int omin = sm.minmax[0], omax = sm.minmax[1];
            const float fmn = __int_as_float(omin >= 0 ? omin : omin ^ 0x7fffffff);
            const float fmx = __int_as_float(omax >= 0 ? omax : omax ^ 0x7fffffff);
            const float guard = REFINE ? tau : 0.0f;
            live = valid && !(fmn > guard || fmx < -guard || px < 2 || py < 2 || pz < 2);
        }
        if (tid == 0 && live) sm.slot = atomicAdd(&ctr->n_slots, 1u);
        __syncthreads();
        const unsigned int slot = sm.slot;
        live = live && (slot < L.slot_cap);                   // always true: one slot per work item

        // ------------------------------------------------------------ classification
        // (1) one thread per z-line: sign / uncertainty bit masks of the line's <= 33 samples
        const int nlines = live ? px * py : 0;
        for (int ln = tid; ln < nlines; ln += THREADS) {
            const float* p = &sm.tile[ln * pz];
            unsigned long long pm = 0ull, um = 0ull;
            for (int j = 0; j < pz; ++j) {
                const float v = p[j];
                pm |= (unsigned long long)(v > 0.0f ? 1u : 0u) << j;
                if (REFINE) um |= (unsigned long long)(fabsf(v) < tau ? 1u : 0u) << j;
            }
            sm.posm[ln] = pm; sm.uncm[ln] = um; sm.refm[ln] = um;
        }
        __syncthreads();
        // (2) one thread per (cx, cy) column of cubes: 32 cubes classified with a few word operations.
        //     cube (cx,cy,cz) is a candidate when its corners are not all > 0, not all <= 0, or one of
        //     them is uncertain; PREC_EXACT queues every corner of a candidate for float64.
        const int cxn = px - 1, cyn = py - 1, czn = pz - 1;
        for (int col = tid; col < TS * TS; col += THREADS) {
            const int cx = col >> 5, cy = col & 31;          // TS == 32
            unsigned int W = 0u;
            if (live && cx < cxn && cy < cyn) {
                const int l00 = cx * py + cy, l01 = l00 + 1, l10 = l00 + py, l11 = l10 + 1;
                const unsigned long long p0 = sm.posm[l00], p1 = sm.posm[l01], p2 = sm.posm[l10], p3 = sm.posm[l11];
                const unsigned long long allp = (p0 & (p0 >> 1)) & (p1 & (p1 >> 1)) & (p2 & (p2 >> 1)) & (p3 & (p3 >> 1));
                const unsigned long long anyp = (p0 | (p0 >> 1)) | (p1 | (p1 >> 1)) | (p2 | (p2 >> 1)) | (p3 | (p3 >> 1));
                unsigned long long c64 = anyp & ~allp;
                if (REFINE) {
                    const unsigned long long u0 = sm.uncm[l00], u1 = sm.uncm[l01], u2 = sm.uncm[l10], u3 = sm.uncm[l11];
                    c64 |= (u0 | (u0 >> 1)) | (u1 | (u1 >> 1)) | (u2 | (u2 >> 1)) | (u3 | (u3 >> 1));
                }
                W = (unsigned int)c64 & (czn >= 32 ? 0xffffffffu : ((1u << czn) - 1u));
            }
            sm.cand[col] = W;
        }
        __syncthreads();
        // ---------------------------------------------- compaction of the two bitmaps
        // one prefix sum over the per-thread popcounts (cubes in the low, samples in the high half
        // word), ONE pair of global atomics per tile, then every thread writes its own entries
        {
            const int nwp = (live && REFINE) ? nlines : 0;
            unsigned int cnt = 0;
            for (int wi = tid; wi < TS * TS; wi += THREADS) cnt += (unsigned)__popc(sm.cand[wi]);
            for (int wi = tid; wi < nwp; wi += THREADS) cnt += ((unsigned)__popcll(sm.refm[wi])) << 16;
            unsigned int incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned int nb = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += nb;
            }
            if (lane == 31) sm.warp_sums[tid >> 5] = incl;
            __syncthreads();
            unsigned int warp_off = 0, total = 0;
#pragma unroll
            for (int wi = 0; wi < THREADS / 32; ++wi) {
                const unsigned int sv = sm.warp_sums[wi];
                if (wi < (tid >> 5)) warp_off += sv;
                total += sv;
            }
            if (tid == 0) {
                sm.any_cand = total & 0xffffu;
                sm.cube_gbase = (total & 0xffffu) ? atomicAdd(&ctr->n_cubes, total & 0xffffu) : 0u;
                sm.pt_gbase = (total >> 16) ? atomicAdd(&ctr->n_pts, total >> 16) : 0u;
            }
            __syncthreads();
            const unsigned int excl = warp_off + incl - cnt;
            unsigned int cat = sm.cube_gbase + (excl & 0xffffu);
            unsigned int pat = sm.pt_gbase + (excl >> 16);
            for (int wi = tid; wi < TS * TS; wi += THREADS) {
                unsigned int word = sm.cand[wi];
                while (word) {
                    const int bpos = __ffs(word) - 1;
                    word &= word - 1u;
                    // cube id with fixed strides: cx << 10 | cy << 5 | cz
                    if (cat < L.cube_cap) L.cubes[cat] = make_uint2(slot, (unsigned int)((wi << 5) | bpos));
                    ++cat;
                }
            }
            for (int wi = tid; wi < nwp; wi += THREADS) {
                unsigned long long word = sm.refm[wi];
                while (word) {
                    const int bpos = __ffsll((long long)word) - 1;
                    word &= word - 1ull;
                    if (pat < L.pt_cap) L.pts[pat] = make_uint2(slot, (unsigned int)(wi * pz + bpos));
                    ++pat;
                }
            }
        }
        __syncthreads();
        // ------------------------------------------------------- store the tile volume
        if (live && sm.any_cand) {
            float* dst = L.vol + (size_t)slot * VOL_STRIDE;
            for (int i = tid; i < npts; i += THREADS) dst[i] = sm.tile[i];
            if (tid == 0) L.tile_meta[slot] = make_tile_meta(t, bid);
        }
        __syncthreads();                                      // everyone is done with this tile's state
        if (tid == 0) {
            sm.work_item = work_begin + atomicAdd(&ctr->work_cursor, 1u);
            sm.minmax[0] = 0x7fffffff;
            sm.minmax[1] = (int)0x80000000;
            sm.any_cand = 0;
        }
        __syncthreads();
    }
    if (tid == 0) atomicAdd(&ctr->pts_f32, my_pts);
}

// ---------------------------------------------------------------------------------------
// (A') PREC_EXACT: evaluate + classify WITHOUT a value tile.
//
// In `exact` mode every float32 value that could reach the output is re-evaluated in float64 by (B), so
// the fast pass only has to answer two questions per sample: "is it > 0" and "is |d| < tau".  Both are
// single predicate bits, and with the lanes of a warp laid along one lattice axis a __ballot_sync turns
// 32 samples into one mask word while they are produced -- no 144 KB shared-memory value tile, no
// per-sample index division, no serial mask loop, no shared atomics, no min/max reduction (a tile
// whose masks hold no mixed cube simply yields no candidates: skimage's "level within data range"
// check, core.py:53-56), no tile store.
//
// (A'1) tile_sign_body -- evaluation by autonomous warps: no CTA barrier anywhere, every WARP is its own
// worker.  It takes batches of SIGN_BATCH thread-runs of one tile from a global cursor (a tile = 26
// batches: the sparse, uneven work balances at 1/26 tile per warp; the next batch is claimed while the
// current one computes, so the cursor's round trip is off the critical path).  In a thread-run a lane evaluates N
// consecutive samples along the share axis S (they share their other two coordinates, i.e. registers:
// the axis-sharing of the generated evaluator); the lane index is the coordinate along the lane axis U.
// Lanes 0..31 cover samples 0..31 of a U-line; the 33rd sample of every line (the plane shared with the
// next tile) is evaluated by separate thread-runs that write one byte per line.  Masks go to a per-tile
// scratch block in global memory (10 KB, L2 resident).
// (A'2) tile_classify_body -- one CTA per tile reads the masks back, classifies the 32^3 cubes with 64-bit
// word operations (one thread per column of 32 cubes), queues the candidate cubes and every sample
// float64 must confirm (bitmap compaction: one CTA prefix sum, one slot + one pair of list ranges per tile).
// All tiles classify side by side in a few microseconds; inside the evaluation kernel the same work
// was a serial, latency-bound tail (measured: profiles/r2_tile_gearlike_autonomous_v1.txt).
#define SIGN_TP 33
#define SIGN_LINES (SIGN_TP * SIGN_TP)
#define SIGN_BATCH 4                                            // thread-runs per work item
#define SIGN_STRIDE 9856                                        // bytes of scratch per tile: 1089 x {u32, u32} + 1089 u8, padded
#define SIGN_THREADS 256
#define SIGN_BLOCKS_PER_SM 4
#define CLASSIFY_THREADS 256

template <int N> struct SignShape {
    static constexpr int RUNS_S = (SIGN_TP + N - 1) / N;        // thread-runs per S-line
    static constexpr int R_MAIN = SIGN_TP * RUNS_S;             // run r = iw * RUNS_S + sr, lanes along U
    static constexpr int R_EXTRA = (R_MAIN + 31) / 32;          // runs of the 33rd U-layer: lane = (iw, sr) pair
    static constexpr int R_TILE = R_MAIN + R_EXTRA;
    static constexpr int NB = (R_TILE + SIGN_BATCH - 1) / SIGN_BATCH;       // batches (work items) per tile
};

// lattice axis roles of the sign kernels, from the evaluator's share axis (-1: no preference -> z)
template <int SHARE> struct SignAxes {
    static constexpr int AS = SHARE < 0 ? 2 : SHARE;           // thread-runs
    static constexpr int AU = AS == 2 ? 1 : 2;                  // lanes
    static constexpr int AW = AS == 0 ? 1 : 0;                  // batches of runs
};

template <int TS, int THREADS, class EV>
__device__ __forceinline__ void tile_sign_body(const EV& ev, const GridDesc& g, const float tau,
                                               const unsigned int* __restrict__ active_list,
                                               const unsigned int work_begin_arg, const unsigned int work_end_arg,
                                               Counters* __restrict__ ctr, unsigned char* __restrict__ scratch) {
    static_assert(TS == 32, "one mask word per line");
    constexpr int TP = SIGN_TP;
    constexpr int N = EV::kN;
    typedef SignAxes<EV::kShare> AX;
    constexpr int AS = AX::AS, AU = AX::AU, AW = AX::AW;
    typedef SignShape<N> SH;
    __shared__ float s_axis[THREADS / 32][TP + 15];             // warp-private: the tile's S-axis values (clamped past the end)
    const int lane = (int)(threadIdx.x & 31u);
    float* const sax = s_axis[threadIdx.x >> 5];
    const unsigned int subs3 = (unsigned int)(g.sub * g.sub * g.sub);
    unsigned int work_begin, work_end;
    work_range(ctr, subs3, work_begin_arg, work_end_arg, &work_begin, &work_end);
    const unsigned int n_items = (work_end > work_begin ? work_end - work_begin : 0u) * (unsigned int)SH::NB;

    unsigned int q_next = 0;
    if (lane == 0) q_next = atomicAdd(&ctr->work_cursor, 1u);
    for (;;) {
        const unsigned int q = __shfl_sync(0xffffffffu, q_next, 0);
        if (q >= n_items) break;
        if (lane == 0) q_next = atomicAdd(&ctr->work_cursor, 1u);       // the next batch: consumed at the top of the loop
        const unsigned int tw = q / (unsigned int)SH::NB;
        const int bb = (int)(q - tw * (unsigned int)SH::NB);
        const unsigned int w = work_begin + tw;
        const unsigned int bid = active_list[w / subs3];
        const int sidx = (int)(w % subs3);
        TileGeom t;
        const bool valid = tile_geom<TS>(g, bid, sidx, t);
        // a tile that is one sample thick along some axis holds no cube (the reference still evaluates it,
        // skimage raises, the batch counts as empty: core.py:53-56): nothing to evaluate
        if (!valid || t.px < 2 || t.py < 2 || t.pz < 2) continue;
        const int ext[3] = {t.px, t.py, t.pz};
        const int pu = ext[AU], ps = ext[AS], pw = ext[AW];
        const float* const axf[3] = {g.xf + (t.bx0 + t.ox), g.yf + (t.by0 + t.oy), g.zf + (t.bz0 + t.oz)};
        uint2* const masks = reinterpret_cast<uint2*>(scratch + (size_t)tw * SIGN_STRIDE);      // per line {pos, unc}, bits 0..31
        unsigned char* const hi = reinterpret_cast<unsigned char*>(masks + SIGN_LINES);       // per line bit 32: pos | unc << 1

        __syncwarp();                                             // (the previous batch is done with sax)
        sax[lane] = __ldg(axf[AS] + min(lane, ps - 1));
        if (lane < TP + 15 - 32) sax[32 + lane] = __ldg(axf[AS] + min(32 + lane, ps - 1));
        const float cu_main = __ldg(axf[AU] + min(lane, pu - 1));
        const float cu_last = __ldg(axf[AU] + (pu - 1));
        __syncwarp();
        const int r_end = min((bb + 1) * SIGN_BATCH, SH::R_TILE);
        for (int r = bb * SIGN_BATCH; r < r_end; ++r) {
            const bool main_run = r < SH::R_MAIN;                       // warp uniform
            const int rr = main_run ? r : (r - SH::R_MAIN) * 32 + lane;
            const int iw = rr / SH::RUNS_S, is0 = (rr - iw * SH::RUNS_S) * N;
            const bool on = rr < SH::R_MAIN && iw < pw && is0 < ps && (main_run || pu == TP);
            if (!__any_sync(0xffffffffu, on)) continue;                // (ragged tile / nothing left in this run)
            const float cw = __ldg(axf[AW] + min(iw, pw - 1));
            const float* const sp = sax + min(is0, TP - 1);            // (is0 + k <= 32 + N - 1 < TP + 15: filled, clamped)
            float c[3][N], d[N];
#pragma unroll
            for (int k = 0; k < N; ++k) {
                c[AU][k] = main_run ? cu_main : cu_last; c[AW][k] = cw;
                c[AS][k] = sp[k];
            }
            ev.template run<float, N, true>(c[0], c[1], c[2], d);
            if (main_run) {
#pragma unroll
                for (int k = 0; k < N; ++k) {
                    const unsigned int bp = __ballot_sync(0xffffffffu, d[k] > 0.0f);         // (bits >= pu: clamped duplicates,
                    const unsigned int bu = __ballot_sync(0xffffffffu, fabsf(d[k]) < tau);   //  masked off by the classifier)
                    if (lane == 0 && is0 + k < ps) masks[(is0 + k) * TP + iw] = make_uint2(bp, bu);
                }
            } else if (on) {
#pragma unroll
                for (int k = 0; k < N; ++k)
                    if (is0 + k < ps)
                        hi[(is0 + k) * TP + iw] = (unsigned char)((d[k] > 0.0f ? 1u : 0u) | (fabsf(d[k]) < tau ? 2u : 0u));
            }
        }
    }
}

struct ClassifySmem {
    uint4 block[SIGN_STRIDE / 16];                              // the tile's mask block, staged by one coalesced copy
    unsigned int cand[32 * 32];                                 // per (cs, cw): bit cu <=> cube is a candidate
    unsigned long long refm[SIGN_LINES];                        // per (is, iw) line: bit iu <=> sample queued for float64
    unsigned int warp_sums[CLASSIFY_THREADS / 32];
    unsigned int slot, cube_gbase, pt_gbase;
};

// (A'2) one CTA per tile of the round (grid-stride); SHARE = the share axis the sign kernel ran with
template <int TS, int SHARE>
__device__ __forceinline__ void tile_classify_body(const GridDesc& g, const unsigned int* __restrict__ active_list,
                                                   const unsigned int work_begin_arg, const unsigned int work_end_arg,
                                                   const TileLists L, Counters* __restrict__ ctr,
                                                   const unsigned char* __restrict__ scratch) {
    constexpr int TP = SIGN_TP;
    constexpr int THREADS = CLASSIFY_THREADS, NW = THREADS / 32;
    typedef SignAxes<SHARE> AX;
    constexpr int AS = AX::AS, AU = AX::AU, AW = AX::AW;
    __shared__ ClassifySmem sm;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned int subs3 = (unsigned int)(g.sub * g.sub * g.sub);
    unsigned int work_begin, work_end;
    work_range(ctr, subs3, work_begin_arg, work_end_arg, &work_begin, &work_end);
    const unsigned int n_tiles = work_end > work_begin ? work_end - work_begin : 0u;
    for (unsigned int tw = blockIdx.x; tw < n_tiles; tw += gridDim.x) {
        const unsigned int w = work_begin + tw;
        const unsigned int bid = active_list[w / subs3];
        const int sidx = (int)(w % subs3);
        TileGeom t;
        const bool valid = tile_geom<TS>(g, bid, sidx, t);
        if (!valid || t.px < 2 || t.py < 2 || t.pz < 2) continue;        // (CTA uniform) thin tile: no cube, not evaluated
        const int ext[3] = {t.px, t.py, t.pz};
        const int pu = ext[AU], ps = ext[AS], pw = ext[AW];
        // the mask block comes in with ONE coalesced pass (616 x 16 B) instead of five dependent, scattered loads per thread
        {
            const uint4* src = reinterpret_cast<const uint4*>(scratch + (size_t)tw * SIGN_STRIDE);
            for (int i = tid; i < SIGN_STRIDE / 16; i += THREADS) sm.block[i] = src[i];
        }
        __syncthreads();
        const uint2* const masks = reinterpret_cast<const uint2*>(sm.block);
        const unsigned char* const hi = reinterpret_cast<const unsigned char*>(masks + SIGN_LINES);
        const bool has_hi = (pu == TP);
        const unsigned int cumask = (pu - 1 >= 32) ? 0xffffffffu : ((1u << (pu - 1)) - 1u);
        const unsigned int umask = pu >= 32 ? 0xffffffffu : ((1u << pu) - 1u);       // valid samples of a U-line
        // (1) one thread per (cs, cw) column of cubes: 32 cubes along U classified with word operations.
        //     A cube is a candidate when its corners are not all > 0 and not all <= 0, or one is uncertain.
        for (int col = tid; col < TS * TS; col += THREADS) {
            const int cs = col >> 5, cw = col & 31;
            unsigned int W = 0u;
            if (cs < ps - 1 && cw < pw - 1) {
                const int l00 = cs * TP + cw;
                const int ls[4] = {l00, l00 + 1, l00 + TP, l00 + TP + 1};
                unsigned long long allp = ~0ull, anyp = 0ull, anyu = 0ull;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const unsigned int h8 = has_hi ? (unsigned int)hi[ls[j]] : 0u;
                    const uint2 mk = masks[ls[j]];
                    const unsigned long long p = (unsigned long long)(mk.x & umask) | ((unsigned long long)(h8 & 1u) << 32);
                    const unsigned long long u = (unsigned long long)(mk.y & umask) | ((unsigned long long)((h8 >> 1) & 1u) << 32);
                    allp &= p & (p >> 1); anyp |= p | (p >> 1); anyu |= u | (u >> 1);
                }
                W = (unsigned int)((anyp & ~allp) | anyu) & cumask;
            }
            sm.cand[col] = W;
        }
        __syncthreads();
        // (2) one thread per line: queue every corner of a candidate cube (and every uncertain sample)
        for (int ln = tid; ln < SIGN_LINES; ln += THREADS) {
            const int is = ln / TP, iw = ln - is * TP;
            unsigned long long ref = 0ull;
            if (is < ps && iw < pw) {
                unsigned int Wl = 0u;
                if (is < TS && iw < TS) Wl |= sm.cand[(is << 5) | iw];
                if (is < TS && iw > 0) Wl |= sm.cand[(is << 5) | (iw - 1)];
                if (is > 0 && iw < TS) Wl |= sm.cand[((is - 1) << 5) | iw];
                if (is > 0 && iw > 0) Wl |= sm.cand[((is - 1) << 5) | (iw - 1)];
                const unsigned int h8 = has_hi ? (unsigned int)hi[ln] : 0u;
                ref = (unsigned long long)Wl | ((unsigned long long)Wl << 1) | (unsigned long long)(masks[ln].y & umask) |
                      ((unsigned long long)((h8 >> 1) & 1u) << 32);
            }
            sm.refm[ln] = ref;
        }
        __syncthreads();
        // (3) compaction: one prefix sum over the per-thread popcounts (cubes in the low, samples in the
        //     high half word), one slot + one pair of list ranges per tile, every thread writes its entries
        unsigned int cnt = 0;
        for (int wi = tid; wi < TS * TS; wi += THREADS) cnt += (unsigned)__popc(sm.cand[wi]);
        for (int wi = tid; wi < SIGN_LINES; wi += THREADS) cnt += ((unsigned)__popcll(sm.refm[wi])) << 16;
        unsigned int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned int nb = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += nb;
        }
        if (lane == 31) sm.warp_sums[warp] = incl;
        __syncthreads();
        unsigned int warp_off = 0, total = 0;
#pragma unroll
        for (int wi = 0; wi < NW; ++wi) {
            const unsigned int sv = sm.warp_sums[wi];
            if (wi < warp) warp_off += sv;
            total += sv;
        }
        if (tid == 0) {
            const unsigned int nc = total & 0xffffu, np = total >> 16;
            unsigned int slot = 0;
            if (nc | np) {
                slot = atomicAdd(&ctr->n_slots, 1u);
                if (slot < L.slot_cap) L.tile_meta[slot] = make_tile_meta(t, bid);
            }
            sm.slot = slot;
            sm.cube_gbase = nc ? atomicAdd(&ctr->n_cubes, nc) : 0u;
            sm.pt_gbase = np ? atomicAdd(&ctr->n_pts, np) : 0u;
            atomicAdd(&ctr->pts_f32, (unsigned long long)(t.px * t.py * t.pz));
        }
        __syncthreads();
        const unsigned int slot = sm.slot;
        const unsigned int excl = warp_off + incl - cnt;
        unsigned int cat = sm.cube_gbase + (excl & 0xffffu);
        unsigned int pat = sm.pt_gbase + (excl >> 16);
        if (total != 0u && slot < L.slot_cap) {
            for (int wi = tid; wi < TS * TS; wi += THREADS) {
                unsigned int word = sm.cand[wi];
                int cc[3];
                cc[AS] = wi >> 5; cc[AW] = wi & 31;
                while (word) {
                    cc[AU] = __ffs(word) - 1;
                    word &= word - 1u;
                    // cube id with fixed strides: cx << 10 | cy << 5 | cz
                    if (cat < L.cube_cap) L.cubes[cat] = make_uint2(slot, (unsigned int)((cc[0] << 10) | (cc[1] << 5) | cc[2]));
                    ++cat;
                }
            }
            for (int wi = tid; wi < SIGN_LINES; wi += THREADS) {
                unsigned long long word = sm.refm[wi];
                int ii[3];
                ii[AS] = wi / TP; ii[AW] = wi - ii[AS] * TP;
                while (word) {
                    ii[AU] = __ffsll((long long)word) - 1;
                    word &= word - 1ull;
                    if (pat < L.pt_cap) L.pts[pat] = make_uint2(slot, (unsigned int)((ii[0] * t.py + ii[1]) * t.pz + ii[2]));
                    ++pat;
                }
            }
        }
        __syncthreads();                                      // the next tile overwrites cand / refm
    }
}

// (B) float64 re-evaluation of the queued samples, flat over the whole list.
template <int TS, class EV>
__device__ __forceinline__ void refine_points_body(const EV& ev, const GridDesc& g, const TileLists L,
                                                   Counters* __restrict__ ctr) {
    const unsigned int n = min(ctr->n_pts, L.pt_cap);
    for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint2 e = L.pts[i];
        const TileMeta m = load_tile_meta(L.tile_meta + e.x);
        const unsigned int py = (m.ext >> 8) & 0xffu, pz = (m.ext >> 16) & 0xffu;
        const unsigned int pyz = py * pz;
        const unsigned int ix = e.y / pyz, rem = e.y - ix * pyz, iy = rem / pz, iz = rem - iy * pz;
        double x[1] = {g.xs[m.bx0 + m.ox + (int)ix]}, y[1] = {g.ys[m.by0 + m.oy + (int)iy]},
               z[1] = {g.zs[m.bz0 + m.oz + (int)iz]}, d[1];
        ev.template run<double, 1, false>(x, y, z, d);
        L.vol[(size_t)e.x * VOL_STRIDE + e.y] = (float)d[0];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&ctr->pts_f64, (unsigned long long)n);
}

// (C) Lewiner marching cubes over the candidate list, two phases per CTA iteration of 256 cubes:
//   1. one thread per CUBE: gather the 8 corner values (float64-confirmed by B), cube index, tiling
//      selection (decision tables in shared memory; test_face / test_interior in non-contracted float64),
//      triangle count; a CTA prefix sum reserves the output range with one global atomic;
//   2. one thread per TRIANGLE (its cube looked up in a small owner map): its three edge vertices
//      (inverse-|value| weights, float64, rounded like the C restatement), world mapping (core.py:58-60),
//      staged per warp in shared memory and written out as one contiguous, fully coalesced run -- to
//      local HBM, or to the root GPU over NVLink in the multi-GPU case.
// Splitting by triangle balances the lanes (a cube yields 0..12 triangles) and keeps the corner values
// in shared memory, where the data-dependent corner indices of the tables cost one load instead of a
// register-select chain or a local-memory array.
#define EXTRACT_THREADS 256
#ifndef EXTRACT_BLOCKS_PER_SM
#define EXTRACT_BLOCKS_PER_SM 4
#endif

struct ExtractSmem {
    McTablesSmem mc;
    float cv[8][EXTRACT_THREADS];                   // corner values of this iteration's cubes
    int org[6][EXTRACT_THREADS];                    // batch origin (bx0, by0, bz0), Lewiner cube position (lx, ly, lz)
    unsigned int pre[EXTRACT_THREADS];              // exclusive prefix of the triangle counts
    unsigned char owner[EXTRACT_THREADS * 12];      // triangle -> cube of this iteration (a cube yields <= 12 triangles)
    unsigned short toff[EXTRACT_THREADS];           // offset of the cube's tiling in tile_data
    float stage[EXTRACT_THREADS / 32][32 * 9];
    unsigned int warp_sums[EXTRACT_THREADS / 32];
    unsigned long long tri_base;
    uint64_t tables_bar;
    unsigned int last_cta;
};

// corner k of the thread's cube: cvp[k * EXTRACT_THREADS]
__device__ __forceinline__ double cvd(const float* cvp, int k) { return (double)cvp[k * EXTRACT_THREADS]; }

__device__ __forceinline__ bool mc_test_face_s(const float* cvp, int face) {
    const int af = face < 0 ? -face : face;
    const double A = cvd(cvp, c_face_corners[af][0]), B = cvd(cvp, c_face_corners[af][1]);
    const double C = cvd(cvp, c_face_corners[af][2]), D = cvd(cvp, c_face_corners[af][3]);
    const double acbd = xsub(xmul(A, C), xmul(B, D));
    if (acbd > -(double)FLT_EPSILON && acbd < (double)FLT_EPSILON) return face >= 0;
    return xmul(xmul((double)face, A), acbd) >= 0;
}

__device__ __forceinline__ int mc_select_tiling_s(const McTablesSmem* t, int index, const float* cvp) {
    const int nf = t->nface[index];
    const int sgn = (__popc(index) <= 4) ? 1 : -1;
    int mask = 0;
    for (int j = 0; j < nf; ++j)
        if (mc_test_face_s(cvp, sgn * (int)t->faces[index * 6 + j])) mask |= 1 << j;
    const int sub = t->sub_base[index] + mask;
    const int mode = t->sub_interior[sub];
    if (mode < 0) return t->sub_true[sub];
    double v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = cvd(cvp, k);
    return mc_test_interior(v, mode, 7 * sgn) ? t->sub_true[sub] : t->sub_false[sub];
}

// Vertex on edge e (12 = centre) of the cube whose corner values are cvp[k * EXTRACT_THREADS], Lewiner
// cube-local coordinates added to (x,y,z).  With w = 1/(eps+|v|) per corner the coordinate along an edge
// is x + w_far/(w0+w1); across the edge both corners share the coordinate c, where the generic form
// (c*w0 + c*w1)/(w0+w1) gives 0 or exactly 1 whenever w0 + w1 is a positive number -- the remaining inputs (a NaN corner,
// two infinite corners) are handled below.  1/x is the correctly rounded reciprocal (== 1.0/x).
__device__ __forceinline__ void mc_vertex_s(const float* cvp, int e, int x, int y, int z, float* out) {
    const double eps = (double)FLT_EPSILON;
    double fx, fy, fz;
    if (e < 12) {
        const int k0 = edge_corner(e, 0), k1 = edge_corner(e, 1);
        const double w0 = __drcp_rn(xadd(eps, fabs(cvd(cvp, k0))));
        const double w1 = __drcp_rn(xadd(eps, fabs(cvd(cvp, k1))));
        const double ff = xadd(w0, w1);
        const int dx = corner_x(k0) ^ corner_x(k1), dy = corner_y(k0) ^ corner_y(k1);
        const int far_is_k1 = dx ? corner_x(k1) : (dy ? corner_y(k1) : corner_z(k1));
        const double q = xdiv(far_is_k1 ? w1 : w0, ff);
        fx = dx ? q : (double)corner_x(k0);
        fy = dx ? (double)corner_y(k0) : (dy ? q : (double)corner_y(k0));
        fz = (dx | dy) ? (double)corner_z(k0) : q;
        // a NaN corner (w = NaN) or two infinite ones (w0 + w1 = 0): the generic form is NaN in EVERY coordinate
        // (0 * NaN, 0 / 0), and so is q
        if (!(ff > 0.0)) { fx = q; fy = q; fz = q; }
    } else {
        double sx = 0, sy = 0, sz = 0, ff = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const double w = __drcp_rn(xadd(eps, fabs(cvd(cvp, k))));
            sx = xadd(sx, xmul((double)corner_x(k), w));
            sy = xadd(sy, xmul((double)corner_y(k), w));
            sz = xadd(sz, xmul((double)corner_z(k), w));
            ff = xadd(ff, w);
        }
        fx = xdiv(sx, ff); fy = xdiv(sy, ff); fz = xdiv(sz, ff);
    }
    out[0] = (float)xadd((double)x, fx);
    out[1] = (float)xadd((double)y, fy);
    out[2] = (float)xadd((double)z, fz);
}

// `cursor`: where output ranges are reserved.  Normally &ctr->tri_cursor of this GPU; for the fused
// multi-GPU gather it is a cursor in the ROOT GPU's memory (mapped through CUDA IPC) and `tri_out` is
// the root's buffer: ranges are reserved with one system-scope atomic per CTA iteration over NVLink and
// the triangles leave this GPU as fully coalesced peer stores straight from the staging area -- the
// transfer overlaps the extraction math, no collective follows.
// `stats_flags` != nullptr (the last round of a call): the CTA that finishes last also counts the empty /
// nonempty batches (core.py:135-141) -- "nonempty" = a non-skipped batch that produced at least one triangle.
template <int TS>
__device__ __forceinline__ void extract_body(const GridDesc& g, const TileLists L, Counters* __restrict__ ctr,
                                             unsigned int* __restrict__ batch_tris, float* __restrict__ tri_out,
                                             const unsigned long long tri_capacity,
                                             unsigned long long* __restrict__ ext_cursor,
                                             const McTablesSmem* __restrict__ tables,
                                             const unsigned char* __restrict__ stats_flags,
                                             RoundProgress* __restrict__ progress,
                                             unsigned long long* __restrict__ done_signal) {
    __shared__ __align__(128) ExtractSmem sm;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    load_mc_tables_async(&sm.mc, tables, &sm.tables_bar, tid);
    const unsigned int n = min(ctr->n_cubes, L.cube_cap);
    const unsigned int n_up = (n + EXTRACT_THREADS - 1) / EXTRACT_THREADS * EXTRACT_THREADS;
    unsigned int my_active = 0;
    const float* cvp = &sm.cv[0][tid];
    bool tables_ready = false;
    __syncthreads();                                          // the barrier object is initialised
    for (unsigned int i0 = blockIdx.x * EXTRACT_THREADS; i0 < n_up; i0 += gridDim.x * EXTRACT_THREADS) {
        // ------------------------------------------------------------ phase 1: one thread per cube
        const unsigned int i = i0 + tid;
        unsigned int nt = 0;
        if (i < n) {
            const uint2 e = L.cubes[i];
            const TileMeta m = load_tile_meta(L.tile_meta + e.x);
            const int py = (int)((m.ext >> 8) & 0xffu), pz = (int)((m.ext >> 16) & 0xffu), pyz = py * pz;
            const int cx = (int)(e.y >> 10), cy = (int)((e.y >> 5) & 31u), cz = (int)(e.y & 31u);
            const float* p = L.vol + (size_t)e.x * VOL_STRIDE + (size_t)((cx * py + cy) * pz + cz);
            const float f0 = p[0], f1 = p[1], f2 = p[pz + 1], f3 = p[pz];
            const float f4 = p[pyz], f5 = p[pyz + 1], f6 = p[pyz + pz + 1], f7 = p[pyz + pz];
            sm.cv[0][tid] = f0; sm.cv[1][tid] = f1; sm.cv[2][tid] = f2; sm.cv[3][tid] = f3;
            sm.cv[4][tid] = f4; sm.cv[5][tid] = f5; sm.cv[6][tid] = f6; sm.cv[7][tid] = f7;
            const int index = (f0 > 0.0f ? 1 : 0) | (f1 > 0.0f ? 2 : 0) | (f2 > 0.0f ? 4 : 0) | (f3 > 0.0f ? 8 : 0) |
                              (f4 > 0.0f ? 16 : 0) | (f5 > 0.0f ? 32 : 0) | (f6 > 0.0f ? 64 : 0) | (f7 > 0.0f ? 128 : 0);
            if (!tables_ready) { mbar_wait(&sm.tables_bar, 0u); tables_ready = true; }      // (the gathers above overlapped the copy)
            if (index != 0 && index != 255) {
                ++my_active;
                const int tiling = mc_select_tiling_s(&sm.mc, index, cvp);
                if (tiling >= 0) {
                    nt = sm.mc.tile_ntri[tiling];
                    sm.toff[tid] = sm.mc.tile_off[tiling];
                }
            }
            if (nt != 0) {
                batch_tris[m.bid] = 1u;                 // "this batch produced triangles" (core.py:135-141); all writers agree
                sm.org[0][tid] = m.bx0; sm.org[1][tid] = m.by0; sm.org[2][tid] = m.bz0;
                sm.org[3][tid] = m.oz + cz; sm.org[4][tid] = m.oy + cy; sm.org[5][tid] = m.ox + cx;   // Lewiner x,y,z (batch-local)
            }
        }
        // CTA-wide exclusive scan of nt
        unsigned int incl = nt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned int nb = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += nb;
        }
        if (lane == 31) sm.warp_sums[warp] = incl;
        __syncthreads();
        unsigned int warp_off = 0, total = 0;
#pragma unroll
        for (int wi = 0; wi < EXTRACT_THREADS / 32; ++wi) {
            const unsigned int sv = sm.warp_sums[wi];
            if (wi < warp) warp_off += sv;
            total += sv;
        }
        {
            const unsigned int first = warp_off + incl - nt;
            sm.pre[tid] = first;
            for (unsigned int k = 0; k < nt; ++k) sm.owner[first + k] = (unsigned char)tid;
        }
        if (tid == 0) {
            unsigned long long base = 0ull;
            if (total) {
                if (ext_cursor) {
                    base = atomicAdd_system(ext_cursor, (unsigned long long)total);
                    atomicAdd(&ctr->tri_cursor, (unsigned long long)total);          // local count (statistics)
                } else {
                    base = atomicAdd(&ctr->tri_cursor, (unsigned long long)total);
                }
            }
            sm.tri_base = base;
        }
        __syncthreads();
        // ------------------------------------------------------------ phase 2: one thread per triangle
        if (!tables_ready) { mbar_wait(&sm.tables_bar, 0u); tables_ready = true; }
        const unsigned long long gbase = sm.tri_base;
        const bool writable = tri_out != nullptr && gbase + total <= tri_capacity;      // else: count only / overflow
        if (writable) {
            for (unsigned int j0 = 0; j0 < total; j0 += EXTRACT_THREADS) {
                const unsigned int j = j0 + tid;
                if (j < total) {
                    const int c = (int)sm.owner[j];
                    const unsigned int tri = j - sm.pre[c];
                    const unsigned char* ed = &sm.mc.tile_data[(unsigned int)sm.toff[c] + 3u * tri];
                    const int bx0 = sm.org[0][c], by0 = sm.org[1][c], bz0 = sm.org[2][c];
                    const int lx = sm.org[3][c], ly = sm.org[4][c], lz = sm.org[5][c];
                    // world mapping of this batch (core.py:58-59): scale = X[1]-X[0], offset = X[0] of the slice
                    const double ofx = g.xs[bx0], ofy = g.ys[by0], ofz = g.zs[bz0];
                    const double scx = __dsub_rn(g.xs[min(bx0 + 1, g.nx - 1)], ofx);
                    const double scy = __dsub_rn(g.ys[min(by0 + 1, g.ny - 1)], ofy);
                    const double scz = __dsub_rn(g.zs[min(bz0 + 1, g.nz - 1)], ofz);
                    float* o = &sm.stage[warp][lane * 9];
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        float q[3];
                        mc_vertex_s(&sm.cv[0][c], (int)ed[k], lx, ly, lz, q);
                        // verts flipped to array-axis order (z,y,x)_L, then core.py:58-60; faces flipped (a,b,c) -> (c,b,a)
                        o[(2 - k) * 3 + 0] = (float)__dadd_rn(__dmul_rn((double)q[2], scx), ofx);
                        o[(2 - k) * 3 + 1] = (float)__dadd_rn(__dmul_rn((double)q[1], scy), ofy);
                        o[(2 - k) * 3 + 2] = (float)__dadd_rn(__dmul_rn((double)q[0], scz), ofz);
                    }
                }
                __syncwarp();
                const unsigned int wj0 = j0 + (unsigned int)warp * 32u;
                if (wj0 < total) {
                    const unsigned int nfl = min(32u, total - wj0) * 9u;
                    float* gdst = tri_out + (gbase + wj0) * 9ull;
                    for (unsigned int k = lane; k < nfl; k += 32) gdst[k] = sm.stage[warp][k];
                }
                __syncwarp();
            }
        }
        __syncthreads();                                      // the next iteration overwrites cv / org / pre
    }
    // statistics
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) my_active += __shfl_xor_sync(0xffffffffu, my_active, o);
    if (lane == 0 && my_active) atomicAdd(&ctr->active_cubes, (unsigned long long)my_active);
    if (blockIdx.x == 0 && tid == 0) {
        atomicMax(&ctr->max_cubes, ctr->n_cubes);
        atomicMax(&ctr->max_pts, ctr->n_pts);
    }
    if (!tables_ready) mbar_wait(&sm.tables_bar, 0u);          // never leave with a bulk copy in flight
    if (stats_flags != nullptr || progress != nullptr || done_signal != nullptr) {
        // this CTA's batch_tris flags / triangles are visible device wide -- system wide when they went to a peer GPU
        if (done_signal != nullptr) __threadfence_system(); else __threadfence();
        __syncthreads();
        if (tid == 0) {
            const unsigned int done = atomicAdd(&ctr->extract_done, 1u);
            sm.last_cta = (done == gridDim.x - 1u) ? 1u : 0u;
        }
        __syncthreads();
        if (sm.last_cta && progress != nullptr && tid == 0) {
            // a pipelined call: tell the host (mapped page-locked memory) that this round's triangles are complete
            progress->tri_cursor = *((volatile unsigned long long*)&ctr->tri_cursor);
            progress->max_cubes = *((volatile unsigned int*)&ctr->max_cubes);
            progress->max_pts = *((volatile unsigned int*)&ctr->max_pts);
            __threadfence_system();
            *((volatile unsigned int*)&progress->done) = 1u;
        }
        if (sm.last_cta && done_signal != nullptr && tid == 0) {
            // fused multi-GPU gather: every triangle of this rank is in the root's buffer -- tell the root
            __threadfence_system();
            atomicAdd_system(done_signal, 1ull);
        }
        if (sm.last_cta && stats_flags != nullptr) {
            __threadfence();
            const int total_b = g.ub1 - g.ub0;
            unsigned int ne = 0, nn = 0;
            for (int t = tid; t < total_b; t += EXTRACT_THREADS) {
                int bx, by, bz;
                const unsigned int bid = window_batch(g, t, &bx, &by, &bz);
                if (stats_flags[bid]) continue;
                if (__ldcg(batch_tris + bid)) ++nn; else ++ne;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { ne += __shfl_xor_sync(0xffffffffu, ne, o); nn += __shfl_xor_sync(0xffffffffu, nn, o); }
            if (lane == 0) { if (ne) atomicAdd(&ctr->empty, ne); if (nn) atomicAdd(&ctr->nonempty, nn); }
        }
    }
}

// ---------------------------------------------------------------------------------------
// fused multi-GPU gather: flags in the ROOT GPU's memory instead of a collective per step.
//   ready[p]   sequence number of the step that may use buffer / cursor p (root: += 2 after it has consumed the buffer)
//   done[p]    ranks whose triangles of the current step are all in buffer p (every rank's extraction tail: += 1)
// A rank's step s (p = s & 1) waits for ready[p] == s before its extraction touches the shared cursor; the root
// waits for done[p] == world before it reads the count.  No NCCL call, no host barrier: a rank only synchronises
// with its own stream.  Spins give up (trap) after ~10 s so that a dead peer cannot hang the box.
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ void spin_until_equal(const unsigned long long* p, unsigned long long want) {
    const long long t0 = clock64();
    while (ld_acquire_sys_u64(p) != want) {
        __nanosleep(256);
        if (clock64() - t0 > 20000000000ll) __trap();          // ~10 s at 2 GHz
    }
}

__device__ __forceinline__ void shared_wait_ready_body(const unsigned long long* ready, unsigned long long* my_seq) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        const unsigned long long s = *my_seq;
        spin_until_equal(ready, s);
        *my_seq = s + 2ull;
    }
}

__device__ __forceinline__ void shared_wait_done_body(const unsigned long long* done, unsigned long long world) {
    if (threadIdx.x == 0 && blockIdx.x == 0) spin_until_equal(done, world);
}

__device__ __forceinline__ void shared_release_body(unsigned long long* cursor, unsigned long long* done, unsigned long long* ready) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        *cursor = 0ull;
        *done = 0ull;
        __threadfence_system();
        atomicAdd_system(ready, 2ull);
    }
}

// skipped / empty / nonempty exactly as core.generate counts them (core.py:135-141): "nonempty" = a
// non-skipped batch that produced at least one triangle
__device__ __forceinline__ void batch_stats_body(const GridDesc& g, const unsigned char* __restrict__ flags,
                                                 const unsigned int* __restrict__ batch_tris, Counters* __restrict__ ctr) {
    const int total = g.ub1 - g.ub0;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    int bx, by, bz;
    const unsigned int bid = window_batch(g, t, &bx, &by, &bz);
    if (flags[bid]) return;
    if (batch_tris[bid]) atomicAdd(&ctr->nonempty, 1u); else atomicAdd(&ctr->empty, 1u);
}

// ---------------------------------------------------------------------------------------
// binary-STL records (sdf/stl.py:4-24): per triangle {normal f32x3, vertices f32x9, attr u16 = 0},
// 50 bytes, the normal computed in float32 operation by operation like numpy does
// (np.cross, np.linalg.norm, divide) -- no FMA contraction, so the bytes equal the reference writer's.
#define STL_THREADS 128

__device__ __forceinline__ void stl_records_body(const float* __restrict__ tris, long long n, unsigned char* __restrict__ out) {
    __shared__ __align__(16) unsigned short rec[STL_THREADS * 25];
    const long long t0 = (long long)blockIdx.x * STL_THREADS;
    const long long t = t0 + threadIdx.x;
    if (t < n) {
        float p[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) p[k] = tris[t * 9 + k];
        const float ax = __fsub_rn(p[3], p[0]), ay = __fsub_rn(p[4], p[1]), az = __fsub_rn(p[5], p[2]);
        const float bx = __fsub_rn(p[6], p[0]), by = __fsub_rn(p[7], p[1]), bz = __fsub_rn(p[8], p[2]);
        float nx = __fsub_rn(__fmul_rn(ay, bz), __fmul_rn(az, by));
        float ny = __fsub_rn(__fmul_rn(az, bx), __fmul_rn(ax, bz));
        float nz = __fsub_rn(__fmul_rn(ax, by), __fmul_rn(ay, bx));
        const float len = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(nx, nx), __fmul_rn(ny, ny)), __fmul_rn(nz, nz)));
        nx = __fdiv_rn(nx, len); ny = __fdiv_rn(ny, len); nz = __fdiv_rn(nz, len);
        unsigned short* r = &rec[threadIdx.x * 25];
        const float v[12] = {nx, ny, nz, p[0], p[1], p[2], p[3], p[4], p[5], p[6], p[7], p[8]};
#pragma unroll
        for (int k = 0; k < 12; ++k) {
            const unsigned int b = __float_as_uint(v[k]);
            r[2 * k] = (unsigned short)(b & 0xffffu);
            r[2 * k + 1] = (unsigned short)(b >> 16);
        }
        r[24] = 0;
    }
    __syncthreads();
    const long long cnt = min((long long)STL_THREADS, n - t0);
    if (cnt <= 0) return;
    const int n16 = (int)cnt * 25;                       // 16-bit units of this block, contiguous in the file
    unsigned short* dst = reinterpret_cast<unsigned short*>(out + t0 * 50);
    for (int i = threadIdx.x; i < n16; i += STL_THREADS) dst[i] = rec[i];
}

// ---------------------------------------------------------------------------------------
// streamed point evaluation (SDF3.__call__): float3 AoS in HBM -> smem by TMA bulk copy

#define STREAM_THREADS 256
#define STREAM_PTS_PER_THREAD 4
#define STREAM_TILE (STREAM_THREADS * STREAM_PTS_PER_THREAD)      // 1024 points = 12288 B
#define STREAM_STAGES 3

// Persistent CTAs; each owns tiles blockIdx.x, blockIdx.x + gridDim.x, ...; a 3-stage ring of
// 12 KB point tiles is kept in flight by one elected thread issuing cp.async.bulk.
template <bool FAST, class EV>
__device__ __forceinline__ void eval_stream_f32_body(const EV& img, const float* __restrict__ xyz, long long n,
                                                     float* __restrict__ out) {
    __shared__ __align__(128) float stage[STREAM_STAGES][STREAM_TILE * 3];
    __shared__ __align__(8) uint64_t full[STREAM_STAGES];
    const int tid = threadIdx.x;
    const long long n_tiles = (n + STREAM_TILE - 1) / STREAM_TILE;
    const bool aligned = ((reinterpret_cast<uintptr_t>(xyz) & 15u) == 0);
    if (tid == 0) {
        for (int s = 0; s < STREAM_STAGES; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto tile_bytes = [&](long long t) -> unsigned {
        const long long pts = min((long long)STREAM_TILE, n - t * STREAM_TILE);
        return (unsigned)(pts * 12);
    };
    auto issue = [&](long long t, int s) {
        const unsigned bytes = tile_bytes(t);
        if (aligned && (bytes & 15u) == 0) {
            mbar_expect_tx(&full[s], bytes);
            tma_bulk_g2s(&stage[s][0], xyz + t * (long long)STREAM_TILE * 3, bytes, &full[s]);
        }
    };
    // prologue
    if (tid == 0) {
        for (int s = 0; s < STREAM_STAGES; ++s) {
            const long long t = blockIdx.x + (long long)s * gridDim.x;
            if (t < n_tiles) issue(t, s);
        }
    }
    int it = 0;
    for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
        const int s = it % STREAM_STAGES;
        const unsigned parity = (unsigned)((it / STREAM_STAGES) & 1);
        const unsigned bytes = tile_bytes(t);
        const bool via_tma = aligned && (bytes & 15u) == 0;
        const long long p0 = t * (long long)STREAM_TILE;
        const int pts = (int)min((long long)STREAM_TILE, n - p0);
        if (via_tma) {
            mbar_wait(&full[s], parity);
        } else {   // ragged / unaligned tail: plain coalesced loads into the same stage
            for (int i = tid; i < pts * 3; i += STREAM_THREADS) stage[s][i] = xyz[p0 * 3 + i];
            __syncthreads();
        }
        float x[STREAM_PTS_PER_THREAD], y[STREAM_PTS_PER_THREAD], z[STREAM_PTS_PER_THREAD], d[STREAM_PTS_PER_THREAD];
#pragma unroll
        for (int k = 0; k < STREAM_PTS_PER_THREAD; ++k) {
            int i = k * STREAM_THREADS + tid;
            if (i >= pts) i = pts - 1;
            x[k] = stage[s][3 * i + 0]; y[k] = stage[s][3 * i + 1]; z[k] = stage[s][3 * i + 2];
        }
        img.template run<float, STREAM_PTS_PER_THREAD, FAST>(x, y, z, d);
#pragma unroll
        for (int k = 0; k < STREAM_PTS_PER_THREAD; ++k) {
            const int i = k * STREAM_THREADS + tid;
            if (i < pts) out[p0 + i] = d[k];
        }
        __syncthreads();                      // everyone is done reading stage s
        if (tid == 0) {
            const long long tn = t + (long long)STREAM_STAGES * gridDim.x;
            if (tn < n_tiles) issue(tn, s);
        }
    }
}

template <class EV>
__device__ __forceinline__ void eval_stream_f64_body(const EV& img, const double* __restrict__ xyz, long long n,
                                                     double* __restrict__ out) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        double x[1] = {xyz[3 * i]}, y[1] = {xyz[3 * i + 1]}, z[1] = {xyz[3 * i + 2]}, d[1];
        img.template run<double, 1, false>(x, y, z, d);
        out[i] = d[0];
    }
}

