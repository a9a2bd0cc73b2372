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

#define MC_TABLE_QUAL __device__ const
#include "mc_tables.inc"

// ---------------------------------------------------------------------------------------
// device-side job description

struct GridDesc {
    const double* xs;
    const double* ys;
    const double* zs;
    int nx, ny, nz;        // lattice points per axis
    int bs;                // cubes per batch side (core.py BATCH_SIZE)
    int nbx, nby, nbz;     // batches per axis = ceil(n / bs)   (core.py:115-117)
    int zb0, zb1;          // Z batch layers handled by this call
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
    unsigned int pad_;
};

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

// skimage test_face (asymptotic decider with the FLT_EPSILON dead band)
__device__ __forceinline__ bool mc_test_face(const double* v, int face) {
    const int af = face < 0 ? -face : face;
    const double A = v[c_face_corners[af][0]], B = v[c_face_corners[af][1]];
    const double C = v[c_face_corners[af][2]], D = v[c_face_corners[af][3]];
    const double acbd = xsub(xmul(A, C), xmul(B, D));
    if (acbd > -(double)FLT_EPSILON && acbd < (double)FLT_EPSILON) return face >= 0;
    return xmul(xmul((double)face, A), acbd) >= 0;
}

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
// shared-memory copies of the decision tables ("edge table in shared memory")

struct McTablesSmem {
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

__device__ __forceinline__ void load_mc_tables(McTablesSmem* t, int tid, int nthreads) {
    for (int i = tid; i < 256; i += nthreads) { t->sub_base[i] = MC_SUB_BASE[i]; t->nface[i] = MC_NFACE[i]; }
    for (int i = tid; i < 256 * 6; i += nthreads) t->faces[i] = MC_FACES[i];
    for (int i = tid; i < MC_TILE_OFF_LEN; i += nthreads) { t->tile_off[i] = MC_TILE_OFF[i]; t->tile_ntri[i] = MC_TILE_NTRI[i]; }
    for (int i = tid; i < MC_SUB_TRUE_LEN; i += nthreads) {
        t->sub_true[i] = MC_SUB_TRUE[i]; t->sub_false[i] = MC_SUB_FALSE[i]; t->sub_interior[i] = MC_SUB_INTERIOR[i];
    }
    for (int i = tid; i < MC_TILE_DATA_LEN; i += nthreads) t->tile_data[i] = MC_TILE_DATA[i];
}

__device__ __forceinline__ int mc_select_tiling(const McTablesSmem* t, int index, const double* v) {
    const int nf = t->nface[index];
    const int sgn = (__popc(index) <= 4) ? 1 : -1;
    int mask = 0;
    for (int j = 0; j < nf; ++j)
        if (mc_test_face(v, sgn * (int)t->faces[index * 6 + j])) mask |= 1 << j;
    const int sub = t->sub_base[index] + mask;
    const int mode = t->sub_interior[sub];
    if (mode < 0) return t->sub_true[sub];
    return mc_test_interior(v, mode, 7 * sgn) ? t->sub_true[sub] : t->sub_false[sub];
}

// vertex on edge e (12 = centre), Lewiner cube-local coordinates added to (x,y,z)
__device__ __forceinline__ void mc_edge_vertex(const double* v, int e, int x, int y, int z, float* out) {
    double fx = 0, fy = 0, fz = 0, ff = 0;
    const double eps = (double)FLT_EPSILON;
    if (e == 12) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const double w = xdiv(1.0, xadd(eps, fabs(v[k])));
            fx = xadd(fx, xmul((double)corner_x(k), w));
            fy = xadd(fy, xmul((double)corner_y(k), w));
            fz = xadd(fz, xmul((double)corner_z(k), w));
            ff = xadd(ff, w);
        }
    } else {
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            const int k = edge_corner(e, s);
            double vk = v[0];
#pragma unroll
            for (int j = 1; j < 8; ++j) vk = (k == j) ? v[j] : vk;      // keeps v[] in registers
            const double w = xdiv(1.0, xadd(eps, fabs(vk)));
            fx = xadd(fx, xmul((double)corner_x(k), w));
            fy = xadd(fy, xmul((double)corner_y(k), w));
            fz = xadd(fz, xmul((double)corner_z(k), w));
            ff = xadd(ff, w);
        }
    }
    out[0] = (float)xadd((double)x, xdiv(fx, ff));
    out[1] = (float)xadd((double)y, xdiv(fy, ff));
    out[2] = (float)xadd((double)z, xdiv(fz, ff));
}

// ---------------------------------------------------------------------------------------
// skip pre-pass (core.py:28-43), one thread per batch, float64

template <class EV>
__device__ __forceinline__ double eval1_f64(const EV& ev, double px, double py, double pz) {
    double x[1] = {px}, y[1] = {py}, z[1] = {pz}, d[1];
    ev.template run<double, 1, false>(x, y, z, d);
    return d[0];
}

template <class EV>
__device__ __forceinline__ void skip_prepass_body(const EV& img, const GridDesc& g, int sparse,
                                                  unsigned char* __restrict__ flags,
                                                  unsigned int* __restrict__ active_list,
                                                  Counters* __restrict__ ctr) {
    const int nzl = g.zb1 - g.zb0;
    const int total = g.nbx * g.nby * nzl;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int bzl = t % nzl, by = (t / nzl) % g.nby, bx = t / (nzl * g.nby);
    const int bz = g.zb0 + bzl;
    const unsigned int bid = (unsigned int)((bx * g.nby + by) * g.nbz + bz);
    bool skipped = false;
    if (sparse) {
        const int ix0 = bx * g.bs, iy0 = by * g.bs, iz0 = bz * g.bs;
        const int ix1 = min(ix0 + g.bs, g.nx - 1), iy1 = min(iy0 + g.bs, g.ny - 1), iz1 = min(iz0 + g.bs, g.nz - 1);
        const double x0 = g.xs[ix0], x1 = g.xs[ix1], y0 = g.ys[iy0], y1 = g.ys[iy1], z0 = g.zs[iz0], z1 = g.zs[iz1];
        const double x = __ddiv_rn(__dadd_rn(x0, x1), 2.0), y = __ddiv_rn(__dadd_rn(y0, y1), 2.0),
                     z = __ddiv_rn(__dadd_rn(z0, z1), 2.0);
        const double r = fabs(eval1_f64(img, x, y, z));
        const double ex = __dsub_rn(x, x0), ey = __dsub_rn(y, y0), ez = __dsub_rn(z, z0);
        const double dd = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(ex, ex), __dmul_rn(ey, ey)), __dmul_rn(ez, ez)));
        if (!(r <= dd)) {
            // corners in itertools.product((x0,x1),(y0,y1),(z0,z1)) order; values[0] decides the sign
            const double v0 = eval1_f64(img, x0, y0, z0);
            const bool pos = v0 > 0;
            bool same = pos ? (v0 > 0) : (v0 < 0);
#pragma unroll 1
            for (int k = 1; k < 8 && same; ++k) {
                const double v = eval1_f64(img, (k & 4) ? x1 : x0, (k & 2) ? y1 : y0, (k & 1) ? z1 : z0);
                same = pos ? (v > 0) : (v < 0);
            }
            skipped = same;
        }
    }
    flags[bid] = skipped ? 1 : 0;
    if (!skipped) {
        const unsigned int slot = atomicAdd(&ctr->n_active, 1u);
        active_list[slot] = bid;
    } else {
        atomicAdd(&ctr->skipped, 1u);
    }
}

// ---------------------------------------------------------------------------------------
// fused tile kernel

#define FUSED_THREADS 512
#define REFINE_CAP 4096
#define CHUNK_CUBES 4096
#define CORNER_ROUND 1024      // (cube, corner) pairs examined per refinement round of PREC_EXACT

template <int TS>
struct FusedSmem {
    static constexpr int TP = TS + 1;
    static constexpr int NPTS = TP * TP * TP;
    float tile[(NPTS + 3) & ~3];
    McTablesSmem mc;
    double axd[3][TP];
    float axf[3][TP];
    // phase-aliased scratch: refine list during evaluation, cube lists during extraction
    union {
        unsigned short refine[REFINE_CAP];
        struct {
            unsigned short cube[CHUNK_CUBES];
            short tiling[CHUNK_CUBES];
            unsigned short scan[CHUNK_CUBES];
        } mcw;
    } u;
    unsigned int refined_bits[(NPTS + 31) / 32];     // samples whose tile value already is float64-derived
    unsigned short corner_list[CORNER_ROUND];
    unsigned int warp_sums[FUSED_THREADS / 32];
    unsigned int n_refine;
    unsigned int n_cubes;
    unsigned int n_tris_chunk;
    unsigned long long tri_base;
    unsigned int work_item;
    int minmax[2];       // order-preserving int encoding of float min / max
    unsigned int tile_tris;
};

__device__ __forceinline__ int float_to_ordered(float f) {
    const int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7fffffff;
}

template <int TS, int PREC, class EV>
__device__ __forceinline__ void fused_tile_body(const EV& img, const GridDesc& g, const float tau,
                                                const unsigned int* __restrict__ active_list,
                                                unsigned int* __restrict__ batch_tris, Counters* __restrict__ ctr,
                                                float* __restrict__ tri_out, const unsigned long long tri_capacity) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    using S = FusedSmem<TS>;
    S& sm = *reinterpret_cast<S*>(smem_raw);
    constexpr int TP = S::TP;
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    constexpr bool REFINE = (PREC == SDFB200_PREC_MIXED || PREC == SDFB200_PREC_EXACT);
    constexpr bool EXACT = (PREC == SDFB200_PREC_EXACT);

    load_mc_tables(&sm.mc, tid, FUSED_THREADS);
    const unsigned int n_active = ctr->n_active;           // written by the pre-pass (previous kernel)
    const unsigned int subs3 = (unsigned int)(g.sub * g.sub * g.sub);
    const unsigned int n_work = n_active * subs3;
    unsigned long long my_f32 = 0, my_f64 = 0, my_cubes = 0;

    for (;;) {
        __syncthreads();
        if (tid == 0) {
            sm.work_item = atomicAdd(&ctr->work_cursor, 1u);
            sm.n_refine = 0;
            sm.minmax[0] = 0x7fffffff;
            sm.minmax[1] = (int)0x80000000;
            sm.tile_tris = 0;
        }
        __syncthreads();
        const unsigned int w = sm.work_item;
        if (w >= n_work) break;
        const unsigned int bid = active_list[w / subs3];
        const int sidx = (int)(w % subs3);
        const int sz = sidx % g.sub, sy = (sidx / g.sub) % g.sub, sx = sidx / (g.sub * g.sub);
        const int bz = (int)(bid % (unsigned)g.nbz), by = (int)((bid / (unsigned)g.nbz) % (unsigned)g.nby),
                  bx = (int)(bid / (unsigned)(g.nbz * g.nby));
        // batch extent in lattice points (core.py:115-117: slices [i, i+s+1) clipped)
        const int bx0 = bx * g.bs, by0 = by * g.bs, bz0 = bz * g.bs;
        const int blx = min(g.bs + 1, g.nx - bx0), bly = min(g.bs + 1, g.ny - by0), blz = min(g.bs + 1, g.nz - bz0);
        // sub-tile extent inside the batch (local point indices)
        const int ox = sx * TS, oy = sy * TS, oz = sz * TS;
        const int px = min(TP, blx - ox), py = min(TP, bly - oy), pz = min(TP, blz - oz);
        if (px <= 0 || py <= 0 || pz <= 0) continue;
        if (sidx != 0 && (px < 2 || py < 2 || pz < 2)) continue;   // a 1-thick remainder holds no cube
        const int npts = px * py * pz;

        if (EXACT) { for (int i = tid; i < (S::NPTS + 31) / 32; i += FUSED_THREADS) sm.refined_bits[i] = 0u; }
        // stage the axis slices
        for (int i = tid; i < px; i += FUSED_THREADS) { const double v = g.xs[bx0 + ox + i]; sm.axd[0][i] = v; sm.axf[0][i] = (float)v; }
        for (int i = tid; i < py; i += FUSED_THREADS) { const double v = g.ys[by0 + oy + i]; sm.axd[1][i] = v; sm.axf[1][i] = (float)v; }
        for (int i = tid; i < pz; i += FUSED_THREADS) { const double v = g.zs[bz0 + oz + i]; sm.axd[2][i] = v; sm.axf[2][i] = (float)v; }
        __syncthreads();

        // ---------------------------------------------------------------- evaluation
        const unsigned int pyz = (unsigned int)(py * pz);
        float vmin = FLT_MAX, vmax = -FLT_MAX;
        if (PREC == SDFB200_PREC_F64) {
            for (int idx = tid; idx < npts; idx += FUSED_THREADS) {
                const int ix = idx / pyz, rem = idx - ix * pyz, iy = rem / pz, iz = rem - iy * pz;
                double x[1] = {sm.axd[0][ix]}, y[1] = {sm.axd[1][iy]}, z[1] = {sm.axd[2][iz]}, d[1];
                img.template run<double, 1, false>(x, y, z, d);
                const float f = (float)d[0];
                sm.tile[idx] = f;
                vmin = fminf(vmin, f); vmax = fmaxf(vmax, f);
            }
            my_f32 += 0;
        } else {
            constexpr int N = 2;
            for (int base = 0; base < npts; base += FUSED_THREADS * N) {
                float x[N], y[N], z[N], d[N];
                int idxs[N];
#pragma unroll
                for (int k = 0; k < N; ++k) {
                    int idx = base + k * FUSED_THREADS + tid;
                    idxs[k] = idx;
                    if (idx >= npts) idx = npts - 1;
                    const int ix = idx / pyz, rem = idx - ix * pyz, iy = rem / pz, iz = rem - iy * pz;
                    x[k] = sm.axf[0][ix]; y[k] = sm.axf[1][iy]; z[k] = sm.axf[2][iz];
                }
                img.template run<float, N, (PREC == SDFB200_PREC_EXACT)>(x, y, z, d);
#pragma unroll
                for (int k = 0; k < N; ++k) {
                    if (idxs[k] < npts) {
                        sm.tile[idxs[k]] = d[k];
                        vmin = fminf(vmin, d[k]); vmax = fmaxf(vmax, d[k]);
                        if (REFINE && fabsf(d[k]) < tau) {
                            const unsigned int slot = atomicAdd(&sm.n_refine, 1u);
                            if (slot < REFINE_CAP) sm.u.refine[slot] = (unsigned short)idxs[k];
                        }
                    }
                }
            }
        }
        if (tid == 0) my_f32 += (unsigned long long)npts;
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

        // ------------------------------------- float64 re-evaluation of sign-critical samples
        if (REFINE) {
            const unsigned int nref = sm.n_refine;
            if (nref > 0) {
                if (nref <= REFINE_CAP) {
                    for (unsigned int r = tid; r < nref; r += FUSED_THREADS) {
                        const int idx = sm.u.refine[r];
                        if (EXACT) atomicOr(&sm.refined_bits[idx >> 5], 1u << (idx & 31));
                        const int ix = idx / pyz, rem = idx - ix * pyz, iy = rem / pz, iz = rem - iy * pz;
                        double x[1] = {sm.axd[0][ix]}, y[1] = {sm.axd[1][iy]}, z[1] = {sm.axd[2][iz]}, d[1];
                        img.template run<double, 1, false>(x, y, z, d);
                        sm.tile[idx] = (float)d[0];
                    }
                } else {   // list overflow: sweep the tile
                    for (int idx = tid; idx < npts; idx += FUSED_THREADS) {
                        if (fabsf(sm.tile[idx]) < tau) {
                            if (EXACT) atomicOr(&sm.refined_bits[idx >> 5], 1u << (idx & 31));
                            const int ix = idx / pyz, rem = idx - ix * pyz, iy = rem / pz, iz = rem - iy * pz;
                            double x[1] = {sm.axd[0][ix]}, y[1] = {sm.axd[1][iy]}, z[1] = {sm.axd[2][iz]}, d[1];
                            img.template run<double, 1, false>(x, y, z, d);
                            sm.tile[idx] = (float)d[0];
                        }
                    }
                }
                if (tid == 0) my_f64 += nref;
                __syncthreads();
            }
        }

        // -------------------------------------------------------------- early out
        // No sample can be > 0 (or none <= 0): every cube index is 0 or 255.  (skimage raises
        // ValueError / RuntimeError here; the reference maps both to an empty batch.)
        {
            const int omin = sm.minmax[0], omax = sm.minmax[1];
            const float fmn = __int_as_float(omin >= 0 ? omin : omin ^ 0x7fffffff);
            const float fmx = __int_as_float(omax >= 0 ? omax : omax ^ 0x7fffffff);
            const float guard = REFINE ? tau : 0.0f;
            if (fmn > guard || fmx < -guard || px < 2 || py < 2 || pz < 2) continue;
        }

        // -------------------------------------------------------- marching cubes
        const int cxn = px - 1, cyn = py - 1, czn = pz - 1;
        const int ncubes = cxn * cyn * czn;
        const unsigned int cyz = (unsigned int)(cyn * czn);
        // world mapping of this batch (core.py:58-59): scale = X[1]-X[0], offset = X[0] of the slice
        const double scx = __dsub_rn(g.xs[min(bx0 + 1, g.nx - 1)], g.xs[bx0]), ofx = g.xs[bx0];
        const double scy = __dsub_rn(g.ys[min(by0 + 1, g.ny - 1)], g.ys[by0]), ofy = g.ys[by0];
        const double scz = __dsub_rn(g.zs[min(bz0 + 1, g.nz - 1)], g.zs[bz0]), ofz = g.zs[bz0];

        for (int c0 = 0; c0 < ncubes; c0 += CHUNK_CUBES) {
            __syncthreads();
            if (tid == 0) { sm.n_cubes = 0; }
            __syncthreads();
            // 1. classify: which cubes cross the level set
            const int cend = min(ncubes, c0 + CHUNK_CUBES);
            for (int cb = c0; cb < cend; cb += FUSED_THREADS) {
                const int c = cb + tid;
                bool act = false;
                if (c < cend) {
                    const int cx = c / cyz, rem = c - cx * cyz, cy = rem / czn, cz = rem - cy * czn;
                    const float* p = &sm.tile[(cx * py + cy) * pz + cz];
                    // Lewiner corner k -> (x_L, y_L, z_L) = (array axis 2, axis 1, axis 0)
                    int index = 0;
                    index |= (p[0] > 0.0f) ? 1 : 0;
                    index |= (p[1] > 0.0f) ? 2 : 0;
                    index |= (p[pz + 1] > 0.0f) ? 4 : 0;
                    index |= (p[pz] > 0.0f) ? 8 : 0;
                    index |= (p[pyz] > 0.0f) ? 16 : 0;
                    index |= (p[pyz + 1] > 0.0f) ? 32 : 0;
                    index |= (p[pyz + pz + 1] > 0.0f) ? 64 : 0;
                    index |= (p[pyz + pz] > 0.0f) ? 128 : 0;
                    act = (index != 0 && index != 255);
                }
                const unsigned int m = __ballot_sync(0xffffffffu, act);
                if (act) {
                    const unsigned int leader = __ffs(m) - 1;
                    unsigned int basep = 0;
                    if ((unsigned)lane == leader) basep = atomicAdd(&sm.n_cubes, (unsigned)__popc(m));
                    basep = __shfl_sync(m, basep, leader);
                    sm.u.mcw.cube[basep + __popc(m & ((1u << lane) - 1u))] = (unsigned short)(c - c0);
                }
            }
            __syncthreads();
            const int nact = (int)sm.n_cubes;
            if (nact == 0) continue;
            if (tid == 0) my_cubes += (unsigned long long)nact;
            if (EXACT) {
                // 1b. every corner of an active cube feeds the output: replace its float32 value by
                // the float64 evaluation (once per sample: refined_bits), CORNER_ROUND pairs a round
                const int npairs = nact * 8;
                for (int p0 = 0; p0 < npairs; p0 += CORNER_ROUND) {
                    if (tid == 0) sm.n_refine = 0;
                    __syncthreads();
                    const int pend = min(npairs, p0 + CORNER_ROUND);
                    for (int q = p0 + tid; q < pend; q += FUSED_THREADS) {
                        const int c = c0 + (int)sm.u.mcw.cube[q >> 3];
                        const int k = q & 7;
                        const int cx = c / cyz, rem = c - cx * cyz, cy = rem / czn, cz = rem - cy * czn;
                        const int idx = ((cx + corner_z(k)) * py + cy + corner_y(k)) * pz + cz + corner_x(k);
                        const unsigned int bit = 1u << (idx & 31);
                        const unsigned int old = atomicOr(&sm.refined_bits[idx >> 5], bit);
                        if (!(old & bit)) sm.corner_list[atomicAdd(&sm.n_refine, 1u)] = (unsigned short)idx;
                    }
                    __syncthreads();
                    const unsigned int nref = sm.n_refine;
                    for (unsigned int r = tid; r < nref; r += FUSED_THREADS) {
                        const int idx = sm.corner_list[r];
                        const int ix = idx / pyz, rem = idx - ix * pyz, iy = rem / pz, iz = rem - iy * pz;
                        double x[1] = {sm.axd[0][ix]}, y[1] = {sm.axd[1][iy]}, z[1] = {sm.axd[2][iz]}, d[1];
                        img.template run<double, 1, false>(x, y, z, d);
                        sm.tile[idx] = (float)d[0];
                    }
                    if (tid == 0) my_f64 += nref;
                    __syncthreads();
                }
            }
            // 2. choose the tiling of every active cube, 3. exclusive scan of triangle counts
            unsigned int running = 0;
            for (int a0 = 0; a0 < nact; a0 += FUSED_THREADS) {
                const int ai = a0 + tid;
                unsigned int nt = 0;
                if (ai < nact) {
                    const int c = c0 + (int)sm.u.mcw.cube[ai];
                    const int cx = c / cyz, rem = c - cx * cyz, cy = rem / czn, cz = rem - cy * czn;
                    const float* p = &sm.tile[(cx * py + cy) * pz + cz];
                    double v[8];
                    v[0] = (double)p[0]; v[1] = (double)p[1]; v[2] = (double)p[pz + 1]; v[3] = (double)p[pz];
                    v[4] = (double)p[pyz]; v[5] = (double)p[pyz + 1]; v[6] = (double)p[pyz + pz + 1]; v[7] = (double)p[pyz + pz];
                    int index = 0;
#pragma unroll
                    for (int k = 0; k < 8; ++k) index |= (v[k] > 0.0) ? (1 << k) : 0;
                    const int tiling = mc_select_tiling(&sm.mc, index, v);
                    sm.u.mcw.tiling[ai] = (short)tiling;
                    nt = tiling >= 0 ? sm.mc.tile_ntri[tiling] : 0;
                }
                // block-wide exclusive scan of nt (warp shuffle + warp sums)
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
                for (int wi = 0; wi < FUSED_THREADS / 32; ++wi) {
                    const unsigned int s = sm.warp_sums[wi];
                    if (wi < warp) warp_off += s;
                    total += s;
                }
                if (ai < nact) sm.u.mcw.scan[ai] = (unsigned short)(running + warp_off + incl - nt);
                running += total;
                __syncthreads();
            }
            // 4. reserve output space: one global atomic per chunk
            if (tid == 0) {
                sm.n_tris_chunk = running;
                sm.tri_base = atomicAdd(&ctr->tri_cursor, (unsigned long long)running);
                sm.tile_tris += running;
            }
            __syncthreads();
            const unsigned int ntris = sm.n_tris_chunk;
            const unsigned long long tbase = sm.tri_base;
            if (tri_out == nullptr || tbase + ntris > tri_capacity) continue;   // count only / overflow
            // 5. one thread per triangle: find its cube (binary search in the scan), emit 9 floats
            for (unsigned int t = tid; t < ntris; t += FUSED_THREADS) {
                int lo = 0, hi = nact - 1;
                while (lo < hi) {
                    const int mid = (lo + hi + 1) >> 1;
                    if (sm.u.mcw.scan[mid] <= t) lo = mid; else hi = mid - 1;
                }
                // skip cubes that emit nothing but share the scan value
                const int ai = lo;
                const int tiling = sm.u.mcw.tiling[ai];
                const int tri = (int)(t - sm.u.mcw.scan[ai]);
                const int c = c0 + (int)sm.u.mcw.cube[ai];
                const int cx = c / cyz, rem = c - cx * cyz, cy = rem / czn, cz = rem - cy * czn;
                const float* p = &sm.tile[(cx * py + cy) * pz + cz];
                double v[8];
                v[0] = (double)p[0]; v[1] = (double)p[1]; v[2] = (double)p[pz + 1]; v[3] = (double)p[pz];
                v[4] = (double)p[pyz]; v[5] = (double)p[pyz + 1]; v[6] = (double)p[pyz + pz + 1]; v[7] = (double)p[pyz + pz];
                const unsigned char* ed = &sm.mc.tile_data[sm.mc.tile_off[tiling] + 3 * tri];
                float* dst = tri_out + (tbase + t) * 9ull;
                // batch-local voxel coordinates (the reference runs MC per batch)
                const int lx = oz + cz, ly = oy + cy, lz = ox + cx;   // Lewiner x,y,z
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    float q[3];
                    mc_edge_vertex(v, ed[k], lx, ly, lz, q);
                    // verts flipped to array-axis order (z,y,x)_L; faces flipped (a,b,c)->(c,b,a)
                    float* o = dst + (2 - k) * 3;
                    o[0] = (float)__dadd_rn(__dmul_rn((double)q[2], scx), ofx);
                    o[1] = (float)__dadd_rn(__dmul_rn((double)q[1], scy), ofy);
                    o[2] = (float)__dadd_rn(__dmul_rn((double)q[0], scz), ofz);
                }
            }
        }
        __syncthreads();
        if (tid == 0 && sm.tile_tris) atomicAdd(&batch_tris[bid], sm.tile_tris);
    }
    if (tid == 0) {
        atomicAdd(&ctr->pts_f32, my_f32);
        atomicAdd(&ctr->pts_f64, my_f64);
        atomicAdd(&ctr->active_cubes, my_cubes);
    }
}

__device__ __forceinline__ void batch_stats_body(const GridDesc& g, const unsigned char* __restrict__ flags,
                                   const unsigned int* __restrict__ batch_tris, Counters* __restrict__ ctr) {
    const int nzl = g.zb1 - g.zb0;
    const int total = g.nbx * g.nby * nzl;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int bzl = t % nzl, by = (t / nzl) % g.nby, bx = t / (nzl * g.nby);
    const unsigned int bid = (unsigned int)((bx * g.nby + by) * g.nbz + g.zb0 + bzl);
    if (flags[bid]) return;
    if (batch_tris[bid]) atomicAdd(&ctr->nonempty, 1u); else atomicAdd(&ctr->empty, 1u);
}

// ---------------------------------------------------------------------------------------
// streamed point evaluation (SDF3.__call__): float3 AoS in HBM -> smem by TMA bulk copy

#define STREAM_THREADS 256
#define STREAM_PTS_PER_THREAD 4
#define STREAM_TILE (STREAM_THREADS * STREAM_PTS_PER_THREAD)      // 1024 points = 12288 B
#define STREAM_STAGES 3

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

