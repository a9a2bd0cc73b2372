// sdf_interp.cuh -- fused SDF expression-tree evaluator (device side).
//
// Replaces the reference's recursive numpy closure tree (sdf/d3.py:24-25 calling
// into sdf/d3.py:92-502 and sdf/dn.py:7-114) with ONE pass of a linear program
// over each sample point.  The program (sdf_b200/ir.py) lives in the kernel's
// parameter constant bank (ProgramImage, __grid_constant__), every instruction is
// warp-uniform, and all per-point state stays in registers except the spill
// levels of the value stack / point slots (per-thread local memory, L1-resident).
//
// Templated on the scalar type: float is the throughput path, double restates the
// reference's float64 arithmetic (used for the skip test, for SDF3.__call__, and
// for the sign-critical re-evaluation of near-zero samples).  N points per thread
// are interpreted together so that instruction fetch/dispatch is amortised and
// the math pipes see N independent dependency chains.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "sdf_opcodes.h"

struct ProgramImage {
    uint32_t n_code;
    uint32_t pad_;
    uint32_t code[SDF_MAX_CODE];
    float cf[SDF_MAX_CONSTS];
    double cd[SDF_MAX_CONSTS];
};

namespace sdfm {
// scalar math overloads ------------------------------------------------------
__device__ __forceinline__ float m_sqrt(float x) { return sqrtf(x); }
__device__ __forceinline__ double m_sqrt(double x) { return sqrt(x); }
__device__ __forceinline__ float m_abs(float x) { return fabsf(x); }
__device__ __forceinline__ double m_abs(double x) { return fabs(x); }
__device__ __forceinline__ float m_min(float a, float b) { return fminf(a, b); }
__device__ __forceinline__ double m_min(double a, double b) { return fmin(a, b); }
__device__ __forceinline__ float m_max(float a, float b) { return fmaxf(a, b); }
__device__ __forceinline__ double m_max(double a, double b) { return fmax(a, b); }
__device__ __forceinline__ void m_sincos(float a, float* s, float* c) { sincosf(a, s, c); }
__device__ __forceinline__ void m_sincos(double a, double* s, double* c) { sincos(a, s, c); }
__device__ __forceinline__ float m_atan2(float y, float x) { return atan2f(y, x); }
__device__ __forceinline__ double m_atan2(double y, double x) { return atan2(y, x); }
__device__ __forceinline__ float m_hypot(float x, float y) { return sqrtf(x * x + y * y); }
__device__ __forceinline__ double m_hypot(double x, double y) { return hypot(x, y); }
__device__ __forceinline__ float m_fmod(float a, float b) { return fmodf(a, b); }
__device__ __forceinline__ double m_fmod(double a, double b) { return fmod(a, b); }
__device__ __forceinline__ float m_rint(float a) { return rintf(a); }
__device__ __forceinline__ double m_rint(double a) { return rint(a); }

template <typename T> __device__ __forceinline__ T m_clip01(T v) { return m_min(m_max(v, T(0)), T(1)); }
template <typename T> __device__ __forceinline__ T m_sign(T v) { return v > T(0) ? T(1) : (v < T(0) ? T(-1) : T(0)); }
template <typename T> __device__ __forceinline__ T m_len2(T a, T b) { return m_sqrt(a * a + b * b); }
template <typename T> __device__ __forceinline__ T m_len3(T a, T b, T c) { return m_sqrt(a * a + b * b + c * c); }
// numpy's floored modulo (np.mod / %), divisor b > 0 or < 0
template <typename T> __device__ __forceinline__ T m_pymod(T a, T b) {
    T m = m_fmod(a, b);
    if (m != T(0)) { if ((b < T(0)) != (m < T(0))) m += b; }
    else m = (b < T(0)) ? -T(0) : T(0);
    return m;
}

template <typename T> struct ConstsOf;
template <> struct ConstsOf<float> {
    static __device__ __forceinline__ const float* get(const ProgramImage& img) { return img.cf; }
};
template <> struct ConstsOf<double> {
    static __device__ __forceinline__ const double* get(const ProgramImage& img) { return img.cd; }
};
}  // namespace sdfm

#define SDF_FOR_N _Pragma("unroll") for (int i = 0; i < N; ++i)

// Evaluate the program at N points.  On return d[i] = sdf(x[i], y[i], z[i]);
// x/y/z are clobbered.
template <typename T, int N>
__device__ __forceinline__ void sdf_interp(const ProgramImage& img, T (&x)[N], T (&y)[N], T (&z)[N], T (&d)[N]) {
    using namespace sdfm;
    T s0[N], r0x[N], r0y[N], r0z[N];
    T vs[SDF_MAX_VSTACK][N];
    T ps[SDF_MAX_PSLOTS][3][N];
    SDF_FOR_N { s0[i] = T(0); r0x[i] = r0y[i] = r0z[i] = T(0); d[i] = T(0); }
    const T* cbase = ConstsOf<T>::get(img);
    const int n_code = (int)img.n_code;
    for (int pc = 0; pc < n_code; ++pc) {
        const uint32_t w = img.code[pc];
        const int a = (int)((w >> 8) & 0xffu);
        const T* c = cbase + (w >> 16);
        switch (w & 0xffu) {
        // ------------------------------------------------------------ primitives
        case OP_SPHERE: {                                   // d3.py:94-95
            const T cx = c[0], cy = c[1], cz = c[2], r = c[3];
            SDF_FOR_N d[i] = m_len3(x[i] - cx, y[i] - cy, z[i] - cz) - r;
        } break;
        case OP_PLANE: {                                    // d3.py:101-102
            const T nx = c[0], ny = c[1], nz = c[2], px = c[3], py = c[4], pz = c[5];
            SDF_FOR_N d[i] = (px - x[i]) * nx + (py - y[i]) * ny + (pz - z[i]) * nz;
        } break;
        case OP_BOX: {                                      // d3.py:131-133
            const T cx = c[0], cy = c[1], cz = c[2], hx = c[3], hy = c[4], hz = c[5];
            SDF_FOR_N {
                const T qx = m_abs(x[i] - cx) - hx, qy = m_abs(y[i] - cy) - hy, qz = m_abs(z[i] - cz) - hz;
                d[i] = m_len3(m_max(qx, T(0)), m_max(qy, T(0)), m_max(qz, T(0))) +
                       m_min(m_max(qx, m_max(qy, qz)), T(0));
            }
        } break;
        case OP_ROUNDED_BOX: {                              // d3.py:139-141
            const T hx = c[0], hy = c[1], hz = c[2], r = c[3];
            SDF_FOR_N {
                const T qx = m_abs(x[i]) - hx + r, qy = m_abs(y[i]) - hy + r, qz = m_abs(z[i]) - hz + r;
                d[i] = m_len3(m_max(qx, T(0)), m_max(qy, T(0)), m_max(qz, T(0))) +
                       m_min(m_max(qx, m_max(qy, qz)), T(0)) - r;
            }
        } break;
        case OP_WIREFRAME_BOX: {                            // d3.py:147-154
            const T hx = c[0], hy = c[1], hz = c[2], t = c[3], ht = c[3] / T(2);
            SDF_FOR_N {
                const T px = m_abs(x[i]) - hx - ht, py = m_abs(y[i]) - hy - ht, pz = m_abs(z[i]) - hz - ht;
                const T qx = m_abs(px + ht) - ht, qy = m_abs(py + ht) - ht, qz = m_abs(pz + ht) - ht;
                (void)t;
                const T g1 = m_len3(m_max(px, T(0)), m_max(qy, T(0)), m_max(qz, T(0))) + m_min(m_max(px, m_max(qy, qz)), T(0));
                const T g2 = m_len3(m_max(qx, T(0)), m_max(py, T(0)), m_max(qz, T(0))) + m_min(m_max(qx, m_max(py, qz)), T(0));
                const T g3 = m_len3(m_max(qx, T(0)), m_max(qy, T(0)), m_max(pz, T(0))) + m_min(m_max(qx, m_max(qy, pz)), T(0));
                d[i] = m_min(m_min(g1, g2), g3);
            }
        } break;
        case OP_TORUS: {                                    // d3.py:159-164
            const T r1 = c[0], r2 = c[1];
            SDF_FOR_N { const T q = m_len2(x[i], y[i]) - r1; d[i] = m_len2(q, z[i]) - r2; }
        } break;
        case OP_CAPSULE: {                                  // d3.py:171-175
            const T ax = c[0], ay = c[1], az = c[2], r = c[6];
            const T bx = c[3] - ax, by = c[4] - ay, bz = c[5] - az;
            const T baba = bx * bx + by * by + bz * bz;
            SDF_FOR_N {
                const T px = x[i] - ax, py = y[i] - ay, pz = z[i] - az;
                const T h = m_clip01((px * bx + py * by + pz * bz) / baba);
                d[i] = m_len3(px - bx * h, py - by * h, pz - bz * h) - r;
            }
        } break;
        case OP_CYLINDER: {                                 // d3.py:180-181
            const T r = c[0];
            SDF_FOR_N d[i] = m_len2(x[i], y[i]) - r;
        } break;
        case OP_CAPPED_CYLINDER: {                          // d3.py:188-203
            const T ax = c[0], ay = c[1], az = c[2], r = c[6];
            const T bx = c[3] - ax, by = c[4] - ay, bz = c[5] - az;
            const T baba = bx * bx + by * by + bz * bz;
            SDF_FOR_N {
                const T px = x[i] - ax, py = y[i] - ay, pz = z[i] - az;
                const T paba = px * bx + py * by + pz * bz;
                const T xx = m_len3(px * baba - bx * paba, py * baba - by * paba, pz * baba - bz * paba) - r * baba;
                const T yy = m_abs(paba - baba * T(0.5)) - baba * T(0.5);
                const T x2 = xx * xx, y2 = yy * yy * baba;
                const T dd = (m_max(xx, yy) < T(0)) ? -m_min(x2, y2)
                                                      : ((xx > T(0) ? x2 : T(0)) + (yy > T(0) ? y2 : T(0)));
                d[i] = m_sign(dd) * m_sqrt(m_abs(dd)) / baba;
            }
        } break;
        case OP_ROUNDED_CYLINDER: {                         // d3.py:208-214
            const T ra = c[0], rb = c[1], hh = c[2] / T(2);
            SDF_FOR_N {
                const T d0 = m_len2(x[i], y[i]) - ra + rb;
                const T d1 = m_abs(z[i]) - hh + rb;
                d[i] = m_min(m_max(d0, d1), T(0)) + m_len2(m_max(d0, T(0)), m_max(d1, T(0))) - rb;
            }
        } break;
        case OP_CAPPED_CONE: {                              // d3.py:221-236
            const T ax = c[0], ay = c[1], az = c[2], ra = c[6], rb = c[7];
            const T bx = c[3] - ax, by = c[4] - ay, bz = c[5] - az;
            const T rba = rb - ra;
            const T baba = bx * bx + by * by + bz * bz;
            const T k = rba * rba + baba;
            SDF_FOR_N {
                const T px = x[i] - ax, py = y[i] - ay, pz = z[i] - az;
                const T papa = px * px + py * py + pz * pz;
                const T paba = (px * bx + py * by + pz * bz) / baba;
                const T xx = m_sqrt(papa - paba * paba * baba);
                const T cax = m_max(T(0), xx - (paba < T(0.5) ? ra : rb));
                const T cay = m_abs(paba - T(0.5)) - T(0.5);
                const T f = m_clip01((rba * (xx - ra) + paba * baba) / k);
                const T cbx = xx - ra - f * rba;
                const T cby = paba - f;
                const T s = (cbx < T(0) && cay < T(0)) ? T(-1) : T(1);
                d[i] = s * m_sqrt(m_min(cax * cax + cay * cay * baba, cbx * cbx + cby * cby * baba));
            }
        } break;
        case OP_ROUNDED_CONE: {                             // d3.py:241-249
            const T r1 = c[0], r2 = c[1], h = c[2];
            const T b = (r1 - r2) / h;
            const T aa = m_sqrt(T(1) - b * b);
            SDF_FOR_N {
                const T qx = m_len2(x[i], y[i]), qy = z[i];
                const T k = qx * (-b) + qy * aa;
                const T c1 = m_len2(qx, qy) - r1;
                const T c2 = m_len2(qx, qy - h) - r2;
                const T c3 = (qx * aa + qy * b) - r1;
                d[i] = k < T(0) ? c1 : (k > aa * h ? c2 : c3);
            }
        } break;
        case OP_ELLIPSOID: {                                // d3.py:255-258
            const T sx = c[0], sy = c[1], sz = c[2];
            const T sxx = sx * sx, syy = sy * sy, szz = sz * sz;
            SDF_FOR_N {
                const T k0 = m_len3(x[i] / sx, y[i] / sy, z[i] / sz);
                const T k1 = m_len3(x[i] / sxx, y[i] / syy, z[i] / szz);
                d[i] = k0 * (k0 - T(1)) / k1;
            }
        } break;
        case OP_PYRAMID: {                                  // d3.py:263-281
            const T h = c[0];
            const T m2 = h * h + T(0.25);
            SDF_FOR_N {
                T a0 = m_abs(x[i]) - T(0.5), a1 = m_abs(y[i]) - T(0.5);
                if (a1 > a0) { const T t = a0; a0 = a1; a1 = t; }
                const T px = a0, py = z[i], pz = a1;
                const T qx = pz, qy = h * py - T(0.5) * px, qz = h * px + T(0.5) * py;
                const T s = m_max(-qx, T(0));
                const T t = m_clip01((qy - T(0.5) * pz) / (m2 + T(0.25)));
                const T aa = m2 * (qx + s) * (qx + s) + qy * qy;
                const T bb = m2 * (qx + T(0.5) * t) * (qx + T(0.5) * t) + (qy - m2 * t) * (qy - m2 * t);
                const T d2 = (m_min(qy, -qx * m2 - qy * T(0.5)) > T(0)) ? T(0) : m_min(aa, bb);
                d[i] = m_sqrt((d2 + qz * qz) / m2) * m_sign(m_max(qz, -py));
            }
        } break;
        case OP_TETRAHEDRON: {                              // d3.py:288-292
            const T r = c[0];
            const T s3 = m_sqrt(T(3));
            SDF_FOR_N d[i] = (m_max(m_abs(x[i] + y[i]) - z[i], m_abs(x[i] - y[i]) + z[i]) - r) / s3;
        } break;
        case OP_OCTAHEDRON: {                               // d3.py:297-298
            const T r = c[0], t30 = c[1];
            SDF_FOR_N d[i] = (m_abs(x[i]) + m_abs(y[i]) + m_abs(z[i]) - r) * t30;
        } break;
        case OP_DODECAHEDRON: {                             // d3.py:303-310
            const T r = c[0], vx = c[1], vy = c[2], vz = c[3];
            SDF_FOR_N {
                const T px = m_abs(x[i] / r), py = m_abs(y[i] / r), pz = m_abs(z[i] / r);
                const T aa = px * vx + py * vy + pz * vz;
                const T bb = px * vz + py * vx + pz * vy;
                const T cc = px * vy + py * vz + pz * vx;
                d[i] = (m_max(m_max(aa, bb), cc) - vx) * r;
            }
        } break;
        case OP_ICOSAHEDRON: {                              // d3.py:318-324
            const T r = c[0], vx = c[1], vy = c[2], vz = c[3], vw = c[4];
            SDF_FOR_N {
                const T px = m_abs(x[i] / r), py = m_abs(y[i] / r), pz = m_abs(z[i] / r);
                const T aa = px * vx + py * vy + pz * vz;
                const T bb = px * vz + py * vx + pz * vy;
                const T cc = px * vy + py * vz + pz * vx;
                const T dd = (px * vw + py * vw + pz * vw) - vx;
                d[i] = m_max(m_max(m_max(aa, bb), cc) - vx, dd) * r;
            }
        } break;
        // ------------------------------------------------------- point transforms
        case OP_TRANSLATE: {                                // d3.py:331-332
            const T ox = c[0], oy = c[1], oz = c[2];
            SDF_FOR_N { x[i] -= ox; y[i] -= oy; z[i] -= oz; }
        } break;
        case OP_SCALE_P: {                                  // d3.py:343-344
            const T sx = c[0], sy = c[1], sz = c[2];
            SDF_FOR_N { x[i] /= sx; y[i] /= sy; z[i] /= sz; }
        } break;
        case OP_ROTATE: {                                   // d3.py:358-359  (p @ M)
            const T m00 = c[0], m01 = c[1], m02 = c[2], m10 = c[3], m11 = c[4], m12 = c[5],
                    m20 = c[6], m21 = c[7], m22 = c[8];
            SDF_FOR_N {
                const T px = x[i], py = y[i], pz = z[i];
                x[i] = px * m00 + py * m10 + pz * m20;
                y[i] = px * m01 + py * m11 + pz * m21;
                z[i] = px * m02 + py * m12 + pz * m22;
            }
        } break;
        case OP_TWIST: {                                    // d3.py:409-418
            const T k = c[0];
            SDF_FOR_N {
                T sn, cs; m_sincos(k * z[i], &sn, &cs);
                const T px = x[i], py = y[i];
                x[i] = cs * px - sn * py;
                y[i] = sn * px + cs * py;
            }
        } break;
        case OP_BEND: {                                     // d3.py:423-432
            const T k = c[0];
            SDF_FOR_N {
                T sn, cs; m_sincos(k * x[i], &sn, &cs);
                const T px = x[i], py = y[i];
                x[i] = cs * px - sn * py;
                y[i] = sn * px + cs * py;
            }
        } break;
        case OP_ELONGATE: {                                 // d3.py:398-404 (pushes w)
            const T sx = c[0], sy = c[1], sz = c[2];
            if (a > 0) { SDF_FOR_N vs[a - 1][i] = s0[i]; }
            SDF_FOR_N {
                const T qx = m_abs(x[i]) - sx, qy = m_abs(y[i]) - sy, qz = m_abs(z[i]) - sz;
                s0[i] = m_min(m_max(qx, m_max(qy, qz)), T(0));
                x[i] = m_max(qx, T(0)); y[i] = m_max(qy, T(0)); z[i] = m_max(qz, T(0));
            }
        } break;
        case OP_CIRCULAR_PRE: {                             // d3.py:385-390
            const T da = c[0];
            SDF_FOR_N {
                const T r = m_hypot(x[i], y[i]);
                const T ang = m_pymod(m_atan2(y[i], x[i]), da);
                T s1, c1, s2, c2;
                m_sincos(ang - da, &s1, &c1);
                m_sincos(ang, &s2, &c2);
                const T qx = c2 * r, qy = s2 * r;
                if (a == 0) { r0x[i] = qx; r0y[i] = qy; r0z[i] = z[i]; }
                else { ps[a][0][i] = qx; ps[a][1][i] = qy; ps[a][2][i] = z[i]; }
                x[i] = c1 * r; y[i] = s1 * r;
            }
        } break;
        case OP_REPEAT_CELL: {                              // dn.py:101-109
            SDF_FOR_N {
                T* pp[3] = {&x[i], &y[i], &z[i]};
                _Pragma("unroll") for (int k = 0; k < 3; ++k) {
                    const T sp = c[k];
                    T q = (sp != T(0)) ? (*pp[k] / sp) : T(0);
                    T idx = m_rint(q);
                    if (c[6] != T(0)) idx = m_min(m_max(idx, -c[3 + k]), c[3 + k]);
                    *pp[k] = *pp[k] - sp * (idx + c[7 + k]);
                }
            }
        } break;
        // ------------------------------------------------------------- value ops
        case OP_MULC: { const T m = c[0]; SDF_FOR_N d[i] *= m; } break;                 // d3.py:344
        case OP_NEGATE: { SDF_FOR_N d[i] = -d[i]; } break;                              // dn.py:64
        case OP_ADDC: { const T v = c[0]; SDF_FOR_N d[i] += v; } break;                 // dn.py:69,74
        case OP_SHELL: { const T h = c[0] / T(2); SDF_FOR_N d[i] = m_abs(d[i]) - h; } break;  // dn.py:79
        case OP_ADD_POP: {                                                              // d3.py:404
            SDF_FOR_N d[i] = d[i] + s0[i];
            if (a > 0) { SDF_FOR_N s0[i] = vs[a - 1][i]; }
        } break;
        // ----------------------------------------------------------------- folds
        case OP_UNION: {                                    // dn.py:14
            SDF_FOR_N d[i] = m_min(s0[i], d[i]);
            if (a > 0) { SDF_FOR_N s0[i] = vs[a - 1][i]; }
        } break;
        case OP_UNION_K: {                                  // dn.py:16-18
            const T K = c[0];
            SDF_FOR_N {
                const T d1 = s0[i], d2 = d[i];
                const T h = m_clip01(T(0.5) + T(0.5) * (d2 - d1) / K);
                const T m = d2 + (d1 - d2) * h;
                d[i] = m - K * h * (T(1) - h);
            }
            if (a > 0) { SDF_FOR_N s0[i] = vs[a - 1][i]; }
        } break;
        case OP_DIFFERENCE: {                               // dn.py:29
            SDF_FOR_N d[i] = m_max(s0[i], -d[i]);
            if (a > 0) { SDF_FOR_N s0[i] = vs[a - 1][i]; }
        } break;
        case OP_DIFFERENCE_K: {                             // dn.py:31-33
            const T K = c[0];
            SDF_FOR_N {
                const T d1 = s0[i], d2 = d[i];
                const T h = m_clip01(T(0.5) - T(0.5) * (d2 + d1) / K);
                const T m = d1 + (-d2 - d1) * h;
                d[i] = m + K * h * (T(1) - h);
            }
            if (a > 0) { SDF_FOR_N s0[i] = vs[a - 1][i]; }
        } break;
        case OP_INTERSECTION: {                             // dn.py:44
            SDF_FOR_N d[i] = m_max(s0[i], d[i]);
            if (a > 0) { SDF_FOR_N s0[i] = vs[a - 1][i]; }
        } break;
        case OP_INTERSECTION_K: {                           // dn.py:46-48
            const T K = c[0];
            SDF_FOR_N {
                const T d1 = s0[i], d2 = d[i];
                const T h = m_clip01(T(0.5) - T(0.5) * (d2 - d1) / K);
                const T m = d2 + (d1 - d2) * h;
                d[i] = m + K * h * (T(1) - h);
            }
            if (a > 0) { SDF_FOR_N s0[i] = vs[a - 1][i]; }
        } break;
        case OP_BLEND: {                                    // dn.py:58
            const T K = c[0];
            SDF_FOR_N d[i] = K * d[i] + (T(1) - K) * s0[i];
            if (a > 0) { SDF_FOR_N s0[i] = vs[a - 1][i]; }
        } break;
        // --------------------------------------------------------- stack / slots
        case OP_PUSH: {
            if (a > 0) { SDF_FOR_N vs[a - 1][i] = s0[i]; }
            SDF_FOR_N s0[i] = d[i];
        } break;
        case OP_SAVEP: {
            if (a == 0) { SDF_FOR_N { r0x[i] = x[i]; r0y[i] = y[i]; r0z[i] = z[i]; } }
            else { SDF_FOR_N { ps[a][0][i] = x[i]; ps[a][1][i] = y[i]; ps[a][2][i] = z[i]; } }
        } break;
        case OP_LOADP: {
            if (a == 0) { SDF_FOR_N { x[i] = r0x[i]; y[i] = r0y[i]; z[i] = r0z[i]; } }
            else { SDF_FOR_N { x[i] = ps[a][0][i]; y[i] = ps[a][1][i]; z[i] = ps[a][2][i]; } }
        } break;
        default: break;
        }
    }
}
