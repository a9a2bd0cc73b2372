// sdf_ops.cuh -- the SDF node library on the device + the interpreter that runs a lowered tree.
//
// Replaces the reference's recursive numpy closure tree (sdf/d3.py:24-25 calling into
// sdf/d3.py:92-502 and sdf/dn.py:7-114).  Every node is ONE inline function SdfOps<T,N,FAST>::NAME
// operating on N sample points held in registers; two front ends drive them:
//
//   * sdf_interp(): a warp-uniform interpreter over the lowered program (sdf_b200/ir.py) that sits
//     in the kernel-parameter constant bank (ProgramImage, __grid_constant__) -- always available;
//   * generated straight-line code (one call per program word with literal constants and literal
//     stack/slot operands), compiled at run time with NVRTC for sm_100a (sdfb200.cu: Jit) -- same
//     functions, so both front ends are the same arithmetic; the compiler then keeps the value
//     stack and the point slots in registers and folds every constant into an immediate.
//
// Templated on the scalar type: float is the throughput path, double restates the reference's
// float64 arithmetic (skip test, SDF3.__call__, re-evaluation of everything that reaches the
// output).  This header compiles under nvcc and under NVRTC (no host headers in the RTC case).
#pragma once
#ifndef __CUDACC_RTC__
#include <cuda_runtime.h>
#include <stdint.h>
#else
typedef unsigned char uint8_t;
typedef unsigned int uint32_t;
typedef unsigned long long uint64_t;
typedef long long int64_t;
#endif
#include "sdf_opcodes.h"

#pragma nv_diag_suppress 177      // SDF_BIND_STATE binds every state array in every op

struct ProgramImage {
    uint32_t n_code;
    uint32_t pad_;
    uint32_t code[SDF_MAX_CODE];
    float cf[SDF_MAX_CONSTS];
    double cd[SDF_MAX_CONSTS];
};

namespace sdfm {
// numpy's minimum / maximum (dn.py:14,29,44) return NaN when either operand is NaN; fmin / fmax would return
// the other operand.  The accurate policies (float64, accurate float32) follow numpy; the FAST float32
// policy keeps the single-instruction fminf / fmaxf -- everything it decides is re-evaluated in float64.
#define SDF_MINMAX(T, F, a, b) (((a) != (a) || (b) != (b)) ? ((a) + (b)) : F((a), (b)))
// ----------------------------------------------------------------------------------------------
// Math policies.  MathP<double, *> and MathP<float, false> use IEEE-rounded division / square
// root and the accurate libdevice transcendentals.  MathP<float, true> is the throughput path:
// approximate division / square root (MUFU.RCP / MUFU.RSQ class, <= 2 ulp), MUFU sine/cosine
// after one Cody-Waite reduction step, polynomial atan2.  It is only ever used where every
// value that can influence the output is re-evaluated afterwards (see fused_tile_kernel).
template <typename T, bool FAST> struct MathP;

template <bool FAST> struct MathP<double, FAST> {
    typedef double T;
    static __device__ __forceinline__ T sqrt_(T x) { return sqrt(x); }
    static __device__ __forceinline__ T div(T a, T b) { return a / b; }
    static __device__ __forceinline__ T abs_(T x) { return fabs(x); }
    static __device__ __forceinline__ T min_(T a, T b) { return SDF_MINMAX(T, fmin, a, b); }
    static __device__ __forceinline__ T max_(T a, T b) { return SDF_MINMAX(T, fmax, a, b); }
    static __device__ __forceinline__ void sincos_(T a, T* s, T* c) { sincos(a, s, c); }
    static __device__ __forceinline__ T atan2_(T y, T x) { return atan2(y, x); }
    static __device__ __forceinline__ T hypot_(T x, T y) { return hypot(x, y); }
    static __device__ __forceinline__ T rint_(T a) { return rint(a); }
    static __device__ __forceinline__ T exp2_(T a) { return exp2(a); }
    // numpy's floored modulo (np.mod / %): fmod, then shift into the divisor's sign
    static __device__ __forceinline__ T pymod(T a, T b) {
        T m = fmod(a, b);
        if (m != T(0)) { if ((b < T(0)) != (m < T(0))) m += b; }
        else m = (b < T(0)) ? -T(0) : T(0);
        return m;
    }
};

template <> struct MathP<float, false> {
    typedef float T;
    static __device__ __forceinline__ T sqrt_(T x) { return sqrtf(x); }
    static __device__ __forceinline__ T div(T a, T b) { return a / b; }
    static __device__ __forceinline__ T abs_(T x) { return fabsf(x); }
    static __device__ __forceinline__ T min_(T a, T b) { return SDF_MINMAX(T, fminf, a, b); }
    static __device__ __forceinline__ T max_(T a, T b) { return SDF_MINMAX(T, fmaxf, a, b); }
    static __device__ __forceinline__ void sincos_(T a, T* s, T* c) { sincosf(a, s, c); }
    static __device__ __forceinline__ T atan2_(T y, T x) { return atan2f(y, x); }
    static __device__ __forceinline__ T hypot_(T x, T y) { return sqrtf(x * x + y * y); }
    static __device__ __forceinline__ T rint_(T a) { return rintf(a); }
    static __device__ __forceinline__ T exp2_(T a) { return exp2f(a); }
    static __device__ __forceinline__ T pymod(T a, T b) {
        T m = fmodf(a, b);
        if (m != T(0)) { if ((b < T(0)) != (m < T(0))) m += b; }
        else m = (b < T(0)) ? -T(0) : T(0);
        return m;
    }
};

template <> struct MathP<float, true> {
    typedef float T;
    static __device__ __forceinline__ T sqrt_(T x) {
        T r;
        asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
        return r;
    }
    static __device__ __forceinline__ T div(T a, T b) { return __fdividef(a, b); }
    static __device__ __forceinline__ T abs_(T x) { return fabsf(x); }
    static __device__ __forceinline__ T min_(T a, T b) { return fminf(a, b); }
    static __device__ __forceinline__ T max_(T a, T b) { return fmaxf(a, b); }
    static __device__ __forceinline__ void sincos_(T a, T* s, T* c) {
        // one Cody-Waite step to [-pi, pi], then the MUFU units (abs error ~4e-7 there)
        const T k = rintf(a * 0.15915494309189535f);
        T r = fmaf(k, -6.2831854820251465f, a);            // 2*pi hi (float)
        r = fmaf(k, 1.7484555e-07f, r);                    // -(2*pi - hi)
        __sincosf(r, s, c);
    }
    static __device__ __forceinline__ T atan2_(T y, T x) {
        const T ax = fabsf(x), ay = fabsf(y);
        const T mx = fmaxf(ax, ay), mn = fminf(ax, ay);
        const T tq = __fdividef(mn, mx);                   // unconditional: keeps the function branch-free, so that
        const T t = (mx > 0.0f) ? tq : 0.0f;               // identical calls merge (tile kernel, axis sharing)
        const T q = t * t;
        // near-minimax fit of atan(t)/t in q = t*t on [0,1], max error 8.9e-8
        T p = -4.7804095924e-03f;
        p = fmaf(p, q, 2.4556965949e-02f);
        p = fmaf(p, q, -5.9904500254e-02f);
        p = fmaf(p, q, 9.9427446283e-02f);
        p = fmaf(p, q, -1.4029414772e-01f);
        p = fmaf(p, q, 1.9971374339e-01f);
        p = fmaf(p, q, -3.3332093464e-01f);
        p = fmaf(p, q, 9.9999991141e-01f);
        T r = p * t;
        r = (ay > ax) ? 1.5707963267948966f - r : r;
        r = (x < 0.0f) ? 3.141592653589793f - r : r;
        return (y < 0.0f) ? -r : r;
    }
    static __device__ __forceinline__ T hypot_(T x, T y) { return sqrt_(x * x + y * y); }
    static __device__ __forceinline__ T rint_(T a) { return rintf(a); }
    static __device__ __forceinline__ T exp2_(T a) { return exp2f(a); }
    static __device__ __forceinline__ T pymod(T a, T b) {
        const T m = fmaf(-floorf(__fdividef(a, b)), b, a);
        return m;
    }
};

template <typename T> struct ConstsOf;
template <> struct ConstsOf<float> {
    static __device__ __forceinline__ const float* get(const ProgramImage& img) { return img.cf; }
};
template <> struct ConstsOf<double> {
    static __device__ __forceinline__ const double* get(const ProgramImage& img) { return img.cd; }
};
}  // namespace sdfm

#define SDF_FOR_N _Pragma("unroll") for (int i = 0; i < N; ++i)

#define SDF_BIND_STATE                                                                             \
    T(&x)[N] = st.x; T(&y)[N] = st.y; T(&z)[N] = st.z; T(&d)[N] = st.d; T(&s0)[N] = st.s0;         \
    T(&r0x)[N] = st.r0x; T(&r0y)[N] = st.r0y; T(&r0z)[N] = st.r0z;                                 \
    T(&vs)[SDF_MAX_VSTACK][N] = st.vs; T(&ps)[SDF_MAX_PSLOTS][3][N] = st.ps

// Machine state of N sample points (sdf_b200/ir.py "machine model"): P = (x,y,z), D = d, cached top
// of the value stack s0, cached point slot 0 (r0*), spill levels vs[] / ps[].
template <typename T, int N>
struct SdfState {
    T (&x)[N]; T (&y)[N]; T (&z)[N]; T (&d)[N];
    T s0[N], r0x[N], r0y[N], r0z[N];
    T vs[SDF_MAX_VSTACK][N];
    T ps[SDF_MAX_PSLOTS][3][N];
    __device__ __forceinline__ SdfState(T (&x_)[N], T (&y_)[N], T (&z_)[N], T (&d_)[N]) : x(x_), y(y_), z(z_), d(d_) {
        SDF_FOR_N { s0[i] = T(0); r0x[i] = r0y[i] = r0z[i] = T(0); d[i] = T(0); }
    }
};

template <typename T, int N, bool FAST>
struct SdfOps {
    typedef sdfm::MathP<T, FAST> M;
    typedef SdfState<T, N> State;
    static __device__ __forceinline__ T m_sqrt(T v) { return M::sqrt_(v); }
    static __device__ __forceinline__ T m_abs(T v) { return M::abs_(v); }
    static __device__ __forceinline__ T m_min(T u, T v) { return M::min_(u, v); }
    static __device__ __forceinline__ T m_max(T u, T v) { return M::max_(u, v); }
    static __device__ __forceinline__ T m_clip01(T v) { return M::min_(M::max_(v, T(0)), T(1)); }
    static __device__ __forceinline__ T m_sign(T v) { return v > T(0) ? T(1) : (v < T(0) ? T(-1) : T(0)); }
    static __device__ __forceinline__ T m_len2(T u, T v) { return M::sqrt_(u * u + v * v); }
    static __device__ __forceinline__ T m_len3(T u, T v, T w) { return M::sqrt_(u * u + v * v + w * w); }
    static __device__ __forceinline__ T m_div(T u, T v) { return M::div(u, v); }
    // u*a + v*b + w*c.  In the generated float32 throughput code (FAST, constants are literals) a term
    // whose coefficient is zero relative to `scale` (below 2^-28 of it: 1/16 ulp of the sum) is left
    // out, so that the result does not depend on that input at all -- what lets the tile kernel's
    // axis sharing see through axis-aligned rotations and planes (cos(pi/2) = 6e-17 is not 0.0).
    // Everything that can reach the output is re-evaluated without this shortcut.
    static __device__ __forceinline__ T dot3z(T u, T v, T w, T a, T b, T c, T scale) {
#ifdef __CUDACC_RTC__
        if (FAST) {
            const T eps = scale * T(3.7252902984619141e-09);
            const bool ka = m_abs(a) > eps, kb = m_abs(b) > eps, kc = m_abs(c) > eps;
            T acc = T(0);
            if (ka) acc = u * a;
            if (kb) acc = ka ? acc + v * b : v * b;
            if (kc) acc = (ka || kb) ? acc + w * c : w * c;
            return acc;
        }
#endif
        return u * a + v * b + w * c;
    }

    static __device__ __forceinline__ void SPHERE(State& st, const T* c, const int a) {
        SDF_BIND_STATE;                               // d3.py:94-95
        const T cx = c[0], cy = c[1], cz = c[2], r = c[3];
        SDF_FOR_N d[i] = m_len3(x[i] - cx, y[i] - cy, z[i] - cz) - r;
        }

    static __device__ __forceinline__ void PLANE(State& st, const T* c, const int a) {
        SDF_BIND_STATE;                                // d3.py:101-102
        const T nx = c[0], ny = c[1], nz = c[2], px = c[3], py = c[4], pz = c[5];
        const T sc = m_max(m_abs(nx), m_max(m_abs(ny), m_abs(nz)));
        SDF_FOR_N d[i] = dot3z(px - x[i], py - y[i], pz - z[i], nx, ny, nz, sc);
        }

    static __device__ __forceinline__ void BOX(State& st, const T* c, const int a) {
        SDF_BIND_STATE;                                  // d3.py:131-133
        const T cx = c[0], cy = c[1], cz = c[2], hx = c[3], hy = c[4], hz = c[5];
        SDF_FOR_N {
            const T qx = m_abs(x[i] - cx) - hx, qy = m_abs(y[i] - cy) - hy, qz = m_abs(z[i] - cz) - hz;
            d[i] = m_len3(m_max(qx, T(0)), m_max(qy, T(0)), m_max(qz, T(0))) +
                   m_min(m_max(qx, m_max(qy, qz)), T(0));
        }
        }

    static __device__ __forceinline__ void ROUNDED_BOX(State& st, const T* c, const int a) {
        SDF_BIND_STATE;                          // d3.py:139-141
        const T hx = c[0], hy = c[1], hz = c[2], r = c[3];
        SDF_FOR_N {
            const T qx = m_abs(x[i]) - hx + r, qy = m_abs(y[i]) - hy + r, qz = m_abs(z[i]) - hz + r;
            d[i] = m_len3(m_max(qx, T(0)), m_max(qy, T(0)), m_max(qz, T(0))) +
                   m_min(m_max(qx, m_max(qy, qz)), T(0)) - r;
        }
        }

    static __device__ __forceinline__ void WIREFRAME_BOX(State& st, const T* c, const int a) {
        SDF_BIND_STATE;                        // d3.py:147-154
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
        }

    static __device__ __forceinline__ void TORUS(State& st, const T* c, const int a) {
        SDF_BIND_STATE;                                // d3.py:159-164
        const T r1 = c[0], r2 = c[1];
        SDF_FOR_N { const T q = m_len2(x[i], y[i]) - r1; d[i] = m_len2(q, z[i]) - r2; }
        }

    static __device__ __forceinline__ void CAPSULE(State& st, const T* c, const int a) {
        SDF_BIND_STATE;                              // d3.py:171-175
        const T ax = c[0], ay = c[1], az = c[2], r = c[6];
        const T bx = c[3] - ax, by = c[4] - ay, bz = c[5] - az;
        const T baba = bx * bx + by * by + bz * bz;
        SDF_FOR_N {
            const T px = x[i] - ax, py = y[i] - ay, pz = z[i] - az;
            const T h = m_clip01(m_div(px * bx + py * by + pz * bz, baba));
            d[i] = m_len3(px - bx * h, py - by * h, pz - bz * h) - r;
        }
        }

    static __device__ __forceinline__ void CYLINDER(State& st, const T* c, const int a) {
        SDF_BIND_STATE;                             // d3.py:180-181
        const T r = c[0];
        SDF_FOR_N d[i] = m_len2(x[i], y[i]) - r;
        }

    static __device__ __forceinline__ void CAPPED_CYLINDER(State& st, const T* c, const int a) {
        SDF_BIND_STATE;                      // d3.py:188-203
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
            d[i] = m_div(m_sign(dd) * m_sqrt(m_abs(dd)), baba);
        }
        }

    static __device__ __forceinline__ void ROUNDED_CYLINDER(State& st, const T* c, const int a) {
        SDF_BIND_STATE;                     // d3.py:208-214
        const T ra = c[0], rb = c[1], hh = c[2] / T(2);
        SDF_FOR_N {
            const T d0 = m_len2(x[i], y[i]) - ra + rb;
            const T d1 = m_abs(z[i]) - hh + rb;
            d[i] = m_min(m_max(d0, d1), T(0)) + m_len2(m_max(d0, T(0)), m_max(d1, T(0))) - rb;
        }
        }

    static __device__ __forceinline__ void CAPPED_CONE(State& st, const T* c, const int a) {
        SDF_BIND_STATE;                          // d3.py:221-236
        const T ax = c[0], ay = c[1], az = c[2], ra = c[6], rb = c[7];
        const T bx = c[3] - ax, by = c[4] - ay, bz = c[5] - az;
        const T rba = rb - ra;
        const T baba = bx * bx + by * by + bz * bz;
        const T k = rba * rba + baba;
        SDF_FOR_N {
            const T px = x[i] - ax, py = y[i] - ay, pz = z[i] - az;
            const T papa = px * px + py * py + pz * pz;
            const T paba = m_div(px * bx + py * by + pz * bz, baba);
            const T xx = m_sqrt(papa - paba * paba * baba);
            const T cax = m_max(T(0), xx - (paba < T(0.5) ? ra : rb));
            const T cay = m_abs(paba - T(0.5)) - T(0.5);
            const T f = m_clip01(m_div(rba * (xx - ra) + paba * baba, k));
            const T cbx = xx - ra - f * rba;
            const T cby = paba - f;
            const T s = (cbx < T(0) && cay < T(0)) ? T(-1) : T(1);
            d[i] = s * m_sqrt(m_min(cax * cax + cay * cay * baba, cbx * cbx + cby * cby * baba));
        }
        }

    static __device__ __forceinline__ void ROUNDED_CONE(State& st, const T* c, const int a) {
        SDF_BIND_STATE;                         // d3.py:241-249
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
        }

    static __device__ __forceinline__ void ELLIPSOID(State& st, const T* c, const int a) {
        SDF_BIND_STATE;                            // d3.py:255-258
        const T sx = c[0], sy = c[1], sz = c[2];
        const T sxx = sx * sx, syy = sy * sy, szz = sz * sz;
        SDF_FOR_N {
            const T k0 = m_len3(m_div(x[i], sx), m_div(y[i], sy), m_div(z[i], sz));
            const T k1 = m_len3(m_div(x[i], sxx), m_div(y[i], syy), m_div(z[i], szz));
            d[i] = m_div(k0 * (k0 - T(1)), k1);
        }
        }

    static __device__ __forceinline__ void PYRAMID(State& st, const T* c, const int a) {
        SDF_BIND_STATE;                              // d3.py:263-281
        const T h = c[0];
        const T m2 = h * h + T(0.25);
        SDF_FOR_N {
            T a0 = m_abs(x[i]) - T(0.5), a1 = m_abs(y[i]) - T(0.5);
            if (a1 > a0) { const T t = a0; a0 = a1; a1 = t; }
            const T px = a0, py = z[i], pz = a1;
            const T qx = pz, qy = h * py - T(0.5) * px, qz = h * px + T(0.5) * py;
            const T s = m_max(-qx, T(0));
            const T t = m_clip01(m_div(qy - T(0.5) * pz, m2 + T(0.25)));
            const T aa = m2 * (qx + s) * (qx + s) + qy * qy;
            const T bb = m2 * (qx + T(0.5) * t) * (qx + T(0.5) * t) + (qy - m2 * t) * (qy - m2 * t);
            const T d2 = (m_min(qy, -qx * m2 - qy * T(0.5)) > T(0)) ? T(0) : m_min(aa, bb);
            d[i] = m_sqrt(m_div(d2 + qz * qz, m2)) * m_sign(m_max(qz, -py));
        }
        }

    static __device__ __forceinline__ void TETRAHEDRON(State& st, const T* c, const int a) {
        SDF_BIND_STATE;                          // d3.py:288-292
        const T r = c[0];
        const T s3 = m_sqrt(T(3));
        SDF_FOR_N d[i] = m_div(m_max(m_abs(x[i] + y[i]) - z[i], m_abs(x[i] - y[i]) + z[i]) - r, s3);
        }

    static __device__ __forceinline__ void OCTAHEDRON(State& st, const T* c, const int a) {
        SDF_BIND_STATE;                           // d3.py:297-298
        const T r = c[0], t30 = c[1];
        SDF_FOR_N d[i] = (m_abs(x[i]) + m_abs(y[i]) + m_abs(z[i]) - r) * t30;
        }

    static __device__ __forceinline__ void DODECAHEDRON(State& st, const T* c, const int a) {
        SDF_BIND_STATE;                         // d3.py:303-310
        const T r = c[0], vx = c[1], vy = c[2], vz = c[3];
        SDF_FOR_N {
            const T px = m_abs(m_div(x[i], r)), py = m_abs(m_div(y[i], r)), pz = m_abs(m_div(z[i], r));
            const T aa = px * vx + py * vy + pz * vz;
            const T bb = px * vz + py * vx + pz * vy;
            const T cc = px * vy + py * vz + pz * vx;
            d[i] = (m_max(m_max(aa, bb), cc) - vx) * r;
        }
        }

    static __device__ __forceinline__ void ICOSAHEDRON(State& st, const T* c, const int a) {  // ------------------------------------------------------- point transforms
        SDF_BIND_STATE;                          // d3.py:318-324
        const T r = c[0], vx = c[1], vy = c[2], vz = c[3], vw = c[4];
        SDF_FOR_N {
            const T px = m_abs(m_div(x[i], r)), py = m_abs(m_div(y[i], r)), pz = m_abs(m_div(z[i], r));
            const T aa = px * vx + py * vy + pz * vz;
            const T bb = px * vz + py * vx + pz * vy;
            const T cc = px * vy + py * vz + pz * vx;
            const T dd = (px * vw + py * vw + pz * vw) - vx;
            d[i] = m_max(m_max(m_max(aa, bb), cc) - vx, dd) * r;
        }
        }

    static __device__ __forceinline__ void TRANSLATE(State& st, const T* c, const int a) {
        SDF_BIND_STATE;                            // d3.py:331-332
        const T ox = c[0], oy = c[1], oz = c[2];
        SDF_FOR_N { x[i] -= ox; y[i] -= oy; z[i] -= oz; }
        }

    static __device__ __forceinline__ void SCALE_P(State& st, const T* c, const int a) {
        SDF_BIND_STATE;                              // d3.py:343-344
        const T sx = c[0], sy = c[1], sz = c[2];
        SDF_FOR_N { x[i] = m_div(x[i], sx); y[i] = m_div(y[i], sy); z[i] = m_div(z[i], sz); }
        }

    static __device__ __forceinline__ void ROTATE(State& st, const T* c, const int a) {
        SDF_BIND_STATE;                               // d3.py:358-359  (p @ M)
        const T m00 = c[0], m01 = c[1], m02 = c[2], m10 = c[3], m11 = c[4], m12 = c[5],
                m20 = c[6], m21 = c[7], m22 = c[8];
        const T sc = m_max(m_max(m_max(m_abs(m00), m_abs(m01)), m_max(m_abs(m02), m_abs(m10))),
                           m_max(m_max(m_abs(m11), m_abs(m12)), m_max(m_abs(m20), m_max(m_abs(m21), m_abs(m22)))));
        SDF_FOR_N {
            const T px = x[i], py = y[i], pz = z[i];
            x[i] = dot3z(px, py, pz, m00, m10, m20, sc);
            y[i] = dot3z(px, py, pz, m01, m11, m21, sc);
            z[i] = dot3z(px, py, pz, m02, m12, m22, sc);
        }
        }

    static __device__ __forceinline__ void TWIST(State& st, const T* c, const int a) {
        SDF_BIND_STATE;                                // d3.py:409-418
        const T k = c[0];
        SDF_FOR_N {
            T sn, cs; M::sincos_(k * z[i], &sn, &cs);
            const T px = x[i], py = y[i];
            x[i] = cs * px - sn * py;
            y[i] = sn * px + cs * py;
        }
        }

    static __device__ __forceinline__ void BEND(State& st, const T* c, const int a) {
        SDF_BIND_STATE;                                 // d3.py:423-432
        const T k = c[0];
        SDF_FOR_N {
            T sn, cs; M::sincos_(k * x[i], &sn, &cs);
            const T px = x[i], py = y[i];
            x[i] = cs * px - sn * py;
            y[i] = sn * px + cs * py;
        }
        }

    static __device__ __forceinline__ void ELONGATE(State& st, const T* c, const int a) {
        SDF_BIND_STATE;                             // d3.py:398-404 (pushes w)
        const T sx = c[0], sy = c[1], sz = c[2];
        if (a > 0) { SDF_FOR_N vs[a - 1][i] = s0[i]; }
        SDF_FOR_N {
            const T qx = m_abs(x[i]) - sx, qy = m_abs(y[i]) - sy, qz = m_abs(z[i]) - sz;
            s0[i] = m_min(m_max(qx, m_max(qy, qz)), T(0));
            x[i] = m_max(qx, T(0)); y[i] = m_max(qy, T(0)); z[i] = m_max(qz, T(0));
        }
        }

    static __device__ __forceinline__ void CIRCULAR_PRE(State& st, const T* c, const int a) {
        SDF_BIND_STATE;                         // d3.py:385-390
        const T da = c[0];
        if (FAST) {
            // Same two copies, written as rotations of (x, y): with k = floor(theta / da) the
            // reference's (cos(a) d, sin(a) d) is (x, y) rotated by -k*da and the other copy is
            // that rotated by -da... one polynomial atan2 (only k depends on it), one MUFU
            // sincos, two 2x2 rotations; no hypot, no fmod.
            T sd, cd;
            M::sincos_(da, &sd, &cd);
            SDF_FOR_N {
                const T th = M::atan2_(y[i], x[i]);
                const T k = floorf((float)M::div(th, da));
                T sk, ck;
                M::sincos_(k * da, &sk, &ck);
                const T qx = ck * x[i] + sk * y[i], qy = ck * y[i] - sk * x[i];      // angle a
                if (a == 0) { r0x[i] = qx; r0y[i] = qy; r0z[i] = z[i]; }
                else { ps[a][0][i] = qx; ps[a][1][i] = qy; ps[a][2][i] = z[i]; }
                x[i] = cd * qx + sd * qy; y[i] = cd * qy - sd * qx;                  // angle a - da
            }
        } else {
            SDF_FOR_N {
                const T r = M::hypot_(x[i], y[i]);
                const T ang = M::pymod(M::atan2_(y[i], x[i]), da);
                T s1, c1, s2, c2;
                M::sincos_(ang - da, &s1, &c1);
                M::sincos_(ang, &s2, &c2);
                const T qx = c2 * r, qy = s2 * r;
                if (a == 0) { r0x[i] = qx; r0y[i] = qy; r0z[i] = z[i]; }
                else { ps[a][0][i] = qx; ps[a][1][i] = qy; ps[a][2][i] = z[i]; }
                x[i] = c1 * r; y[i] = s1 * r;
            }
        }
        }

    static __device__ __forceinline__ void REPEAT_CELL(State& st, const T* c, const int a) {  // ------------------------------------------------------------- value ops
        SDF_BIND_STATE;                          // dn.py:101-109
        SDF_FOR_N {
            T* pp[3] = {&x[i], &y[i], &z[i]};
            _Pragma("unroll") for (int k = 0; k < 3; ++k) {
                const T sp = c[k];
                T q = (sp != T(0)) ? m_div(*pp[k], sp) : T(0);
                T idx = M::rint_(q);
                if (c[6] != T(0)) idx = m_min(m_max(idx, -c[3 + k]), c[3 + k]);
                *pp[k] = *pp[k] - sp * (idx + c[7 + k]);
            }
        }
        }

    static __device__ __forceinline__ void MULC(State& st, const T* c, const int a) {  // d3.py:344
        SDF_BIND_STATE; const T m = c[0]; SDF_FOR_N d[i] *= m;     }

    static __device__ __forceinline__ void NEGATE(State& st, const T* c, const int a) {  // dn.py:64
        SDF_BIND_STATE; SDF_FOR_N d[i] = -d[i];     }

    static __device__ __forceinline__ void ADDC(State& st, const T* c, const int a) {  // dn.py:69,74
        SDF_BIND_STATE; const T v = c[0]; SDF_FOR_N d[i] += v;     }

    static __device__ __forceinline__ void SHELL(State& st, const T* c, const int a) {  // dn.py:79
        SDF_BIND_STATE; const T h = c[0] / T(2); SDF_FOR_N d[i] = m_abs(d[i]) - h;     }

    static __device__ __forceinline__ void ADD_POP(State& st, const T* c, const int a) {  // ----------------------------------------------------------------- folds
        SDF_BIND_STATE;                                                          // d3.py:404
        SDF_FOR_N d[i] = d[i] + s0[i];
        if (a > 0) { SDF_FOR_N s0[i] = vs[a - 1][i]; }
        }

    static __device__ __forceinline__ void UNION(State& st, const T* c, const int a) {
        SDF_BIND_STATE;                                // dn.py:14
        SDF_FOR_N d[i] = m_min(s0[i], d[i]);
        if (a > 0) { SDF_FOR_N s0[i] = vs[a - 1][i]; }
        }

    static __device__ __forceinline__ void UNION_K(State& st, const T* c, const int a) {
        SDF_BIND_STATE;                              // dn.py:16-18
        const T K = c[0];
        SDF_FOR_N {
            const T d1 = s0[i], d2 = d[i];
            const T h = m_clip01(T(0.5) + m_div(T(0.5) * (d2 - d1), K));
            const T m = d2 + (d1 - d2) * h;
            d[i] = m - K * h * (T(1) - h);
        }
        if (a > 0) { SDF_FOR_N s0[i] = vs[a - 1][i]; }
        }

    static __device__ __forceinline__ void DIFFERENCE(State& st, const T* c, const int a) {
        SDF_BIND_STATE;                           // dn.py:29
        SDF_FOR_N d[i] = m_max(s0[i], -d[i]);
        if (a > 0) { SDF_FOR_N s0[i] = vs[a - 1][i]; }
        }

    static __device__ __forceinline__ void DIFFERENCE_K(State& st, const T* c, const int a) {
        SDF_BIND_STATE;                         // dn.py:31-33
        const T K = c[0];
        SDF_FOR_N {
            const T d1 = s0[i], d2 = d[i];
            const T h = m_clip01(T(0.5) - m_div(T(0.5) * (d2 + d1), K));
            const T m = d1 + (-d2 - d1) * h;
            d[i] = m + K * h * (T(1) - h);
        }
        if (a > 0) { SDF_FOR_N s0[i] = vs[a - 1][i]; }
        }

    static __device__ __forceinline__ void INTERSECTION(State& st, const T* c, const int a) {
        SDF_BIND_STATE;                         // dn.py:44
        SDF_FOR_N d[i] = m_max(s0[i], d[i]);
        if (a > 0) { SDF_FOR_N s0[i] = vs[a - 1][i]; }
        }

    static __device__ __forceinline__ void INTERSECTION_K(State& st, const T* c, const int a) {
        SDF_BIND_STATE;                       // dn.py:46-48
        const T K = c[0];
        SDF_FOR_N {
            const T d1 = s0[i], d2 = d[i];
            const T h = m_clip01(T(0.5) - m_div(T(0.5) * (d2 - d1), K));
            const T m = d2 + (d1 - d2) * h;
            d[i] = m + K * h * (T(1) - h);
        }
        if (a > 0) { SDF_FOR_N s0[i] = vs[a - 1][i]; }
        }

    static __device__ __forceinline__ void BLEND(State& st, const T* c, const int a) {  // --------------------------------------------------------- stack / slots
        SDF_BIND_STATE;                                // dn.py:58
        const T K = c[0];
        SDF_FOR_N d[i] = K * d[i] + (T(1) - K) * s0[i];
        if (a > 0) { SDF_FOR_N s0[i] = vs[a - 1][i]; }
        }

    static __device__ __forceinline__ void PUSH(State& st, const T* c, const int a) {
        SDF_BIND_STATE;
        if (a > 0) { SDF_FOR_N vs[a - 1][i] = s0[i]; }
        SDF_FOR_N s0[i] = d[i];
        }

    static __device__ __forceinline__ void SAVEP(State& st, const T* c, const int a) {
        SDF_BIND_STATE;
        if (a == 0) { SDF_FOR_N { r0x[i] = x[i]; r0y[i] = y[i]; r0z[i] = z[i]; } }
        else { SDF_FOR_N { ps[a][0][i] = x[i]; ps[a][1][i] = y[i]; ps[a][2][i] = z[i]; } }
        }

    static __device__ __forceinline__ void LOADP(State& st, const T* c, const int a) {
        SDF_BIND_STATE;
        if (a == 0) { SDF_FOR_N { x[i] = r0x[i]; y[i] = r0y[i]; z[i] = r0z[i]; } }
        else { SDF_FOR_N { x[i] = ps[a][0][i]; y[i] = ps[a][1][i]; z[i] = ps[a][2][i]; } }
        }

    // ------------------------------------------------------------------ 2-D primitives (P.xy)
    static __device__ __forceinline__ void CIRCLE(State& st, const T* c, const int a) {   // d2.py:78-79
        SDF_BIND_STATE;
        const T cx = c[0], cy = c[1], r = c[2];
        SDF_FOR_N d[i] = m_len2(x[i] - cx, y[i] - cy) - r;
    }
    static __device__ __forceinline__ void LINE2(State& st, const T* c, const int a) {    // d2.py:85-86
        SDF_BIND_STATE;
        const T nx = c[0], ny = c[1], px = c[2], py = c[3];
        SDF_FOR_N d[i] = (px - x[i]) * nx + (py - y[i]) * ny;
    }
    static __device__ __forceinline__ void RECTANGLE(State& st, const T* c, const int a) {    // d2.py:111-113
        SDF_BIND_STATE;
        const T cx = c[0], cy = c[1], hx = c[2], hy = c[3];
        SDF_FOR_N {
            const T qx = m_abs(x[i] - cx) - hx, qy = m_abs(y[i] - cy) - hy;
            d[i] = m_len2(m_max(qx, T(0)), m_max(qy, T(0))) + m_min(m_max(qx, qy), T(0));
        }
    }
    static __device__ __forceinline__ void ROUNDED_RECTANGLE(State& st, const T* c, const int a) {    // d2.py:122-134
        SDF_BIND_STATE;
        const T hx = c[0], hy = c[1], r0 = c[2], r1 = c[3], r2 = c[4], r3 = c[5];
        SDF_FOR_N {
            const bool xp = x[i] > T(0), yp = y[i] > T(0);
            const T r = xp ? (yp ? r0 : r1) : (yp ? r3 : r2);
            const T qx = m_abs(x[i]) - hx + r, qy = m_abs(y[i]) - hy + r;
            d[i] = m_min(m_max(qx, qy), T(0)) + m_len2(m_max(qx, T(0)), m_max(qy, T(0))) - r;
        }
    }
    static __device__ __forceinline__ void EQUILATERAL_TRIANGLE(State& st, const T* c, const int a) {   // d2.py:139-152
        SDF_BIND_STATE;
        const T k = m_sqrt(T(3));
        SDF_FOR_N {
            T px = m_abs(x[i]) - T(1), py = y[i] + m_div(T(1), k);
            {   // both candidates computed, then selected: no branch, so identical calls can merge (axis sharing)
                const T qx = m_div(px - k * py, T(2)), qy = m_div(-k * px - py, T(2));
                const bool fold = px + k * py > T(0);
                px = fold ? qx : px; py = fold ? qy : py;
            }
            px = px - m_min(m_max(px, T(-2)), T(0));
            d[i] = -m_len2(px, py) * m_sign(py);
        }
    }
    static __device__ __forceinline__ void HEXAGON(State& st, const T* c, const int a) {   // d2.py:157-166
        SDF_BIND_STATE;
        const T r = c[0];
        const T k0 = m_div(m_sqrt(T(3)), T(-2)), k1 = T(0.5), k2 = T(0.57735026918962573);   // tan(pi/6)
        SDF_FOR_N {
            T px = m_abs(x[i]), py = m_abs(y[i]);
            const T h = m_min(k0 * px + k1 * py, T(0));
            px -= T(2) * k0 * h; py -= T(2) * k1 * h;
            px -= m_min(m_max(px, -k2 * r), k2 * r);
            py -= r;
            d[i] = m_len2(px, py) * m_sign(py);
        }
    }
    static __device__ __forceinline__ void ROUNDED_X(State& st, const T* c, const int a) {   // d2.py:171-174
        SDF_BIND_STATE;
        const T w = c[0], r = c[1];
        SDF_FOR_N {
            const T px = m_abs(x[i]), py = m_abs(y[i]);
            const T q = m_min(px + py, w) * T(0.5);
            d[i] = m_len2(px - q, py - q) - r;
        }
    }
    static __device__ __forceinline__ void POLYGON(State& st, const T* c, const int a) {   // d2.py:180-197 (a = n)
        SDF_BIND_STATE;
        const int n = a;
        SDF_FOR_N {
            const T px = x[i], py = y[i];
            T dd = (px - c[0]) * (px - c[0]) + (py - c[1]) * (py - c[1]);
            T s = T(1);
            for (int q = 0; q < n; ++q) {
                const int j = (q + n - 1) % n;
                const T vix = c[2 * q], viy = c[2 * q + 1], vjx = c[2 * j], vjy = c[2 * j + 1];
                const T ex = vjx - vix, ey = vjy - viy;
                const T wx = px - vix, wy = py - viy;
                const T h = m_clip01(m_div(wx * ex + wy * ey, ex * ex + ey * ey));
                const T bx = wx - ex * h, by = wy - ey * h;
                dd = m_min(dd, bx * bx + by * by);
                const bool c1 = py >= viy, c2 = py < vjy, c3 = ex * wy > ey * wx;
                if ((c1 && c2 && c3) || (!c1 && !c2 && !c3)) s = -s;
            }
            d[i] = s * m_sqrt(dd);
        }
    }
    static __device__ __forceinline__ void VESICA(State& st, const T* c, const int a) {   // d2.py:202-208
        SDF_BIND_STATE;
        const T r = c[0], dd = c[1];
        const T b = m_sqrt(r * r - dd * dd);
        SDF_FOR_N {
            const T px = m_abs(x[i]), py = m_abs(y[i]);
            const T va = m_len2(px, py - b), vb = m_len2(px + dd, py) - r;      // both arms, then a select (no branch)
            d[i] = ((py - b) * dd > px * b) ? va : vb;
        }
    }
    // ------------------------------------------------------------------ 2-D <-> 3-D
    static __device__ __forceinline__ void REVOLVE(State& st, const T* c, const int a) {   // d2.py:282-285
        SDF_BIND_STATE;
        const T off = c[0];
        SDF_FOR_N { const T q = m_len2(x[i], y[i]) - off; x[i] = q; y[i] = z[i]; }
    }
    static __device__ __forceinline__ void EXTRUDE_PRE(State& st, const T* c, const int a) {   // d2.py:264 (pushes w)
        SDF_BIND_STATE;
        const T hh = c[0] / T(2);
        if (a > 0) { SDF_FOR_N vs[a - 1][i] = s0[i]; }
        SDF_FOR_N s0[i] = m_abs(z[i]) - hh;
    }
    static __device__ __forceinline__ void EXTRUDE_POST(State& st, const T* c, const int a) {   // d2.py:265
        SDF_BIND_STATE;
        SDF_FOR_N {
            const T w0 = d[i], w1 = s0[i];
            d[i] = m_min(m_max(w0, w1), T(0)) + m_len2(m_max(w0, T(0)), m_max(w1, T(0)));
        }
        if (a > 0) { SDF_FOR_N s0[i] = vs[a - 1][i]; }
    }
    static __device__ __forceinline__ void SLICE_PRE(State& st, const T* c, const int a) {   // d3.py:514
        SDF_BIND_STATE;
        SDF_FOR_N z[i] = T(0);
    }
    static __device__ __forceinline__ void SLICE_SELECT(State& st, const T* c, const int a) {   // d3.py:515-519
        SDF_BIND_STATE;
        SDF_FOR_N { const T A = s0[i], B = -d[i]; d[i] = (A <= T(0)) ? B : A; }
        if (a > 0) { SDF_FOR_N s0[i] = vs[a - 1][i]; }
    }
    // ------------------------------------------------------------------ eased warps / transitions
    static __device__ __forceinline__ void TPARAM_LINEAR(State& st, const T* c, const int a) {   // d3.py:442,464
        SDF_BIND_STATE;
        const T ax = c[3], ay = c[4], az = c[5];
        const T abab = ax * ax + ay * ay + az * az;
        SDF_FOR_N d[i] = m_clip01(m_div((x[i] - c[0]) * ax + (y[i] - c[1]) * ay + (z[i] - c[2]) * az, abab));
    }
    static __device__ __forceinline__ void TPARAM_RADIAL(State& st, const T* c, const int a) {   // d3.py:453-454
        SDF_BIND_STATE;
        const T r0 = c[0], dr = c[1] - c[0];
        SDF_FOR_N d[i] = m_clip01(m_div(M::hypot_(x[i], y[i]) - r0, dr));
    }
    static __device__ __forceinline__ void TPARAM_Z(State& st, const T* c, const int a) {   // d2.py:273
        SDF_BIND_STATE;
        const T h = c[0];
        SDF_FOR_N d[i] = m_min(m_max(m_div(z[i], h), T(-0.5)), T(0.5)) + T(0.5);
    }
    static __device__ __forceinline__ T ease_bounce_out(T t) {                             // ease.py:131-140
        const T aa = m_div(T(121) * t * t, T(16));
        const T bb = (T(363) / T(40) * t * t) - (T(99) / T(10) * t) + T(17) / T(5);
        const T cc = (T(4356) / T(361) * t * t) - (T(35442) / T(1805) * t) + T(16061) / T(1805);
        const T dd = (T(54) / T(5) * t * t) - (T(513) / T(25) * t) + T(268) / T(25);
        return t < T(4) / T(11) ? aa : (t < T(8) / T(11) ? bb : (t < T(9) / T(10) ? cc : dd));
    }
    static __device__ __forceinline__ T ease_eval(int id, T t) {                          // ease.py:3-162
        const T PI = T(3.141592653589793);
        T sn, cs;
        switch (id) {
            case 0: return t;                                                              // linear
            case 1: return t * t;                                                          // in_quad
            case 2: return -t * (t - T(2));                                                // out_quad
            case 3: { const T u = T(2) * t - T(1); return t < T(0.5) ? T(2) * t * t : T(-0.5) * (u * (u - T(2)) - T(1)); }
            case 4: return t * t * t;                                                      // in_cubic
            case 5: { const T u = t - T(1); return u * u * u + T(1); }
            case 6: { const T u = t * T(2), v = u - T(2); return u < T(1) ? T(0.5) * u * u * u : T(0.5) * (v * v * v + T(2)); }
            case 7: return t * t * t * t;                                                  // in_quart
            case 8: { const T u = t - T(1); return -(u * u * u * u - T(1)); }
            case 9: { const T u = t * T(2), v = u - T(2); return u < T(1) ? T(0.5) * u * u * u * u : T(-0.5) * (v * v * v * v - T(2)); }
            case 10: return t * t * t * t * t;                                             // in_quint
            case 11: { const T u = t - T(1); return u * u * u * u * u + T(1); }
            case 12: { const T u = t * T(2), v = u - T(2); return u < T(1) ? T(0.5) * u * u * u * u * u : T(0.5) * (v * v * v * v * v + T(2)); }
            case 13: M::sincos_(m_div(t * PI, T(2)), &sn, &cs); return -cs + T(1);         // in_sine
            case 14: M::sincos_(m_div(t * PI, T(2)), &sn, &cs); return sn;                 // out_sine
            case 15: M::sincos_(PI * t, &sn, &cs); return T(-0.5) * (cs - T(1));           // in_out_sine
            case 16: return t == T(0) ? T(0) : M::exp2_(T(10) * (t - T(1)));               // in_expo
            case 17: return t == T(1) ? T(1) : T(1) - M::exp2_(T(-10) * t);                // out_expo
            case 18: return t == T(0) ? T(0) : (t == T(1) ? T(1) : (t < T(0.5) ? T(0.5) * M::exp2_(T(20) * t - T(10))
                                                                              : T(1) - T(0.5) * M::exp2_(T(-20) * t + T(10))));
            case 19: return T(-1) * (m_sqrt(T(1) - t * t) - T(1));                         // in_circ
            case 20: { const T u = t - T(1); return m_sqrt(T(1) - u * u); }
            case 21: { const T u = t * T(2), v = u - T(2);
                       return u < T(1) ? T(-0.5) * (m_sqrt(T(1) - u * u) - T(1)) : T(0.5) * (m_sqrt(T(1) - v * v) + T(1)); }
            case 22: { const T u = t - T(1), k = T(0.5);                                   // in_elastic
                       M::sincos_(m_div((u - k / T(4)) * (T(2) * PI), k), &sn, &cs); return T(-1) * (M::exp2_(T(10) * u) * sn); }
            case 23: { const T k = T(0.5);                                                 // out_elastic
                       M::sincos_((t - k / T(4)) * m_div(T(2) * PI, k), &sn, &cs); return M::exp2_(T(-10) * t) * sn + T(1); }
            case 24: { const T u = t * T(2), v = u - T(1), k = T(0.5);                     // in_out_elastic
                       M::sincos_(m_div((v - k / T(4)) * T(2) * PI, k), &sn, &cs);
                       return u < T(1) ? T(-0.5) * (M::exp2_(T(10) * v) * sn) : M::exp2_(T(-10) * v) * sn * T(0.5) + T(1); }
            case 25: { const T k = T(1.70158); return t * t * ((k + T(1)) * t - k); }      // in_back
            case 26: { const T k = T(1.70158), u = t - T(1); return u * u * ((k + T(1)) * u + k) + T(1); }
            case 27: { const T k = T(1.70158) * T(1.525), u = t * T(2), v = u - T(2);
                       return u < T(1) ? T(0.5) * (u * u * ((k + T(1)) * u - k)) : T(0.5) * (v * v * ((k + T(1)) * v + k) + T(2)); }
            case 28: return T(1) - ease_bounce_out(T(1) - t);                              // in_bounce
            case 29: return ease_bounce_out(t);                                            // out_bounce
            case 30: return t < T(0.5) ? (T(1) - ease_bounce_out(T(1) - T(2) * t)) * T(0.5)
                                       : ease_bounce_out(T(2) * t - T(1)) * T(0.5) + T(0.5);
            case 31: return t < T(1) ? T(0) : T(1);                                        // in_square
            case 32: return t > T(0) ? T(1) : T(0);                                        // out_square
            case 33: return t < T(0.5) ? T(0) : T(1);                                      // in_out_square
            default: return t;
        }
    }
    static __device__ __forceinline__ void EASE(State& st, const T* c, const int a) {
        SDF_BIND_STATE;
        if (a != 0) { SDF_FOR_N d[i] = ease_eval(a, d[i]); }
    }
    static __device__ __forceinline__ void DISPLACE(State& st, const T* c, const int a) {   // d3.py:444,455
        SDF_BIND_STATE;
        const T vx = c[0], vy = c[1], vz = c[2];
        SDF_FOR_N { x[i] = x[i] + d[i] * vx; y[i] = y[i] + d[i] * vy; z[i] = z[i] + d[i] * vz; }
    }
    static __device__ __forceinline__ void LERP_POP2(State& st, const T* c, const int a) {   // d3.py:466,479
        SDF_BIND_STATE;
        SDF_FOR_N { const T t = d[i], d2 = s0[i], d1 = vs[a][i]; d[i] = t * d2 + (T(1) - t) * d1; }
        if (a > 0) { SDF_FOR_N s0[i] = vs[a - 1][i]; }
    }
    static __device__ __forceinline__ void MIX_POP2(State& st, const T* c, const int a) {   // d2.py:274
        SDF_BIND_STATE;
        SDF_FOR_N { const T t = d[i], d2 = s0[i], d1 = vs[a][i]; d[i] = d1 + (d2 - d1) * t; }
        if (a > 0) { SDF_FOR_N s0[i] = vs[a - 1][i]; }
    }
    static __device__ __forceinline__ void WRAP_PRE(State& st, const T* c, const int a) {   // d3.py:490-497
        SDF_BIND_STATE;
        const T r = c[0];
        const T PI = T(3.141592653589793);
        SDF_FOR_N {
            const T dd = M::hypot_(x[i], y[i]) - r;
            const T ang = M::atan2_(y[i], x[i]);
            d[i] = m_div(ang + PI, T(2) * PI);
            y[i] = -dd;
        }
    }
    static __device__ __forceinline__ void WRAP_POST(State& st, const T* c, const int a) {   // d3.py:498-499
        SDF_BIND_STATE;
        const T x0 = c[0], dx = c[1] - c[0];
        SDF_FOR_N x[i] = x0 + dx * d[i];
    }
};

// Interpreter front end: evaluate the program at N points.  On return d[i] = sdf(x[i], y[i], z[i]);
// x/y/z are clobbered.
template <typename T, int N, bool FAST = false>
__device__ __forceinline__ void sdf_interp(const ProgramImage& img, T (&x)[N], T (&y)[N], T (&z)[N], T (&d)[N]) {
    typedef SdfOps<T, N, FAST> O;
    typename O::State st(x, y, z, d);
    const T* cbase = sdfm::ConstsOf<T>::get(img);
    const int n_code = (int)img.n_code;
    for (int pc = 0; pc < n_code; ++pc) {
        const uint32_t w = img.code[pc];
        const int a = (int)((w >> 8) & 0xffu);
        const T* c = cbase + (w >> 16);
        switch (w & 0xffu) {
#define SDF_CASE(NAME) case OP_##NAME: O::NAME(st, c, a); break;
            SDF_OP_LIST(SDF_CASE)
#undef SDF_CASE
            default: break;
        }
    }
}

// Evaluator handed to the kernels (sdf_kernels.cuh): the interpreter over a ProgramImage.  The
// run-time-generated evaluator (struct SpecEval in the JIT translation unit) has the same shape.
struct InterpEval {
    static constexpr int kN = 2;             // samples per thread in the tile kernel
    static constexpr int kShare = -1;        // no axis sharing: nothing to gain without straight-line code
    const ProgramImage* img;
    template <typename T, int N, bool FAST>
    __device__ __forceinline__ void run(T (&x)[N], T (&y)[N], T (&z)[N], T (&d)[N]) const {
        sdf_interp<T, N, FAST>(*img, x, y, z, d);
    }
};
