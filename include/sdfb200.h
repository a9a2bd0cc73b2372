/*
 * sdfb200.h -- C-ABI of libsdfb200.so: the B200-native replacement for the
 * mesh-generation hot path of fogleman/sdf.
 *
 * The reference is pure Python; the "FFI" a maintainer would bind is therefore a
 * ctypes stub (INTEGRATION.md shows it).  Each entry point names the reference
 * interface it replaces (paths relative to /root/reference).
 *
 * Conventions: every function returns 0 on success or a negative SDFB200_E_*
 * code; sdfb200_last_error() returns a thread-local message for the last
 * failure on the calling thread.  No C++ exception crosses this boundary.
 * Pointers may be host or device pointers (detected with
 * cudaPointerGetAttributes); host buffers are staged through device memory
 * inside the call.  `stream` is a cudaStream_t (NULL = default stream); calls
 * are re-entrant per (program, stream).
 */
#ifndef SDFB200_H
#define SDFB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SDFB200_ABI_VERSION 2

#define SDFB200_OK 0
#define SDFB200_E_INVALID (-1)   /* bad argument / malformed program            */
#define SDFB200_E_CUDA (-2)      /* CUDA runtime error (message has the detail)  */
#define SDFB200_E_CAPACITY (-3)  /* tri_out too small; *n_tri = required count   */
#define SDFB200_E_NOGPU (-4)     /* no CUDA device: there is NO CPU fallback     */

/* precision modes of sdfb200_generate */
#define SDFB200_PREC_F32 0    /* float32 evaluation only                                      */
#define SDFB200_PREC_MIXED 1  /* float32, then float64 re-evaluation of samples with |d| < tau */
#define SDFB200_PREC_F64 2    /* float64 evaluation of every sample (reference arithmetic)    */
#define SDFB200_PREC_EXACT 3  /* fast float32 classification; float64 for samples with |d| < tau and
                                 for every corner of a surface-crossing cube: output == PREC_F64    */

typedef struct sdfb200_prog sdfb200_prog;

int sdfb200_abi_version(void);
const char* sdfb200_last_error(void);
int sdfb200_device_count(int* count);

/*
 * Page-locked, device-mapped host memory for output buffers (cudaHostAlloc).  Replaces nothing in
 * the reference (numpy owns its buffers, core.py:128); a tri_out allocated here is filled by one asynchronous
 * copy at PCIe rate (default) or, with SDFB200_ZEROCOPY=1, written directly by the kernel over PCIe --
 * measured slower on B200 (scattered 4-byte stores) -- instead of the slow pageable path.
 */
int sdfb200_host_alloc(size_t bytes, void** out);
void sdfb200_host_free(void* p);

/*
 * Multi-GPU host output (replaces nothing in the reference): page-lock caller-owned host memory -- sdf_b200/multi.py
 * uses one POSIX shared-memory segment that every rank process maps -- and copy a rank's slab into it at its offset:
 * N ranks move their triangles to the host over N PCIe links at once instead of funnelling them through one GPU.
 */
int sdfb200_host_register(void* p, size_t bytes);
void sdfb200_host_unregister(void* p);
int sdfb200_copy_to_host(void* dst_host, const void* src_device, size_t bytes, int device, void* stream);

/*
 * Replaces: construction of the closure tree by the @sdf3/@op3 factories
 * (sdf/d3.py:48-63) -- the lowered program of sdf_b200/ir.py.
 *   code[n_code]     instruction words  op | a << 8 | const_index << 16
 *   consts[n_consts] float64 constants (converted once to float32 for the f32 path)
 */
int sdfb200_program_create(const uint32_t* code, int n_code, const double* consts, int n_consts,
                           sdfb200_prog** out);
void sdfb200_program_free(sdfb200_prog* prog);

/*
 * Replaces: SDF3.__call__(p) (sdf/d3.py:24-25) on a point block.
 *   xyz: n x 3 points (AoS, the reference's P[N,3] layout of core.py:20-26), d_out: n distances.
 * f32: 12 B/point read + 4 B/point written, streamed through shared memory by
 * bulk async copies.  f64: the reference's float64 arithmetic.
 */
int sdfb200_eval_f32(sdfb200_prog* prog, const float* xyz, int64_t n, float* d_out, int device, void* stream);
/* same as sdfb200_eval_f32 with the approximate-math float32 evaluator (MUFU division / sqrt /
 * sin / cos, polynomial atan2): the classification pass of SDFB200_PREC_EXACT, exposed for benchmarking */
int sdfb200_eval_f32_fast(sdfb200_prog* prog, const float* xyz, int64_t n, float* d_out, int device, void* stream);
int sdfb200_eval_f64(sdfb200_prog* prog, const double* xyz, int64_t n, double* d_out, int device, void* stream);

/*
 * Replaces: the batch loop of core.generate (sdf/core.py:114-141), i.e. per batch
 * _skip (core.py:28-43), _cartesian_product + sdf(P) (core.py:51-52),
 * _marching_cubes (core.py:16-18) and the world mapping (core.py:58-60).
 *
 *   xs/ys/zs     the float64 sample axes exactly as np.arange produced them (core.py:110-112)
 *   batch_size   cubes per batch side (core.py:14)
 *   sparse       core.py:47
 *   precision    SDFB200_PREC_*
 *   tau          PREC_MIXED: samples with |d_f32| < tau are re-evaluated in float64 (sign safety)
 *   zb_begin/end this call's range of Z batch layers (multi-GPU slab split); 0, -1 = all
 *   tri_out      9 float32 per triangle (3 world-space vertices, verts[faces] order), caller-owned,
 *                host or device; NULL = count only
 *   n_tri        triangles produced (or required, with SDFB200_E_CAPACITY)
 *   stats[6]     skipped, empty, nonempty batches (core.py:135-141), samples evaluated in f32,
 *                samples re-evaluated in f64, active cubes
 *   timings_ms[8] optional, CUDA events on `stream`: [0] setup + skip pre-pass, [1] tile pipeline (A+B+C),
 *                [2] whole device section, [3] statistics + result copy, [4] tile evaluation kernel (A; for
 *                precision `exact` the sign-mask evaluation alone), [5] float64 refinement kernel (B),
 *                [6] extraction kernel (C), [7] tile classification kernel (precision `exact`, else 0)
 */
int sdfb200_generate(sdfb200_prog* prog,
                     const double* xs, int64_t nx, const double* ys, int64_t ny, const double* zs, int64_t nz,
                     int batch_size, int sparse, int precision, double tau, int zb_begin, int zb_end,
                     float* tri_out, int64_t tri_capacity, int64_t* n_tri, int64_t stats[6],
                     double timings_ms[8], int device, void* stream);

/*
 * Replaces: core.save() for '.stl' paths (sdf/core.py:152-156) = generate() followed by
 * write_binary_stl's record building (sdf/stl.py:7-20): same as sdfb200_generate, but the result is
 * the file payload -- tri records of 50 bytes {normal f32x3, 3 vertices f32x3, attr u16 = 0}, the
 * normal computed on the device operation by operation like numpy (cross, norm, divide in float32).
 * records_out: host (ideally sdfb200_host_alloc) or device, 50 * tri_capacity bytes.
 */
int sdfb200_generate_stl(sdfb200_prog* prog,
                         const double* xs, int64_t nx, const double* ys, int64_t ny, const double* zs, int64_t nz,
                         int batch_size, int sparse, int precision, double tau,
                         void* records_out, int64_t tri_capacity, int64_t* n_tri, int64_t stats[6],
                         double timings_ms[8], int device, void* stream);

/*
 * Fused multi-GPU gather (replaces nothing in the reference, which is single process): the ROOT rank creates
 * two triangle buffers (step parity) + a few flag words in its HBM and hands the 64-byte CUDA-IPC handle to the
 * other ranks (any transport; sdf_b200/multi.py uses torch.distributed), which map it.  Every rank then runs
 * sdfb200_generate_window / _shared on its share with `which` = step & 1: its extraction kernel reserves output
 * ranges with system-scope atomics on the root's cursor and stores its triangles into the root's buffer over NVLink
 * as coalesced peer writes.  There is NO collective and no host barrier per step -- the ranks meet through flag
 * words in the root's memory:
 *   - before a rank's extraction touches the buffer of parity p it waits (on the device) until the root has released
 *     that buffer for this step (`ready[p]` equals the rank's own step counter for p, kept in its GPU's memory);
 *   - when all of a rank's triangles are written its extraction kernel adds one to `done[p]`;
 *   - the root: sdfb200_shared_wait (device wait for done[p] == world, then the triangle count), consume
 *     sdfb200_shared_buffer(s, p), then sdfb200_shared_release (cursor and done reset, `ready[p]` += 2).
 * A rank can therefore run at most two steps ahead of the root's consumption and never overwrites data in use.
 * Waits give up (the kernel traps, the call returns SDFB200_E_CUDA) after about ten seconds.
 * n_tri_local = triangles this rank contributed.  sdfb200_shared_reset / _count (ABI v1) are gone.
 */
typedef struct sdfb200_shared sdfb200_shared;
int sdfb200_shared_create(int64_t tri_capacity, int device, sdfb200_shared** out, unsigned char handle_out[64]);
int sdfb200_shared_open(const unsigned char handle[64], int64_t tri_capacity, int device, sdfb200_shared** out);
void sdfb200_shared_free(sdfb200_shared* s);
void* sdfb200_shared_buffer(sdfb200_shared* s, int which);
int sdfb200_shared_wait(sdfb200_shared* s, int which, int world, int64_t* n, void* stream);
int sdfb200_shared_release(sdfb200_shared* s, int which, void* stream);
int sdfb200_generate_shared(sdfb200_prog* prog,
                            const double* xs, int64_t nx, const double* ys, int64_t ny, const double* zs, int64_t nz,
                            int batch_size, int sparse, int precision, double tau, int zb_begin, int zb_end,
                            sdfb200_shared* dst, int which, int64_t* n_tri_local, int64_t stats[6],
                            double timings_ms[8], int device, void* stream);

/*
 * Multi-GPU with slabs cut INSIDE a batch layer (replaces nothing in the reference, whose only parallelism is
 * the ThreadPool over independent batches, sdf/core.py:131-133): like sdfb200_generate (dst == NULL) or
 * sdfb200_generate_shared (dst != NULL, `which` = cursor), but the batches of the call are the window
 * [ub_begin, ub_end) of the Z-MAJOR batch order  u = bz * (nbx * nby) + bx * nby + by  instead of whole Z batch
 * layers -- a Z-slab with ragged first and last layers, so that N ranks get equal numbers of ACTIVE batches
 * even when the surface lives in fewer layers than there are ranks.  Batches stay the reference's batches (same
 * lattice slices, same _skip decisions): the union of the windows' triangles is the single-call result.
 */
int sdfb200_generate_window(sdfb200_prog* prog,
                            const double* xs, int64_t nx, const double* ys, int64_t ny, const double* zs, int64_t nz,
                            int batch_size, int sparse, int precision, double tau, int64_t ub_begin, int64_t ub_end,
                            float* tri_out, int64_t tri_capacity, sdfb200_shared* dst, int which,
                            int64_t* n_tri, int64_t stats[6], double timings_ms[8], int device, void* stream);

/*
 * Replaces: core._marching_cubes + the world mapping of core._worker (sdf/core.py:16-18, 58-60) for ONE batch whose
 * values the CALLER computed -- the route for a tree that holds a user-written numpy closure (README "writing your
 * own SDFs"), which cannot be lowered to the CUDA evaluator: the host evaluates that closure on the batch's points
 * exactly like the reference does (children that are library nodes still evaluate on the GPU), and hands the
 * volume over.  volume: float32 [nx][ny][nz] in C order (skimage casts to float32 too), host or device,
 * 1..33 samples per axis; xs / ys / zs: the batch's float64 axis slices (scale = xs[1] - xs[0], offset = xs[0]).
 * A volume without a sign change, or thinner than 2 samples, yields 0 triangles (skimage raises; the reference
 * maps that to an empty batch, core.py:53-56).  `prog` only lends its device context (any program).
 */
int sdfb200_mc_volume(sdfb200_prog* prog, const float* volume, int64_t nx, int64_t ny, int64_t nz,
                      const double* xs, const double* ys, const double* zs, float* tri_out, int64_t tri_capacity,
                      int64_t* n_tri, int device, void* stream);

/*
 * Replaces: core._mesh's  np.unique(points, axis=0, return_inverse=True)  (sdf/core.py:160-164), the welding
 * between generate()'s vertex soup and an indexed mesh (meshio formats).  tris: n_tri x 9 float32 (host or device);
 * verts_out: room for 3 * n_tri x 3 float32, receives the n_verts unique vertices in lexicographic (x, y, z) order
 * (numpy's float comparison: -0.0 welds with +0.0); faces_out: n_tri x 3 int32, the unique-vertex index of every
 * soup vertex (numpy's `inverse`).
 */
int sdfb200_weld(const float* tris, int64_t n_tri, float* verts_out, int32_t* faces_out, int64_t* n_verts,
                 int device, void* stream);

/*
 * Run-time specialisation check (needs no GPU): generate the straight-line evaluator of `prog` for
 * kernel id 0 (skip pre-pass), 1..4 (fused tile kernel, precision 0..3), 5/6 (streamed float32
 * evaluator accurate/fast), 7 (streamed float64) and compile it with NVRTC for sm_100a.  The other
 * entry points do this lazily and fall back to the interpreter kernels when NVRTC is missing or
 * SDFB200_JIT=0.  Replaces nothing in the reference.
 */
int sdfb200_jit_probe(sdfb200_prog* prog, int kernel, int64_t* cubin_bytes);

/*
 * Diagnostic (needs no GPU): how the specialised tile kernel of `precision` (0..3) groups its samples.
 * A thread evaluates `samples_per_thread` consecutive lattice samples along axis `share_axis`
 * (0 = x, 1 = y, 2 = z; -1 = flat mapping), which share the other two coordinates, so every
 * sub-expression of the tree that does not depend on that axis -- a cylinder along it, an extruded
 * 2-D profile, a circular array around it -- is evaluated once per group.  `invariant_fraction` is the
 * estimated share of the evaluation work that does not depend on the best axis.  The grouping only
 * affects speed, never results.  Replaces nothing in the reference (its numpy closures evaluate every
 * node at every point of the batch, sdf/core.py:49-50).
 */
int sdfb200_jit_plan(sdfb200_prog* prog, int precision, int* share_axis, int* samples_per_thread,
                     double* invariant_fraction);

/*
 * Replaces: nothing in the reference (diagnostic / load balancing): the skip
 * flags of every batch (1 = skipped), X-major / Z-fastest batch order of
 * core.py:119.  flags: host pointer, nbx*nby*nbz bytes.
 */
int sdfb200_skip_flags(sdfb200_prog* prog,
                       const double* xs, int64_t nx, const double* ys, int64_t ny, const double* zs, int64_t nz,
                       int batch_size, uint8_t* flags, int device, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SDFB200_H */
