// sdfb200_weld.cu -- vertex welding of a triangle soup on the GPU (sm_100a).
//
// Replaces: core._mesh's  points, cells = np.unique(points, axis=0, return_inverse=True)  (sdf/core.py:160-164),
// the step between generate()'s vertex soup and an indexed mesh for the meshio formats: unique vertices in
// lexicographic (x, y, z) order + for every soup vertex the index of its unique vertex.
//
// 3T soup vertices -> order-preserving 32-bit keys per coordinate (-0.0 welded with +0.0 like numpy's float
// comparison) -> three stable least-significant-digit passes (z, then y, then x) of a key/index radix sort ->
// run boundaries -> exclusive scan -> scatter.  The radix sort and the scan are cub::DeviceRadixSort /
// cub::DeviceScan (the CUDA toolkit's header library -- this is not one of the hot kernels of the path); key
// building, boundary detection and the scatter are the kernels below.
#include <cuda_runtime.h>
#include <stdint.h>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "../../include/sdfb200.h"

int sdfb200_fail_(int code, const char* fmt, ...);      // sdfb200.cu: sets the thread-local error string

#define WELD_TRY(expr)                                                                                      \
    do {                                                                                                    \
        cudaError_t err__ = (expr);                                                                         \
        if (err__ != cudaSuccess) {                                                                         \
            rc = sdfb200_fail_(SDFB200_E_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(err__), __FILE__, __LINE__); \
            goto done;                                                                                      \
        }                                                                                                   \
    } while (0)

__device__ __forceinline__ unsigned int weld_key(float f) {
    const unsigned int b = __float_as_uint(f == 0.0f ? 0.0f : f);          // -0.0 == +0.0
    return b ^ ((b >> 31) ? 0xffffffffu : 0x80000000u);                    // ascending as floats
}

__global__ void weld_iota_kernel(unsigned int* idx, long long n) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) idx[i] = (unsigned int)i;
}

// keys[i] = key of coordinate `axis` of soup vertex idx[i]
__global__ void weld_keys_kernel(const float* __restrict__ pts, const unsigned int* __restrict__ idx, unsigned int* __restrict__ keys,
                                 long long n, int axis) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) keys[i] = weld_key(pts[(size_t)idx[i] * 3 + axis]);
}

// first[i] = 1 when sorted vertex i starts a run of equal rows
__global__ void weld_first_kernel(const float* __restrict__ pts, const unsigned int* __restrict__ idx, unsigned int* __restrict__ first,
                                  long long n) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned int f = 1u;
    if (i > 0) {
        const float* a = pts + (size_t)idx[i] * 3;
        const float* b = pts + (size_t)idx[i - 1] * 3;
        f = (a[0] == b[0] && a[1] == b[1] && a[2] == b[2]) ? 0u : 1u;
    }
    first[i] = f;
}

// uid = inclusive scan of first, minus one; the run's first member writes the unique vertex
__global__ void weld_scatter_kernel(const float* __restrict__ pts, const unsigned int* __restrict__ idx,
                                    const unsigned int* __restrict__ first, const unsigned int* __restrict__ incl,
                                    float* __restrict__ verts, int* __restrict__ faces, long long n) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned int u = incl[i] - 1u, src = idx[i];
    faces[src] = (int)u;
    if (first[i]) {
        verts[(size_t)u * 3 + 0] = pts[(size_t)src * 3 + 0];
        verts[(size_t)u * 3 + 1] = pts[(size_t)src * 3 + 1];
        verts[(size_t)u * 3 + 2] = pts[(size_t)src * 3 + 2];
    }
}

static bool weld_is_device(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

extern "C" int sdfb200_weld(const float* tris, int64_t n_tri, float* verts_out, int32_t* faces_out, int64_t* n_verts,
                            int device, void* stream_) {
    if (!n_verts || n_tri < 0 || (n_tri > 0 && (!tris || !verts_out || !faces_out)))
        return sdfb200_fail_(SDFB200_E_INVALID, "weld: bad argument");
    *n_verts = 0;
    if (n_tri == 0) return 0;
    const long long n = (long long)n_tri * 3;
    if (n >= (1ll << 32)) return sdfb200_fail_(SDFB200_E_INVALID, "weld: more than 2^32 soup vertices");
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
        cudaGetLastError();
        return sdfb200_fail_(SDFB200_E_NOGPU, "no CUDA device available; libsdfb200 has no CPU fallback");
    }
    int rc = 0;
    cudaStream_t stream = (cudaStream_t)stream_;
    float *dpts = nullptr, *dverts = nullptr;
    int* dfaces = nullptr;
    unsigned int *keys[2] = {nullptr, nullptr}, *idx[2] = {nullptr, nullptr}, *first = nullptr, *incl = nullptr;
    void* temp = nullptr;
    size_t temp_sort = 0, temp_scan = 0;
    const bool in_dev = weld_is_device(tris), v_dev = weld_is_device(verts_out), f_dev = weld_is_device(faces_out);
    const unsigned int blocks = (unsigned int)((n + 255) / 256);
    unsigned int last = 0;
    if (cudaSetDevice(device) != cudaSuccess) return sdfb200_fail_(SDFB200_E_INVALID, "weld: bad device %d", device);
    if (!in_dev) {
        WELD_TRY(cudaMalloc(&dpts, (size_t)n * 12));
        WELD_TRY(cudaMemcpyAsync(dpts, tris, (size_t)n * 12, cudaMemcpyHostToDevice, stream));
    }
    {
        const float* pts = in_dev ? tris : dpts;
        for (int k = 0; k < 2; ++k) { WELD_TRY(cudaMalloc(&keys[k], (size_t)n * 4)); WELD_TRY(cudaMalloc(&idx[k], (size_t)n * 4)); }
        WELD_TRY(cudaMalloc(&first, (size_t)n * 4));
        WELD_TRY(cudaMalloc(&incl, (size_t)n * 4));
        if (!v_dev) WELD_TRY(cudaMalloc(&dverts, (size_t)n * 12));
        if (!f_dev) WELD_TRY(cudaMalloc(&dfaces, (size_t)n * 4));
        WELD_TRY(cub::DeviceRadixSort::SortPairs(nullptr, temp_sort, keys[0], keys[1], idx[0], idx[1], (int)n, 0, 32, stream));
        WELD_TRY(cub::DeviceScan::InclusiveSum(nullptr, temp_scan, first, incl, (int)n, stream));
        WELD_TRY(cudaMalloc(&temp, temp_sort > temp_scan ? temp_sort : temp_scan));
        weld_iota_kernel<<<blocks, 256, 0, stream>>>(idx[0], n);
        int cur = 0;
        for (int axis = 2; axis >= 0; --axis) {          // least significant key first: z, y, x (each pass is stable)
            weld_keys_kernel<<<blocks, 256, 0, stream>>>(pts, idx[cur], keys[cur], n, axis);
            size_t tb = temp_sort;
            WELD_TRY(cub::DeviceRadixSort::SortPairs(temp, tb, keys[cur], keys[cur ^ 1], idx[cur], idx[cur ^ 1], (int)n, 0, 32, stream));
            cur ^= 1;
        }
        weld_first_kernel<<<blocks, 256, 0, stream>>>(pts, idx[cur], first, n);
        size_t tb = temp_scan;
        WELD_TRY(cub::DeviceScan::InclusiveSum(temp, tb, first, incl, (int)n, stream));
        float* dv = v_dev ? verts_out : dverts;
        int* df = f_dev ? (int*)faces_out : dfaces;
        weld_scatter_kernel<<<blocks, 256, 0, stream>>>(pts, idx[cur], first, incl, dv, df, n);
        WELD_TRY(cudaGetLastError());
        WELD_TRY(cudaMemcpyAsync(&last, incl + (n - 1), 4, cudaMemcpyDeviceToHost, stream));
        WELD_TRY(cudaStreamSynchronize(stream));
        *n_verts = (int64_t)last;
        if (!v_dev) WELD_TRY(cudaMemcpyAsync(verts_out, dverts, (size_t)last * 12, cudaMemcpyDeviceToHost, stream));
        if (!f_dev) WELD_TRY(cudaMemcpyAsync(faces_out, dfaces, (size_t)n * 4, cudaMemcpyDeviceToHost, stream));
        WELD_TRY(cudaStreamSynchronize(stream));
    }
done:
    cudaFree(dpts); cudaFree(dverts); cudaFree(dfaces); cudaFree(keys[0]); cudaFree(keys[1]); cudaFree(idx[0]); cudaFree(idx[1]);
    cudaFree(first); cudaFree(incl); cudaFree(temp);
    return rc;
}
