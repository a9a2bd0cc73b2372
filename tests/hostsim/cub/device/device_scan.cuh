// Host stand-in for cub::DeviceScan (tests/hostsim).
#pragma once
namespace cub {
struct DeviceScan {
    template <class T>
    static cudaError_t InclusiveSum(void* temp, size_t& temp_bytes, const T* in, T* out, int n, cudaStream_t = nullptr) {
        if (!temp) { temp_bytes = 256; return cudaSuccess; }
        T acc = T(0);
        for (int i = 0; i < n; ++i) { acc = (T)(acc + in[i]); out[i] = acc; }
        return cudaSuccess;
    }
};
}  // namespace cub
