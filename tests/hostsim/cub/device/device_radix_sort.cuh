// Host stand-in for cub::DeviceRadixSort (tests/hostsim): a stable sort of (key, value) pairs by key.
#pragma once
#include <algorithm>
#include <numeric>
#include <vector>
namespace cub {
struct DeviceRadixSort {
    template <class K, class V>
    static cudaError_t SortPairs(void* temp, size_t& temp_bytes, const K* keys_in, K* keys_out, const V* vals_in, V* vals_out,
                                 int n, int begin_bit = 0, int end_bit = (int)sizeof(K) * 8, cudaStream_t = nullptr) {
        if (!temp) { temp_bytes = 256; return cudaSuccess; }
        const K mask = (end_bit - begin_bit >= (int)sizeof(K) * 8) ? ~K(0) : (K)(((K(1) << (end_bit - begin_bit)) - 1) << begin_bit);
        std::vector<int> order((size_t)n);
        std::iota(order.begin(), order.end(), 0);
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return (keys_in[a] & mask) < (keys_in[b] & mask); });
        for (int i = 0; i < n; ++i) { keys_out[i] = keys_in[order[(size_t)i]]; vals_out[i] = vals_in[order[(size_t)i]]; }
        return cudaSuccess;
    }
};
}  // namespace cub
