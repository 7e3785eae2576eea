// TEST INFRASTRUCTURE: host stand-ins for the cub primitives libfplgpu calls (a stable LSD radix sort, exclusive sum, select)
#pragma once
#include "../cuda_runtime.h"
#include <algorithm>
#include <vector>
namespace cub {
struct DeviceRadixSort {
    template <class K, class V>
    static cudaError_t SortPairsDescending(void* tmp, size_t& bytes, const K* kin, K* kout, const V* vin, V* vout, int n, int begin_bit,
                                           int end_bit, cudaStream_t = nullptr) {
        if (!tmp) { bytes = 64; return cudaSuccess; }
        std::vector<int> idx((size_t)n);
        for (int i = 0; i < n; i++) idx[i] = i;
        const K mask = end_bit - begin_bit >= (int)(8 * sizeof(K)) ? ~(K)0 : (K)((((K)1 << (end_bit - begin_bit)) - 1) << begin_bit);
        std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return (kin[a] & mask) > (kin[b] & mask); });
        for (int i = 0; i < n; i++) { kout[i] = kin[idx[i]]; vout[i] = vin[idx[i]]; }
        return cudaSuccess;
    }
};
template <class T> struct CountingInputIterator {
    T base;
    explicit CountingInputIterator(T b) : base(b) {}
    T operator[](int64_t i) const { return base + (T)i; }
};
struct DeviceSelect {
    template <class In, class Out, class Num, class Pred>
    static cudaError_t If(void* tmp, size_t& bytes, In in, Out out, Num num_out, int64_t n, Pred pred, cudaStream_t = nullptr) {
        if (!tmp) { bytes = 64; return cudaSuccess; }
        int64_t k = 0;
        for (int64_t i = 0; i < n; i++) { const auto v = in[i]; if (pred(v)) out[k++] = v; }
        *num_out = k;
        return cudaSuccess;
    }
};
struct DeviceScan {
    template <class I, class O>
    static cudaError_t ExclusiveSum(void* tmp, size_t& bytes, I in, O out, int64_t n, cudaStream_t = nullptr) {
        if (!tmp) { bytes = 64; return cudaSuccess; }
        typename std::remove_reference<decltype(out[0])>::type acc = 0;
        for (int64_t i = 0; i < n; i++) { auto v = in[i]; out[i] = acc; acc += v; }       // in may alias out
        return cudaSuccess;
    }
};
}  // namespace cub
