// TEST INFRASTRUCTURE ONLY -- host stand-ins for the CUB device-wide primitives csrc/ calls (see ../cuda_runtime.h).
// Same call signatures and the same two-phase protocol (d_temp_storage == nullptr -> report temp_storage_bytes).
#pragma once
#include <algorithm>
#include <numeric>
#include <vector>

#include "../cuda_runtime.h"

namespace cub {

namespace detail {
template <class K>
static inline uint64_t key_bits(K k, int begin_bit, int end_bit) {
  // radix order of signed integers: flip the sign bit (CUB's twiddle), then keep bits [begin_bit, end_bit)
  typedef typename std::make_unsigned<K>::type U;
  U u = (U)k;
  if (std::is_signed<K>::value) u ^= (U)1 << (sizeof(K) * 8 - 1);
  uint64_t v = (uint64_t)u;
  if (end_bit < 64) v &= (end_bit <= 0 ? 0ull : ((1ull << end_bit) - 1));
  return begin_bit > 0 ? (v >> begin_bit) : v;
}
}  // namespace detail

struct DeviceRadixSort {
  template <class K, class V>
  static cudaError_t SortPairs(void* temp, size_t& temp_bytes, const K* keys_in, K* keys_out, const V* vals_in, V* vals_out,
                               int n, int begin_bit = 0, int end_bit = sizeof(K) * 8, cudaStream_t = 0) {
    if (!temp) { temp_bytes = 1; return cudaSuccess; }
    std::vector<int> order(n);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
      return detail::key_bits(keys_in[a], begin_bit, end_bit) < detail::key_bits(keys_in[b], begin_bit, end_bit);
    });
    std::vector<K> ko(n);
    std::vector<V> vo(n);
    for (int i = 0; i < n; ++i) { ko[i] = keys_in[order[i]]; vo[i] = vals_in[order[i]]; }
    std::copy(ko.begin(), ko.end(), keys_out);
    std::copy(vo.begin(), vo.end(), vals_out);
    return cudaSuccess;
  }
  template <class K>
  static cudaError_t SortKeys(void* temp, size_t& temp_bytes, const K* keys_in, K* keys_out, int n, int begin_bit = 0,
                              int end_bit = sizeof(K) * 8, cudaStream_t = 0) {
    if (!temp) { temp_bytes = 1; return cudaSuccess; }
    std::vector<K> ko(keys_in, keys_in + n);
    std::stable_sort(ko.begin(), ko.end(), [&](K a, K b) {
      return detail::key_bits(a, begin_bit, end_bit) < detail::key_bits(b, begin_bit, end_bit);
    });
    std::copy(ko.begin(), ko.end(), keys_out);
    return cudaSuccess;
  }
};

struct DeviceSelect {
  template <class InIt, class OutIt, class NumIt, class Pred>
  static cudaError_t If(void* temp, size_t& temp_bytes, InIt in, OutIt out, NumIt num_out, int n, Pred pred, cudaStream_t = 0) {
    if (!temp) { temp_bytes = 1; return cudaSuccess; }
    int m = 0;
    for (int i = 0; i < n; ++i) {
      const auto v = in[i];
      if (pred(v)) out[m++] = v;
    }
    *num_out = m;
    return cudaSuccess;
  }
};

struct DeviceScan {
  template <class InIt, class OutIt>
  static cudaError_t ExclusiveSum(void* temp, size_t& temp_bytes, InIt in, OutIt out, int n, cudaStream_t = 0) {
    if (!temp) { temp_bytes = 1; return cudaSuccess; }
    typename std::remove_reference<decltype(out[0])>::type run = 0;
    for (int i = 0; i < n; ++i) {
      const auto v = in[i];
      out[i] = run;
      run += v;
    }
    return cudaSuccess;
  }
};

}  // namespace cub
