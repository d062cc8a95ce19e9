// CSR / CSC construction, gcn_norm weights and the SpMM split plan.
//
// Replaces torch_sparse.SparseTensor(row, col, sparse_sizes) (reference data/lightgcn_loader.py:65-79),
// its lazily built CSC (csr2csc / colptr used by SPMMSum::backward) and PyG gcn_norm
// (reference model/lightgcn.py:56).  One-off per graph (LightGCN) or once per mini-batch and edge
// type (hetero path); sorting is CUB's device radix sort (stable => deterministic permutations),
// everything else is hand-written.
#include <cub/cub.cuh>
#include <thrust/iterator/counting_iterator.h>
#include <stdarg.h>

#include "common.cuh"

namespace lgb {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static int bits_for(uint64_t max_value_exclusive) {
  int b = 1;
  while (b < 64 && (1ull << b) < max_value_exclusive) ++b;
  return b;
}

// ------------------------------------------------------------------------------------------
// Indices outside [0, n_rows) x [0, n_cols) (e.g. item ids that were not shifted by both_indexes_from_zero) are COUNTED in
// *bad and clamped into range, so that nothing downstream writes out of bounds; the caller turns a non-zero count into the
// error torch_sparse's SparseTensor raises for the same input (lgb_csr_build_check).
__global__ void make_keys_kernel(const int64_t* __restrict__ row, const int64_t* __restrict__ col, int64_t nnz,
                                 int64_t n_rows, int64_t n_cols, int64_t* __restrict__ keys, int32_t* __restrict__ idx,
                                 int32_t* __restrict__ bad) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nnz) {
    int64_t r = row[i], c = col[i];
    if (r < 0 || r >= n_rows || c < 0 || c >= n_cols) {
      atomicAdd(bad, 1);
      r = r < 0 ? 0 : (r >= n_rows ? n_rows - 1 : r);
      c = c < 0 ? 0 : (c >= n_cols ? n_cols - 1 : c);
    }
    keys[i] = r * n_cols + c;
    if (idx) idx[i] = (int32_t)i;
  }
}

// sorted keys -> colidx, rowptr (ind2ptr), optional int64 perm
__global__ void split_keys_kernel(const int64_t* __restrict__ keys, const int32_t* __restrict__ idx, int64_t nnz,
                                  int64_t n_rows, int64_t n_cols, int32_t* __restrict__ rowptr,
                                  int32_t* __restrict__ colidx, int64_t* __restrict__ perm) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i > nnz) return;
  if (nnz == 0) {
    for (int64_t j = 0; j <= n_rows; ++j) rowptr[j] = 0;
    return;
  }
  if (i == nnz) {  // tail: rows after the last non-empty one
    int64_t last = keys[nnz - 1] / n_cols;
    for (int64_t j = last + 1; j <= n_rows; ++j) rowptr[j] = (int32_t)nnz;
    return;
  }
  int64_t k = keys[i];
  int64_t r = k / n_cols;
  colidx[i] = (int32_t)(k - r * n_cols);
  if (perm) perm[i] = idx[i];
  int64_t r_prev = (i == 0) ? -1 : keys[i - 1] / n_cols;
  for (int64_t j = r_prev + 1; j <= r; ++j) rowptr[j] = (int32_t)i;
}

// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int32_t row_of_entry(const int32_t* __restrict__ rowptr, int32_t n_rows, int32_t e) {
  // largest r with rowptr[r] <= e  (upper_bound - 1); empty rows are skipped automatically
  int32_t lo = 0, hi = n_rows;  // invariant: rowptr[lo] <= e < rowptr[hi]
  while (hi - lo > 1) {
    int32_t mid = (lo + hi) >> 1;
    if (rowptr[mid] <= e) lo = mid; else hi = mid;
  }
  return lo;
}

__global__ void iota_colkeys_kernel(const int32_t* __restrict__ colidx, int64_t nnz, int32_t* __restrict__ keys,
                                    int32_t* __restrict__ idx) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nnz) {
    keys[i] = colidx[i];
    idx[i] = (int32_t)i;
  }
}

__global__ void transpose_finish_kernel(const int32_t* __restrict__ keys_sorted, const int32_t* __restrict__ csr2csc,
                                        const int32_t* __restrict__ rowptr, int64_t nnz, int32_t n_rows,
                                        int64_t n_cols, int32_t* __restrict__ colptr, int32_t* __restrict__ rowidx) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i > nnz) return;
  if (nnz == 0) {
    for (int64_t j = 0; j <= n_cols; ++j) colptr[j] = 0;
    return;
  }
  if (i == nnz) {
    for (int64_t j = (int64_t)keys_sorted[nnz - 1] + 1; j <= n_cols; ++j) colptr[j] = (int32_t)nnz;
    return;
  }
  rowidx[i] = row_of_entry(rowptr, n_rows, csr2csc[i]);
  int32_t c = keys_sorted[i];
  int32_t c_prev = (i == 0) ? -1 : keys_sorted[i - 1];
  for (int32_t j = c_prev + 1; j <= c; ++j) colptr[j] = (int32_t)i;
}

__global__ void gather_f32_kernel(const float* __restrict__ src, const int32_t* __restrict__ perm, int64_t n,
                                  float* __restrict__ dst) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[perm[i]];
}

// ------------------------------------------------------------------------------------------
// gcn_norm: dinv = 1/sqrt(deg) (IEEE division and sqrt => bit-identical to ATen's CPU pow(-0.5)), 0 for
// isolated rows; val = (1*dinv[row]) * dinv[col], two successive fp32 multiplies like PyG's mul().
__global__ void dinv_kernel(const int32_t* __restrict__ rowptr, int64_t n, float* __restrict__ dinv) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    float deg = (float)(rowptr[i + 1] - rowptr[i]);
    dinv[i] = deg > 0.f ? __fdiv_rn(1.0f, __fsqrt_rn(deg)) : 0.f;
  }
}
__global__ void gcn_val_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                               const float* __restrict__ dinv, int32_t n, int64_t nnz, float* __restrict__ val) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < nnz) {
    int32_t r = row_of_entry(rowptr, n, (int32_t)e);
    float v = __fmul_rn(1.0f, dinv[r]);
    val[e] = __fmul_rn(v, dinv[colidx[e]]);
  }
}

// ------------------------------------------------------------------------------------------
// SpMM plan
__global__ void plan_count_kernel(const int32_t* __restrict__ rowptr, int64_t n_rows, int32_t chunk,
                                  unsigned long long* __restrict__ counts) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long is_long = 0, tasks = 0;
  if (r < n_rows) {
    int32_t deg = rowptr[r + 1] - rowptr[r];
    if (deg > chunk) {
      is_long = 1;
      tasks = (unsigned long long)((deg + chunk - 1) / chunk);
    }
  }
  // warp-aggregate, then one atomic per warp
  for (int off = 16; off; off >>= 1) {
    is_long += __shfl_xor_sync(FULL_MASK, is_long, off);
    tasks += __shfl_xor_sync(FULL_MASK, tasks, off);
  }
  if ((threadIdx.x & 31) == 0 && is_long) {
    atomicAdd(&counts[0], is_long);
    atomicAdd(&counts[1], tasks);
  }
}

struct IsLongRow {
  const int32_t* rowptr;
  int32_t chunk;
  __device__ __forceinline__ bool operator()(const int32_t& r) const { return rowptr[r + 1] - rowptr[r] > chunk; }
};

__global__ void plan_ntasks_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ long_rows,
                                   int64_t n_long, int32_t chunk, int32_t* __restrict__ ntasks) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_long) {
    int32_t r = long_rows[i];
    ntasks[i] = (rowptr[r + 1] - rowptr[r] + chunk - 1) / chunk;
  } else if (i == n_long) {
    ntasks[i] = 0;
  }
}

__global__ void plan_tasks_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ long_rows,
                                  const int32_t* __restrict__ long_ptr, int64_t n_long, int32_t chunk,
                                  int32_t* __restrict__ task_row, int32_t* __restrict__ task_start,
                                  int32_t* __restrict__ task_end) {
  // one warp per long row
  int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (w >= n_long) return;
  int32_t r = long_rows[w];
  int32_t s = rowptr[r], e = rowptr[r + 1];
  int32_t t0 = long_ptr[w], t1 = long_ptr[w + 1];
  for (int32_t t = t0 + lane; t < t1; t += 32) {
    task_row[t] = r;
    task_start[t] = s + (t - t0) * chunk;
    task_end[t] = min(s + (t - t0 + 1) * chunk, e);
  }
}

__global__ void degree_bucket_kernel(const int32_t* __restrict__ rowptr, int64_t n_rows, int32_t* __restrict__ keys,
                                     int32_t* __restrict__ ids) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r < n_rows) {
    int32_t deg = rowptr[r + 1] - rowptr[r];
    int32_t bucket = deg <= 0 ? 0 : (deg == 1 ? 1 : 33 - __clz(deg - 1));  // 1 + ceil(log2(deg))
    keys[r] = 33 - bucket;  // ascending key == descending degree bucket
    ids[r] = (int32_t)r;
  }
}

static inline unsigned blocks_for(int64_t n, int threads) { return (unsigned)((n + threads - 1) / threads); }

}  // namespace lgb

using namespace lgb;

extern "C" {

int lgb_abi_version(void) { return LGB_ABI_VERSION; }
const char* lgb_last_error(void) { return lgb::g_err; }

int lgb_sm_count(int* out) {
  LGB_REQUIRE(out, LGB_EINVAL, "lgb_sm_count: null out");
  int dev = 0;
  LGB_CUDA(cudaGetDevice(&dev));
  LGB_CUDA(cudaDeviceGetAttribute(out, cudaDevAttrMultiProcessorCount, dev));
  return LGB_OK;
}

// workspace layout: 256 bytes of status (ws[0] = count of out-of-range indices) | keys_in | keys_out | idx_in | idx_out | cub temp
int lgb_csr_build_ws_bytes(int64_t nnz, int64_t n_rows, size_t* bytes) {
  LGB_REQUIRE(bytes && nnz >= 0 && n_rows >= 0, LGB_EINVAL, "lgb_csr_build_ws_bytes: bad argument");
  LGB_REQUIRE(nnz < (1ll << 31) && n_rows < (1ll << 31) - 1, LGB_ERANGE,
              "lgb_csr_build: nnz=%lld / n_rows=%lld exceed the int32 index space", (long long)nnz, (long long)n_rows);
  size_t temp = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, temp, (const int64_t*)nullptr, (int64_t*)nullptr, (const int32_t*)nullptr,
                                  (int32_t*)nullptr, (int)nnz, 0, 64, (cudaStream_t)0);
  *bytes = 256 + 2 * align_up((size_t)nnz * 8) + 2 * align_up((size_t)nnz * 4) + align_up(temp) + 256;
  return LGB_OK;
}

int lgb_csr_build(const int64_t* row, const int64_t* col, int64_t nnz, int64_t n_rows, int64_t n_cols, int32_t* rowptr,
                  int32_t* colidx, int64_t* perm, void* ws, size_t ws_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  LGB_REQUIRE(rowptr && n_rows >= 0 && n_cols >= 0 && nnz >= 0, LGB_EINVAL, "lgb_csr_build: bad argument");
  LGB_REQUIRE(nnz == 0 || (row && col && colidx), LGB_EINVAL, "lgb_csr_build: null row/col/colidx");
  size_t need = 0;
  int rc = lgb_csr_build_ws_bytes(nnz, n_rows, &need);
  if (rc) return rc;
  LGB_REQUIRE(n_cols < (1ll << 31), LGB_ERANGE, "lgb_csr_build: n_cols=%lld exceeds int32", (long long)n_cols);
  LGB_REQUIRE(ws_bytes >= need && (ws || nnz == 0), LGB_EWS, "lgb_csr_build: workspace %zu < %zu", ws_bytes, need);
  LGB_REQUIRE(nnz == 0 || (n_rows > 0 && n_cols > 0), LGB_EINVAL, "lgb_csr_build: %lld entries in an empty %lld x %lld matrix",
              (long long)nnz, (long long)n_rows, (long long)n_cols);
  if (ws) LGB_CUDA(cudaMemsetAsync(ws, 0, 256, stream));   // ws[0]: int32 count of out-of-range indices (lgb_csr_build_check)
  if (nnz == 0) {
    LGB_CUDA(cudaMemsetAsync(rowptr, 0, sizeof(int32_t) * (size_t)(n_rows + 1), stream));
    return LGB_OK;
  }
  int32_t* bad = (int32_t*)ws;
  char* p = (char*)ws + 256;
  int64_t* keys_in = (int64_t*)p;  p += align_up((size_t)nnz * 8);
  int64_t* keys_out = (int64_t*)p; p += align_up((size_t)nnz * 8);
  int32_t* idx_in = (int32_t*)p;   p += align_up((size_t)nnz * 4);
  int32_t* idx_out = (int32_t*)p;  p += align_up((size_t)nnz * 4);
  void* temp = p;
  size_t temp_bytes = ws_bytes - (size_t)(p - (char*)ws);
  const int T = 256;
  make_keys_kernel<<<blocks_for(nnz, T), T, 0, stream>>>(row, col, nnz, n_rows, n_cols, keys_in, idx_in, bad);
  LGB_LAUNCH_CHECK();
  int end_bit = bits_for((uint64_t)n_rows * (uint64_t)(n_cols > 0 ? n_cols : 1));
  LGB_CUDA(cub::DeviceRadixSort::SortPairs(temp, temp_bytes, keys_in, keys_out, idx_in, idx_out, (int)nnz, 0, end_bit,
                                           stream));
  split_keys_kernel<<<blocks_for(nnz + 1, T), T, 0, stream>>>(keys_out, idx_out, nnz, n_rows, n_cols, rowptr, colidx,
                                                               perm);
  LGB_LAUNCH_CHECK();
  return LGB_OK;
}

int lgb_csr_build_check(const void* ws, void* stream_) {
  LGB_REQUIRE(ws, LGB_EINVAL, "lgb_csr_build_check: null workspace");
  int32_t bad = 0;
  LGB_CUDA(cudaMemcpyAsync(&bad, ws, sizeof(bad), cudaMemcpyDeviceToHost, (cudaStream_t)stream_));
  LGB_CUDA(cudaStreamSynchronize((cudaStream_t)stream_));
  LGB_REQUIRE(bad == 0, LGB_ERANGE, "lgb_csr_build: %d entries have a row / column index outside sparse_sizes", (int)bad);
  return LGB_OK;
}

// workspace: keys_in | keys_out | idx_in | cub temp   (idx_out is csr2csc itself)
int lgb_csr_transpose_ws_bytes(int64_t nnz, int64_t n_cols, size_t* bytes) {
  LGB_REQUIRE(bytes && nnz >= 0 && n_cols >= 0, LGB_EINVAL, "lgb_csr_transpose_ws_bytes: bad argument");
  LGB_REQUIRE(nnz < (1ll << 31), LGB_ERANGE, "lgb_csr_transpose: nnz exceeds int32");
  size_t temp = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, temp, (const int32_t*)nullptr, (int32_t*)nullptr, (const int32_t*)nullptr,
                                  (int32_t*)nullptr, (int)nnz, 0, 32, (cudaStream_t)0);
  *bytes = 3 * align_up((size_t)nnz * 4) + align_up(temp) + 256;
  return LGB_OK;
}

int lgb_csr_transpose(const int32_t* rowptr, const int32_t* colidx, int64_t n_rows, int64_t n_cols, int64_t nnz,
                      int32_t* colptr, int32_t* rowidx, int32_t* csr2csc, void* ws, size_t ws_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  LGB_REQUIRE(rowptr && colptr && n_rows >= 0 && n_cols >= 0 && nnz >= 0, LGB_EINVAL, "lgb_csr_transpose: bad argument");
  LGB_REQUIRE(nnz == 0 || (colidx && rowidx && csr2csc), LGB_EINVAL, "lgb_csr_transpose: null array");
  LGB_REQUIRE(n_rows < (1ll << 31) - 1 && n_cols < (1ll << 31) - 1, LGB_ERANGE, "lgb_csr_transpose: size exceeds int32");
  size_t need = 0;
  int rc = lgb_csr_transpose_ws_bytes(nnz, n_cols, &need);
  if (rc) return rc;
  LGB_REQUIRE(ws_bytes >= need && (ws || nnz == 0), LGB_EWS, "lgb_csr_transpose: workspace %zu < %zu", ws_bytes, need);
  if (nnz == 0) {
    LGB_CUDA(cudaMemsetAsync(colptr, 0, sizeof(int32_t) * (size_t)(n_cols + 1), stream));
    return LGB_OK;
  }
  char* p = (char*)ws;
  int32_t* keys_in = (int32_t*)p;  p += align_up((size_t)nnz * 4);
  int32_t* keys_out = (int32_t*)p; p += align_up((size_t)nnz * 4);
  int32_t* idx_in = (int32_t*)p;   p += align_up((size_t)nnz * 4);
  void* temp = p;
  size_t temp_bytes = ws_bytes - (size_t)(p - (char*)ws);
  const int T = 256;
  iota_colkeys_kernel<<<blocks_for(nnz, T), T, 0, stream>>>(colidx, nnz, keys_in, idx_in);
  LGB_LAUNCH_CHECK();
  // stable sort by column keeps the CSR (row-major) order inside every column == argsort(col*M + row)
  LGB_CUDA(cub::DeviceRadixSort::SortPairs(temp, temp_bytes, keys_in, keys_out, idx_in, csr2csc, (int)nnz, 0,
                                           bits_for((uint64_t)(n_cols > 0 ? n_cols : 1)), stream));
  transpose_finish_kernel<<<blocks_for(nnz + 1, T), T, 0, stream>>>(keys_out, csr2csc, rowptr, nnz, (int32_t)n_rows,
                                                                     n_cols, colptr, rowidx);
  LGB_LAUNCH_CHECK();
  return LGB_OK;
}

int lgb_gather_f32(const float* src, const int32_t* perm, int64_t n, float* dst, void* stream) {
  LGB_REQUIRE(n >= 0 && (n == 0 || (src && perm && dst)), LGB_EINVAL, "lgb_gather_f32: bad argument");
  if (n == 0) return LGB_OK;
  gather_f32_kernel<<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(src, perm, n, dst);
  LGB_LAUNCH_CHECK();
  return LGB_OK;
}

int lgb_gcn_norm(const int32_t* rowptr, const int32_t* colidx, int64_t n, int64_t nnz, float* dinv, float* val,
                 void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  LGB_REQUIRE(rowptr && dinv && n >= 0 && nnz >= 0 && (nnz == 0 || (colidx && val)), LGB_EINVAL,
              "lgb_gcn_norm: bad argument");
  LGB_REQUIRE(n < (1ll << 31) - 1 && nnz < (1ll << 31), LGB_ERANGE, "lgb_gcn_norm: size exceeds int32");
  if (n > 0) {
    dinv_kernel<<<blocks_for(n, 256), 256, 0, stream>>>(rowptr, n, dinv);
    LGB_LAUNCH_CHECK();
  }
  if (nnz > 0) {
    gcn_val_kernel<<<blocks_for(nnz, 256), 256, 0, stream>>>(rowptr, colidx, dinv, (int32_t)n, nnz, val);
    LGB_LAUNCH_CHECK();
  }
  return LGB_OK;
}

int lgb_gcn_values(const int32_t* rowptr, const int32_t* colidx, int64_t n, int64_t nnz, const float* dinv, float* val,
                   void* stream) {
  LGB_REQUIRE(rowptr && dinv && n >= 0 && nnz >= 0 && (nnz == 0 || (colidx && val)), LGB_EINVAL,
              "lgb_gcn_values: bad argument");
  LGB_REQUIRE(n < (1ll << 31) - 1 && nnz < (1ll << 31), LGB_ERANGE, "lgb_gcn_values: size exceeds int32");
  if (nnz > 0) {
    gcn_val_kernel<<<blocks_for(nnz, 256), 256, 0, (cudaStream_t)stream>>>(rowptr, colidx, dinv, (int32_t)n, nnz, val);
    LGB_LAUNCH_CHECK();
  }
  return LGB_OK;
}

// ---- plan ---------------------------------------------------------------------------------
int lgb_spmm_plan_ws_bytes(int64_t n_rows, size_t* bytes) {
  LGB_REQUIRE(bytes && n_rows >= 0, LGB_EINVAL, "lgb_spmm_plan_ws_bytes: bad argument");
  LGB_REQUIRE(n_rows < (1ll << 31) - 1, LGB_ERANGE, "lgb_spmm_plan: n_rows exceeds int32");
  size_t t1 = 0, t2 = 0;
  thrust::counting_iterator<int32_t> it(0);
  cub::DeviceSelect::If(nullptr, t1, it, (int32_t*)nullptr, (int32_t*)nullptr, (int)n_rows, IsLongRow{nullptr, 0},
                        (cudaStream_t)0);
  cub::DeviceScan::ExclusiveSum(nullptr, t2, (int32_t*)nullptr, (int32_t*)nullptr, (int)n_rows + 1, (cudaStream_t)0);
  *bytes = 256 + align_up(t1 > t2 ? t1 : t2) + 256;
  return LGB_OK;
}

int lgb_spmm_plan_count(const int32_t* rowptr, int64_t n_rows, int32_t chunk, int64_t* counts_host, void* ws,
                        size_t ws_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  LGB_REQUIRE(rowptr && counts_host && n_rows >= 0 && chunk > 0, LGB_EINVAL, "lgb_spmm_plan_count: bad argument");
  LGB_REQUIRE(ws && ws_bytes >= 256, LGB_EWS, "lgb_spmm_plan_count: workspace too small");
  unsigned long long* counts = (unsigned long long*)ws;
  LGB_CUDA(cudaMemsetAsync(counts, 0, 16, stream));
  if (n_rows > 0) {
    plan_count_kernel<<<blocks_for(n_rows, 256), 256, 0, stream>>>(rowptr, n_rows, chunk, counts);
    LGB_LAUNCH_CHECK();
  }
  unsigned long long h[2] = {0, 0};
  LGB_CUDA(cudaMemcpyAsync(h, counts, 16, cudaMemcpyDeviceToHost, stream));
  LGB_CUDA(cudaStreamSynchronize(stream));
  counts_host[0] = (int64_t)h[0];
  counts_host[1] = (int64_t)h[1];
  return LGB_OK;
}

int lgb_spmm_plan_fill(const int32_t* rowptr, int64_t n_rows, int32_t chunk, int64_t n_long, int64_t n_tasks,
                       int32_t* long_rows, int32_t* long_ptr, int32_t* task_row, int32_t* task_start, int32_t* task_end,
                       void* ws, size_t ws_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  LGB_REQUIRE(rowptr && n_rows >= 0 && chunk > 0 && n_long >= 0 && n_tasks >= 0, LGB_EINVAL,
              "lgb_spmm_plan_fill: bad argument");
  if (n_long == 0) return LGB_OK;
  LGB_REQUIRE(long_rows && long_ptr && task_row && task_start && task_end, LGB_EINVAL, "lgb_spmm_plan_fill: null output");
  size_t need = 0;
  int rc = lgb_spmm_plan_ws_bytes(n_rows, &need);
  if (rc) return rc;
  LGB_REQUIRE(ws && ws_bytes >= need, LGB_EWS, "lgb_spmm_plan_fill: workspace %zu < %zu", ws_bytes, need);
  int32_t* d_num = (int32_t*)ws;
  void* temp = (char*)ws + 256;
  size_t temp_bytes = ws_bytes - 256;
  thrust::counting_iterator<int32_t> it(0);
  LGB_CUDA(cub::DeviceSelect::If(temp, temp_bytes, it, long_rows, d_num, (int)n_rows, IsLongRow{rowptr, chunk}, stream));
  plan_ntasks_kernel<<<blocks_for(n_long + 1, 256), 256, 0, stream>>>(rowptr, long_rows, n_long, chunk, long_ptr);
  LGB_LAUNCH_CHECK();
  LGB_CUDA(cub::DeviceScan::ExclusiveSum(temp, temp_bytes, long_ptr, long_ptr, (int)n_long + 1, stream));
  plan_tasks_kernel<<<blocks_for(n_long * 32, 256), 256, 0, stream>>>(rowptr, long_rows, long_ptr, n_long, chunk,
                                                                       task_row, task_start, task_end);
  LGB_LAUNCH_CHECK();
  return LGB_OK;
}

int lgb_degree_order_ws_bytes(int64_t n_rows, size_t* bytes) {
  LGB_REQUIRE(bytes && n_rows >= 0, LGB_EINVAL, "lgb_degree_order_ws_bytes: bad argument");
  LGB_REQUIRE(n_rows < (1ll << 31) - 1, LGB_ERANGE, "lgb_degree_order: n_rows exceeds int32");
  size_t temp = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, temp, (const int32_t*)nullptr, (int32_t*)nullptr, (const int32_t*)nullptr,
                                  (int32_t*)nullptr, (int)n_rows, 0, 6, (cudaStream_t)0);
  *bytes = 3 * align_up((size_t)n_rows * 4) + align_up(temp) + 256;
  return LGB_OK;
}

int lgb_degree_order(const int32_t* rowptr, int64_t n_rows, int32_t* row_order, void* ws, size_t ws_bytes,
                     void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  LGB_REQUIRE(rowptr && n_rows >= 0 && (n_rows == 0 || row_order), LGB_EINVAL, "lgb_degree_order: bad argument");
  if (n_rows == 0) return LGB_OK;
  size_t need = 0;
  int rc = lgb_degree_order_ws_bytes(n_rows, &need);
  if (rc) return rc;
  LGB_REQUIRE(ws && ws_bytes >= need, LGB_EWS, "lgb_degree_order: workspace %zu < %zu", ws_bytes, need);
  char* p = (char*)ws;
  int32_t* keys_in = (int32_t*)p;  p += align_up((size_t)n_rows * 4);
  int32_t* keys_out = (int32_t*)p; p += align_up((size_t)n_rows * 4);
  int32_t* ids_in = (int32_t*)p;   p += align_up((size_t)n_rows * 4);
  void* temp = p;
  size_t temp_bytes = ws_bytes - (size_t)(p - (char*)ws);
  degree_bucket_kernel<<<blocks_for(n_rows, 256), 256, 0, stream>>>(rowptr, n_rows, keys_in, ids_in);
  LGB_LAUNCH_CHECK();
  LGB_CUDA(cub::DeviceRadixSort::SortPairs(temp, temp_bytes, keys_in, keys_out, ids_in, row_order, (int)n_rows, 0, 6,
                                           stream));
  return LGB_OK;
}

}  // extern "C"
