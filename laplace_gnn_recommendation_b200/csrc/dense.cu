// Weight / bias gradient of a Linear layer over MANY rows and FEW features: dW[o,i] = sum_n dY[n,o] * X[n,i],
// db[o] = sum_n dY[n,o]   (the backward of SAGEConv's lin_l / lin_r and of the edge decoder's Linear layers,
// model/layers.py:9-24, model/encoder_decoder.py:55-72 -> torch.nn.Linear's AddmmBackward).
//
// Why a kernel of its own: torch.profiler on the ranking step (profiles/r2c_hetero_profile_*.txt) shows this contraction as the
// top device-time item of the whole step -- cuBLAS picks cutlass_80_simt_sgemm_64x64_8x5_nt for the fp32 [out x N] x [N x in]
// product, i.e. FOUR CTAs for out, in <= 128 walking N = 10^4..10^5 rows sequentially (66-207 us per call, 37 % of the L batch),
// and the bias gradient is a second full pass over dY (aten::sum, another 8 %).  The shape wants split-K: here the row range is
// cut into S slices so that tiles x S covers the machine, every CTA accumulates a 64 x 64 tile of its slice in registers
// (fp32 FMA on the CUDA cores: TF32 tensor cores would break the rtol 1e-5 parity with the fp32 reference), the bias column sums
// ride along in the tile column 0 CTAs, and a second kernel adds the S partial tiles in a fixed order (deterministic).
#include <algorithm>

#include "common.cuh"

namespace lgb {

constexpr int WG_T = 64;         // tile edge (outputs and inputs)
constexpr int WG_K = 16;         // rows per shared-memory stage
constexpr int WG_THREADS = 256;  // 16 x 16 threads, 4 x 4 results each

template <bool VEC>
__device__ __forceinline__ void wg_load(const float* __restrict__ src, int64_t row, int64_t n1, int ld, int c0, float (*dst)[WG_T]) {
  // 16 rows x 64 columns = 256 threads x 4 consecutive columns
  const int r = threadIdx.x / 16, c = (threadIdx.x % 16) * 4;
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  const int64_t n = row + r;
  if (n < n1) {
    const float* p = src + n * ld + c0 + c;
    if (VEC) {
      if (c0 + c + 3 < ld) v = *reinterpret_cast<const float4*>(p);
      else {
        if (c0 + c < ld) v.x = p[0];
        if (c0 + c + 1 < ld) v.y = p[1];
        if (c0 + c + 2 < ld) v.z = p[2];
      }
    } else {
      if (c0 + c < ld) v.x = p[0];
      if (c0 + c + 1 < ld) v.y = p[1];
      if (c0 + c + 2 < ld) v.z = p[2];
      if (c0 + c + 3 < ld) v.w = p[3];
    }
  }
  *reinterpret_cast<float4*>(&dst[r][c]) = v;
}

template <bool VEC>
__global__ void __launch_bounds__(WG_THREADS) linear_wgrad_kernel(const float* __restrict__ X, const float* __restrict__ dY, int64_t N,
                                                                 int in, int out, int64_t rows_per_split, float* __restrict__ partial,
                                                                 float* __restrict__ bpartial) {
  __shared__ __align__(16) float As[WG_K][WG_T];   // dY stage: [row][output]
  __shared__ __align__(16) float Bs[WG_K][WG_T];   // X stage:  [row][input]
  const int i0 = blockIdx.x * WG_T, o0 = blockIdx.y * WG_T;
  const int64_t n0 = (int64_t)blockIdx.z * rows_per_split;
  const int64_t n1 = n0 + rows_per_split < N ? n0 + rows_per_split : N;
  const int ty = threadIdx.x / 16, tx = threadIdx.x % 16;
  const bool do_bias = bpartial != nullptr && blockIdx.x == 0 && tx == 0;
  float acc[4][4];
  float bacc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;
  for (int64_t n = n0; n < n1; n += WG_K) {
    wg_load<VEC>(dY, n, n1, out, o0, As);
    wg_load<VEC>(X, n, n1, in, i0, Bs);
    __syncthreads();
#pragma unroll
    for (int k = 0; k < WG_K; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int r = 0; r < 4; ++r) {
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = fmaf(av[r], bv[c], acc[r][c]);
        if (do_bias) bacc[r] += av[r];
      }
    }
    __syncthreads();
  }
  float* P = partial + (size_t)blockIdx.z * out * in;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int o = o0 + ty * 4 + r;
    if (o >= out) continue;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int i = i0 + tx * 4 + c;
      if (i < in) P[(size_t)o * in + i] = acc[r][c];
    }
    if (do_bias) bpartial[(size_t)blockIdx.z * out + o] = bacc[r];
  }
}

// out[j] = sum_s partial[s][j], s ascending (fixed order)
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* __restrict__ partial, int S, int64_t n, float* __restrict__ outp) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  float a = 0.f;
  for (int s = 0; s < S; ++s) a += partial[(size_t)s * n + j];
  outp[j] = a;
}

static int wgrad_splits(int64_t N, int in, int out) {
  const int64_t tiles = (int64_t)((in + WG_T - 1) / WG_T) * ((out + WG_T - 1) / WG_T);
  int64_t S = (2 * 148 + tiles - 1) / tiles;                       // ~two CTAs per SM
  S = std::min<int64_t>(S, (N + 255) / 256);                       // at least 256 rows per slice
  return (int)std::max<int64_t>(S, 1);
}

}  // namespace lgb

using namespace lgb;

extern "C" {

int lgb_linear_wgrad_ws_bytes(int64_t N, int32_t in, int32_t out, size_t* bytes) {
  LGB_REQUIRE(bytes && N >= 0 && in > 0 && out > 0, LGB_EINVAL, "lgb_linear_wgrad_ws_bytes: bad argument");
  *bytes = (size_t)wgrad_splits(N, in, out) * ((size_t)out * in + out) * sizeof(float) + 256;
  return LGB_OK;
}

int lgb_linear_wgrad(const float* X, const float* dY, int64_t N, int32_t in, int32_t out, float* dW, float* db, void* ws,
                     size_t ws_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  LGB_REQUIRE(N >= 0 && in > 0 && out > 0 && dW && (N == 0 || (X && dY)), LGB_EINVAL, "lgb_linear_wgrad: bad argument");
  if (N == 0) {
    LGB_CUDA(cudaMemsetAsync(dW, 0, sizeof(float) * (size_t)out * in, stream));
    if (db) LGB_CUDA(cudaMemsetAsync(db, 0, sizeof(float) * (size_t)out, stream));
    return LGB_OK;
  }
  size_t need = 0;
  lgb_linear_wgrad_ws_bytes(N, in, out, &need);
  LGB_REQUIRE(ws && ws_bytes >= need, LGB_EWS, "lgb_linear_wgrad: workspace %zu < %zu", ws_bytes, need);
  const int S = wgrad_splits(N, in, out);
  const int64_t rows = ((N + S - 1) / S + WG_K - 1) / WG_K * WG_K;
  float* partial = (float*)ws;
  float* bpartial = db ? partial + (size_t)S * out * in : nullptr;
  const dim3 grid((in + WG_T - 1) / WG_T, (out + WG_T - 1) / WG_T, S);
  const bool vec = in % 4 == 0 && out % 4 == 0 && ((((uintptr_t)X | (uintptr_t)dY) & 15) == 0);
  if (vec)
    linear_wgrad_kernel<true><<<grid, WG_THREADS, 0, stream>>>(X, dY, N, in, out, rows, partial, bpartial);
  else
    linear_wgrad_kernel<false><<<grid, WG_THREADS, 0, stream>>>(X, dY, N, in, out, rows, partial, bpartial);
  LGB_LAUNCH_CHECK();
  const int64_t n = (int64_t)out * in;
  wgrad_reduce_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(partial, S, n, dW);
  LGB_LAUNCH_CHECK();
  if (db) {
    wgrad_reduce_kernel<<<(unsigned)((out + 255) / 256), 256, 0, stream>>>(bpartial, S, out, db);
    LGB_LAUNCH_CHECK();
  }
  return LGB_OK;
}

}  // extern "C"
