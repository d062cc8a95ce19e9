// Edge decoder kernels for sm_100a (reference model/encoder_decoder.py:55-72).
//
//  * concat: the reference's z = cat(z_user[row], z_item[col]) (two index_selects + cat, then cuBLAS
//    Linear layers) as one gather kernel, with an atomic scatter backward;
//  * dot: score_e = <z_user[row_e], z_item[col_e]>, the decoder BASELINE.json's north_star names,
//    forward and hand-written backward.
// One warp per label edge; lanes stride over 128-bit chunks of the row (scalar path when d % 4 != 0).
#include "common.cuh"

namespace lgb {

constexpr int DEC_WARPS = 8;

__global__ void __launch_bounds__(DEC_WARPS * 32)
edge_concat_fwd_kernel(const float* __restrict__ zu, const float* __restrict__ zi, const int64_t* __restrict__ row,
                       const int64_t* __restrict__ col, int64_t L, int du, int di, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t e = (int64_t)blockIdx.x * DEC_WARPS + (threadIdx.x >> 5);
  if (e >= L) return;
  const int64_t r = row[e], c = col[e];
  float* o = out + e * (du + di);
  if (((du | di) & 3) == 0) {
    const float4* a = (const float4*)(zu + r * du);
    const float4* b = (const float4*)(zi + c * di);
    float4* o4 = (float4*)o;
    for (int f = lane; f < du / 4; f += 32) st_f4(o4 + f, ld_gather_f4(a + f));
    for (int f = lane; f < di / 4; f += 32) st_f4(o4 + du / 4 + f, ld_gather_f4(b + f));
  } else {
    for (int f = lane; f < du; f += 32) o[f] = zu[r * du + f];
    for (int f = lane; f < di; f += 32) o[du + f] = zi[c * di + f];
  }
}

__global__ void __launch_bounds__(DEC_WARPS * 32)
edge_concat_bwd_kernel(const float* __restrict__ gout, const int64_t* __restrict__ row, const int64_t* __restrict__ col,
                       int64_t L, int du, int di, float* __restrict__ dzu, float* __restrict__ dzi) {
  const int lane = threadIdx.x & 31;
  const int64_t e = (int64_t)blockIdx.x * DEC_WARPS + (threadIdx.x >> 5);
  if (e >= L) return;
  const int64_t r = row[e], c = col[e];
  const float* g = gout + e * (du + di);
  if (((du | di) & 3) == 0) {
    const float4* g4 = (const float4*)g;
    if (dzu) for (int f = lane; f < du / 4; f += 32) red_add_f4((float4*)(dzu + r * du) + f, ld_stream_f4(g4 + f));
    if (dzi) for (int f = lane; f < di / 4; f += 32) red_add_f4((float4*)(dzi + c * di) + f, ld_stream_f4(g4 + du / 4 + f));
  } else {
    if (dzu) for (int f = lane; f < du; f += 32) atomicAdd(dzu + r * du + f, g[f]);
    if (dzi) for (int f = lane; f < di; f += 32) atomicAdd(dzi + c * di + f, g[du + f]);
  }
}

__global__ void __launch_bounds__(DEC_WARPS * 32)
edge_dot_fwd_kernel(const float* __restrict__ zu, const float* __restrict__ zi, const int64_t* __restrict__ row,
                    const int64_t* __restrict__ col, int64_t L, int d, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t e = (int64_t)blockIdx.x * DEC_WARPS + (threadIdx.x >> 5);
  if (e >= L) return;
  const int64_t r = row[e], c = col[e];
  float acc = 0.f;
  if ((d & 3) == 0) {
    const float4* a = (const float4*)(zu + r * d);
    const float4* b = (const float4*)(zi + c * d);
    for (int f = lane; f < d / 4; f += 32) {
      const float4 x = ld_gather_f4(a + f), y = ld_gather_f4(b + f);
      acc += x.x * y.x + x.y * y.y + x.z * y.z + x.w * y.w;
    }
  } else {
    for (int f = lane; f < d; f += 32) acc += zu[r * d + f] * zi[c * d + f];
  }
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(FULL_MASK, acc, off);
  if (lane == 0) out[e] = acc;
}

__global__ void __launch_bounds__(DEC_WARPS * 32)
edge_dot_bwd_kernel(const float* __restrict__ zu, const float* __restrict__ zi, const int64_t* __restrict__ row,
                    const int64_t* __restrict__ col, const float* __restrict__ gout, int64_t L, int d,
                    float* __restrict__ dzu, float* __restrict__ dzi) {
  const int lane = threadIdx.x & 31;
  const int64_t e = (int64_t)blockIdx.x * DEC_WARPS + (threadIdx.x >> 5);
  if (e >= L) return;
  const int64_t r = row[e], c = col[e];
  const float g = gout[e];
  if ((d & 3) == 0) {
    const float4* a = (const float4*)(zu + r * d);
    const float4* b = (const float4*)(zi + c * d);
    for (int f = lane; f < d / 4; f += 32) {
      const float4 x = ld_gather_f4(a + f), y = ld_gather_f4(b + f);
      if (dzu) red_add_f4((float4*)(dzu + r * d) + f, make_float4(g * y.x, g * y.y, g * y.z, g * y.w));
      if (dzi) red_add_f4((float4*)(dzi + c * d) + f, make_float4(g * x.x, g * x.y, g * x.z, g * x.w));
    }
  } else {
    for (int f = lane; f < d; f += 32) {
      if (dzu) atomicAdd(dzu + r * d + f, g * zi[c * d + f]);
      if (dzi) atomicAdd(dzi + c * d + f, g * zu[r * d + f]);
    }
  }
}

// Inference form of the two-layer concat-MLP decoder: out[e] = b2 + sum_h w2[h] * relu(pu[row[e], h] + pi[col[e], h]), pu / pi
// the per-node halves of the first Linear (bias folded into pu by the caller).  One warp per label edge, one 128-bit load per
// lane and 32 * 4 hidden units per step (H = 128: exactly one step), shuffle butterfly for the final sum.
__global__ void __launch_bounds__(DEC_WARPS * 32)
edge_mlp2_fwd_kernel(const float* __restrict__ pu, const float* __restrict__ pi, const int64_t* __restrict__ row,
                     const int64_t* __restrict__ col, int64_t L, int H, const float* __restrict__ w2,
                     const float* __restrict__ b2, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t e = (int64_t)blockIdx.x * DEC_WARPS + (threadIdx.x >> 5);
  if (e >= L) return;
  const int64_t r = row[e], c = col[e];
  float acc = 0.f;
  if ((H & 3) == 0) {
    const float4* a = (const float4*)(pu + r * H);
    const float4* b = (const float4*)(pi + c * H);
    const float4* w = (const float4*)w2;
    for (int f = lane; f < H / 4; f += 32) {
      const float4 x = ld_gather_f4(a + f), y = ld_gather_f4(b + f), ww = w[f];
      acc = fmaf(ww.x, fmaxf(x.x + y.x, 0.f), acc);
      acc = fmaf(ww.y, fmaxf(x.y + y.y, 0.f), acc);
      acc = fmaf(ww.z, fmaxf(x.z + y.z, 0.f), acc);
      acc = fmaf(ww.w, fmaxf(x.w + y.w, 0.f), acc);
    }
  } else {
    for (int f = lane; f < H; f += 32) acc = fmaf(w2[f], fmaxf(pu[r * H + f] + pi[c * H + f], 0.f), acc);
  }
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(FULL_MASK, acc, off);
  if (lane == 0) out[e] = acc + (b2 ? b2[0] : 0.f);
}

// the kernels above take the 128-bit path whenever the widths are multiples of four: the tables must then be 16-byte aligned
static inline bool edge_aligned(int wa, int wb, const void* a, const void* b, const void* c, const void* d2) {
  return ((wa | wb) & 3) != 0 || ((((uintptr_t)a | (uintptr_t)b | (uintptr_t)c | (uintptr_t)d2) & 15) == 0);
}

static inline unsigned edge_blocks(int64_t L) { return (unsigned)((L + DEC_WARPS - 1) / DEC_WARPS); }

}  // namespace lgb

using namespace lgb;

extern "C" {

int lgb_edge_concat_fwd(const float* zu, const float* zi, const int64_t* row, const int64_t* col, int64_t L, int32_t du,
                        int32_t di, float* out, void* stream) {
  LGB_REQUIRE(L >= 0 && du > 0 && di > 0 && (L == 0 || (zu && zi && row && col && out)), LGB_EINVAL,
              "lgb_edge_concat_fwd: bad argument");
  if (L == 0) return LGB_OK;
  LGB_REQUIRE(edge_aligned(du, di, zu, zi, out, nullptr), LGB_EINVAL, "lgb_edge_concat_fwd: zu / zi / out must be 16-byte aligned when du, di %% 4 == 0");
  edge_concat_fwd_kernel<<<edge_blocks(L), DEC_WARPS * 32, 0, (cudaStream_t)stream>>>(zu, zi, row, col, L, du, di, out);
  LGB_LAUNCH_CHECK();
  return LGB_OK;
}

int lgb_edge_concat_bwd(const float* gout, const int64_t* row, const int64_t* col, int64_t L, int32_t du, int32_t di,
                        float* dzu, float* dzi, void* stream) {
  LGB_REQUIRE(L >= 0 && du > 0 && di > 0 && (L == 0 || (gout && row && col)), LGB_EINVAL,
              "lgb_edge_concat_bwd: bad argument");
  if (L == 0 || (!dzu && !dzi)) return LGB_OK;
  LGB_REQUIRE(edge_aligned(du, di, gout, dzu, dzi, nullptr), LGB_EINVAL, "lgb_edge_concat_bwd: gout / dzu / dzi must be 16-byte aligned when du, di %% 4 == 0");
  edge_concat_bwd_kernel<<<edge_blocks(L), DEC_WARPS * 32, 0, (cudaStream_t)stream>>>(gout, row, col, L, du, di, dzu, dzi);
  LGB_LAUNCH_CHECK();
  return LGB_OK;
}

int lgb_edge_dot_fwd(const float* zu, const float* zi, const int64_t* row, const int64_t* col, int64_t L, int32_t d,
                     float* out, void* stream) {
  LGB_REQUIRE(L >= 0 && d > 0 && (L == 0 || (zu && zi && row && col && out)), LGB_EINVAL, "lgb_edge_dot_fwd: bad argument");
  if (L == 0) return LGB_OK;
  LGB_REQUIRE(edge_aligned(d, d, zu, zi, nullptr, nullptr), LGB_EINVAL, "lgb_edge_dot_fwd: zu / zi must be 16-byte aligned when d %% 4 == 0");
  edge_dot_fwd_kernel<<<edge_blocks(L), DEC_WARPS * 32, 0, (cudaStream_t)stream>>>(zu, zi, row, col, L, d, out);
  LGB_LAUNCH_CHECK();
  return LGB_OK;
}

int lgb_edge_dot_bwd(const float* zu, const float* zi, const int64_t* row, const int64_t* col, const float* gout,
                     int64_t L, int32_t d, float* dzu, float* dzi, void* stream) {
  LGB_REQUIRE(L >= 0 && d > 0 && (L == 0 || (zu && zi && row && col && gout)), LGB_EINVAL, "lgb_edge_dot_bwd: bad argument");
  if (L == 0 || (!dzu && !dzi)) return LGB_OK;
  LGB_REQUIRE(edge_aligned(d, d, zu, zi, dzu, dzi), LGB_EINVAL, "lgb_edge_dot_bwd: zu / zi / dzu / dzi must be 16-byte aligned when d %% 4 == 0");
  edge_dot_bwd_kernel<<<edge_blocks(L), DEC_WARPS * 32, 0, (cudaStream_t)stream>>>(zu, zi, row, col, gout, L, d, dzu, dzi);
  LGB_LAUNCH_CHECK();
  return LGB_OK;
}

int lgb_edge_mlp2_fwd(const float* pu, const float* pi, const int64_t* row, const int64_t* col, int64_t L, int32_t H,
                      const float* w2, const float* b2, float* out, void* stream) {
  LGB_REQUIRE(L >= 0 && H > 0 && (L == 0 || (pu && pi && row && col && w2 && out)), LGB_EINVAL, "lgb_edge_mlp2_fwd: bad argument");
  LGB_REQUIRE((H & 3) != 0 || ((((uintptr_t)pu | (uintptr_t)pi | (uintptr_t)w2) & 15) == 0), LGB_EINVAL,
              "lgb_edge_mlp2_fwd: pu / pi / w2 must be 16-byte aligned when H %% 4 == 0");
  if (L == 0) return LGB_OK;
  edge_mlp2_fwd_kernel<<<edge_blocks(L), DEC_WARPS * 32, 0, (cudaStream_t)stream>>>(pu, pi, row, col, L, H, w2, b2, out);
  LGB_LAUNCH_CHECK();
  return LGB_OK;
}

}  // extern "C"
