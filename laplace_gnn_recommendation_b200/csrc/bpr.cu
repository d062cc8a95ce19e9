// Fused BPR forward + backward for sm_100a.
//
// Replaces the six row gathers of reference run_pipeline_lightgcn.py:133-144, bpr_loss
// (utils/metrics_lightgcn.py:9-45: ~12 ATen launches) and their autograd backward (six
// index_put_(accumulate) into dense zero tensors) with ONE kernel: a group of G lanes owns one
// (user, pos, neg) triple, fetches the six rows with 128-bit loads, forms the two dot products with a
// shuffle reduction, evaluates softplus / sigmoid exactly like ATen (threshold 20) and, in the same
// pass, scatters the gradient rows with red.global.add.v4.f32.  The scalar loss is reduced
// deterministically: per-CTA partials, then a single-CTA final pass.
#include "common.cuh"

namespace lgb {

constexpr int BPR_THREADS = 256;

struct BprParams {
  lgb_bpr_args a;
  int d4;
  int64_t nblocks;
};

__device__ __forceinline__ float group_sum(float v, int G) {
  for (int off = G >> 1; off > 0; off >>= 1) v += __shfl_xor_sync(FULL_MASK, v, off);
  return v;
}

__device__ __forceinline__ float sumsq(const float4& v) { return v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w; }
__device__ __forceinline__ float dot4(const float4& a, const float4& b) {
  return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
}
__device__ __forceinline__ float4 scale4(const float4& a, float s) { return make_float4(a.x * s, a.y * s, a.z * s, a.w * s); }

template <int G, int VPL, bool INDEXED>
__global__ void __launch_bounds__(BPR_THREADS) bpr_kernel(const BprParams p) {
  constexpr int GROUPS = BPR_THREADS / G;
  __shared__ float s_sp[BPR_THREADS / 32], s_reg[BPR_THREADS / 32];
  const lgb_bpr_args& a = p.a;
  const int lig = threadIdx.x % G;
  const int64_t b = (int64_t)blockIdx.x * GROUPS + threadIdx.x / G;
  bool active = b < a.B;
  const int d4 = p.d4;
  float sp = 0.f, reg = 0.f;

  int64_t ru = 0, rp = 0, rn = 0;
  bool own_u = true;                  // LGB_BPR_FILTER_USER_ROWS_ONLY: foreign triples keep their item rows
  if (active) {
    ru = INDEXED ? a.iu[b] : b;
    rp = INDEXED ? a.ip[b] : b;
    rn = INDEXED ? a.in[b] : b;
    if (INDEXED && a.user_hi > 0) {   // this rank only owns users [user_lo, user_hi)
      own_u = ru >= a.user_lo && ru < a.user_hi;
      if (!(a.flags & LGB_BPR_FILTER_USER_ROWS_ONLY)) active = own_u;
      ru -= a.user_lo;
    }
  }
  float4 uf[VPL], pf[VPL], nf[VPL], u0[VPL], p0[VPL], n0[VPL];
  float pos = 0.f, neg = 0.f, rsq = 0.f;
#pragma unroll
  for (int q = 0; q < VPL; ++q) {
    const int f = lig + q * G;
    const bool ok = active && f < d4;
    uf[q] = (ok && own_u) ? ld_gather_f4((const float4*)a.uf + ru * d4 + f) : f4_zero();
    pf[q] = ok ? ld_gather_f4((const float4*)a.pf + rp * d4 + f) : f4_zero();
    nf[q] = ok ? ld_gather_f4((const float4*)a.nf + rn * d4 + f) : f4_zero();
    u0[q] = (ok && own_u && a.u0) ? ld_gather_f4((const float4*)a.u0 + ru * d4 + f) : f4_zero();
    p0[q] = (ok && a.p0) ? ld_gather_f4((const float4*)a.p0 + rp * d4 + f) : f4_zero();
    n0[q] = (ok && a.n0) ? ld_gather_f4((const float4*)a.n0 + rn * d4 + f) : f4_zero();
    pos += dot4(uf[q], pf[q]);
    neg += dot4(uf[q], nf[q]);
    rsq += sumsq(u0[q]) + sumsq(p0[q]) + sumsq(n0[q]);
  }
  pos = group_sum(pos, G);
  neg = group_sum(neg, G);
  rsq = group_sum(rsq, G);
  const float x = pos - neg;
  // ATen softplus(beta=1, threshold=20) and its backward z/(z+1), z = exp(x)
  const float z = expf(x);
  const float soft = x > 20.f ? x : log1pf(z);
  const float sig = x > 20.f ? 1.f : z / (z + 1.f);
  if (active && lig == 0) { sp = soft; reg = rsq; }

  // ---- backward (same pass) ----
  const bool want_f = a.duf || a.dpf || a.dnf;
  const bool want_0 = a.du0 || a.dp0 || a.dn0;
  if (active && (want_f || want_0)) {
    const float g = a.gout ? *a.gout : 1.f;
    const float cf = -g * sig / (float)(a.B_norm > 0 ? a.B_norm : a.B) * a.gscale;  // d loss / d x, folded with the caller's scale
    const float c0 = 2.f * a.lambda * g;
#pragma unroll
    for (int q = 0; q < VPL; ++q) {
      const int f = lig + q * G;
      if (f >= d4) continue;
      if (want_f) {
        const float4 gu = make_float4(cf * (pf[q].x - nf[q].x), cf * (pf[q].y - nf[q].y), cf * (pf[q].z - nf[q].z),
                                      cf * (pf[q].w - nf[q].w));
        const float4 gp = scale4(uf[q], cf);
        const float4 gn = scale4(uf[q], -cf);
        if (INDEXED) {
          if (a.duf && own_u) red_add_f4((float4*)a.duf + ru * d4 + f, gu);
          if (a.dpf) red_add_f4((float4*)a.dpf + rp * d4 + f, gp);
          if (a.dnf) red_add_f4((float4*)a.dnf + rn * d4 + f, gn);
        } else {
          if (a.duf) st_f4((float4*)a.duf + ru * d4 + f, gu);
          if (a.dpf) st_f4((float4*)a.dpf + rp * d4 + f, gp);
          if (a.dnf) st_f4((float4*)a.dnf + rn * d4 + f, gn);
        }
      }
      if (want_0) {
        if (INDEXED) {
          if (a.du0 && own_u) red_add_f4((float4*)a.du0 + ru * d4 + f, scale4(u0[q], c0));
          if (a.dp0) red_add_f4((float4*)a.dp0 + rp * d4 + f, scale4(p0[q], c0));
          if (a.dn0) red_add_f4((float4*)a.dn0 + rn * d4 + f, scale4(n0[q], c0));
        } else {
          if (a.du0) st_f4((float4*)a.du0 + ru * d4 + f, scale4(u0[q], c0));
          if (a.dp0) st_f4((float4*)a.dp0 + rp * d4 + f, scale4(p0[q], c0));
          if (a.dn0) st_f4((float4*)a.dn0 + rn * d4 + f, scale4(n0[q], c0));
        }
      }
    }
  }

  // ---- deterministic loss reduction: warp -> CTA -> per-CTA partial ----
  if (a.loss) {
    for (int off = 16; off > 0; off >>= 1) {
      sp += __shfl_xor_sync(FULL_MASK, sp, off);
      reg += __shfl_xor_sync(FULL_MASK, reg, off);
    }
    const int warp = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) { s_sp[warp] = sp; s_reg[warp] = reg; }
    __syncthreads();
    if (threadIdx.x == 0) {
      float t1 = 0.f, t2 = 0.f;
      for (int i = 0; i < BPR_THREADS / 32; ++i) { t1 += s_sp[i]; t2 += s_reg[i]; }
      a.ws[blockIdx.x] = t1;
      a.ws[p.nblocks + blockIdx.x] = t2;
    }
  }
}

// Scalar variant for d % 4 != 0: one warp per triple, lanes stride over d (tiny test shapes).
template <bool INDEXED>
__global__ void __launch_bounds__(BPR_THREADS) bpr_scalar_kernel(const BprParams p) {
  __shared__ float s_sp[BPR_THREADS / 32], s_reg[BPR_THREADS / 32];
  const lgb_bpr_args& a = p.a;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t b = (int64_t)blockIdx.x * (BPR_THREADS / 32) + warp;
  bool active = b < a.B;
  bool own_u = true;
  const int d = a.d;
  int64_t ru = 0, rp = 0, rn = 0;
  if (active) {
    ru = INDEXED ? a.iu[b] : b; rp = INDEXED ? a.ip[b] : b; rn = INDEXED ? a.in[b] : b;
    if (INDEXED && a.user_hi > 0) {
      own_u = ru >= a.user_lo && ru < a.user_hi;
      if (!(a.flags & LGB_BPR_FILTER_USER_ROWS_ONLY)) active = own_u;
      ru -= a.user_lo;
    }
  }
  float pos = 0.f, neg = 0.f, rsq = 0.f;
  if (active)
    for (int f = lane; f < d; f += 32) {
      const float u = own_u ? a.uf[ru * d + f] : 0.f;
      pos += u * a.pf[rp * d + f];
      neg += u * a.nf[rn * d + f];
      if (a.u0 && own_u) { const float v = a.u0[ru * d + f]; rsq += v * v; }
      if (a.p0) { const float v = a.p0[rp * d + f]; rsq += v * v; }
      if (a.n0) { const float v = a.n0[rn * d + f]; rsq += v * v; }
    }
  pos = group_sum(pos, 32); neg = group_sum(neg, 32); rsq = group_sum(rsq, 32);
  const float x = pos - neg;
  const float z = expf(x);
  const float soft = x > 20.f ? x : log1pf(z);
  const float sig = x > 20.f ? 1.f : z / (z + 1.f);
  if (active && (a.duf || a.dpf || a.dnf || a.du0 || a.dp0 || a.dn0)) {
    const float g = a.gout ? *a.gout : 1.f;
    const float cf = -g * sig / (float)(a.B_norm > 0 ? a.B_norm : a.B) * a.gscale;
    const float c0 = 2.f * a.lambda * g;
    for (int f = lane; f < d; f += 32) {
      const float u = own_u ? a.uf[ru * d + f] : 0.f, pp = a.pf[rp * d + f], nn = a.nf[rn * d + f];
      if (INDEXED) {
        if (a.duf && own_u) atomicAdd(a.duf + ru * d + f, cf * (pp - nn));
        if (a.dpf) atomicAdd(a.dpf + rp * d + f, cf * u);
        if (a.dnf) atomicAdd(a.dnf + rn * d + f, -cf * u);
        if (a.du0 && own_u) atomicAdd(a.du0 + ru * d + f, c0 * a.u0[ru * d + f]);
        if (a.dp0) atomicAdd(a.dp0 + rp * d + f, c0 * a.p0[rp * d + f]);
        if (a.dn0) atomicAdd(a.dn0 + rn * d + f, c0 * a.n0[rn * d + f]);
      } else {
        if (a.duf) a.duf[ru * d + f] = cf * (pp - nn);
        if (a.dpf) a.dpf[rp * d + f] = cf * u;
        if (a.dnf) a.dnf[rn * d + f] = -cf * u;
        if (a.du0) a.du0[ru * d + f] = c0 * a.u0[ru * d + f];
        if (a.dp0) a.dp0[rp * d + f] = c0 * a.p0[rp * d + f];
        if (a.dn0) a.dn0[rn * d + f] = c0 * a.n0[rn * d + f];
      }
    }
  }
  if (a.loss) {
    if (lane == 0) { s_sp[warp] = active ? soft : 0.f; s_reg[warp] = active ? rsq : 0.f; }
    __syncthreads();
    if (threadIdx.x == 0) {
      float t1 = 0.f, t2 = 0.f;
      for (int i = 0; i < BPR_THREADS / 32; ++i) { t1 += s_sp[i]; t2 += s_reg[i]; }
      a.ws[blockIdx.x] = t1;
      a.ws[p.nblocks + blockIdx.x] = t2;
    }
  }
}

__global__ void __launch_bounds__(1024) bpr_finalize_kernel(const float* __restrict__ ws, int64_t nblocks, int64_t B,
                                                            float lambda, float* __restrict__ loss) {
  __shared__ float s1[32], s2[32];
  float t1 = 0.f, t2 = 0.f;
  for (int64_t i = threadIdx.x; i < nblocks; i += blockDim.x) { t1 += ws[i]; t2 += ws[nblocks + i]; }
  for (int off = 16; off > 0; off >>= 1) {
    t1 += __shfl_xor_sync(FULL_MASK, t1, off);
    t2 += __shfl_xor_sync(FULL_MASK, t2, off);
  }
  if ((threadIdx.x & 31) == 0) { s1[threadIdx.x >> 5] = t1; s2[threadIdx.x >> 5] = t2; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, b = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { a += s1[i]; b += s2[i]; }
    *loss = -(a / (float)B) + lambda * b;
  }
}

// Rows of a user-sharded table for a GLOBAL batch: dst[b] = src[idx[b] - lo] when this rank owns user idx[b], else 0
// (the sum over ranks of these buffers is the gathered batch); and the way back: dst[idx[b] - lo] += src[b] for owned b.
__global__ void __launch_bounds__(256) gather_rows_owned_kernel(const float* __restrict__ src, const int64_t* __restrict__ idx,
                                                                int64_t B, int d, int64_t lo, int64_t hi, float* __restrict__ dst) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * d) return;
  const int64_t b = i / d, u = idx[b];
  dst[i] = (u >= lo && u < hi) ? src[(u - lo) * d + (i - b * d)] : 0.f;
}
__global__ void __launch_bounds__(256) scatter_add_rows_owned_kernel(const float* __restrict__ src, const int64_t* __restrict__ idx,
                                                                     int64_t B, int d, int64_t lo, int64_t hi, float* __restrict__ dst) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * d) return;
  const int64_t b = i / d, u = idx[b];
  if (u >= lo && u < hi) atomicAdd(dst + (u - lo) * d + (i - b * d), src[i]);
}

static int groups_per_block(int d) {
  if (d % 4 != 0) return BPR_THREADS / 32;
  const int d4 = d / 4;
  const int G = d4 <= 8 ? 8 : (d4 <= 16 ? 16 : 32);
  return BPR_THREADS / G;
}

// The vector kernels move rows with 128-bit loads / reductions: every table must start on a 16-byte boundary (rows then do,
// d % 4 == 0); anything else takes the scalar kernel instead of a misaligned-address fault.
static bool bpr_needs_scalar(const lgb_bpr_args& a) {
  const uintptr_t ptrs = (uintptr_t)a.uf | (uintptr_t)a.u0 | (uintptr_t)a.pf | (uintptr_t)a.p0 | (uintptr_t)a.nf | (uintptr_t)a.n0 |
                         (uintptr_t)a.duf | (uintptr_t)a.du0 | (uintptr_t)a.dpf | (uintptr_t)a.dp0 | (uintptr_t)a.dnf | (uintptr_t)a.dn0;
  return a.d % 4 != 0 || (ptrs & 15) != 0;
}

template <bool INDEXED>
static int launch_bpr(const BprParams& p, bool scalar, cudaStream_t stream) {
  const unsigned nb = (unsigned)p.nblocks;
  const int d4 = p.d4;
  if (scalar) bpr_scalar_kernel<INDEXED><<<nb, BPR_THREADS, 0, stream>>>(p);
  else if (d4 <= 8) bpr_kernel<8, 1, INDEXED><<<nb, BPR_THREADS, 0, stream>>>(p);
  else if (d4 <= 16) bpr_kernel<16, 1, INDEXED><<<nb, BPR_THREADS, 0, stream>>>(p);
  else if (d4 <= 32) bpr_kernel<32, 1, INDEXED><<<nb, BPR_THREADS, 0, stream>>>(p);
  else if (d4 <= 64) bpr_kernel<32, 2, INDEXED><<<nb, BPR_THREADS, 0, stream>>>(p);
  else if (d4 <= 128) bpr_kernel<32, 4, INDEXED><<<nb, BPR_THREADS, 0, stream>>>(p);
  else { set_error("lgb_bpr: d=%d > 512 not supported", p.a.d); return LGB_EINVAL; }
  LGB_LAUNCH_CHECK();
  return LGB_OK;
}

}  // namespace lgb

using namespace lgb;

extern "C" {

// The block count only depends on B for a fixed d-class; callers size ws for the worst case (8 triples / CTA).
int64_t lgb_bpr_blocks(int64_t B) { return B <= 0 ? 1 : (B + 7) / 8; }

int lgb_bpr(const lgb_bpr_args* a, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  LGB_REQUIRE(a && a->uf && a->pf && a->nf && a->B >= 0 && a->d > 0, LGB_EINVAL, "lgb_bpr: bad argument");
  const int n_idx = (a->iu != nullptr) + (a->ip != nullptr) + (a->in != nullptr);
  LGB_REQUIRE(n_idx == 0 || n_idx == 3, LGB_EINVAL, "lgb_bpr: index arrays must be all NULL or all set");
  LGB_REQUIRE(!a->loss || a->ws, LGB_EINVAL, "lgb_bpr: loss requested without workspace");
  LGB_REQUIRE(!(a->flags & LGB_BPR_FILTER_USER_ROWS_ONLY) || !(a->loss || a->duf || a->dpf || a->dnf), LGB_EINVAL,
              "lgb_bpr: LGB_BPR_FILTER_USER_ROWS_ONLY serves the layer-0 regulariser gradients only (no loss, no *_f gradients)");
  if (a->B == 0) {
    if (a->loss) LGB_CUDA(cudaMemsetAsync(a->loss, 0, sizeof(float), stream));
    return LGB_OK;
  }
  BprParams p;
  p.a = *a;
  p.d4 = a->d / 4;
  const bool scalar = bpr_needs_scalar(*a);
  const int gpb = scalar ? BPR_THREADS / 32 : groups_per_block(a->d);
  p.nblocks = (a->B + gpb - 1) / gpb;
  LGB_REQUIRE(p.nblocks < (1ll << 31), LGB_ERANGE, "lgb_bpr: grid too large");
  int rc = n_idx ? launch_bpr<true>(p, scalar, stream) : launch_bpr<false>(p, scalar, stream);
  if (rc) return rc;
  if (a->loss) {
    bpr_finalize_kernel<<<1, 1024, 0, stream>>>(a->ws, p.nblocks, a->B_norm > 0 ? a->B_norm : a->B, a->lambda, a->loss);
    LGB_LAUNCH_CHECK();
  }
  return LGB_OK;
}

int lgb_gather_rows_owned(const float* src, const int64_t* idx, int64_t B, int32_t d, int64_t lo, int64_t hi, float* dst,
                          void* stream) {
  LGB_REQUIRE(B >= 0 && d > 0 && lo <= hi && (B == 0 || (idx && dst && (src || lo == hi))), LGB_EINVAL,
              "lgb_gather_rows_owned: bad argument");
  if (B == 0) return LGB_OK;
  gather_rows_owned_kernel<<<(unsigned)((B * d + 255) / 256), 256, 0, (cudaStream_t)stream>>>(src, idx, B, d, lo, hi, dst);
  LGB_LAUNCH_CHECK();
  return LGB_OK;
}

int lgb_scatter_add_rows_owned(const float* src, const int64_t* idx, int64_t B, int32_t d, int64_t lo, int64_t hi, float* dst,
                               void* stream) {
  LGB_REQUIRE(B >= 0 && d > 0 && lo <= hi && (B == 0 || (src && idx && (dst || lo == hi))), LGB_EINVAL,
              "lgb_scatter_add_rows_owned: bad argument");
  if (B == 0) return LGB_OK;
  scatter_add_rows_owned_kernel<<<(unsigned)((B * d + 255) / 256), 256, 0, (cudaStream_t)stream>>>(src, idx, B, d, lo, hi, dst);
  LGB_LAUNCH_CHECK();
  return LGB_OK;
}

}  // extern "C"
