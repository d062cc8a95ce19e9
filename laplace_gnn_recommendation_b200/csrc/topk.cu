// Candidate generation: score + top-k with seen-item exclusion, and the negative-sampling rejection test.
//
// Replaces make_predictions_for_user (reference utils/metrics_lightgcn.py:125-142; a per-user Python loop
// of mv + topk + np.setdiff1d on the CPU, driven from run_pipeline_lightgcn.py:211-222 and
// get_metrics_lightgcn :79-122) and the np.isin membership test inside PyG's
// structured_negative_sampling (reference data/lightgcn_loader.py:105-107).
//
// top-k: one CTA per user.  Scores are fp32 FMA chains in ascending-d order (CUDA cores, no tensor
// cores: ids must not depend on a reduced-precision contraction); seen items are overwritten with
// -inf; an exact 4-pass radix select over the order-preserving uint32 image of the scores finds the
// k-th key; survivors are collected in item-id order (ties resolved towards the smaller id) and
// bitonic-sorted in shared memory.
#include <cub/cub.cuh>

#include "common.cuh"

namespace lgb {

constexpr int TOPK_THREADS = 256;
constexpr int TOPK_MAXK = 1024;
constexpr int TOPK_MAXD = 512;

__device__ __forceinline__ uint32_t f2key(float f) {
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__device__ __forceinline__ int block_excl_scan(int flag, int* warp_tot, int& total) {
  // exclusive prefix of a 0/1 flag over the CTA, in thread order
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned bal = __ballot_sync(FULL_MASK, flag);
  const int within = __popc(bal & ((1u << lane) - 1));
  if (lane == 0) warp_tot[warp] = __popc(bal);
  __syncthreads();
  int before = 0, tot = 0;
  for (int w = 0; w < TOPK_THREADS / 32; ++w) {
    const int c = warp_tot[w];
    if (w < warp) before += c;
    tot += c;
  }
  __syncthreads();
  total = tot;
  return before + within;
}

// SCORED = true: score_ws already holds the scores of this block of users (topk_score_tile_kernel); only mask + select run.
template <bool SCORED>
__global__ void __launch_bounds__(TOPK_THREADS)
topk_exclude_kernel(const float* __restrict__ Wu, const float* __restrict__ Wi, const int64_t* __restrict__ users,
                    int64_t n_items, int d, const int32_t* __restrict__ seen_ptr, const int32_t* __restrict__ seen_idx,
                    int k, int kp, int64_t* __restrict__ out_ids, float* __restrict__ out_scores,
                    float* __restrict__ score_ws) {
  __shared__ float su[TOPK_MAXD];
  __shared__ unsigned hist[256];
  __shared__ uint32_t sel_key[TOPK_MAXK];
  __shared__ int sel_id[TOPK_MAXK];
  __shared__ int warp_tot[TOPK_THREADS / 32];
  __shared__ uint32_t s_prefix, s_mask;
  __shared__ int s_need;

  const int tid = threadIdx.x;
  const int64_t u = users[blockIdx.x];
  float* ws = score_ws + (size_t)blockIdx.x * n_items;
  const int I = (int)n_items;
  const int keff = min(k, I);

  if (!SCORED) {
    for (int j = tid; j < d; j += TOPK_THREADS) su[j] = Wu[u * d + j];
    __syncthreads();
    for (int i = tid; i < I; i += TOPK_THREADS) {
      const float* w = Wi + (size_t)i * d;
      float s = 0.f;
      if ((d & 3) == 0) {
        for (int j = 0; j < d; j += 4) {
          const float4 v = *reinterpret_cast<const float4*>(w + j);
          s = fmaf(su[j], v.x, s); s = fmaf(su[j + 1], v.y, s); s = fmaf(su[j + 2], v.z, s); s = fmaf(su[j + 3], v.w, s);
        }
      } else {
        for (int j = 0; j < d; ++j) s = fmaf(su[j], w[j], s);
      }
      ws[i] = s;
    }
    __syncthreads();
  }
  if (seen_ptr) {
    const int s0 = seen_ptr[u], s1 = seen_ptr[u + 1];
    for (int t = s0 + tid; t < s1; t += TOPK_THREADS) {
      const int it = seen_idx[t];
      if (it >= 0 && it < I) ws[it] = -INFINITY;
    }
  }
  if (tid == 0) { s_prefix = 0; s_mask = 0; s_need = keff; }
  __syncthreads();

  // exact radix select of the keff-th largest key
  for (int pass = 3; pass >= 0; --pass) {
    const int shift = pass * 8;
    hist[tid] = 0;
    __syncthreads();
    const uint32_t prefix = s_prefix, mask = s_mask;
    for (int i = tid; i < I; i += TOPK_THREADS) {
      const uint32_t key = f2key(ws[i]);
      if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 255], 1u);
    }
    __syncthreads();
    if (tid == 0) {
      int need = s_need, cum = 0, b = 255;
      for (; b > 0; --b) {
        const int c = (int)hist[b];
        if (cum + c >= need) break;
        cum += c;
      }
      s_need = need - cum;
      s_prefix = prefix | ((uint32_t)b << shift);
      s_mask = mask | (255u << shift);
    }
    __syncthreads();
  }
  const uint32_t T = s_prefix;
  const int need_eq = s_need;           // how many keys == T are taken (smallest ids first)
  const int n_gt = keff - need_eq;      // keys strictly above T

  // ordered collection
  int done_gt = 0, done_eq = 0;
  for (int base = 0; base < I; base += TOPK_THREADS) {
    const int i = base + tid;
    uint32_t key = 0;
    int gt = 0, eq = 0;
    if (i < I) { key = f2key(ws[i]); gt = key > T; eq = key == T; }
    int tot_gt, tot_eq;
    const int pg = block_excl_scan(gt, warp_tot, tot_gt);
    const int pe = block_excl_scan(eq, warp_tot, tot_eq);
    if (gt) { sel_key[done_gt + pg] = key; sel_id[done_gt + pg] = i; }
    if (eq && done_eq + pe < need_eq) { sel_key[n_gt + done_eq + pe] = key; sel_id[n_gt + done_eq + pe] = i; }
    done_gt += tot_gt;
    done_eq += tot_eq;
  }
  for (int s = keff + tid; s < kp; s += TOPK_THREADS) { sel_key[s] = 0u; sel_id[s] = 0x7fffffff; }
  __syncthreads();

  // bitonic sort: key descending, id ascending
  for (int size = 2; size <= kp; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = tid; t < kp / 2; t += TOPK_THREADS) {
        const int lo = 2 * t - (t & (stride - 1));
        const int hi = lo + stride;
        const bool desc = ((lo & size) == 0);
        const uint32_t ka = sel_key[lo], kb = sel_key[hi];
        const int ia = sel_id[lo], ib = sel_id[hi];
        const bool a_first = (ka > kb) || (ka == kb && ia < ib);  // a belongs before b in the final order
        if (a_first != desc) { sel_key[lo] = kb; sel_key[hi] = ka; sel_id[lo] = ib; sel_id[hi] = ia; }
      }
      __syncthreads();
    }
  }
  for (int s = tid; s < k; s += TOPK_THREADS) {
    int64_t id = -1;
    float sc = -INFINITY;
    if (s < keff) {
      const int i = sel_id[s];
      sc = ws[i];
      if (sc != -INFINITY) id = i;
    }
    out_ids[(size_t)blockIdx.x * k + s] = id;
    if (out_scores) out_scores[(size_t)blockIdx.x * k + s] = sc;
  }
}

// ---- user-tiled scoring ---------------------------------------------------------------------------------------
// The one-CTA-per-user kernel re-reads the whole item table (27 MB at the H&M shape) from L2 for every user: 4 FMAs per
// 16 bytes loaded.  Here a CTA scores TOPK_UT users at once: every item row is loaded ONCE per CTA and multiplied against
// TOPK_UT user rows held in shared memory (broadcast reads), 4*TOPK_UT FMAs per 16 bytes -- the contraction becomes
// FMA-bound on the CUDA cores instead of L2-bound.  Each score is the same fp32 FMA chain in ascending-d order as in the
// per-user kernel (bit-identical scores, hence bit-identical ids); selection then runs per user on the score rows.
constexpr int TOPK_UT = 8;

__global__ void __launch_bounds__(TOPK_THREADS)
topk_score_tile_kernel(const float* __restrict__ Wu, const float* __restrict__ Wi, const int64_t* __restrict__ users,
                       int64_t n_users, int64_t n_items, int d, float* __restrict__ score_ws) {
  __shared__ __align__(16) float su[TOPK_UT][TOPK_MAXD];
  const int tid = threadIdx.x;
  const int64_t u0 = (int64_t)blockIdx.x * TOPK_UT;
  const int nu = (int)min((int64_t)TOPK_UT, n_users - u0);
  for (int t = tid; t < TOPK_UT * d; t += TOPK_THREADS) {
    const int uu = t / d, j = t - uu * d;
    su[uu][j] = uu < nu ? Wu[users[u0 + uu] * d + j] : 0.f;
  }
  __syncthreads();
  const int I = (int)n_items;
  for (int i = tid; i < I; i += TOPK_THREADS) {
    const float* w = Wi + (size_t)i * d;
    float s[TOPK_UT];
#pragma unroll
    for (int uu = 0; uu < TOPK_UT; ++uu) s[uu] = 0.f;
    if ((d & 3) == 0) {
      for (int j = 0; j < d; j += 4) {
        const float4 v = *reinterpret_cast<const float4*>(w + j);
#pragma unroll
        for (int uu = 0; uu < TOPK_UT; ++uu) {
          const float4 a = *reinterpret_cast<const float4*>(&su[uu][j]);
          s[uu] = fmaf(a.x, v.x, s[uu]); s[uu] = fmaf(a.y, v.y, s[uu]); s[uu] = fmaf(a.z, v.z, s[uu]); s[uu] = fmaf(a.w, v.w, s[uu]);
        }
      }
    } else {
      for (int j = 0; j < d; ++j) {
        const float wj = w[j];
#pragma unroll
        for (int uu = 0; uu < TOPK_UT; ++uu) s[uu] = fmaf(su[uu][j], wj, s[uu]);
      }
    }
#pragma unroll
    for (int uu = 0; uu < TOPK_UT; ++uu)
      if (uu < nu) score_ws[(size_t)(u0 + uu) * n_items + i] = s[uu];
  }
}

// ---- negative sampling support --------------------------------------------------------------
__global__ void edge_keys_kernel(const int64_t* __restrict__ row, const int64_t* __restrict__ col, int64_t n,
                                 int64_t num_nodes, int64_t* __restrict__ keys) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) keys[i] = row[i] * num_nodes + col[i];
}

__global__ void neg_reject_kernel(const int64_t* __restrict__ row, const int64_t* __restrict__ cand, int64_t n,
                                  int64_t num_nodes, const int64_t* __restrict__ pos, int64_t n_pos, int with_loops,
                                  uint8_t* __restrict__ mask) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t key = row[i] * num_nodes + cand[i];
  int64_t lo = 0, hi = n_pos;  // first index with pos[idx] >= key
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (pos[mid] < key) lo = mid + 1; else hi = mid;
  }
  bool hit = lo < n_pos && pos[lo] == key;
  if (!hit && with_loops && num_nodes > 0) {
    // self-loop keys i*(num_nodes+1) for 0 <= i < num_nodes
    const int64_t q = key / (num_nodes + 1);
    hit = key >= 0 && q * (num_nodes + 1) == key && q < num_nodes;
  }
  mask[i] = hit ? 1 : 0;
}

}  // namespace lgb

using namespace lgb;

extern "C" {

int lgb_topk_exclude(const float* Wu, const float* Wi, const int64_t* users, int64_t n_users, int64_t n_items, int32_t d,
                     const int32_t* seen_ptr, const int32_t* seen_idx, int32_t k, int64_t* out_ids, float* out_scores,
                     float* score_ws, void* stream) {
  LGB_REQUIRE(n_users >= 0 && n_items > 0 && d > 0 && k > 0, LGB_EINVAL, "lgb_topk_exclude: bad size");
  LGB_REQUIRE(k <= TOPK_MAXK, LGB_EINVAL, "lgb_topk_exclude: k=%d > %d", k, TOPK_MAXK);
  LGB_REQUIRE(d <= TOPK_MAXD, LGB_EINVAL, "lgb_topk_exclude: d=%d > %d", d, TOPK_MAXD);
  LGB_REQUIRE(n_items < (1ll << 31) - 1 && n_users < (1ll << 31) - 1, LGB_ERANGE, "lgb_topk_exclude: size exceeds int32");
  if (n_users == 0) return LGB_OK;
  LGB_REQUIRE(Wu && Wi && users && out_ids && score_ws, LGB_EINVAL, "lgb_topk_exclude: null pointer");
  int kp = 2;
  while (kp < k) kp <<= 1;
  topk_exclude_kernel<false><<<(unsigned)n_users, TOPK_THREADS, 0, (cudaStream_t)stream>>>(
      Wu, Wi, users, n_items, d, seen_ptr, seen_idx, k, kp, out_ids, out_scores, score_ws);
  LGB_LAUNCH_CHECK();
  return LGB_OK;
}

int lgb_topk_exclude_tiled(const float* Wu, const float* Wi, const int64_t* users, int64_t n_users, int64_t n_items,
                           int32_t d, const int32_t* seen_ptr, const int32_t* seen_idx, int32_t k, int64_t* out_ids,
                           float* out_scores, float* score_ws, void* stream) {
  LGB_REQUIRE(n_users >= 0 && n_items > 0 && d > 0 && k > 0, LGB_EINVAL, "lgb_topk_exclude_tiled: bad size");
  LGB_REQUIRE(k <= TOPK_MAXK, LGB_EINVAL, "lgb_topk_exclude_tiled: k=%d > %d", k, TOPK_MAXK);
  LGB_REQUIRE(d <= TOPK_MAXD, LGB_EINVAL, "lgb_topk_exclude_tiled: d=%d > %d", d, TOPK_MAXD);
  LGB_REQUIRE(n_items < (1ll << 31) - 1 && n_users < (1ll << 31) - 1, LGB_ERANGE, "lgb_topk_exclude_tiled: size exceeds int32");
  if (n_users == 0) return LGB_OK;
  LGB_REQUIRE(Wu && Wi && users && out_ids && score_ws, LGB_EINVAL, "lgb_topk_exclude_tiled: null pointer");
  LGB_REQUIRE((((uintptr_t)Wi) & 15) == 0 || (d & 3) != 0, LGB_EINVAL, "lgb_topk_exclude_tiled: item table must be 16-byte aligned");
  int kp = 2;
  while (kp < k) kp <<= 1;
  topk_score_tile_kernel<<<(unsigned)((n_users + TOPK_UT - 1) / TOPK_UT), TOPK_THREADS, 0, (cudaStream_t)stream>>>(
      Wu, Wi, users, n_users, n_items, d, score_ws);
  LGB_LAUNCH_CHECK();
  topk_exclude_kernel<true><<<(unsigned)n_users, TOPK_THREADS, 0, (cudaStream_t)stream>>>(
      Wu, Wi, users, n_items, d, seen_ptr, seen_idx, k, kp, out_ids, out_scores, score_ws);
  LGB_LAUNCH_CHECK();
  return LGB_OK;
}

int lgb_sort_keys_ws_bytes(int64_t n, size_t* bytes) {
  LGB_REQUIRE(bytes && n >= 0, LGB_EINVAL, "lgb_sort_keys_ws_bytes: bad argument");
  LGB_REQUIRE(n < (1ll << 31), LGB_ERANGE, "lgb_sort_keys: n exceeds int32");
  size_t temp = 0;
  cub::DeviceRadixSort::SortKeys(nullptr, temp, (const int64_t*)nullptr, (int64_t*)nullptr, (int)n, 0, 64, (cudaStream_t)0);
  *bytes = align_up((size_t)n * 8) + align_up(temp) + 256;
  return LGB_OK;
}

int lgb_edge_keys_sorted(const int64_t* row, const int64_t* col, int64_t n, int64_t num_nodes, int64_t* keys_out,
                         void* ws, size_t ws_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  LGB_REQUIRE(n >= 0 && (n == 0 || (row && col && keys_out)), LGB_EINVAL, "lgb_edge_keys_sorted: bad argument");
  if (n == 0) return LGB_OK;
  size_t need = 0;
  int rc = lgb_sort_keys_ws_bytes(n, &need);
  if (rc) return rc;
  LGB_REQUIRE(ws && ws_bytes >= need, LGB_EWS, "lgb_edge_keys_sorted: workspace %zu < %zu", ws_bytes, need);
  int64_t* keys_in = (int64_t*)ws;
  void* temp = (char*)ws + align_up((size_t)n * 8);
  size_t temp_bytes = ws_bytes - align_up((size_t)n * 8);
  edge_keys_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(row, col, n, num_nodes, keys_in);
  LGB_LAUNCH_CHECK();
  LGB_CUDA(cub::DeviceRadixSort::SortKeys(temp, temp_bytes, keys_in, keys_out, (int)n, 0, 64, stream));
  return LGB_OK;
}

int lgb_neg_reject_mask(const int64_t* row, const int64_t* cand, int64_t n, int64_t num_nodes,
                        const int64_t* pos_keys_sorted, int64_t n_pos, int32_t with_self_loops, uint8_t* mask,
                        void* stream) {
  LGB_REQUIRE(n >= 0 && n_pos >= 0 && (n == 0 || (row && cand && mask)) && (n_pos == 0 || pos_keys_sorted), LGB_EINVAL,
              "lgb_neg_reject_mask: bad argument");
  if (n == 0) return LGB_OK;
  neg_reject_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(row, cand, n, num_nodes, pos_keys_sorted,
                                                                                  n_pos, with_self_loops, mask);
  LGB_LAUNCH_CHECK();
  return LGB_OK;
}

}  // extern "C"
