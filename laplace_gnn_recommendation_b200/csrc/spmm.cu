// Fused CSR SpMM for sm_100a: Y = A @ X with the LightGCN layer-accumulate / mean / residual epilogues.
//
// Replaces torch_sparse.matmul (reference model/lightgcn.py:85-87), its backward (the same kernel on the
// CSC arrays), the [N,K+1,d] stack + mean read-out (model/lightgcn.py:67-68) and, with val == NULL, PyG
// SAGEConv's gather + scatter-{add,mean} (model/layers.py:9-24).
//
// Design (HBM/L2-bound gather, no tensor cores):
//   * one warp per row (or per <=chunk-sized slice of a long row); the warp is cut into 32/G groups of G
//     lanes, G*VPL float4 = one embedding row, so every gathered row is fetched with 128-bit
//     ld.global.nc loads that cover whole 32-byte sectors;
//   * the 32 (col,val) pairs of a batch are loaded coalesced (one per lane, streaming / no L1 allocate)
//     and broadcast with __shfl_sync; UNROLL independent row gathers per lane are issued before the
//     first FMA to keep enough bytes in flight for HBM latency;
//   * rows longer than `chunk` were split at plan time into fixed slices; slices write partial sums
//     that a second kernel reduces in a fixed order -> deterministic summation, no atomics;
//   * epilogue fuses the mean division, the residual add, the layer accumulate (acc_out = (acc_in + y)/div)
//     and lets the last layer skip writing Y altogether.
#include <algorithm>

#include "common.cuh"

namespace lgb {

constexpr int SPMM_WARPS = 4;  // 128-thread CTAs: fine-grained dynamic balancing across skewed degrees

struct SpmmParams {
  const int32_t* rowptr;
  const int32_t* colidx;
  const float* val;
  const int32_t* row_order;
  const int32_t* task_row;
  const int32_t* task_start;
  const int32_t* task_end;
  const int32_t* long_rows;
  const int32_t* long_ptr;
  int64_t n_rows;
  int64_t n_tasks;
  int64_t n_long;
  int32_t chunk;
  int32_t d4;  // d / 4
  const float* X;
  float* Y;
  const float* resid;
  const float* acc_in;
  float* acc_out;
  float acc_div;
  int32_t mean;
  float* partial;
  int64_t split_row;   // rows >= split_row skip the epilogue and store their raw sums to y_tail[r - split_row] (0 = off)
  float* y_tail;
  // stage-2 tree plan (optional): the partial rows of every long row cut into segments of <= 32
  const int32_t* seg_row;
  const int32_t* seg_t0;
  const int32_t* seg_t1;
  const int32_t* row_seg0;
  int64_t n_seg;
  float* part2;        // [n_seg, d] level-2 partial rows
  int* tickets;        // [n_long], zero between launches
  const int32_t* task_exec;   // [n_tasks] execution order of the slices (NULL = plan order); partial rows stay indexed by slice id
  const uint32_t* filter;     // row-sparse operand (lgb_spmm_rowsparse): bit c set <=> row c of X may be non-zero; NULL = dense X
  const uint32_t* resid_filter;   // same for the rows of resid (NULL = dense): a row that is not flagged is not read
  const int32_t* task_seg;    // fused stage 2 (LGB_SPMM_FUSED_STAGE2): segment of slice t; NULL = stage 2 is a launch of its own
  int* seg_tickets;           // [n_seg], zero between launches
  const int32_t* task_seg_plan;   // host side only: the plan's task_seg when the call allows the fused stage 2 (the sub-warp
                                  // launcher copies it into task_seg; other kernel families leave task_seg NULL)
};

__device__ __forceinline__ void slice_done(const SpmmParams& p, int64_t t, int lane);

// bit test of the row-sparse operand's bitmap (n_cols bits, L1-resident: 185 KB for the H&M-shaped table)
__device__ __forceinline__ bool filter_hit(const uint32_t* __restrict__ filter, int c) {
  return (__ldg(filter + ((uint32_t)c >> 5)) >> ((uint32_t)c & 31u)) & 1u;
}

// STEP = distance between the 32-entry batches this warp takes (32: the whole range; 32*SPMM_WARPS: every SPMM_WARPS-th
// batch, when the warps of a CTA share one slice).
// IT = type of the element index into X: uint32_t whenever n_cols*d/4 < 2^31 (two integer instructions per gather address
// instead of seven for the sign-extended 64-bit form -- the ncu capture of round 1 shows the SM issue slots 58 % busy, so
// the instruction count of this loop is a first-order term), size_t for tables beyond that.
// Loads and FMAs of the entries past the end of a row are predicated off (no zero-filled registers, no weight select).
// D4C = d/4 when it is known at compile time (the row is exactly one float4 per lane of the group: d = 4*G*VPL, e.g. d = 64
// with G = 16): the column-bound predicates disappear and the row offset becomes a shift; 0 = read it from the parameters.
// W256: a lane holds two ADJACENT float4 of the row (2*lig, 2*lig+1) and fetches them with one 256-bit load (LDG.E.256).
// FILTER (lgb_spmm_rowsparse): X is zero outside the rows flagged in p.filter.  Every lane tests the column it loaded; a batch
// without a hit costs one ballot, and the hits of a batch are taken NG at a time straight from the ballot mask (group g pops the
// g-th lowest set bit) -- entries that multiply a zero row are never gathered.  Order per group: ascending, fixed.
template <int G, int VPL, int UNROLL, int STEP = 32, typename IT = uint32_t, int D4C = 0, bool W256 = false, bool FILTER = false>
__device__ __forceinline__ void accumulate_slice(const SpmmParams& p, int s, int e, int lane, float4 (&acc)[VPL]) {
  constexpr int NG = 32 / G;
  static_assert(32 % (NG * UNROLL) == 0, "a batch of 32 entries must be a whole number of unrolled steps");
  static_assert(D4C == 0 || D4C == G * VPL, "a compile-time d/4 must fill the lane group exactly");
  static_assert(!W256 || (VPL == 2 && D4C == 2 * G), "256-bit gathers need two adjacent float4 per lane and an exact fit");
  const int grp = lane / G;
  const int lig = lane % G;
  const float4* __restrict__ X4 = reinterpret_cast<const float4*>(p.X);
  const int d4 = D4C ? D4C : p.d4;
  const IT d4i = (IT)d4;
  for (int base = s; base < e; base += STEP) {
    const int idx = base + lane;
    int c = FILTER ? -1 : 0;
    float w = 0.f;
    if (idx < e) {
      c = ld_stream_i32(p.colidx + idx);
      if (FILTER && !filter_hit(p.filter, c)) c = -1;
      if (!FILTER || c >= 0) w = p.val ? ld_stream_f32(p.val + idx) : 1.f;
    }
    if (FILTER) {
      unsigned m = __ballot_sync(FULL_MASK, c >= 0);
      while (m) {                                   // warp-uniform: every lane holds the same mask
        int k = -1;
#pragma unroll
        for (int g2 = 0; g2 < NG; ++g2) {
          const int b = m ? __ffs((int)m) - 1 : -1;
          if (g2 == grp) k = b;
          m &= m - 1u;                              // 0 stays 0
        }
        const int cc = __shfl_sync(FULL_MASK, c, k & 31);
        const float wk = __shfl_sync(FULL_MASK, w, k & 31);
        if (k >= 0) {
          const IT row = (IT)cc * d4i;
          if (W256) {
            const f4x2 t = ld_gather_f8(X4 + (row + (IT)(lig * 2)));
            f4_fma(acc[0], wk, t.a);
            f4_fma(acc[VPL - 1], wk, t.b);
          } else {
#pragma unroll
            for (int q = 0; q < VPL; ++q) {
              const int f = lig + q * G;
              if (f < d4) f4_fma(acc[q], wk, ld_gather_f4(X4 + (row + (IT)f)));
            }
          }
        }
      }
      continue;
    }
    const int cnt = min(32, e - base);
    for (int j = 0; j < cnt; j += NG * UNROLL) {
      float4 v[UNROLL][VPL];
      float ww[UNROLL];
      bool ok[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        const int k = j + u * NG + grp;                 // < 32: j is a multiple of NG*UNROLL, which divides 32
        const int cc = __shfl_sync(FULL_MASK, c, k);
        ww[u] = __shfl_sync(FULL_MASK, w, k);
        ok[u] = k < cnt;
        const IT row = (IT)cc * d4i;
        if (W256) {
          if (ok[u]) {
            const f4x2 t = ld_gather_f8(X4 + (row + (IT)(lig * 2)));
            v[u][0] = t.a;
            v[u][VPL - 1] = t.b;
          }
        } else {
#pragma unroll
          for (int q = 0; q < VPL; ++q) {
            const int f = lig + q * G;
            if (ok[u] && f < d4) v[u][q] = ld_gather_f4(X4 + (row + (IT)f));
          }
        }
      }
#pragma unroll
      for (int u = 0; u < UNROLL; ++u)
#pragma unroll
        for (int q = 0; q < VPL; ++q)
          if (ok[u] && (W256 || lig + q * G < d4)) f4_fma(acc[q], ww[u], v[u][q]);
    }
  }
  // combine the groups (fixed butterfly order)
#pragma unroll
  for (int off = G; off < 32; off <<= 1)
#pragma unroll
    for (int q = 0; q < VPL; ++q) acc[q] = f4_add(acc[q], f4_shfl_xor(acc[q], off));
}

// EF: the epilogue operands (read once per launch) are loaded with the L2 evict_first policy.
// CONTIG: lane `lig` holds the VPL ADJACENT float4 lig*VPL .. lig*VPL+VPL-1 of the row (the 256-bit-load layout) instead of
// the strided lig, lig+G, ...
template <int G, int VPL, bool EF = false, bool CONTIG = false>
__device__ __forceinline__ void epilogue_row(const SpmmParams& p, int r, int deg, int lig, float4 (&acc)[VPL]) {
  const uint64_t pol = EF ? l2_policy_evict_first() : 0;
  if (p.y_tail && r >= p.split_row) {   // multi-GPU item rows: partial sums go to the exchange buffer untouched
    float4* out = reinterpret_cast<float4*>(p.y_tail) + (size_t)(r - p.split_row) * p.d4;
#pragma unroll
    for (int q = 0; q < VPL; ++q) {
      const int f = CONTIG ? lig * VPL + q : lig + q * G;
      if (f < p.d4) st_f4(out + f, acc[q]);
    }
    return;
  }
  const size_t rowoff = (size_t)r * p.d4;
  const bool has_resid = p.resid && (!p.resid_filter || filter_hit(p.resid_filter, r));
  if (CONTIG && VPL == 2) {
    // the 256-bit layout (callers guarantee d/4 == 2*G): both float4 of the lane move with ONE load / store each
    const size_t o = rowoff + (size_t)lig * 2;
    float4 y0 = acc[0], y1 = acc[VPL - 1];
    if (p.mean) {
      const float c = (float)max(deg, 1);
      y0.x = __fdiv_rn(y0.x, c); y0.y = __fdiv_rn(y0.y, c); y0.z = __fdiv_rn(y0.z, c); y0.w = __fdiv_rn(y0.w, c);
      y1.x = __fdiv_rn(y1.x, c); y1.y = __fdiv_rn(y1.y, c); y1.z = __fdiv_rn(y1.z, c); y1.w = __fdiv_rn(y1.w, c);
    }
    if (has_resid) {
      const f4x2 t = ld_once_f8(reinterpret_cast<const float4*>(p.resid) + o);
      y0 = f4_add(y0, t.a); y1 = f4_add(y1, t.b);
    }
    if (p.Y) st_f8(reinterpret_cast<float4*>(p.Y) + o, y0, y1);
    if (p.acc_out) {
      float4 a0 = y0, a1 = y1;
      if (p.acc_in) {
        const f4x2 t = ld_once_f8(reinterpret_cast<const float4*>(p.acc_in) + o);
        a0 = f4_add(t.a, y0); a1 = f4_add(t.b, y1);
      }
      if (p.acc_div != 1.0f) {
        a0.x = __fdiv_rn(a0.x, p.acc_div); a0.y = __fdiv_rn(a0.y, p.acc_div); a0.z = __fdiv_rn(a0.z, p.acc_div); a0.w = __fdiv_rn(a0.w, p.acc_div);
        a1.x = __fdiv_rn(a1.x, p.acc_div); a1.y = __fdiv_rn(a1.y, p.acc_div); a1.z = __fdiv_rn(a1.z, p.acc_div); a1.w = __fdiv_rn(a1.w, p.acc_div);
      }
      st_f8(reinterpret_cast<float4*>(p.acc_out) + o, a0, a1);
    }
    return;
  }
#pragma unroll
  for (int q = 0; q < VPL; ++q) {
    const int f = CONTIG ? lig * VPL + q : lig + q * G;
    if (f >= p.d4) continue;
    float4 y = acc[q];
    if (p.mean) {
      const float c = (float)max(deg, 1);
      y.x = __fdiv_rn(y.x, c); y.y = __fdiv_rn(y.y, c); y.z = __fdiv_rn(y.z, c); y.w = __fdiv_rn(y.w, c);
    }
    if (has_resid) {
      const float4* src = reinterpret_cast<const float4*>(p.resid) + rowoff + f;
      y = f4_add(y, EF ? ld_once_f4_hint(src, pol) : ld_once_f4(src));
    }
    if (p.Y) st_f4(reinterpret_cast<float4*>(p.Y) + rowoff + f, y);
    if (p.acc_out) {
      float4 a = y;
      if (p.acc_in) {
        const float4* src = reinterpret_cast<const float4*>(p.acc_in) + rowoff + f;
        a = f4_add(EF ? ld_once_f4_hint(src, pol) : ld_once_f4(src), y);
      }
      if (p.acc_div != 1.0f) {
        a.x = __fdiv_rn(a.x, p.acc_div); a.y = __fdiv_rn(a.y, p.acc_div);
        a.z = __fdiv_rn(a.z, p.acc_div); a.w = __fdiv_rn(a.w, p.acc_div);
      }
      st_f4(reinterpret_cast<float4*>(p.acc_out) + rowoff + f, a);
    }
  }
}

// Stage 1: slices of long rows first (heaviest work is scheduled first), then one warp per ordinary row.
template <int G, int VPL, int UNROLL, int MINB, typename IT = uint32_t, int D4C = 0, bool W256 = false>
__global__ void __launch_bounds__(SPMM_WARPS * 32, MINB) spmm_rows_kernel(const SpmmParams p) {
  const int lane = threadIdx.x & 31;
  const int64_t w = (int64_t)blockIdx.x * SPMM_WARPS + (threadIdx.x >> 5);
  float4 acc[VPL];
#pragma unroll
  for (int q = 0; q < VPL; ++q) acc[q] = f4_zero();

  if (w < p.n_tasks) {
    const int r = p.task_row[w];
    const int s = p.task_start[w];
    const int e = min(s + p.chunk, p.rowptr[r + 1]);
    accumulate_slice<G, VPL, UNROLL, 32, IT, D4C, W256>(p, s, e, lane, acc);
    if (lane < G) {
      float4* out = reinterpret_cast<float4*>(p.partial) + (size_t)w * p.d4;
#pragma unroll
      for (int q = 0; q < VPL; ++q) {
        const int f = W256 ? lane * VPL + q : lane + q * G;
        if (f < p.d4) st_f4(out + f, acc[q]);
      }
    }
    return;
  }
  int64_t ri = w - p.n_tasks;
  if (ri >= p.n_rows) return;
  const int r = p.row_order ? p.row_order[ri] : (int)ri;
  const int s = p.rowptr[r];
  const int e = p.rowptr[r + 1];
  if (p.chunk > 0 && e - s > p.chunk) return;  // long row: handled by its slices + stage 2
  accumulate_slice<G, VPL, UNROLL, 32, IT, D4C, W256>(p, s, e, lane, acc);
  if (lane < G) epilogue_row<G, VPL, false, W256>(p, r, e - s, lane, acc);
}


// ---- software-pipelined persistent variant ---------------------------------------------------------
// ncu on the first version (profiles/r1a_*) showed the kernel latency-bound, not HBM-bound: a short row costs four
// DEPENDENT memory round trips (rowptr -> col/val -> gathered rows -> acc/resid) and the SM only holds ~28 such
// chains.  Here every warp is persistent and walks work items w, w+W, w+2W, ... (W = warps in the grid) with a
// 3-deep software pipeline: while the gathers of item i are in flight, the (col,val) batch of item i+1, the row
// pointers of item i+2 and the epilogue operands (acc_in / resid rows) of item i are already loading, so an
// item costs ~one round trip instead of four.  Long-row slices carry their end offset in the plan (task_end), so
// no stage has a dependent load.  Summation order is identical to spmm_rows_kernel (bit-identical results).
enum { ITEM_SKIP = 0, ITEM_ROW = 1, ITEM_TASK = 2 };

__device__ __forceinline__ void load_item(const SpmmParams& p, int64_t w, int64_t total, int& r, int& s, int& e, int& kind) {
  r = 0; s = 0; e = 0; kind = ITEM_SKIP;
  if (w >= total) return;
  if (w < p.n_tasks) {
    r = ld_stream_i32(p.task_row + w);
    s = ld_stream_i32(p.task_start + w);
    e = ld_stream_i32(p.task_end + w);
    kind = ITEM_TASK;
    return;
  }
  const int64_t ri = w - p.n_tasks;
  r = p.row_order ? p.row_order[ri] : (int)ri;
  s = p.rowptr[r];
  e = p.rowptr[r + 1];
  kind = ITEM_ROW;
}

__device__ __forceinline__ void load_first_batch(const SpmmParams& p, int s, int e, int lane, int& c, float& w) {
  c = 0; w = 0.f;
  const int idx = s + lane;
  if (idx < e) {
    c = ld_stream_i32(p.colidx + idx);
    w = p.val ? ld_stream_f32(p.val + idx) : 1.f;
  }
}

template <int G, int VPL, int UNROLL>
__device__ __forceinline__ void accumulate_slice_pf(const SpmmParams& p, int s, int e, int lane, int c, float w,
                                                    float4 (&acc)[VPL]) {
  constexpr int NG = 32 / G;
  const int grp = lane / G;
  const int lig = lane % G;
  const float4* __restrict__ X4 = reinterpret_cast<const float4*>(p.X);
  const int d4 = p.d4;
  for (int base = s; base < e; base += 32) {
    // next (col,val) batch of this item is requested before the current one is consumed
    int cn = 0;
    float wn = 0.f;
    const int nidx = base + 32 + lane;
    if (nidx < e) {
      cn = ld_stream_i32(p.colidx + nidx);
      wn = p.val ? ld_stream_f32(p.val + nidx) : 1.f;
    }
    const int cnt = min(32, e - base);
    for (int j = 0; j < cnt; j += NG * UNROLL) {
      float4 v[UNROLL][VPL];
      float ww[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        const int k = j + u * NG + grp;
        const int cc = __shfl_sync(FULL_MASK, c, k & 31);
        const float wk = __shfl_sync(FULL_MASK, w, k & 31);
        const bool ok = k < cnt;
        ww[u] = ok ? wk : 0.f;
#pragma unroll
        for (int q = 0; q < VPL; ++q) {
          const int f = lig + q * G;
          v[u][q] = (ok && f < d4) ? ld_gather_f4(X4 + (size_t)cc * d4 + f) : f4_zero();
        }
      }
#pragma unroll
      for (int u = 0; u < UNROLL; ++u)
#pragma unroll
        for (int q = 0; q < VPL; ++q) f4_fma(acc[q], ww[u], v[u][q]);
    }
    c = cn;
    w = wn;
  }
#pragma unroll
  for (int off = G; off < 32; off <<= 1)
#pragma unroll
    for (int q = 0; q < VPL; ++q) acc[q] = f4_add(acc[q], f4_shfl_xor(acc[q], off));
}

template <int G, int VPL, int UNROLL>
__global__ void __launch_bounds__(SPMM_WARPS * 32) spmm_pipe_kernel(const SpmmParams p) {
  const int lane = threadIdx.x & 31;
  const int64_t nw = (int64_t)gridDim.x * SPMM_WARPS;
  const int64_t total = p.n_tasks + p.n_rows;
  int64_t w = (int64_t)blockIdx.x * SPMM_WARPS + (threadIdx.x >> 5);

  int r0, s0, e0, k0, r1, s1, e1, k1;
  load_item(p, w, total, r0, s0, e0, k0);
  load_item(p, w + nw, total, r1, s1, e1, k1);
  int c0;
  float v0;
  load_first_batch(p, s0, e0, lane, c0, v0);

  for (; w < total; w += nw) {
    int r2, s2, e2, k2;
    load_item(p, w + 2 * nw, total, r2, s2, e2, k2);   // row pointers two items ahead
    if (k1 == ITEM_ROW && p.chunk > 0 && e1 - s1 > p.chunk) k1 = ITEM_SKIP;   // long row: its slices do the work
    int c1 = 0;
    float v1 = 0.f;
    if (k1 != ITEM_SKIP) load_first_batch(p, s1, e1, lane, c1, v1);           // (col,val) one item ahead
    if (k0 == ITEM_ROW && p.chunk > 0 && e0 - s0 > p.chunk) k0 = ITEM_SKIP;

    if (k0 != ITEM_SKIP) {
      // epilogue operands of THIS row are requested now and consumed after the gathers
      float4 pre_acc[VPL], pre_res[VPL];
      const bool row_out = (k0 == ITEM_ROW) && lane < G;
#pragma unroll
      for (int q = 0; q < VPL; ++q) {
        const int f = lane + q * G;
        const bool ok = row_out && f < p.d4;
        pre_acc[q] = (ok && p.acc_in) ? ld_once_f4(reinterpret_cast<const float4*>(p.acc_in) + (size_t)r0 * p.d4 + f) : f4_zero();
        pre_res[q] = (ok && p.resid) ? ld_once_f4(reinterpret_cast<const float4*>(p.resid) + (size_t)r0 * p.d4 + f) : f4_zero();
      }
      float4 acc[VPL];
#pragma unroll
      for (int q = 0; q < VPL; ++q) acc[q] = f4_zero();
      accumulate_slice_pf<G, VPL, UNROLL>(p, s0, e0, lane, c0, v0, acc);
      if (lane < G) {
        if (k0 == ITEM_TASK) {
          float4* out = reinterpret_cast<float4*>(p.partial) + (size_t)w * p.d4;
#pragma unroll
          for (int q = 0; q < VPL; ++q) {
            const int f = lane + q * G;
            if (f < p.d4) st_f4(out + f, acc[q]);
          }
        } else {
          const size_t rowoff = (size_t)r0 * p.d4;
          const int deg = e0 - s0;
#pragma unroll
          for (int q = 0; q < VPL; ++q) {
            const int f = lane + q * G;
            if (f >= p.d4) continue;
            float4 y = acc[q];
            if (p.mean) {
              const float c = (float)max(deg, 1);
              y.x = __fdiv_rn(y.x, c); y.y = __fdiv_rn(y.y, c); y.z = __fdiv_rn(y.z, c); y.w = __fdiv_rn(y.w, c);
            }
            if (p.resid) y = f4_add(y, pre_res[q]);
            if (p.Y) st_f4(reinterpret_cast<float4*>(p.Y) + rowoff + f, y);
            if (p.acc_out) {
              float4 a = p.acc_in ? f4_add(pre_acc[q], y) : y;
              if (p.acc_div != 1.0f) {
                a.x = __fdiv_rn(a.x, p.acc_div); a.y = __fdiv_rn(a.y, p.acc_div);
                a.z = __fdiv_rn(a.z, p.acc_div); a.w = __fdiv_rn(a.w, p.acc_div);
              }
              st_f4(reinterpret_cast<float4*>(p.acc_out) + rowoff + f, a);
            }
          }
        }
      }
    }
    r0 = r1; s0 = s1; e0 = e1; k0 = k1; c0 = c1; v0 = v1;
    r1 = r2; s1 = s2; e1 = e2; k1 = k2;
  }
}



// ---- sub-warp rows: one lane GROUP per short row ------------------------------------------------------------
// Probe (tools/dist_probe.py, tools/spmm_probe.py): with many short rows the launch time is (#rows / resident warps) x
// the per-row dependent chain rowptr -> (col,val) -> gathers -> epilogue, NOT bytes.  For d <= 64 a row is only G = d/4
// lanes wide, so a warp can run 32/G independent row chains at once: warp w owns rows NG*w .. NG*w+NG-1, one per lane
// group, when all of them are short (<= SUBW_MAX non-zeros); otherwise it walks them one by one with the whole warp
// (the spmm_rows_kernel path).  No cross-group reduction in sub-warp mode; per-row summation order is plain ascending.
//
// WIDE (variant 16; measured in round 2 -- every plan the autotune picks on the H&M graph uses it): a slice of a long row is shared by the SPMM_WARPS warps of ONE CTA (warp i takes
// the 32-entry batches i, i+4, ...; the four sums are added in warp order through shared memory).  Scaling measurements
// (profiles/README.md r1c) fit  t_launch = 0.09 ms + nnz / 34 G/s : the constant is the dependent-iteration chain of one
// chunk-sized slice at gather-unroll 2, which bounds the launch from below once the graph is sharded 4-8 ways.  Sharing
// the slice cuts that chain by SPMM_WARPS without more registers, more partial rows or more long rows.
constexpr int SUBW_MAX = 64;

// PF (variant 18; measured in round 2: within 2 % of the plain form, never the winner on the H&M graph): shortens the dependent chain of a short row by two memory round trips -- the
// epilogue operands (acc_in / resid rows, streamed from DRAM) are requested into L2 as soon as the row id is known, and the
// second batch of (col,val) pairs of rows with more than G entries is loaded before the first batch is consumed.
// VPL > 1 (variants 20/21; measured in round 2: the winner on item-row views, profiles/README.md r2a): a lane holds VPL float4 of the row, so a row needs only G = d/(4*VPL) lanes
// and a warp runs 32/G row chains at once (d = 64: G = 8, VPL = 2 -> four rows per warp, half the shuffles and address
// arithmetic per non-zero).
// W256 (variants 23-25, d = 64 only; measured in round 2: variant 23 is the plan that ships at N = 1, r2d / r2n): the short-row path fetches a lane's two float4 with ONE 256-bit
// load (sm_100's LDG.E.256) -- a 256-byte row = one load instruction of 8 lanes, four rows per warp-level load.
template <int G, int UNROLL, int MINB, bool WIDE = false, int D4C = 0, bool PF = false, int VPL = 1, bool W256 = false, bool FILTER = false>
__global__ void __launch_bounds__(SPMM_WARPS * 32, MINB) spmm_subwarp_kernel(const SpmmParams p) {
  static_assert(G % UNROLL == 0, "the unrolled gather step must divide the lane-group width");
  static_assert(D4C == 0 || D4C == G * VPL, "a compile-time d/4 must fill the lane group exactly");
  static_assert(!W256 || (VPL == 2 && D4C == 2 * G), "256-bit gathers need two adjacent float4 per lane and an exact fit");
  constexpr int NG = 32 / G;
  const int lane = threadIdx.x & 31;
  const int grp = lane / G, lig = lane % G;
  const int d4 = D4C ? D4C : p.d4;
  const float4* __restrict__ X4 = reinterpret_cast<const float4*>(p.X);
  int64_t w;
  if (WIDE) {
    if ((int64_t)blockIdx.x < p.n_tasks) {   // one CTA per slice
      __shared__ float4 wsum[SPMM_WARPS][G * VPL];
      const int wi = threadIdx.x >> 5;
      const int64_t t = p.task_exec ? p.task_exec[blockIdx.x] : (int64_t)blockIdx.x;
      const int s = p.task_start[t], e = p.task_end[t];
      float4 acc[VPL];
#pragma unroll
      for (int q = 0; q < VPL; ++q) acc[q] = f4_zero();
      accumulate_slice<G, VPL, UNROLL, 32 * SPMM_WARPS, uint32_t, D4C, W256, FILTER>(p, s + 32 * wi, e, lane, acc);
      if (lane < G) {
#pragma unroll
        for (int q = 0; q < VPL; ++q) wsum[wi][W256 ? lane * VPL + q : lane + q * G] = acc[q];   // indexed by float4 of the row
      }
      __syncthreads();
      if (wi == 0 && lane < G) {
#pragma unroll
        for (int q = 0; q < VPL; ++q) {
          const int f = lane + q * G;
          if (f >= d4) continue;
          float4 sum = wsum[0][f];
#pragma unroll
          for (int k = 1; k < SPMM_WARPS; ++k) sum = f4_add(sum, wsum[k][f]);
          st_f4(reinterpret_cast<float4*>(p.partial) + (size_t)t * d4 + f, sum);
        }
      }
      if (wi == 0 && p.task_seg) slice_done(p, t, lane);     // fused stage 2: the warp that stored the partial row
      return;
    }
    w = p.n_tasks + ((int64_t)blockIdx.x - p.n_tasks) * SPMM_WARPS + (threadIdx.x >> 5);
  } else {
    w = (int64_t)blockIdx.x * SPMM_WARPS + (threadIdx.x >> 5);
  }

  if (!WIDE && w < p.n_tasks) {   // slices of long rows: whole warp, partial sums
    float4 acc[VPL];
#pragma unroll
    for (int q = 0; q < VPL; ++q) acc[q] = f4_zero();
    const int64_t t = p.task_exec ? p.task_exec[w] : w;
    const int s = p.task_start[t], e = p.task_end[t];
    accumulate_slice<G, VPL, UNROLL, 32, uint32_t, D4C, W256, FILTER>(p, s, e, lane, acc);
    if (lane < G) {
#pragma unroll
      for (int q = 0; q < VPL; ++q) {
        const int f = W256 ? lane * VPL + q : lane + q * G;
        if (f < d4) st_f4(reinterpret_cast<float4*>(p.partial) + (size_t)t * d4 + f, acc[q]);
      }
    }
    if (p.task_seg) slice_done(p, t, lane);
    return;
  }
  const int64_t first = (w - p.n_tasks) * NG;
  if (first >= p.n_rows) return;
  // every group fetches the bounds of its own row
  const int64_t ri = first + grp;
  int r = -1, s = 0, e = 0;
  if (ri < p.n_rows) {
    r = p.row_order ? p.row_order[ri] : (int)ri;
    s = p.rowptr[r];
    e = p.rowptr[r + 1];
  }
  int deg = e - s;
  const bool is_long = p.chunk > 0 && deg > p.chunk;   // handled by its slices + stage 2
  if (is_long) { deg = 0; }
  if (PF && r >= 0 && !(p.y_tail && r >= p.split_row)) {   // one request per 128-byte line of the epilogue operands
#pragma unroll
    for (int q = 0; q < VPL; ++q) {
      const int f = W256 ? lig * VPL + q : lig + q * G;
      if ((f & 7) == 0 && f < d4) {
        const size_t o = (size_t)r * d4 + f;
        if (p.acc_in) prefetch_l2(reinterpret_cast<const float4*>(p.acc_in) + o);
        if (p.resid) prefetch_l2(reinterpret_cast<const float4*>(p.resid) + o);
      }
    }
  }
  const bool all_short = __all_sync(FULL_MASK, deg <= SUBW_MAX);

  if (all_short) {
    int maxdeg = deg;
#pragma unroll
    for (int off = G; off < 32; off <<= 1) maxdeg = max(maxdeg, __shfl_xor_sync(FULL_MASK, maxdeg, off));
    float4 acc[VPL];
#pragma unroll
    for (int q = 0; q < VPL; ++q) acc[q] = f4_zero();
    const uint32_t d4u = (uint32_t)d4;
    int c = 0, cn = 0;
    float wv = 0.f, wn = 0.f;
    const uint64_t pol = PF ? l2_policy_evict_first() : 0;      // PF variants: streamed operands leave L2 to the gathered rows
    if (PF && lig < deg) {                         // batch 0 of the (col,val) pairs
      c = ld_stream_i32_hint(p.colidx + s + lig, pol);
      wv = p.val ? ld_stream_f32_hint(p.val + s + lig, pol) : 1.f;
    }
    for (int base = 0; FILTER && base < maxdeg; base += G) {
      // row-sparse operand: every lane tests its own column; the groups pop their hits from the ballot mask, one per step
      c = -1; wv = 0.f;
      if (base + lig < deg) {
        c = ld_stream_i32(p.colidx + s + base + lig);
        if (!filter_hit(p.filter, c)) c = -1;
        else wv = p.val ? ld_stream_f32(p.val + s + base + lig) : 1.f;
      }
      unsigned gm = (__ballot_sync(FULL_MASK, c >= 0) >> (grp * G)) & (G == 32 ? 0xffffffffu : ((1u << G) - 1u));
      while (__any_sync(FULL_MASK, gm != 0)) {
        const int k = gm ? __ffs((int)gm) - 1 : -1;
        gm &= gm - 1u;
        const int cc = __shfl_sync(FULL_MASK, c, k & (G - 1), G);
        const float wk = __shfl_sync(FULL_MASK, wv, k & (G - 1), G);
        if (k >= 0) {
          const uint32_t row = (uint32_t)cc * d4u;
          if (W256) {
            const f4x2 t = ld_gather_f8(X4 + (row + (uint32_t)(lig * 2)));
            f4_fma(acc[0], wk, t.a);
            f4_fma(acc[VPL - 1], wk, t.b);
          } else {
#pragma unroll
            for (int q = 0; q < VPL; ++q) {
              const int f = lig + q * G;
              if (f < d4) f4_fma(acc[q], wk, ld_gather_f4(X4 + (row + (uint32_t)f)));
            }
          }
        }
      }
    }
    for (int base = 0; !FILTER && base < maxdeg; base += G) {
      if (PF) {                                    // the NEXT batch is requested before this one is consumed
        cn = 0; wn = 0.f;
        if (base + G + lig < deg) {
          cn = ld_stream_i32_hint(p.colidx + s + base + G + lig, pol);
          wn = p.val ? ld_stream_f32_hint(p.val + s + base + G + lig, pol) : 1.f;
        }
      } else {
        c = 0; wv = 0.f;
        if (base + lig < deg) {
          c = ld_stream_i32(p.colidx + s + base + lig);
          wv = p.val ? ld_stream_f32(p.val + s + base + lig) : 1.f;
        }
      }
      const int cnt = min(G, deg - base);          // may be <= 0 for the shorter rows of the warp
      const int cntmax = min(G, maxdeg - base);
      for (int j = 0; j < cntmax; j += UNROLL) {    // j + u < G: UNROLL divides G
        float4 v[UNROLL][VPL];
        float ww[UNROLL];
        bool ok[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
          const int k = j + u;
          const int cc = __shfl_sync(FULL_MASK, c, k, G);
          ww[u] = __shfl_sync(FULL_MASK, wv, k, G);
          ok[u] = k < cnt;
          const uint32_t row = (uint32_t)cc * d4u;
          if (W256) {
            if (ok[u]) {
              const f4x2 t = ld_gather_f8(X4 + (row + (uint32_t)(lig * 2)));
              v[u][0] = t.a;
              v[u][VPL - 1] = t.b;
            }
          } else {
#pragma unroll
            for (int q = 0; q < VPL; ++q) {
              const int f = lig + q * G;
              if (ok[u] && f < d4) v[u][q] = ld_gather_f4(X4 + (row + (uint32_t)f));
            }
          }
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u)
#pragma unroll
          for (int q = 0; q < VPL; ++q)
            if (ok[u] && (W256 || lig + q * G < d4)) f4_fma(acc[q], ww[u], v[u][q]);
      }
      if (PF) { c = cn; wv = wn; }
    }
    if (r >= 0 && !is_long) epilogue_row<G, VPL, PF, W256>(p, r, e - s, lig, acc);
    return;
  }
  // mixed / longer rows: the whole warp walks the NG rows one after the other
#pragma unroll 1
  for (int g2 = 0; g2 < NG; ++g2) {
    const int rr = __shfl_sync(FULL_MASK, r, g2 * G);
    const int ss = __shfl_sync(FULL_MASK, s, g2 * G);
    const int ee = __shfl_sync(FULL_MASK, e, g2 * G);
    if (rr < 0 || (p.chunk > 0 && ee - ss > p.chunk)) continue;
    float4 acc[VPL];
#pragma unroll
    for (int q = 0; q < VPL; ++q) acc[q] = f4_zero();
    accumulate_slice<G, VPL, UNROLL, 32, uint32_t, D4C, W256, FILTER>(p, ss, ee, lane, acc);
    if (lane < G) epilogue_row<G, VPL, false, W256>(p, rr, ee - ss, lane, acc);
  }
}

// Threads of one CTA of the flat stage-2 kernel (the fallback of the tree kernel below; 1024-thread CTAs were measured and
// lose on plans with many short long-rows: their final combine walks 128 groups whatever the row holds).
template <int G, int VPL>
constexpr int reduce_threads() { return 256; }

template <int G, int VPL>
static int launch_stage2(const SpmmParams& p, cudaStream_t stream);

template <int G, int VPL>
__global__ void spmm_long_reduce_kernel(const SpmmParams p);

template <int G, int UNROLL, int MINB, bool WIDE = false, int D4C = 0, bool PF = false, int VPL = 1, bool W256 = false, bool FILTER = false>
static int launch_subwarp_impl(const SpmmParams& p, cudaStream_t stream);

// d/4 == G*VPL (d = 64 with G = 16, d = 32 with G = 8) gets the kernel specialised on that constant
template <int G, int UNROLL, int MINB, bool WIDE = false, bool PF = false, int VPL = 1>
static int launch_subwarp(const SpmmParams& p, cudaStream_t stream) {
  return p.d4 == G * VPL ? launch_subwarp_impl<G, UNROLL, MINB, WIDE, G * VPL, PF, VPL>(p, stream)
                         : launch_subwarp_impl<G, UNROLL, MINB, WIDE, 0, PF, VPL>(p, stream);
}

template <int G, int UNROLL, int MINB, bool WIDE, int D4C, bool PF, int VPL, bool W256, bool FILTER>
static int launch_subwarp_impl(const SpmmParams& p_in, cudaStream_t stream) {
  constexpr int NG = 32 / G;
  SpmmParams p = p_in;
  p.task_seg = p_in.task_seg_plan;                   // this family runs the fused stage 2 when the call allows it
  const int64_t row_warps = (p.n_rows + NG - 1) / NG;
  const int64_t warps = p.n_tasks + row_warps;
  if (warps > 0) {
    const int64_t blocks = WIDE ? p.n_tasks + (row_warps + SPMM_WARPS - 1) / SPMM_WARPS : (warps + SPMM_WARPS - 1) / SPMM_WARPS;
    LGB_REQUIRE(blocks < (1ll << 31), LGB_ERANGE, "lgb_spmm: grid too large");
    spmm_subwarp_kernel<G, UNROLL, MINB, WIDE, D4C, PF, VPL, W256, FILTER><<<(unsigned)blocks, SPMM_WARPS * 32, 0, stream>>>(p);
    LGB_LAUNCH_CHECK();
  }
  if (p.task_seg) return LGB_OK;                     // stage 2 ran inside the launch (slice_done)
  { const int rc2 = launch_stage2<G, VPL>(p, stream); if (rc2) return rc2; }
  return LGB_OK;
}

// ---- async-copy (LDGSTS) gather pipeline --------------------------------------------------------------
// The ncu captures (profiles/r1b_*) show the gather latency-bound even at 64 resident warps: a warp can only keep
// UNROLL register-backed loads in flight.  Here every lane owns S 16-byte slots of shared memory and streams its
// piece of the gathered rows through them with cp.async (SASS LDGSTS): S gathers per lane in flight at NO
// register cost, the FMA reads the slot back with a conflict-free LDS.128.  Only the issuing lane ever touches its
// slots, so cp.async.wait_group is the only synchronisation.  (col,val) batches are double-buffered in registers.
// Summation order per row is the same as spmm_rows_kernel (entries ascending within a lane group).
__device__ __forceinline__ void cp_async_16(void* smem_dst, const void* gmem_src) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int G, int S, int MINB>
__global__ void __launch_bounds__(SPMM_WARPS * 32, MINB) spmm_async_kernel(const SpmmParams p) {
  constexpr int NG = 32 / G;
  constexpr int STEPS_PER_BATCH = 32 / NG;
  static_assert(S <= STEPS_PER_BATCH, "pipeline may run at most one (col,val) batch ahead");
  __shared__ float4 ring[SPMM_WARPS][S][32];
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const int grp = lane / G, lig = lane % G;
  const int64_t w = (int64_t)blockIdx.x * SPMM_WARPS + wib;
  const int d4 = p.d4;
  const float4* __restrict__ X4 = reinterpret_cast<const float4*>(p.X);
  float4* my = &ring[wib][0][lane];

  int r, s, e;
  bool is_task = false;
  if (w < p.n_tasks) {
    r = p.task_row[w]; s = p.task_start[w]; e = p.task_end[w]; is_task = true;
  } else {
    const int64_t ri = w - p.n_tasks;
    if (ri >= p.n_rows) return;
    r = p.row_order ? p.row_order[ri] : (int)ri;
    s = p.rowptr[r]; e = p.rowptr[r + 1];
    if (p.chunk > 0 && e - s > p.chunk) return;
  }
  const int n_ent = e - s;
  const int T = (n_ent + NG - 1) / NG;   // pipeline steps of this item

  // (col,val) of batch 0 and batch 1
  int c_cur = 0, c_nxt = 0;
  float w_cur = 0.f, w_nxt = 0.f;
  if (lane < n_ent) { c_cur = ld_stream_i32(p.colidx + s + lane); w_cur = p.val ? ld_stream_f32(p.val + s + lane) : 1.f; }
  if (32 + lane < n_ent) { c_nxt = ld_stream_i32(p.colidx + s + 32 + lane); w_nxt = p.val ? ld_stream_f32(p.val + s + 32 + lane) : 1.f; }

  auto issue = [&](int ti, int consume_batch) {
    if (ti < T) {
      const int k = ti * NG + grp;
      const int kk = k & 31;
      const int cc = ((k >> 5) == consume_batch) ? __shfl_sync(FULL_MASK, c_cur, kk) : __shfl_sync(FULL_MASK, c_nxt, kk);
      if (k < n_ent && lig < d4) cp_async_16(my + (ti % S) * 32, X4 + (size_t)cc * d4 + lig);
    }
    cp_async_commit();
  };

  float4 acc = f4_zero();
#pragma unroll
  for (int ti = 0; ti < S - 1; ++ti) issue(ti, 0);
  for (int t = 0; t < T; ++t) {
    const int cb = (t * NG) >> 5;
    issue(t + S - 1, cb);
    cp_async_wait<S - 1>();
    const int k = t * NG + grp;
    const float wk = __shfl_sync(FULL_MASK, w_cur, k & 31);
    if (k < n_ent && lig < d4) {
      const float4 v = my[(t % S) * 32];
      f4_fma(acc, wk, v);
    }
    if (((t + 1) * NG & 31) == 0) {   // consume side crosses into the next batch: rotate and prefetch the one after
      c_cur = c_nxt; w_cur = w_nxt;
      c_nxt = 0; w_nxt = 0.f;
      const int idx = (cb + 2) * 32 + lane;
      if (idx < n_ent) { c_nxt = ld_stream_i32(p.colidx + s + idx); w_nxt = p.val ? ld_stream_f32(p.val + s + idx) : 1.f; }
    }
  }
  cp_async_wait<0>();
#pragma unroll
  for (int off = G; off < 32; off <<= 1) acc = f4_add(acc, f4_shfl_xor(acc, off));
  float4 a1[1] = {acc};
  if (lane < G) {
    if (is_task) {
      if (lane < d4) st_f4(reinterpret_cast<float4*>(p.partial) + (size_t)w * d4 + lane, acc);
    } else {
      epilogue_row<G, 1>(p, r, n_ent, lane, a1);
    }
  }
}

template <int G, int VPL>
__global__ void spmm_long_reduce_kernel(const SpmmParams p);

template <int G, int S, int MINB>
static int launch_async(const SpmmParams& p, cudaStream_t stream) {
  const int64_t items = p.n_tasks + p.n_rows;
  if (items > 0) {
    const int64_t blocks = (items + SPMM_WARPS - 1) / SPMM_WARPS;
    LGB_REQUIRE(blocks < (1ll << 31), LGB_ERANGE, "lgb_spmm: grid too large");
    static bool configured = false;
    if (!configured) {
      LGB_CUDA(cudaFuncSetAttribute(spmm_async_kernel<G, S, MINB>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
      configured = true;
    }
    spmm_async_kernel<G, S, MINB><<<(unsigned)blocks, SPMM_WARPS * 32, 0, stream>>>(p);
    LGB_LAUNCH_CHECK();
  }
  { const int rc2 = launch_stage2<G, 1>(p, stream); if (rc2) return rc2; }
  return LGB_OK;
}

// Stage 2: one CTA per long row sums that row's partials in a fixed order, then runs the epilogue.
template <int G, int VPL>
__global__ void __launch_bounds__(reduce_threads<G, VPL>()) spmm_long_reduce_kernel(const SpmmParams p) {
  constexpr int NGRP = reduce_threads<G, VPL>() / G;
  __shared__ float4 sm[NGRP][G * VPL];
  const int L = blockIdx.x;
  const int grp = threadIdx.x / G;
  const int lig = threadIdx.x % G;
  const int t0 = p.long_ptr[L], t1 = p.long_ptr[L + 1];
  const float4* part = reinterpret_cast<const float4*>(p.partial);
  float4 acc[VPL];
#pragma unroll
  for (int q = 0; q < VPL; ++q) acc[q] = f4_zero();
  // RU independent loads in flight per lane: the heaviest row of a power-law graph has thousands of partial rows, and a loop
  // with one load in flight made this kernel 85 us (7 % of the launch, ncu r2a) -- the order of the additions stays fixed
  constexpr int RU = 4;
  for (int t = t0 + grp; t < t1; t += NGRP * RU) {
    float4 v[RU][VPL];
#pragma unroll
    for (int u = 0; u < RU; ++u) {
      const int tt = t + u * NGRP;
#pragma unroll
      for (int q = 0; q < VPL; ++q) {
        const int f = lig + q * G;
        v[u][q] = (tt < t1 && f < p.d4) ? ld_stream_f4(part + (size_t)tt * p.d4 + f) : f4_zero();
      }
    }
#pragma unroll
    for (int u = 0; u < RU; ++u)
#pragma unroll
      for (int q = 0; q < VPL; ++q) pin_f4(v[u][q]);       // all RU * VPL loads are in flight before the first add
#pragma unroll
    for (int u = 0; u < RU; ++u)
#pragma unroll
      for (int q = 0; q < VPL; ++q) acc[q] = f4_add(acc[q], v[u][q]);
  }
#pragma unroll
  for (int q = 0; q < VPL; ++q) sm[grp][lig + q * G] = acc[q];
  __syncthreads();
  if (grp == 0) {
#pragma unroll
    for (int q = 0; q < VPL; ++q) {
      float4 t = sm[0][lig + q * G];
      for (int g2 = 1; g2 < NGRP; ++g2) t = f4_add(t, sm[g2][lig + q * G]);
      acc[q] = t;
    }
    const int r = p.long_rows[L];
    epilogue_row<G, VPL>(p, r, p.rowptr[r + 1] - p.rowptr[r], lig, acc);
  }
}


// ---- stage 2 as a tree (default when the plan carries segments) -------------------------------------------------------------
// The flat kernel above gives one CTA per long row and walks its partial rows with one or two loads in flight: 85-145 us per
// call on the H&M graph (the most popular item has 4 000-8 000 partial rows), 13-40 us on a 1/8 shard -- 7-13 % of every
// lgb_spmm call (ncu launch lists r2d / r2f).  Here the partial rows of a long row are cut at plan time into segments of 32;
// one WARP per segment: lane l loads partial row t0 + l, column by column (16 independent 128-bit loads in flight per lane --
// no add depends on another load), a shuffle butterfly adds the 32 rows.  A row of one segment goes straight to the
// epilogue; otherwise the warp stores a level-2 partial row, takes a ticket, and the LAST warp of the row to arrive adds the
// level-2 rows (same procedure, fixed order) and runs the epilogue.  Deterministic: which warp does the final pass depends on
// timing, what it computes does not.  Tickets reset themselves.
__device__ __forceinline__ float4 ld_cg_f4(const float4* p) {    // level-2 rows were written by other SMs: bypass the L1
  float4 v;
  asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}

constexpr int TREE_CB = 16;   // float4 columns handled per pass (64 floats: the whole row at d = 64)

template <bool CG>
__device__ __forceinline__ void tree_sum32(const float4* __restrict__ rows, int t0, int t1, int d4, int c0, int lane, float4 (&v)[TREE_CB]) {
  const bool have = t0 + lane < t1;
  const float4* src = rows + (size_t)(t0 + lane) * d4 + c0;
#pragma unroll
  for (int f = 0; f < TREE_CB; ++f)
    v[f] = (have && c0 + f < d4) ? (CG ? ld_cg_f4(src + f) : ld_stream_f4(src + f)) : f4_zero();
#pragma unroll
  for (int off = 16; off > 0; off >>= 1)
#pragma unroll
    for (int f = 0; f < TREE_CB; ++f) v[f] = f4_add(v[f], f4_shfl_xor(v[f], off));
}

__device__ __forceinline__ float4 pick_col(const float4 (&v)[TREE_CB], int lane) {
  float4 m = v[0];
#pragma unroll
  for (int f = 1; f < TREE_CB; ++f)
    if ((lane & (TREE_CB - 1)) == f) m = v[f];
  return m;
}

// the epilogue of epilogue_row for ONE float4 column f of row r
__device__ __forceinline__ void epilogue_col(const SpmmParams& p, int r, int deg, int f, float4 y) {
  if (p.y_tail && r >= p.split_row) {
    st_f4(reinterpret_cast<float4*>(p.y_tail) + (size_t)(r - p.split_row) * p.d4 + f, y);
    return;
  }
  const size_t o = (size_t)r * p.d4 + f;
  if (p.mean) {
    const float c = (float)max(deg, 1);
    y.x = __fdiv_rn(y.x, c); y.y = __fdiv_rn(y.y, c); y.z = __fdiv_rn(y.z, c); y.w = __fdiv_rn(y.w, c);
  }
  if (p.resid) y = f4_add(y, ld_once_f4(reinterpret_cast<const float4*>(p.resid) + o));
  if (p.Y) st_f4(reinterpret_cast<float4*>(p.Y) + o, y);
  if (p.acc_out) {
    float4 a = y;
    if (p.acc_in) a = f4_add(ld_once_f4(reinterpret_cast<const float4*>(p.acc_in) + o), y);
    if (p.acc_div != 1.0f) {
      a.x = __fdiv_rn(a.x, p.acc_div); a.y = __fdiv_rn(a.y, p.acc_div); a.z = __fdiv_rn(a.z, p.acc_div); a.w = __fdiv_rn(a.w, p.acc_div);
    }
    st_f4(reinterpret_cast<float4*>(p.acc_out) + o, a);
  }
}

__global__ void __launch_bounds__(128) spmm_long_reduce_tree_kernel(const SpmmParams p) {
  const int lane = threadIdx.x & 31;
  const int64_t seg = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
  if (seg >= p.n_seg) return;
  const int L = p.seg_row[seg];
  const int t0 = p.seg_t0[seg], t1 = p.seg_t1[seg];
  const int s0 = p.row_seg0[L], nseg = p.row_seg0[L + 1] - s0;
  const int r = p.long_rows[L];
  const int deg = p.rowptr[r + 1] - p.rowptr[r];
  const int d4 = p.d4;
  const float4* part = reinterpret_cast<const float4*>(p.partial);
  float4* part2 = reinterpret_cast<float4*>(p.part2);
  float4 v[TREE_CB];
  if (nseg == 1) {                                   // the whole row in one segment: sum and finish
    for (int c0 = 0; c0 < d4; c0 += TREE_CB) {
      tree_sum32<false>(part, t0, t1, d4, c0, lane, v);
      if (lane < TREE_CB && c0 + lane < d4) epilogue_col(p, r, deg, c0 + lane, pick_col(v, lane));
    }
    return;
  }
  for (int c0 = 0; c0 < d4; c0 += TREE_CB) {         // level-2 partial row of this segment
    tree_sum32<false>(part, t0, t1, d4, c0, lane, v);
    if (lane < TREE_CB && c0 + lane < d4) st_f4(part2 + (size_t)seg * d4 + c0 + lane, pick_col(v, lane));
  }
  __threadfence();
  __syncwarp();                                      // EVERY lane's stores are fenced before lane 0 takes the ticket
  int ticket = 0;
  if (lane == 0) ticket = atomicAdd(p.tickets + L, 1);
  ticket = __shfl_sync(FULL_MASK, ticket, 0);
  if (ticket != nseg - 1) return;
  __threadfence();                                   // every other segment of the row has published its level-2 row
  for (int c0 = 0; c0 < d4; c0 += TREE_CB) {
    float4 tot = f4_zero();
    for (int b = s0; b < s0 + nseg; b += 32) {       // fixed order: segment blocks ascending, butterfly inside a block
      tree_sum32<true>(part2, b, min(b + 32, s0 + nseg), d4, c0, lane, v);
      tot = f4_add(tot, pick_col(v, lane));
    }
    if (lane < TREE_CB && c0 + lane < d4) epilogue_col(p, r, deg, c0 + lane, tot);
  }
  if (lane == 0) p.tickets[L] = 0;
}

// Fused stage 2 (LGB_SPMM_FUSED_STAGE2): called by the ONE warp that has just stored partial row t.  The last slice of a
// segment to arrive adds the segment's partial rows, the last segment of a row adds the level-2 rows and runs the epilogue --
// the tree kernel above, executed by whoever finishes last.  Four float4 columns per pass (16 registers of payload: this code
// shares the 32-register budget of the gather kernel it is linked into, and it is off the hot path -- one call in 32 does work).
constexpr int FUSE_CB = 4;

__device__ __forceinline__ void fuse_sum32(const float4* __restrict__ rows, int t0, int t1, int d4, int c0, int lane, float4 (&v)[FUSE_CB]) {
  const bool have = t0 + lane < t1;
  const float4* src = rows + (size_t)(t0 + lane) * d4 + c0;
#pragma unroll
  for (int f = 0; f < FUSE_CB; ++f) v[f] = (have && c0 + f < d4) ? ld_cg_f4(src + f) : f4_zero();   // written by other SMs: L2
#pragma unroll
  for (int off = 16; off > 0; off >>= 1)
#pragma unroll
    for (int f = 0; f < FUSE_CB; ++f) v[f] = f4_add(v[f], f4_shfl_xor(v[f], off));
}

__device__ __forceinline__ float4 fuse_pick(const float4 (&v)[FUSE_CB], int lane) {
  float4 m = v[0];
#pragma unroll
  for (int f = 1; f < FUSE_CB; ++f)
    if ((lane & (FUSE_CB - 1)) == f) m = v[f];
  return m;
}

__device__ __forceinline__ void slice_done(const SpmmParams& p, int64_t t, int lane) {
  __threadfence();                                   // partial row t is visible device-wide before the ticket is taken ...
  __syncwarp();                                      // ... by lane 0, on behalf of EVERY lane that stored a piece of it
  const int seg = p.task_seg[t];
  const int t0 = p.seg_t0[seg], t1 = p.seg_t1[seg];
  int ticket = 0;
  if (lane == 0) ticket = atomicAdd(p.seg_tickets + seg, 1);
  ticket = __shfl_sync(FULL_MASK, ticket, 0);
  if (ticket != t1 - t0 - 1) return;
  __threadfence();                                   // every other slice of the segment has published its partial row
  if (lane == 0) p.seg_tickets[seg] = 0;             // nobody takes this ticket again in this launch
  const int L = p.seg_row[seg];
  const int s0 = p.row_seg0[L], nseg = p.row_seg0[L + 1] - s0;
  const int r = p.long_rows[L];
  const int deg = p.rowptr[r + 1] - p.rowptr[r];
  const int d4 = p.d4;
  const float4* part = reinterpret_cast<const float4*>(p.partial);
  float4* part2 = reinterpret_cast<float4*>(p.part2);
  float4 v[FUSE_CB];
  if (nseg == 1) {
    for (int c0 = 0; c0 < d4; c0 += FUSE_CB) {
      fuse_sum32(part, t0, t1, d4, c0, lane, v);
      if (lane < FUSE_CB && c0 + lane < d4) epilogue_col(p, r, deg, c0 + lane, fuse_pick(v, lane));
    }
    return;
  }
  for (int c0 = 0; c0 < d4; c0 += FUSE_CB) {
    fuse_sum32(part, t0, t1, d4, c0, lane, v);
    if (lane < FUSE_CB && c0 + lane < d4) st_f4(part2 + (size_t)seg * d4 + c0 + lane, fuse_pick(v, lane));
  }
  __threadfence();
  __syncwarp();
  if (lane == 0) ticket = atomicAdd(p.tickets + L, 1);
  ticket = __shfl_sync(FULL_MASK, ticket, 0);
  if (ticket != nseg - 1) return;
  __threadfence();
  for (int c0 = 0; c0 < d4; c0 += FUSE_CB) {
    float4 tot = f4_zero();
    for (int b = s0; b < s0 + nseg; b += 32) {       // fixed order: segment blocks ascending, butterfly inside a block
      fuse_sum32(part2, b, min(b + 32, s0 + nseg), d4, c0, lane, v);
      tot = f4_add(tot, fuse_pick(v, lane));
    }
    if (lane < FUSE_CB && c0 + lane < d4) epilogue_col(p, r, deg, c0 + lane, tot);
  }
  if (lane == 0) p.tickets[L] = 0;
}

// Stage 2 of every kernel family: the tree when the plan carries segments (and the caller sized the scratch for it), else flat.
template <int G, int VPL>
static int launch_stage2(const SpmmParams& p, cudaStream_t stream) {
  if (p.n_long <= 0) return LGB_OK;
  if (p.n_seg > 0 && p.seg_row && p.part2 && p.tickets) {
    spmm_long_reduce_tree_kernel<<<(unsigned)((p.n_seg + 3) / 4), 128, 0, stream>>>(p);
  } else {
    constexpr int RT = reduce_threads<G, VPL>();
    spmm_long_reduce_kernel<G, VPL><<<(unsigned)p.n_long, RT, 0, stream>>>(p);
  }
  LGB_LAUNCH_CHECK();
  return LGB_OK;
}

// Scalar path for d % 4 != 0 (tiny test shapes): one warp per row, lane strides over the columns.
__global__ void __launch_bounds__(SPMM_WARPS * 32) spmm_scalar_kernel(const SpmmParams p, int d) {
  const int lane = threadIdx.x & 31;
  const int64_t w = (int64_t)blockIdx.x * SPMM_WARPS + (threadIdx.x >> 5);
  if (w >= p.n_rows) return;
  const int r = (int)w;
  const int s = p.rowptr[r], e = p.rowptr[r + 1];
  for (int f = lane; f < d; f += 32) {
    float acc = 0.f;
    for (int i = s; i < e; ++i) acc = fmaf(p.val ? p.val[i] : 1.f, p.X[(size_t)p.colidx[i] * d + f], acc);
    float y = acc;
    if (p.mean) y = __fdiv_rn(y, (float)max(e - s, 1));
    const size_t o = (size_t)r * d + f;
    if (p.resid) y += p.resid[o];
    if (p.Y) p.Y[o] = y;
    if (p.acc_out) {
      float a = p.acc_in ? p.acc_in[o] + y : y;
      if (p.acc_div != 1.0f) a = __fdiv_rn(a, p.acc_div);
      p.acc_out[o] = a;
    }
  }
}

static int g_sm_count = 0;
static int sm_count() {
  if (g_sm_count == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      g_sm_count = n;
    else
      return 148;
  }
  return g_sm_count;
}

template <int G, int VPL, int UNROLL, int MINB = 1, typename IT = uint32_t, int D4C = 0, bool W256 = false>
static int launch_vec(const SpmmParams& p, int variant, cudaStream_t stream) {
  const int64_t items = p.n_tasks + p.n_rows;
  if (items > 0) {
    const int64_t blocks = (items + SPMM_WARPS - 1) / SPMM_WARPS;
    LGB_REQUIRE(blocks < (1ll << 31), LGB_ERANGE, "lgb_spmm: grid too large");
    if (variant == 1 || (p.n_tasks > 0 && !p.task_end)) {
      spmm_rows_kernel<G, VPL, UNROLL, MINB, IT, D4C, W256><<<(unsigned)blocks, SPMM_WARPS * 32, 0, stream>>>(p);
    } else {
      // persistent grid: every SM filled to the occupancy limit, each warp walks items w, w+W, ...
      static int occ = 0;
      if (occ == 0) {
        LGB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, spmm_pipe_kernel<G, VPL, UNROLL>, SPMM_WARPS * 32, 0));
        if (occ <= 0) occ = 1;
      }
      const int64_t resident = (int64_t)sm_count() * occ;
      spmm_pipe_kernel<G, VPL, UNROLL><<<(unsigned)std::min<int64_t>(blocks, resident), SPMM_WARPS * 32, 0, stream>>>(p);
    }
    LGB_LAUNCH_CHECK();
  }
  { const int rc2 = launch_stage2<G, VPL>(p, stream); if (rc2) return rc2; }
  return LGB_OK;
}


// ---- hot-column cache (variant 30): the H most-referenced operand rows live in shared memory ------------------------------
// ncu on the shipped kernel (profiles/r2a_*): 10.3 GB cross the L2 -> SM path per launch at an L1 hit rate of 33 %, long
// scoreboard 26 per issue -- every gathered row pays an L2 (or DRAM) round trip in a dependent chain.  On a power-law graph a
// handful of columns carries most of the non-zeros (H&M shape: the 256 most popular items are referenced by 65 % of the
// user-row entries), far more than the L1 keeps alive next to the streamed rows.  Here every CTA first copies the plan's
// n_hot hottest rows of X into shared memory (64 KB: 256 rows at d = 64, two 1024-thread CTAs per SM; 128 KB: 512 rows, one
// CTA); the plan's recoded column array marks a hot column with ~slot, and such an entry is served by one LDS.128 instead of
// an L2 round trip.  CTAs are persistent (the copy is paid once per CTA): warp w of the grid walks the work items
// w, w + W, ... in the plan's order (slices of long rows first, then row groups) -- the same per-row summation order as the
// sub-warp kernel, hence bit-identical results.
constexpr int HOT_THREADS = 1024;
constexpr int HOT_WARPS = HOT_THREADS / 32;
constexpr int HOT_GRAB = 8;        // row groups a warp takes per visit of the work counter
constexpr int HOT_CTR_FLOATS = 64; // the work counters live in 64 floats behind the n_tasks * d partial sums (caller-zeroed once)

template <int G, int UNROLL, int D4C>
__device__ __forceinline__ float4 gather_hot(const float4* __restrict__ X4, const float4* hot, int cc, int lig) {
  return cc < 0 ? hot[(~cc) * D4C + lig] : ld_gather_f4(X4 + ((uint32_t)cc * (uint32_t)D4C + (uint32_t)lig));
}

template <int G, int UNROLL, int D4C>
__device__ __forceinline__ void accumulate_slice_hot(const SpmmParams& p, const int32_t* __restrict__ colh, const float4* hot, int s,
                                                     int e, int lane, float4& acc) {
  constexpr int NG = 32 / G;
  static_assert(32 % (NG * UNROLL) == 0, "a batch of 32 entries must be a whole number of unrolled steps");
  const int grp = lane / G, lig = lane % G;
  const float4* __restrict__ X4 = reinterpret_cast<const float4*>(p.X);
  for (int base = s; base < e; base += 32) {
    const int idx = base + lane;
    int c = 0;
    float w = 0.f;
    if (idx < e) {
      c = ld_stream_i32(colh + idx);
      w = p.val ? ld_stream_f32(p.val + idx) : 1.f;
    }
    const int cnt = min(32, e - base);
    for (int j = 0; j < cnt; j += NG * UNROLL) {
      float4 v[UNROLL];
      float ww[UNROLL];
      bool ok[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        const int k = j + u * NG + grp;
        const int cc = __shfl_sync(FULL_MASK, c, k);
        ww[u] = __shfl_sync(FULL_MASK, w, k);
        ok[u] = k < cnt;
        if (ok[u]) v[u] = gather_hot<G, UNROLL, D4C>(X4, hot, cc, lig);
      }
#pragma unroll
      for (int u = 0; u < UNROLL; ++u)
        if (ok[u]) f4_fma(acc, ww[u], v[u]);
    }
  }
#pragma unroll
  for (int off = G; off < 32; off <<= 1) acc = f4_add(acc, f4_shfl_xor(acc, off));
}

template <int G, int UNROLL, int D4C, int MINB>
__global__ void __launch_bounds__(HOT_THREADS, MINB) spmm_hot_kernel(const SpmmParams p, const int32_t* __restrict__ colh,
                                                                    const int32_t* __restrict__ hot_cols, int n_hot) {
  static_assert(D4C == G, "one float4 per lane of the group");
  constexpr int NG = 32 / G;
  float4* hot = dyn_smem_f4();
  const float4* __restrict__ X4 = reinterpret_cast<const float4*>(p.X);
  for (int i = threadIdx.x; i < n_hot * D4C; i += HOT_THREADS)
    hot[i] = ld_gather_f4(X4 + ((size_t)hot_cols[i / D4C] * D4C + (i % D4C)));
  __syncthreads();

  const int lane = threadIdx.x & 31;
  const int grp = lane / G, lig = lane % G;
  // 32-bit work-item arithmetic (the launcher checks that the item count fits): keeps the persistent loop inside the register budget
  const int n_tasks = (int)p.n_tasks, n_rows = (int)p.n_rows;
  const int row_items = (n_rows + NG - 1) / NG;
  // Dynamic work distribution (a static round-robin over persistent warps loses 1.5x on a power-law graph: the warp that
  // draws a 1000-entry row still has its whole share of other rows to walk).  Two global counters behind the partial-sum
  // scratch: slices are taken one at a time, row groups HOT_GRAB at a time; the last CTA to finish resets them, so a launch
  // leaves them zero for the next one on the same scratch buffer.
  int* ctr = reinterpret_cast<int*>(p.partial + (size_t)n_tasks * (D4C * 4));
  for (;;) {          // ---- phase 1: slices of long rows (whole warp, partial sums) ----
    int w = 0;
    if (lane == 0) w = n_tasks > 0 ? atomicAdd(ctr, 1) : 0;
    w = __shfl_sync(FULL_MASK, w, 0);
    if (w >= n_tasks) break;
    float4 acc = f4_zero();
    accumulate_slice_hot<G, UNROLL, D4C>(p, colh, hot, p.task_start[w], p.task_end[w], lane, acc);
    if (lane < G) st_f4(reinterpret_cast<float4*>(p.partial) + (size_t)w * D4C + lane, acc);
  }
  for (;;) {          // ---- phase 2: row groups ----
    int base = 0;
    if (lane == 0) base = atomicAdd(ctr + 1, HOT_GRAB);
    base = __shfl_sync(FULL_MASK, base, 0);
    if (base >= row_items) break;
    const int last = min(base + HOT_GRAB, row_items);
#pragma unroll 1
    for (int it = base; it < last; ++it) {
      const int ri = it * NG + grp;      // every lane group fetches the bounds of its own row
      int r = -1, s = 0, e = 0;
      if (ri < n_rows) {
        r = p.row_order ? p.row_order[ri] : ri;
        s = p.rowptr[r];
        e = p.rowptr[r + 1];
      }
      int deg = e - s;
      const bool is_long = p.chunk > 0 && deg > p.chunk;   // handled by its slices + stage 2
      if (is_long) deg = 0;
      if (__all_sync(FULL_MASK, deg <= SUBW_MAX)) {
        int maxdeg = deg;
#pragma unroll
        for (int off = G; off < 32; off <<= 1) maxdeg = max(maxdeg, __shfl_xor_sync(FULL_MASK, maxdeg, off));
        float4 acc[1] = {f4_zero()};
        for (int b0 = 0; b0 < maxdeg; b0 += G) {
          int c = 0;
          float wv = 0.f;
          if (b0 + lig < deg) {
            c = ld_stream_i32(colh + s + b0 + lig);
            wv = p.val ? ld_stream_f32(p.val + s + b0 + lig) : 1.f;
          }
          const int cnt = min(G, deg - b0);
          const int cntmax = min(G, maxdeg - b0);
          for (int j = 0; j < cntmax; j += UNROLL) {
            float4 v[UNROLL];
            float ww[UNROLL];
            bool ok[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
              const int k = j + u;
              const int cc = __shfl_sync(FULL_MASK, c, k, G);
              ww[u] = __shfl_sync(FULL_MASK, wv, k, G);
              ok[u] = k < cnt;
              if (ok[u]) v[u] = gather_hot<G, UNROLL, D4C>(X4, hot, cc, lig);
            }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u)
              if (ok[u]) f4_fma(acc[0], ww[u], v[u]);
          }
        }
        if (r >= 0 && !is_long) epilogue_row<G, 1>(p, r, e - s, lig, acc);
        continue;
      }
      // mixed / longer rows: the whole warp walks the NG rows one after the other
#pragma unroll 1
      for (int g2 = 0; g2 < NG; ++g2) {
        const int rr = __shfl_sync(FULL_MASK, r, g2 * G);
        const int ss = __shfl_sync(FULL_MASK, s, g2 * G);
        const int ee = __shfl_sync(FULL_MASK, e, g2 * G);
        if (rr < 0 || (p.chunk > 0 && ee - ss > p.chunk)) continue;
        float4 acc[1] = {f4_zero()};
        accumulate_slice_hot<G, UNROLL, D4C>(p, colh, hot, ss, ee, lane, acc[0]);
        if (lane < G) epilogue_row<G, 1>(p, rr, ee - ss, lane, acc);
      }
    }
  }
  // the last CTA to get here puts the counters back to zero (every warp of every CTA has left both loops by then)
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(ctr + 2, 1) == (int)gridDim.x - 1) {
      ctr[0] = 0; ctr[1] = 0; ctr[2] = 0;
      __threadfence();
    }
  }
}

// MINB = 2: two 1024-thread CTAs per SM (32 registers per thread, 64 resident warps, <= 100 KB of hot rows each);
// MINB = 1: one CTA per SM (up to 64 registers, 32 resident warps, up to 200 KB of hot rows).
template <int G, int UNROLL, int MINB>
static int launch_hot(const SpmmParams& p, const lgb_csr* g, cudaStream_t stream) {
  constexpr int NG = 32 / G;
  const size_t smem = (size_t)g->n_hot * G * sizeof(float4);
  const size_t cap = (MINB == 2 ? 100 : 200) * 1024;
  LGB_REQUIRE(smem <= cap, LGB_EINVAL, "lgb_spmm: hot plan of %d rows needs %zu bytes of shared memory (this variant holds %zu)",
              (int)g->n_hot, smem, cap);
  static bool configured = false;
  if (!configured) {
    LGB_CUDA(cudaFuncSetAttribute(spmm_hot_kernel<G, UNROLL, G, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cap));
    configured = true;
  }
  const int64_t total = p.n_tasks + (p.n_rows + NG - 1) / NG;
  LGB_REQUIRE(total < (1ll << 31) - (1 << 20), LGB_ERANGE, "lgb_spmm: too many work items for the hot-column kernel");
  LGB_REQUIRE(p.partial, LGB_EINVAL, "lgb_spmm: variants 30 / 31 need partial_ws of n_tasks*d + %d floats, zero before the first launch "
              "(the work counters live behind the partial sums)", HOT_CTR_FLOATS);
  if (total > 0) {
    const int64_t resident = (int64_t)sm_count() * MINB;
    const int64_t blocks = std::min<int64_t>(resident, (total + HOT_WARPS - 1) / HOT_WARPS);
    spmm_hot_kernel<G, UNROLL, G, MINB><<<(unsigned)blocks, HOT_THREADS, smem, stream>>>(p, g->colidx_hot, g->hot_cols, g->n_hot);
    LGB_LAUNCH_CHECK();
  }
  { const int rc2 = launch_stage2<G, 1>(p, stream); if (rc2) return rc2; }
  return LGB_OK;
}

// ---- segment max (PyG aggr="max") ----------------------------------------------------------
__global__ void segment_max_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                                   const float* __restrict__ X, int64_t n_rows, int d, float* __restrict__ Y,
                                   int32_t* __restrict__ argmax) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= n_rows) return;
  const int s = rowptr[r], e = rowptr[r + 1];
  for (int f = lane; f < d; f += 32) {
    float best = 0.f;
    int arg = -1;
    for (int i = s; i < e; ++i) {
      const int c = colidx[i];
      const float v = X[(size_t)c * d + f];
      if (arg < 0 || v > best) { best = v; arg = c; }
    }
    Y[(size_t)r * d + f] = best;
    if (argmax) argmax[(size_t)r * d + f] = arg;
  }
}
__global__ void segment_max_bwd_kernel(const float* __restrict__ gY, const int32_t* __restrict__ argmax, int64_t total,
                                       int d, float* __restrict__ gX) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int a = argmax[i];
  if (a >= 0) atomicAdd(gX + (size_t)a * d + (i % d), gY[i]);
}

__global__ void row_div_kernel(const float* __restrict__ X, const int32_t* __restrict__ rowptr, int64_t n_rows, int d,
                               float* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rows * d) return;
  const int64_t r = i / d;
  out[i] = __fdiv_rn(X[i], (float)max(rowptr[r + 1] - rowptr[r], 1));
}

// y, acc and out may alias (in-place accumulate): no __restrict__ on them
__global__ void accumulate_kernel(const float* y, const float* acc, const float* __restrict__ resid, int64_t n,
                                  float div, float* out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    float v = y[i];
    if (resid) v += resid[i];
    if (acc) v = acc[i] + v;
    if (div != 1.0f) v = __fdiv_rn(v, div);
    out[i] = v;
  }
}

// out = (((s0 + s1) + s2) + ...) / div over up to MEAN_MAX equally-shaped buffers: the layer mean of LightGCN
// (model/lightgcn.py:67-68) in ONE pass when the layer outputs are kept (the sharded engine's two independent layer chains),
// same left-to-right fp32 order as the fused accumulate epilogue.
constexpr int MEAN_MAX = 8;
struct MeanSrcs { const float* s[MEAN_MAX]; };
__global__ void __launch_bounds__(256) mean_rows_kernel(const MeanSrcs srcs, int count, int64_t n, float div, float* __restrict__ out, int vec) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  if (vec) {
    const int64_t n4 = n / 4;
    for (; i < n4; i += stride) {
      float4 a = ld_stream_f4(reinterpret_cast<const float4*>(srcs.s[0]) + i);
      for (int k = 1; k < count; ++k) a = f4_add(a, ld_stream_f4(reinterpret_cast<const float4*>(srcs.s[k]) + i));
      if (div != 1.0f) { a.x = __fdiv_rn(a.x, div); a.y = __fdiv_rn(a.y, div); a.z = __fdiv_rn(a.z, div); a.w = __fdiv_rn(a.w, div); }
      st_f4(reinterpret_cast<float4*>(out) + i, a);
    }
  } else {
    for (; i < n; i += stride) {
      float a = srcs.s[0][i];
      for (int k = 1; k < count; ++k) a += srcs.s[k][i];
      out[i] = div != 1.0f ? __fdiv_rn(a, div) : a;
    }
  }
}

__global__ void scale_concat_kernel(const float4* __restrict__ a, int64_t na4, const float4* __restrict__ b, int64_t nb4,
                                    float scale, float4* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < na4 + nb4; i += stride) {
    float4 v = i < na4 ? (a ? ld_stream_f4(a + i) : f4_zero()) : (b ? ld_stream_f4(b + (i - na4)) : f4_zero());
    v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale;
    st_f4(out + i, v);
  }
}
__global__ void scale_concat_scalar_kernel(const float* __restrict__ a, int64_t na, const float* __restrict__ b,
                                           int64_t nb, float scale, float* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < na + nb; i += stride) {
    float v = i < na ? (a ? a[i] : 0.f) : (b ? b[i - na] : 0.f);
    out[i] = v * scale;
  }
}

// out = scale * a, and which rows of a hold a non-zero: bit r of bitmap, count[0] += rows flagged.  G = d/4 lanes per row (a
// power of two <= 32), so a warp covers 32/G whole rows per step and one ballot finds them.
template <int G>
__global__ void __launch_bounds__(256) scale_rows_nonzero_kernel(const float4* __restrict__ a, int64_t n4, float scale,
                                                                 float4* __restrict__ out, int64_t row_offset,
                                                                 uint32_t* __restrict__ bitmap, int* __restrict__ count) {
  const int lane = threadIdx.x & 31;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t n4_up = (n4 + 31) / 32 * 32;                      // warp-uniform trip count (256-thread CTAs, 32-aligned starts)
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4_up; i += stride) {
    bool nz = false;
    if (i < n4) {
      float4 v = ld_stream_f4(a + i);
      nz = v.x != 0.f || v.y != 0.f || v.z != 0.f || v.w != 0.f;
      v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale;
      st_f4(out + i, v);
    }
    const unsigned m = __ballot_sync(FULL_MASK, nz);
    const unsigned gm = (m >> ((lane / G) * G)) & (G == 32 ? 0xffffffffu : ((1u << G) - 1u));
    if (gm != 0 && lane % G == 0) {
      const int64_t row = row_offset + i / G;
      atomicOr(bitmap + (row >> 5), 1u << (row & 31));
      atomicAdd(count, 1);
    }
  }
}

}  // namespace lgb

namespace lgb {
__global__ void __launch_bounds__(256) rows_bitmap_kernel(const int64_t* __restrict__ idx, int64_t n, int64_t offset, int64_t n_bits,
                                                          uint32_t* __restrict__ bitmap) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t c = idx[i] + offset;
  if (c >= 0 && c < n_bits) atomicOr(bitmap + (c >> 5), 1u << (c & 31));
}
}  // namespace lgb

using namespace lgb;

extern "C" {

int lgb_rows_bitmap(const int64_t* idx, int64_t n, int64_t offset, int64_t n_bits, uint32_t* bitmap, void* stream_) {
  LGB_REQUIRE(n >= 0 && n_bits >= 0 && (n == 0 || (idx && bitmap)), LGB_EINVAL, "lgb_rows_bitmap: bad argument");
  if (n == 0) return LGB_OK;
  rows_bitmap_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream_>>>(idx, n, offset, n_bits, bitmap);
  LGB_LAUNCH_CHECK();
  return LGB_OK;
}

int lgb_scale_rows_nonzero(const float* a, int64_t n_rows, int32_t d, float scale, float* out, int64_t row_offset, uint32_t* bitmap,
                           int32_t* count, void* stream_) {
  LGB_REQUIRE(n_rows >= 0 && d > 0 && row_offset >= 0 && (n_rows == 0 || (a && out && bitmap && count)), LGB_EINVAL, "lgb_scale_rows_nonzero: bad argument");
  const int d4 = d / 4;
  LGB_REQUIRE(d % 4 == 0 && d4 <= 32 && (d4 & (d4 - 1)) == 0 && ((((uintptr_t)a | (uintptr_t)out) & 15) == 0), LGB_EINVAL,
              "lgb_scale_rows_nonzero: d must be 4, 8, 16, 32, 64 or 128 and the rows 16-byte aligned");
  if (n_rows == 0) return LGB_OK;
  const int64_t n4 = n_rows * d4;
  const unsigned blocks = (unsigned)std::min<int64_t>((n4 + 255) / 256, (int64_t)sm_count() * 16);
  cudaStream_t st = (cudaStream_t)stream_;
  const float4* a4 = (const float4*)a;
  float4* o4 = (float4*)out;
  int* cnt = (int*)count;
  switch (d4) {
    case 1: scale_rows_nonzero_kernel<1><<<blocks, 256, 0, st>>>(a4, n4, scale, o4, row_offset, bitmap, cnt); break;
    case 2: scale_rows_nonzero_kernel<2><<<blocks, 256, 0, st>>>(a4, n4, scale, o4, row_offset, bitmap, cnt); break;
    case 4: scale_rows_nonzero_kernel<4><<<blocks, 256, 0, st>>>(a4, n4, scale, o4, row_offset, bitmap, cnt); break;
    case 8: scale_rows_nonzero_kernel<8><<<blocks, 256, 0, st>>>(a4, n4, scale, o4, row_offset, bitmap, cnt); break;
    case 16: scale_rows_nonzero_kernel<16><<<blocks, 256, 0, st>>>(a4, n4, scale, o4, row_offset, bitmap, cnt); break;
    default: scale_rows_nonzero_kernel<32><<<blocks, 256, 0, st>>>(a4, n4, scale, o4, row_offset, bitmap, cnt); break;
  }
  LGB_LAUNCH_CHECK();
  return LGB_OK;
}

static int spmm_impl(const lgb_csr* g, const float* X, int32_t d, float* Y, const float* resid, const float* acc_in,
                     float* acc_out, float acc_div, int32_t flags, float* partial_ws, int64_t split_row, float* y_tail,
                     void* stream_, const uint32_t* filter = nullptr, const uint32_t* resid_filter = nullptr);

int lgb_spmm_rowsparse(const lgb_csr* g, const float* X, const uint32_t* x_row_bitmap, int32_t d, float* Y, const float* resid,
                       const uint32_t* resid_row_bitmap, const float* acc_in, float* acc_out, float acc_div, int32_t flags,
                       float* partial_ws, void* stream_) {
  LGB_REQUIRE(x_row_bitmap, LGB_EINVAL, "lgb_spmm_rowsparse: null bitmap");
  return spmm_impl(g, X, d, Y, resid, acc_in, acc_out, acc_div, flags, partial_ws, 0, nullptr, stream_, x_row_bitmap, resid_row_bitmap);
}

int lgb_spmm(const lgb_csr* g, const float* X, int32_t d, float* Y, const float* resid, const float* acc_in,
             float* acc_out, float acc_div, int32_t flags, float* partial_ws, void* stream_) {
  return spmm_impl(g, X, d, Y, resid, acc_in, acc_out, acc_div, flags, partial_ws, 0, nullptr, stream_);
}

int lgb_spmm_split(const lgb_csr* g, const float* X, int32_t d, float* Y, const float* resid, const float* acc_in,
                   float* acc_out, float acc_div, int32_t flags, float* partial_ws, int64_t split_row, float* y_tail,
                   void* stream_) {
  LGB_REQUIRE(g && y_tail && split_row >= 0 && split_row <= g->n_rows, LGB_EINVAL, "lgb_spmm_split: bad split_row / y_tail");
  LGB_REQUIRE(y_tail != X, LGB_EINVAL, "lgb_spmm_split: y_tail aliases the gathered operand");
  const int variant = (flags >> LGB_SPMM_VARIANT_SHIFT) & 0xFF;
  LGB_REQUIRE(variant != 2 && variant != 3, LGB_EINVAL, "lgb_spmm_split: the software-pipelined variants have no split epilogue");
  LGB_REQUIRE(d % 4 == 0, LGB_EINVAL, "lgb_spmm_split: d %% 4 != 0");
  return spmm_impl(g, X, d, Y, resid, acc_in, acc_out, acc_div, flags, partial_ws, split_row, y_tail, stream_);
}

}  // extern "C"

static int spmm_impl(const lgb_csr* g, const float* X, int32_t d, float* Y, const float* resid, const float* acc_in,
                     float* acc_out, float acc_div, int32_t flags, float* partial_ws, int64_t split_row, float* y_tail,
                     void* stream_, const uint32_t* filter, const uint32_t* resid_filter) {
  cudaStream_t stream = (cudaStream_t)stream_;
  LGB_REQUIRE(g && g->rowptr && X && d > 0, LGB_EINVAL, "lgb_spmm: null graph/X or d <= 0");
  LGB_REQUIRE(g->nnz == 0 || g->colidx, LGB_EINVAL, "lgb_spmm: null colidx");
  LGB_REQUIRE(Y || acc_out, LGB_EINVAL, "lgb_spmm: no output requested");
  LGB_REQUIRE(acc_div != 0.f, LGB_EINVAL, "lgb_spmm: acc_div == 0");
  LGB_REQUIRE(g->n_rows < (1ll << 31) - 1 && g->n_cols < (1ll << 31) - 1 && g->nnz < (1ll << 31), LGB_ERANGE,
              "lgb_spmm: size exceeds int32");
  LGB_REQUIRE(Y != X && acc_out != X, LGB_EINVAL, "lgb_spmm: output aliases the gathered operand");
  if (g->n_rows == 0) return LGB_OK;
  SpmmParams p;
  p.rowptr = g->rowptr; p.colidx = g->colidx; p.val = g->val; p.row_order = g->row_order;
  p.task_row = g->task_row; p.task_start = g->task_start; p.task_end = g->task_end; p.long_rows = g->long_rows; p.long_ptr = g->long_ptr;
  p.n_rows = g->n_rows; p.n_tasks = g->chunk > 0 ? g->n_tasks : 0; p.n_long = g->chunk > 0 ? g->n_long : 0;
  p.chunk = g->chunk; p.d4 = d / 4;
  p.X = X; p.Y = Y; p.resid = resid; p.acc_in = acc_in; p.acc_out = acc_out; p.acc_div = acc_div;
  p.mean = (flags & LGB_SPMM_MEAN) ? 1 : 0; p.partial = partial_ws;
  p.split_row = split_row; p.y_tail = y_tail;
  // stage-2 tree: segments from the plan, level-2 rows and tickets behind the partial sums and the 64 counter floats
  p.seg_row = g->seg_row; p.seg_t0 = g->seg_t0; p.seg_t1 = g->seg_t1; p.row_seg0 = g->row_seg0;
  p.n_seg = (g->chunk > 0 && g->seg_row && g->seg_t0 && g->seg_t1 && g->row_seg0 && (flags & LGB_SPMM_TREE_WS)) ? g->n_seg : 0;
  p.part2 = p.n_seg > 0 && partial_ws ? partial_ws + (size_t)p.n_tasks * d + 64 : nullptr;
  p.tickets = p.part2 ? reinterpret_cast<int*>(p.part2 + (size_t)p.n_seg * d) : nullptr;
  p.task_exec = p.n_tasks > 0 ? g->task_exec : nullptr;
  p.filter = nullptr; p.resid_filter = nullptr;
  // fused stage 2: only the sub-warp family runs it; the tickets of the segments sit behind the tree's row tickets
  p.task_seg = nullptr; p.seg_tickets = nullptr;
  const bool want_fused2 = (flags & LGB_SPMM_FUSED_STAGE2) && p.n_seg > 0 && p.tickets && g->task_seg;
  p.task_seg_plan = want_fused2 ? g->task_seg : nullptr;
  p.seg_tickets = want_fused2 ? p.tickets + p.n_long : nullptr;
  if (p.n_tasks > 0) {
    LGB_REQUIRE(partial_ws && g->task_row && g->task_start && g->task_end && g->long_rows && g->long_ptr, LGB_EINVAL,
                "lgb_spmm: plan has %lld tasks but plan arrays (task_row/start/end, long_rows/ptr) / partial workspace missing",
                (long long)p.n_tasks);
  }
  // vector paths: every row pointer the kernels touch with 128-bit (256-bit: variants 23-27) accesses must be aligned
  const uintptr_t all_ptrs = (uintptr_t)X | (uintptr_t)Y | (uintptr_t)resid | (uintptr_t)acc_in | (uintptr_t)acc_out |
                             (uintptr_t)partial_ws | (uintptr_t)y_tail;
  LGB_REQUIRE(d % 4 != 0 || (all_ptrs & 15) == 0, LGB_EINVAL,
              "lgb_spmm: X / Y / resid / acc_in / acc_out / partial_ws / y_tail must be 16-byte aligned when d %% 4 == 0");
  if (d % 4 != 0) {
    // scalar path ignores the plan (tiny shapes only)
    const int64_t blocks = (p.n_rows + SPMM_WARPS - 1) / SPMM_WARPS;
    spmm_scalar_kernel<<<(unsigned)blocks, SPMM_WARPS * 32, 0, stream>>>(p, d);
    LGB_LAUNCH_CHECK();
    return LGB_OK;
  }
  const int d4 = d / 4;
  // Variant 0 is the tuned default.  Measured on B200 (profiles/r1b_*): the kernel is latency-bound, so the
  // configuration that keeps 64 warps resident per SM (<= 32 registers: gather unroll 2, __launch_bounds__(128,16))
  // beats deeper unrolls (72 regs -> 28 warps) by 1.3x and the software-pipelined persistent variant by 1.5x.
  int variant = (flags >> LGB_SPMM_VARIANT_SHIFT) & 0xFF;
  if (variant >= 23 && variant <= 27 && ((all_ptrs & 31) != 0 || (d * 4) % 32 != 0))
    variant = 0;   // the 256-bit forms need 32-byte aligned rows: run the default kernel instead of faulting
  if ((int64_t)g->n_cols * d4 > 0x7fffffffll || variant == 17) {
    // element index into X does not fit 32 bits (a table of > 34 GB): the one kernel family compiled with 64-bit indexing
    // (variant 17 forces it, so the tests can reach it with small tables)
    if (d4 <= 8) return launch_vec<8, 1, 2, 16, size_t>(p, 1, stream);
    if (d4 <= 16) return launch_vec<16, 1, 2, 16, size_t>(p, 1, stream);
    if (d4 <= 32) return launch_vec<32, 1, 2, 16, size_t>(p, 1, stream);
    if (d4 <= 64) return launch_vec<32, 2, 2, 12, size_t>(p, 1, stream);
    if (d4 <= 128) return launch_vec<32, 4, 1, 8, size_t>(p, 1, stream);
    set_error("lgb_spmm: d=%d > 512 not supported", d);
    return LGB_EINVAL;
  }
  if (filter && d4 <= 16) {
    // row-sparse operand (lgb_spmm_rowsparse): the filtered forms of the sub-warp kernel with CTA-wide slices; other widths run
    // the dense kernels below (same result: the flagged rows are the only non-zero ones)
    p.filter = filter;
    p.resid_filter = resid ? resid_filter : nullptr;
    if (d4 == 16 && (all_ptrs & 31) == 0) return launch_subwarp_impl<8, 1, 16, true, 16, false, 2, true, true>(p, stream);
    if (d4 <= 8) return launch_subwarp_impl<8, 2, 16, true, 0, false, 1, false, true>(p, stream);
    return launch_subwarp_impl<16, 2, 16, true, 0, false, 1, false, true>(p, stream);
  }
  if (variant == 30 || variant == 31) {   // hot-column cache: needs the plan's recoded column array and an exact one-float4-per-lane width
    if (g->colidx_hot && g->hot_cols && g->n_hot > 0 && (d4 == 8 || d4 == 16) && !y_tail) {
      if (variant == 30) return d4 == 8 ? launch_hot<8, 1, 2>(p, g, stream) : launch_hot<16, 1, 2>(p, g, stream);
      return d4 == 8 ? launch_hot<8, 2, 1>(p, g, stream) : launch_hot<16, 2, 1>(p, g, stream);
    }
    variant = 0;
  }
  if (d4 <= 8) {
    if (variant == 1) return launch_vec<8, 1, 2, 16>(p, 1, stream);
    if (variant == 16) return launch_subwarp<8, 2, 16, true>(p, stream);
    if (variant == 18) return launch_subwarp<8, 2, 16, false, true>(p, stream);
    if (variant == 19) return launch_subwarp<8, 2, 16, true, true>(p, stream);
    return launch_subwarp<8, 2, 16>(p, stream);
  }
  if (d4 <= 16) {
    switch (variant) {
      case 0: return launch_subwarp<16, 2, 16>(p, stream);     // default: sub-warp rows, 64 resident warps
      case 12: return launch_vec<16, 1, 2, 16>(p, 1, stream);  // warp per row, unroll 2, 64 resident warps (r1b best)
      case 1: return launch_vec<16, 1, 8>(p, 1, stream);       // first version (r1a)
      case 2: return launch_vec<16, 1, 8>(p, 0, stream);       // software-pipelined persistent warps, unroll 8
      case 3: return launch_vec<16, 1, 4>(p, 0, stream);       // software-pipelined persistent warps, unroll 4
      case 4: return launch_vec<16, 1, 4>(p, 1, stream);
      case 5: return launch_vec<16, 1, 4, 12>(p, 1, stream);
      case 6: return launch_vec<16, 1, 1, 16>(p, 1, stream);
      case 7: return launch_async<16, 4, 16>(p, stream);
      case 8: return launch_async<16, 6, 16>(p, stream);
      case 9: return launch_async<16, 8, 12>(p, stream);
      case 10: return launch_async<16, 12, 8>(p, stream);
      case 11: return launch_async<16, 16, 6>(p, stream);
      case 13: return launch_subwarp<16, 4, 12>(p, stream);
      case 14: return launch_subwarp<16, 4, 10>(p, stream);
      case 15: return launch_subwarp<16, 8, 8>(p, stream);
      case 16: return launch_subwarp<16, 2, 16, true>(p, stream);   // sub-warp rows + one CTA per slice
      case 18: return launch_subwarp<16, 2, 16, false, true>(p, stream);   // sub-warp rows + chain-shortening prefetches
      case 19: return launch_subwarp<16, 2, 16, true, true>(p, stream);    // 16 + 18
      case 20: return launch_subwarp<8, 1, 16, true, false, 2>(p, stream);  // four rows per warp (8 lanes x 2 float4), unroll 1
      case 21: return launch_subwarp<8, 2, 12, true, false, 2>(p, stream);  // four rows per warp, unroll 2 (48 warps / SM)
      case 22: return launch_subwarp<8, 1, 16, true, true, 2>(p, stream);   // 20 + chain-shortening prefetches
      // 23-25: the same with ONE 256-bit load per lane and non-zero (LDG.E.256, new on sm_100); d = 64 exactly
      case 23: return d4 == 16 ? launch_subwarp_impl<8, 1, 16, true, 16, false, 2, true>(p, stream) : launch_subwarp<16, 2, 16>(p, stream);
      case 24: return d4 == 16 ? launch_subwarp_impl<8, 2, 12, true, 16, false, 2, true>(p, stream) : launch_subwarp<16, 2, 16>(p, stream);
      case 25: return d4 == 16 ? launch_subwarp_impl<8, 1, 16, true, 16, true, 2, true>(p, stream) : launch_subwarp<16, 2, 16>(p, stream);
      default: return launch_subwarp<16, 2, 16>(p, stream);
    }
  }
  if (d4 <= 32) {
    if (variant == 1) return launch_vec<32, 1, 8>(p, 1, stream);
    // 26: d = 128 with 256-bit gathers -- 16 lanes x 32 bytes per row, two rows per warp-level load (r2a sweep: d = 128 picks 26 / 27 on some shards)
    if (variant == 26 && d4 == 32) return launch_vec<16, 2, 1, 16, uint32_t, 32, true>(p, 1, stream);
    if (variant == 27 && d4 == 32) return launch_vec<16, 2, 2, 10, uint32_t, 32, true>(p, 1, stream);
    return launch_vec<32, 1, 2, 16>(p, 1, stream);
  }
  if (d4 <= 64) return launch_vec<32, 2, 2, 12>(p, 1, stream);
  if (d4 <= 128) return launch_vec<32, 4, 1, 8>(p, 1, stream);
  set_error("lgb_spmm: d=%d > 512 not supported", d);
  return LGB_EINVAL;
}

extern "C" {

int lgb_segment_max(const lgb_csr* g, const float* X, int32_t d, float* Y, int32_t* argmax, void* stream) {
  LGB_REQUIRE(g && g->rowptr && X && Y && d > 0, LGB_EINVAL, "lgb_segment_max: bad argument");
  if (g->n_rows == 0) return LGB_OK;
  const int warps = 4;
  const int64_t blocks = (g->n_rows + warps - 1) / warps;
  segment_max_kernel<<<(unsigned)blocks, warps * 32, 0, (cudaStream_t)stream>>>(g->rowptr, g->colidx, X, g->n_rows, d, Y,
                                                                                argmax);
  LGB_LAUNCH_CHECK();
  return LGB_OK;
}

int lgb_segment_max_bwd(const float* gY, const int32_t* argmax, int64_t n_rows, int32_t d, float* gX, void* stream) {
  LGB_REQUIRE(n_rows >= 0 && d > 0 && (n_rows == 0 || (gY && argmax && gX)), LGB_EINVAL, "lgb_segment_max_bwd: bad argument");
  const int64_t total = n_rows * d;
  if (total == 0) return LGB_OK;
  segment_max_bwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(gY, argmax, total, d, gX);
  LGB_LAUNCH_CHECK();
  return LGB_OK;
}

int lgb_row_div_by_degree(const float* X, const int32_t* rowptr, int64_t n_rows, int32_t d, float* out, void* stream) {
  LGB_REQUIRE(n_rows >= 0 && d > 0 && (n_rows == 0 || (X && rowptr && out)), LGB_EINVAL,
              "lgb_row_div_by_degree: bad argument");
  const int64_t total = n_rows * d;
  if (total == 0) return LGB_OK;
  row_div_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(X, rowptr, n_rows, d, out);
  LGB_LAUNCH_CHECK();
  return LGB_OK;
}

int lgb_accumulate(const float* y, const float* acc, const float* resid, int64_t n, float div, float* out, void* stream) {
  LGB_REQUIRE(n >= 0 && div != 0.f && (n == 0 || (y && out)), LGB_EINVAL, "lgb_accumulate: bad argument");
  if (n == 0) return LGB_OK;
  const unsigned blocks = (unsigned)std::min<int64_t>((n + 255) / 256, 148 * 16);
  accumulate_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(y, acc, resid, n, div, out);
  LGB_LAUNCH_CHECK();
  return LGB_OK;
}

int lgb_mean_rows(const float* const* srcs_host, int32_t count, int64_t n, float div, float* out, void* stream) {
  LGB_REQUIRE(srcs_host && count >= 1 && count <= MEAN_MAX && n >= 0 && div != 0.f && (n == 0 || out), LGB_EINVAL,
              "lgb_mean_rows: bad argument (1 <= count <= %d)", MEAN_MAX);
  if (n == 0) return LGB_OK;
  MeanSrcs srcs;
  uintptr_t bits = (uintptr_t)out;
  for (int k = 0; k < MEAN_MAX; ++k) {
    srcs.s[k] = k < count ? srcs_host[k] : nullptr;
    LGB_REQUIRE(k >= count || srcs.s[k], LGB_EINVAL, "lgb_mean_rows: source %d is NULL", k);
    LGB_REQUIRE(k >= count || srcs.s[k] != out, LGB_EINVAL, "lgb_mean_rows: out aliases source %d", k);
    bits |= (uintptr_t)srcs.s[k];
  }
  const int vec = (n % 4 == 0) && ((bits & 15) == 0);
  const int64_t work = vec ? n / 4 : n;
  const unsigned blocks = (unsigned)std::min<int64_t>((work + 255) / 256, 148 * 16);
  mean_rows_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(srcs, count, n, div, out, vec);
  LGB_LAUNCH_CHECK();
  return LGB_OK;
}

int lgb_zero(void* p, size_t bytes, void* stream) {
  LGB_REQUIRE(p || bytes == 0, LGB_EINVAL, "lgb_zero: null pointer");
  if (bytes) LGB_CUDA(cudaMemsetAsync(p, 0, bytes, (cudaStream_t)stream));
  return LGB_OK;
}

int lgb_scale_concat(const float* a, int64_t na, const float* b, int64_t nb, int32_t d, float scale, float* out,
                     void* stream) {
  LGB_REQUIRE(na >= 0 && nb >= 0 && d > 0 && out, LGB_EINVAL, "lgb_scale_concat: bad argument");
  const int64_t ea = na * d, eb = nb * d;
  if (ea + eb == 0) return LGB_OK;
  int sms = 148;
  const bool vec = (ea % 4 == 0) && (eb % 4 == 0) && ((((uintptr_t)a | (uintptr_t)b | (uintptr_t)out) & 15) == 0);
  if (vec) {
    const int64_t n4 = (ea + eb) / 4;
    const unsigned blocks = (unsigned)std::min<int64_t>((n4 + 255) / 256, (int64_t)sms * 16);
    scale_concat_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((const float4*)a, ea / 4, (const float4*)b, eb / 4,
                                                                  scale, (float4*)out);
  } else {
    const unsigned blocks = (unsigned)std::min<int64_t>((ea + eb + 255) / 256, (int64_t)sms * 16);
    scale_concat_scalar_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(a, ea, b, eb, scale, out);
  }
  LGB_LAUNCH_CHECK();
  return LGB_OK;
}

}  // extern "C"
