// Item-block exchange over NVLink peer memory / NVSwitch multicast (multi-GPU LightGCN, dist.py).
//
// The per-layer exchange of the sharded engine is a sum-all-reduce of the replicated [I, d] item block.  Instead
// of calling NCCL, these kernels run on buffers from CUDA symmetric memory (every rank maps every peer's copy, and
// -- with NVLS -- one multicast address that fans out to all copies):
//
//   multimem variant : rank g owns slice g.  One multimem.ld_reduce per 16 bytes returns the SUM over all ranks'
//                      copies (reduced inside the NVSwitch), one multimem.st writes it back to ALL copies.
//                      Per rank only n/G values cross its links in each direction.
//   peer variant     : no multicast: rank g reads slice g from every peer with plain P2P loads (fixed order
//                      0..G-1 => deterministic, identical on all ranks because each slice has ONE reducer) and
//                      stores the result into every peer's copy.
//
// Both are bracketed on the host by symmetric-memory barriers (all partial sums written before / all results
// visible after), so the kernels themselves never spin on another rank.
#include "common.cuh"

namespace lgb {

__device__ __forceinline__ float4 multimem_ld_reduce_f4(const float4* mc) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(mc)
               : "memory");
  return v;
}
__device__ __forceinline__ void multimem_st_f4(float4* mc, const float4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}

__global__ void __launch_bounds__(256) multimem_allreduce_kernel(float4* mc, int64_t lo4, int64_t hi4) {
  int64_t i = lo4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < hi4; i += stride) multimem_st_f4(mc + i, multimem_ld_reduce_f4(mc + i));
}

constexpr int PEER_MAX = 16;
struct PeerPtrs {
  float4* p[PEER_MAX];
};

__global__ void __launch_bounds__(256) peer_allreduce_kernel(PeerPtrs peers, int world, int64_t lo4, int64_t hi4) {
  int64_t i = lo4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < hi4; i += stride) {
    float4 acc = peers.p[0][i];
    for (int r = 1; r < world; ++r) acc = f4_add(acc, peers.p[r][i]);   // fixed order: deterministic
    for (int r = 0; r < world; ++r) peers.p[r][i] = acc;
  }
}

// ---- the exchange the sharded engine ships: barrier + reduce + republish + barrier in ONE launch ----------------------------
// Every rank launches the same grid.  Block b of rank r pairs with block b of every peer through a symmetric array of
// 32-bit signal slots (slot [b][src] on the destination rank), toggled with compare-and-swap so that a slot is consumed
// exactly once per use -- no epoch counter, hence replayable from a CUDA graph:
//   put : spin CAS(peer slot, 0 -> 1)          wait : spin CAS(own slot, 1 -> 0)
// Entry barrier: every rank's partial sums (written by the kernel that precedes this one in stream order) are complete.
// Exit barrier (after a system-scope fence): every slice has been republished to every rank before any rank's next kernel
// reads its buffer, and nobody overwrites a buffer a peer is still reading.
// The body is the slice loop of the kernels above, four 16-byte accesses in flight per thread.  A spin that lasts longer
// than ~4 s (a peer died, mismatched launch order) traps instead of hanging the GPU.
constexpr int XCHG_THREADS = 512;
constexpr int XCHG_MAX_BLOCKS = 128;      // signal slots are sized for this; the launch uses min(this, blocks asked for, work)
constexpr int XCHG_DEFAULT_BLOCKS = 64;
constexpr int XCHG_UNROLL = 4;
constexpr long long XCHG_SPIN_LIMIT = 8000000000ll;   // clock64 ticks (~4 s at 1.9 GHz)

struct XchgParams {
  float4* mc;                      // multicast address of the buffer (NULL: peer loads / stores)
  float4* peer[PEER_MAX];          // the buffer on every rank (UVA)
  uint32_t* pad[PEER_MAX];         // the signal-slot array on every rank
  int rank, world;
  int64_t lo4, hi4;                // this rank's slice, in float4
  int barriers;                    // 0: skip both barriers (single-process tests where the ranks run one after the other)
};

__device__ __forceinline__ uint32_t cas_sys(uint32_t* p, uint32_t cmp, uint32_t val, bool release) {
  uint32_t old;
  if (release)
    asm volatile("atom.release.sys.global.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "l"(p), "r"(cmp), "r"(val) : "memory");
  else
    asm volatile("atom.acquire.sys.global.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "l"(p), "r"(cmp), "r"(val) : "memory");
  return old;
}
__device__ __forceinline__ void fence_sys() { asm volatile("fence.acq_rel.sys;" ::: "memory"); }
__device__ __forceinline__ void trap_now() { asm volatile("trap;"); }

__device__ __forceinline__ void xchg_barrier(const XchgParams& p) {
  // thread t < world of every block: tell peer t "block b of rank `rank` is here", then wait for peer t's block b
  if ((int)threadIdx.x < p.world) {
    const int t = threadIdx.x;
    uint32_t* put = p.pad[t] + (size_t)blockIdx.x * p.world + p.rank;
    uint32_t* get = p.pad[p.rank] + (size_t)blockIdx.x * p.world + t;
    const long long t0 = clock64();
    while (cas_sys(put, 0u, 1u, true) != 0u)
      if (clock64() - t0 > XCHG_SPIN_LIMIT) { printf("lgb exchange: rank %d block %d timed out signalling rank %d\n", p.rank, (int)blockIdx.x, t); trap_now(); }
    while (cas_sys(get, 1u, 0u, false) != 1u)
      if (clock64() - t0 > XCHG_SPIN_LIMIT) { printf("lgb exchange: rank %d block %d timed out waiting for rank %d\n", p.rank, (int)blockIdx.x, t); trap_now(); }
  }
}

template <bool MC>
__global__ void __launch_bounds__(XCHG_THREADS) exchange_allreduce_kernel(const XchgParams p) {
  if (p.barriers) {
    xchg_barrier(p);
    __syncthreads();
  }
  const int64_t stride = (int64_t)gridDim.x * XCHG_THREADS;
  for (int64_t base = p.lo4 + (int64_t)blockIdx.x * XCHG_THREADS + threadIdx.x; base < p.hi4; base += stride * XCHG_UNROLL) {
    float4 v[XCHG_UNROLL];
#pragma unroll
    for (int u = 0; u < XCHG_UNROLL; ++u) {
      const int64_t i = base + u * stride;
      if (i < p.hi4) {
        if (MC) {
          v[u] = multimem_ld_reduce_f4(p.mc + i);
        } else {
          float4 acc = p.peer[0][i];
          for (int r = 1; r < p.world; ++r) acc = f4_add(acc, p.peer[r][i]);   // fixed order: deterministic
          v[u] = acc;
        }
      }
    }
#pragma unroll
    for (int u = 0; u < XCHG_UNROLL; ++u) {
      const int64_t i = base + u * stride;
      if (i < p.hi4) {
        if (MC) {
          multimem_st_f4(p.mc + i, v[u]);
        } else {
          for (int r = 0; r < p.world; ++r) p.peer[r][i] = v[u];
        }
      }
    }
  }
  if (p.barriers) {
    fence_sys();            // this thread's republished values are visible system-wide before the block signals
    __syncthreads();
    xchg_barrier(p);
  }
}

static void slice_of(int64_t n4, int rank, int world, int64_t& lo, int64_t& hi) {
  const int64_t per = (n4 + world - 1) / world;
  lo = per * rank < n4 ? per * rank : n4;
  hi = lo + per < n4 ? lo + per : n4;
}

}  // namespace lgb

using namespace lgb;

extern "C" {

int lgb_multimem_allreduce_f32(void* multicast_ptr, int64_t n_floats, int32_t rank, int32_t world, void* stream) {
  LGB_REQUIRE(multicast_ptr && n_floats >= 0 && world > 0 && rank >= 0 && rank < world, LGB_EINVAL,
              "lgb_multimem_allreduce_f32: bad argument");
  LGB_REQUIRE(n_floats % 4 == 0 && (((uintptr_t)multicast_ptr) & 15) == 0, LGB_EINVAL,
              "lgb_multimem_allreduce_f32: buffer must be 16-byte aligned with n %% 4 == 0");
  int64_t lo, hi;
  slice_of(n_floats / 4, rank, world, lo, hi);
  if (hi <= lo) return LGB_OK;
  const int64_t blocks = (hi - lo + 255) / 256;
  multimem_allreduce_kernel<<<(unsigned)(blocks < 1184 ? blocks : 1184), 256, 0, (cudaStream_t)stream>>>(
      (float4*)multicast_ptr, lo, hi);
  LGB_LAUNCH_CHECK();
  return LGB_OK;
}

int lgb_peer_allreduce_f32(const uint64_t* peer_ptrs_host, int64_t n_floats, int32_t rank, int32_t world, void* stream) {
  LGB_REQUIRE(peer_ptrs_host && n_floats >= 0 && world > 0 && world <= PEER_MAX && rank >= 0 && rank < world, LGB_EINVAL,
              "lgb_peer_allreduce_f32: bad argument (world <= %d)", PEER_MAX);
  LGB_REQUIRE(n_floats % 4 == 0, LGB_EINVAL, "lgb_peer_allreduce_f32: n %% 4 != 0");
  PeerPtrs pp;
  for (int r = 0; r < PEER_MAX; ++r) pp.p[r] = r < world ? (float4*)(uintptr_t)peer_ptrs_host[r] : nullptr;
  for (int r = 0; r < world; ++r)
    LGB_REQUIRE(pp.p[r] && (((uintptr_t)pp.p[r]) & 15) == 0, LGB_EINVAL, "lgb_peer_allreduce_f32: peer %d pointer null/unaligned", r);
  int64_t lo, hi;
  slice_of(n_floats / 4, rank, world, lo, hi);
  if (hi <= lo) return LGB_OK;
  const int64_t blocks = (hi - lo + 255) / 256;
  peer_allreduce_kernel<<<(unsigned)(blocks < 1184 ? blocks : 1184), 256, 0, (cudaStream_t)stream>>>(pp, world, lo, hi);
  LGB_LAUNCH_CHECK();
  return LGB_OK;
}

int lgb_exchange_pad_words(int32_t world) { return XCHG_MAX_BLOCKS * (world > 0 ? world : 1); }

int lgb_exchange_allreduce_f32(const lgb_exchange* x, int64_t byte_offset, int64_t n_floats, int32_t channel, int32_t flags,
                               void* stream) {
  LGB_REQUIRE(x && x->world > 0 && x->world <= PEER_MAX && x->rank >= 0 && x->rank < x->world, LGB_EINVAL,
              "lgb_exchange_allreduce_f32: bad rank / world (world <= %d)", PEER_MAX);
  LGB_REQUIRE(n_floats >= 0 && n_floats % 4 == 0 && byte_offset >= 0 && byte_offset % 16 == 0, LGB_EINVAL,
              "lgb_exchange_allreduce_f32: offset must be 16-byte aligned and n %% 4 == 0");
  LGB_REQUIRE(channel >= 0 && channel < x->n_channels, LGB_EINVAL, "lgb_exchange_allreduce_f32: channel %d of %d", (int)channel,
              (int)x->n_channels);
  const bool barriers = !(flags & LGB_EXCHANGE_NO_BARRIER);
  XchgParams p;
  p.mc = (x->multicast_base && !(flags & LGB_EXCHANGE_PEER)) ? (float4*)((char*)x->multicast_base + byte_offset) : nullptr;
  for (int r = 0; r < PEER_MAX; ++r) {
    p.peer[r] = r < x->world && x->peer_base[r] ? (float4*)((char*)x->peer_base[r] + byte_offset) : nullptr;
    p.pad[r] = r < x->world && x->pad_base[r] ? (uint32_t*)x->pad_base[r] + (size_t)channel * lgb_exchange_pad_words(x->world) : nullptr;
    if (r < x->world) {
      LGB_REQUIRE(p.mc || p.peer[r], LGB_EINVAL, "lgb_exchange_allreduce_f32: neither a multicast address nor peer %d's pointer", r);
      LGB_REQUIRE(!barriers || p.pad[r], LGB_EINVAL, "lgb_exchange_allreduce_f32: signal pad of rank %d missing", r);
    }
  }
  p.rank = x->rank; p.world = x->world; p.barriers = barriers ? 1 : 0;
  slice_of(n_floats / 4, x->rank, x->world, p.lo4, p.hi4);
  // the SAME grid on every rank (blocks pair up across ranks): sized from the slice length, which is rank-independent
  const int64_t per = (n_floats / 4 + x->world - 1) / x->world;
  int64_t blocks = (per + (int64_t)XCHG_THREADS * XCHG_UNROLL - 1) / ((int64_t)XCHG_THREADS * XCHG_UNROLL);
  int64_t cap = (flags >> LGB_EXCHANGE_BLOCKS_SHIFT) & 0xFF;
  cap = cap == 0 ? XCHG_DEFAULT_BLOCKS : (cap > XCHG_MAX_BLOCKS ? XCHG_MAX_BLOCKS : cap);
  blocks = blocks < 1 ? 1 : (blocks > cap ? cap : blocks);
  if (p.mc)
    exchange_allreduce_kernel<true><<<(unsigned)blocks, XCHG_THREADS, 0, (cudaStream_t)stream>>>(p);
  else
    exchange_allreduce_kernel<false><<<(unsigned)blocks, XCHG_THREADS, 0, (cudaStream_t)stream>>>(p);
  LGB_LAUNCH_CHECK();
  return LGB_OK;
}

}  // extern "C"
