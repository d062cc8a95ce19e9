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

}  // extern "C"
