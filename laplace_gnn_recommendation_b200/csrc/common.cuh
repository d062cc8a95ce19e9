// Shared helpers for liblaplace_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/laplace_b200.h"

namespace lgb {

void set_error(const char* fmt, ...);

#define LGB_REQUIRE(cond, code, ...)   \
  do {                                 \
    if (!(cond)) {                     \
      ::lgb::set_error(__VA_ARGS__);   \
      return (code);                   \
    }                                  \
  } while (0)

#define LGB_CUDA(expr)                                                                  \
  do {                                                                                  \
    cudaError_t _e = (expr);                                                            \
    if (_e != cudaSuccess) {                                                            \
      ::lgb::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                       __LINE__);                                                       \
      return LGB_ECUDA;                                                                 \
    }                                                                                   \
  } while (0)

#define LGB_LAUNCH_CHECK() LGB_CUDA(cudaGetLastError())

static inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

constexpr unsigned FULL_MASK = 0xffffffffu;

// ---- memory-access primitives -------------------------------------------------------------
// Streamed, read-once data (CSR arrays): read-only path, do not allocate in L1.
__device__ __forceinline__ int ld_stream_i32(const int32_t* p) {
  int v;
  asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ float ld_stream_f32(const float* p) {
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ float4 ld_stream_f4(const float4* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}
// Gathered embedding rows: 128-bit read-only loads (rows are re-used across the grid -> keep default L2 policy).
__device__ __forceinline__ float4 ld_gather_f4(const float4* p) {
  float4 v;
  asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}
// The same streamed loads with an L2 eviction policy (createpolicy descriptor, folded into the load's memory descriptor):
// read-once CSR arrays / epilogue operands marked evict_first leave the L2 to the gathered embedding rows.
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ int ld_stream_i32_hint(const int32_t* p, uint64_t pol) {
  int v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ float ld_stream_f32_hint(const float* p, uint64_t pol) {
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ float4 ld_stream_f4_hint(const float4* p, uint64_t pol) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p), "l"(pol));
  return v;
}
// Epilogue operands (acc_in / resid rows): read once, but the caller may update them IN PLACE (acc_out == acc_in), so they
// must not go through the read-only (.nc) path, whose data has to stay constant for the whole kernel.
__device__ __forceinline__ float4 ld_once_f4(const float4* p) {
  float4 v;
  asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p)
               : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_once_f4_hint(const float4* p, uint64_t pol) {
  float4 v;
  asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p), "l"(pol)
               : "memory");
  return v;
}
// Ask L2 for a line that will be read later in this thread's dependent chain (no register, no scoreboard entry).
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// 256-bit per-lane global accesses (new on sm_100: SASS LDG.E.256 / STG.E.256): 32 contiguous bytes of an embedding row in
// ONE instruction, i.e. a 256-byte row with 8 lanes.  The address must be 32-byte aligned.
struct f4x2 { float4 a, b; };
__device__ __forceinline__ f4x2 ld_gather_f8(const float4* p) {
  f4x2 v;
  asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(v.a.x), "=f"(v.a.y), "=f"(v.a.z), "=f"(v.a.w), "=f"(v.b.x), "=f"(v.b.y), "=f"(v.b.z), "=f"(v.b.w)
               : "l"(p));
  return v;
}
__device__ __forceinline__ f4x2 ld_stream_f8(const float4* p) {   // read-once 32 bytes: no L1 allocation
  f4x2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(v.a.x), "=f"(v.a.y), "=f"(v.a.z), "=f"(v.a.w), "=f"(v.b.x), "=f"(v.b.y), "=f"(v.b.z), "=f"(v.b.w)
               : "l"(p));
  return v;
}
__device__ __forceinline__ f4x2 ld_once_f8(const float4* p) {     // see ld_once_f4: may alias an output of the same kernel
  f4x2 v;
  asm volatile("ld.global.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(v.a.x), "=f"(v.a.y), "=f"(v.a.z), "=f"(v.a.w), "=f"(v.b.x), "=f"(v.b.y), "=f"(v.b.z), "=f"(v.b.w)
               : "l"(p)
               : "memory");
  return v;
}
__device__ __forceinline__ void st_f8(float4* p, const float4& a, const float4& b) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w), "f"(b.x),
               "f"(b.y), "f"(b.z), "f"(b.w)
               : "memory");
}
// Streaming 128-bit store (written once, read by a later kernel).
__device__ __forceinline__ void st_f4(float4* p, const float4& v) {
  asm volatile("st.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}
// Vector reduction into global memory (sm_90+): one 16-byte atomic instead of four scalar ones.
__device__ __forceinline__ void red_add_f4(float4* p, const float4& v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}

// Base of the dynamic shared memory of the launch (helper so that the CPU emulator can substitute a plain buffer).
__device__ __forceinline__ float4* dyn_smem_f4() {
  extern __shared__ float4 lgb_dyn_smem[];
  asm volatile("" ::: "memory");
  return lgb_dyn_smem;
}

// Scheduling fence for a batch of independent loads: placed after the LAST load of the batch, once per loaded value.  volatile
// asm statements keep their program order, so every load of the batch is issued before the first use of any of them -- without
// it ptxas interleaves each load with the add that consumes the previous one and the in-order issue stalls on the first add.
__device__ __forceinline__ void pin_f4(float4& v) { asm volatile("" : "+f"(v.x), "+f"(v.y), "+f"(v.z), "+f"(v.w)); }

__device__ __forceinline__ float4 f4_zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ void f4_fma(float4& a, float w, const float4& v) {
  a.x = fmaf(w, v.x, a.x);
  a.y = fmaf(w, v.y, a.y);
  a.z = fmaf(w, v.z, a.z);
  a.w = fmaf(w, v.w, a.w);
}
__device__ __forceinline__ float4 f4_add(const float4& a, const float4& b) {
  return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}
__device__ __forceinline__ float4 f4_shfl_xor(const float4& v, int off) {
  return make_float4(__shfl_xor_sync(FULL_MASK, v.x, off), __shfl_xor_sync(FULL_MASK, v.y, off),
                     __shfl_xor_sync(FULL_MASK, v.z, off), __shfl_xor_sync(FULL_MASK, v.w, off));
}

}  // namespace lgb
