// TEST INFRASTRUCTURE ONLY -- a warp-lockstep CPU emulation of the CUDA subset liblaplace_b200's kernels use.
//
// tests/emu/build_emu.py compiles the UNMODIFIED kernel sources of laplace_gnn_recommendation_b200/csrc/ against
// this header (instead of the CUDA toolkit's cuda_runtime.h) with g++, so `pytest -m "not gpu"` can execute the
// real kernel logic -- index arithmetic, warp shuffles, split plans, epilogues -- on a machine without a GPU.
// Every CUDA thread is a fiber; warp collectives (__shfl*_sync, __ballot_sync, __all_sync) and __syncthreads block a
// fiber until every live lane of its warp / thread of its block has arrived, and the engine reports divergent
// collectives, collectives entered after part of the warp exited, and deadlocks as launch errors.
// Nothing in the product package knows about this file: the shipped library is built by nvcc for sm_100a only and
// has no CPU path (see tests/test_abi.py::test_no_cpu_fallback).
#pragma once
#define LGB_CPU_EMU 1

#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <type_traits>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__ __attribute__((noinline))
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))

// ---- vector types ---------------------------------------------------------------------------------------
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(16) int4 { int x, y, z, w; };
struct alignas(8) float2 { float x, y; };
struct uint3 { unsigned x, y, z; };
struct dim3 {
  unsigned x, y, z;
  dim3() : x(1), y(1), z(1) {}
  template <class T, class = typename std::enable_if<std::is_integral<T>::value>::type>
  dim3(T x_) : x((unsigned)x_), y(1), z(1) {}
  dim3(unsigned x_, unsigned y_, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
static inline float4 make_float4(float x, float y, float z, float w) { float4 v; v.x = x; v.y = y; v.z = z; v.w = w; return v; }
static inline int4 make_int4(int x, int y, int z, int w) { int4 v; v.x = x; v.y = y; v.z = z; v.w = w; return v; }
static inline float2 make_float2(float x, float y) { float2 v; v.x = x; v.y = y; return v; }

// ---- runtime API subset ---------------------------------------------------------------------------------
typedef int cudaError_t;
typedef void* cudaStream_t;
enum { cudaSuccess = 0, cudaErrorLaunchFailure = 719, cudaErrorInvalidValue = 1 };
enum { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3, cudaMemcpyDefault = 4 };
enum { cudaDevAttrMultiProcessorCount = 16 };
enum { cudaFuncAttributePreferredSharedMemoryCarveout = 9, cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };

namespace emu {
// engine (tests/emu/engine.cpp)
typedef void (*body_fn)(void*);
void launch(dim3 grid, dim3 block, body_fn fn, void* arg);
uint64_t warp_collect(int op, unsigned mask, uint64_t v, uint64_t* all /*[32]*/, unsigned* arrived_mask);
void block_barrier();
cudaError_t take_error();            // returns and clears the sticky launch error
const char* error_string();
void note_inactive_read(const char* what);
float4 mc_ld_reduce(const float4* mc);   // emulated NVSwitch multicast: see emu_multicast_bind() in engine.cpp
void mc_st(float4* mc, const float4& v);
extern uint3 g_threadIdx, g_blockIdx;
extern dim3 g_blockDim, g_gridDim;

template <class F>
static void call_body(void* p) { (*static_cast<F*>(p))(); }
template <class F>
static inline void launch(dim3 grid, dim3 block, size_t /*smem*/, cudaStream_t /*stream*/, F&& f) {
  typedef typename std::remove_reference<F>::type Fn;
  launch(grid, block, &call_body<Fn>, (void*)&f);
}
enum { OP_SHFL = 1, OP_SHFL_XOR, OP_SHFL_DOWN, OP_SHFL_UP, OP_BALLOT, OP_ALL, OP_ANY, OP_SYNCWARP };

template <class T>
static inline uint64_t to_bits(T v) {
  static_assert(sizeof(T) <= 8, "shuffle operand wider than 64 bits");
  uint64_t b = 0;
  memcpy(&b, &v, sizeof(T));
  return b;
}
template <class T>
static inline T from_bits(uint64_t b) {
  T v;
  memcpy(&v, &b, sizeof(T));
  return v;
}
}  // namespace emu

#define threadIdx (emu::g_threadIdx)
#define blockIdx (emu::g_blockIdx)
#define blockDim (emu::g_blockDim)
#define gridDim (emu::g_gridDim)
#define warpSize 32

static inline cudaError_t cudaGetLastError() { return emu::take_error(); }
static inline cudaError_t cudaPeekAtLastError() { return emu::take_error(); }
static inline const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : emu::error_string(); }
static inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { memset(p, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, int, cudaStream_t) { memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline long long clock64() { return 0; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
static inline cudaError_t cudaDeviceGetAttribute(int* v, int attr, int) {
  *v = (attr == cudaDevAttrMultiProcessorCount) ? 148 : 0;
  return cudaSuccess;
}
template <class F>
static inline cudaError_t cudaFuncSetAttribute(F, int, int) { return cudaSuccess; }
template <class F>
static inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int* occ, F, int, size_t) { *occ = 2; return cudaSuccess; }

// ---- device intrinsics ----------------------------------------------------------------------------------
template <class A, class B>
static inline typename std::common_type<A, B>::type min(A a, B b) { return b < a ? b : a; }
template <class A, class B>
static inline typename std::common_type<A, B>::type max(A a, B b) { return a < b ? b : a; }

static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsqrt_rn(float a) { return sqrtf(a); }
static inline float __frcp_rn(float a) { return 1.0f / a; }
static inline unsigned __float_as_uint(float f) { return emu::to_bits(f) & 0xffffffffu; }
static inline int __float_as_int(float f) { return (int)__float_as_uint(f); }
static inline float __uint_as_float(unsigned u) { return emu::from_bits<float>(u); }
static inline float __int_as_float(int i) { return emu::from_bits<float>((unsigned)i); }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
static inline int __clz(int v) { return v == 0 ? 32 : __builtin_clz((unsigned)v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
template <class T>
static inline T __ldg(const T* p) { return *p; }
static inline size_t __cvta_generic_to_shared(const void* p) { return (size_t)p; }
static inline void __threadfence() {}
static inline void __threadfence_block() {}
static inline void __syncthreads() { emu::block_barrier(); }

template <class T, class U>
static inline T atomicAdd(T* p, U v) { T old = *p; *p = (T)(old + (T)v); return old; }
template <class T, class U>
static inline T atomicMax(T* p, U v) { T old = *p; if ((T)v > old) *p = (T)v; return old; }
template <class T, class U>
static inline T atomicMin(T* p, U v) { T old = *p; if ((T)v < old) *p = (T)v; return old; }
template <class T, class U>
static inline T atomicExch(T* p, U v) { T old = *p; *p = (T)v; return old; }
template <typename T, typename U>
static inline T atomicOr(T* p, U v) { T old = *p; *p = (T)(old | (T)v); return old; }
template <class T, class U, class V>
static inline T atomicCAS(T* p, U cmp, V v) { T old = *p; if (old == (T)cmp) *p = (T)v; return old; }

namespace emu {
template <class T>
static inline T shfl_common(int op, unsigned mask, T v, int src_lane_abs, const char* what) {
  uint64_t all[32];
  unsigned arrived = 0;
  warp_collect(op, mask, to_bits(v), all, &arrived);
  if (!((arrived >> src_lane_abs) & 1u)) {   // CUDA: reading an inactive / exited lane is undefined
    note_inactive_read(what);
    return v;
  }
  return from_bits<T>(all[src_lane_abs]);
}
static inline int lane_id() { return (int)(g_threadIdx.x + g_blockDim.x * (g_threadIdx.y + g_blockDim.y * g_threadIdx.z)) & 31; }
}  // namespace emu

template <class T>
static inline T __shfl_sync(unsigned mask, T v, int src, int width = 32) {
  const int lane = emu::lane_id();
  const int base = lane & ~(width - 1);
  return emu::shfl_common(emu::OP_SHFL, mask, v, base + (src & (width - 1)), "__shfl_sync");
}
template <class T>
static inline T __shfl_xor_sync(unsigned mask, T v, int lane_mask, int width = 32) {
  const int lane = emu::lane_id();
  const int src = lane ^ lane_mask;
  // CUDA: a source outside the caller's width-segment returns the caller's own value
  const bool same_seg = (src & ~(width - 1)) == (lane & ~(width - 1));
  return emu::shfl_common(emu::OP_SHFL_XOR, mask, v, same_seg ? (src & 31) : lane, "__shfl_xor_sync");
}
template <class T>
static inline T __shfl_down_sync(unsigned mask, T v, unsigned delta, int width = 32) {
  const int lane = emu::lane_id();
  const int src = lane + (int)delta;
  const bool ok = (src & ~(width - 1)) == (lane & ~(width - 1)) && src < 32;
  return emu::shfl_common(emu::OP_SHFL_DOWN, mask, v, ok ? src : lane, "__shfl_down_sync");
}
template <class T>
static inline T __shfl_up_sync(unsigned mask, T v, unsigned delta, int width = 32) {
  const int lane = emu::lane_id();
  const int src = lane - (int)delta;
  const bool ok = src >= 0 && (src & ~(width - 1)) == (lane & ~(width - 1));
  return emu::shfl_common(emu::OP_SHFL_UP, mask, v, ok ? src : lane, "__shfl_up_sync");
}
static inline unsigned __ballot_sync(unsigned mask, int pred) {
  uint64_t all[32];
  unsigned arrived = 0;
  emu::warp_collect(emu::OP_BALLOT, mask, pred ? 1 : 0, all, &arrived);
  unsigned r = 0;
  for (int i = 0; i < 32; ++i)
    if (((arrived >> i) & 1u) && all[i]) r |= 1u << i;
  return r & mask;
}
static inline int __all_sync(unsigned mask, int pred) {
  uint64_t all[32];
  unsigned arrived = 0;
  emu::warp_collect(emu::OP_ALL, mask, pred ? 1 : 0, all, &arrived);
  for (int i = 0; i < 32; ++i)
    if (((arrived >> i) & 1u) && !all[i]) return 0;
  return 1;
}
static inline int __any_sync(unsigned mask, int pred) {
  uint64_t all[32];
  unsigned arrived = 0;
  emu::warp_collect(emu::OP_ANY, mask, pred ? 1 : 0, all, &arrived);
  for (int i = 0; i < 32; ++i)
    if (((arrived >> i) & 1u) && all[i]) return 1;
  return 0;
}
static inline void __syncwarp(unsigned mask = 0xffffffffu) {
  uint64_t all[32];
  unsigned arrived = 0;
  emu::warp_collect(emu::OP_SYNCWARP, mask, 0, all, &arrived);
}
