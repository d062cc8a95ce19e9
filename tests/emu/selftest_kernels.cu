// TEST INFRASTRUCTURE ONLY -- kernels that check the emulator itself (tests/emu/selftest.py): correct shuffle
// semantics on well-formed code, and an ERROR for the three things hardware leaves undefined.
#include <cuda_runtime.h>

__global__ void ok_kernel(int* out) {
  const int lane = threadIdx.x & 31;
  int v = lane;
  for (int off = 16; off; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);      // 496 in every lane
  const int seg = __shfl_sync(0xffffffffu, lane, 3, 8);                               // lane 3 of my 8-lane segment
  const unsigned bal = __ballot_sync(0xffffffffu, lane % 3 == 0);
  __shared__ int sm[64];
  sm[threadIdx.x] = v + seg;
  __syncthreads();
  const int other = sm[(threadIdx.x + 32) % 64];
  out[blockIdx.x * 64 + threadIdx.x] = other + __popc(bal) + (__all_sync(0xffffffffu, v == 496) ? 1000 : 0);
}

__global__ void divergent_kernel(int* out) {          // half the warp shuffles, the other half votes
  const int lane = threadIdx.x & 31;
  int v;
  if (lane < 16) v = __shfl_sync(0xffffffffu, lane, 0);
  else v = __all_sync(0xffffffffu, 1);
  out[threadIdx.x] = v;
}

__global__ void exited_lane_kernel(int* out) {        // lanes >= 20 leave, the rest shuffle under a full mask
  const int lane = threadIdx.x & 31;
  if (lane >= 20) return;
  out[threadIdx.x] = __shfl_xor_sync(0xffffffffu, lane, 1);
}

__global__ void deadlock_kernel(int* out) {           // half of each warp waits in a shuffle, the other half in a barrier
  const int lane = threadIdx.x & 31;
  int v = 0;
  if (lane < 16) v = __shfl_xor_sync(0xffffffffu, lane, 1);
  __syncthreads();
  out[threadIdx.x] = v;
}

extern "C" int selftest_run(int which, int* out) {
  (void)cudaGetLastError();
  switch (which) {
    case 0: ok_kernel<<<3, 64>>>(out); break;
    case 1: divergent_kernel<<<1, 32>>>(out); break;
    case 2: exited_lane_kernel<<<1, 32>>>(out); break;
    case 3: deadlock_kernel<<<1, 64>>>(out); break;
  }
  return (int)cudaGetLastError();
}
extern "C" const char* selftest_error() { return cudaGetErrorString(1); }
