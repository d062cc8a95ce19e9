// Fused Adam step for the embedding tables (reference run_pipeline_lightgcn.py:103,159: optim.Adam over (U+I) x d
// every iteration = ~12 ATen passes over the tables with the default foreach implementation).  One streaming kernel:
// 128-bit loads of p, g, m, v, one pass, 128-bit stores of p, m, v  (7 * n * 4 bytes, HBM-bound).
// Arithmetic follows torch.optim.Adam's single-tensor path operation by operation (lerp, addcmul, sqrt / sqrt(bc2) + eps,
// addcdiv with step_size = lr / bc1); the bias corrections AND the lerp / addcmul weights 1 - beta1, 1 - beta2 are computed
// on the host in double and rounded once, like torch does (1 - 0.999f in fp32 would be 1.3e-5 off 0.001).
#include "common.cuh"

namespace lgb {

__device__ __forceinline__ void adam1(float& p, float g, float& m, float& v, float omb1, float beta2, float omb2, float eps,
                                      float step_size, float bc2_sqrt) {
  m = m + omb1 * (g - m);                      // exp_avg.lerp_(grad, 1 - beta1)
  v = v * beta2 + omb2 * g * g;                // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
  const float denom = sqrtf(v) / bc2_sqrt + eps;
  p = p - step_size * (m / denom);             // param.addcdiv_(exp_avg, denom, value=-step_size)
}

__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                   float* __restrict__ v, int64_t n, float omb1, float beta2, float omb2, float eps,
                                                   float step_size, float bc2_sqrt, int vec) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  if (vec) {
    const int64_t n4 = n / 4;
    for (; i < n4; i += stride) {
      float4 P = reinterpret_cast<float4*>(p)[i], M = reinterpret_cast<float4*>(m)[i], V = reinterpret_cast<float4*>(v)[i];
      const float4 G = ld_stream_f4(reinterpret_cast<const float4*>(g) + i);
      adam1(P.x, G.x, M.x, V.x, omb1, beta2, omb2, eps, step_size, bc2_sqrt);
      adam1(P.y, G.y, M.y, V.y, omb1, beta2, omb2, eps, step_size, bc2_sqrt);
      adam1(P.z, G.z, M.z, V.z, omb1, beta2, omb2, eps, step_size, bc2_sqrt);
      adam1(P.w, G.w, M.w, V.w, omb1, beta2, omb2, eps, step_size, bc2_sqrt);
      reinterpret_cast<float4*>(p)[i] = P; reinterpret_cast<float4*>(m)[i] = M; reinterpret_cast<float4*>(v)[i] = V;
    }
  } else {
    for (; i < n; i += stride) adam1(p[i], g[i], m[i], v[i], omb1, beta2, omb2, eps, step_size, bc2_sqrt);
  }
}

}  // namespace lgb

using namespace lgb;

extern "C" int lgb_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                             float eps, int32_t step, void* stream) {
  LGB_REQUIRE(n >= 0 && step >= 1 && (n == 0 || (p && g && m && v)), LGB_EINVAL, "lgb_adam_step: bad argument");
  if (n == 0) return LGB_OK;
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  const float step_size = (float)((double)lr / bc1);
  const float bc2_sqrt = (float)sqrt(bc2);
  const float omb1 = (float)(1.0 - (double)beta1), omb2 = (float)(1.0 - (double)beta2);
  const int vec = (n % 4 == 0) && ((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) == 0);
  const int64_t work = vec ? n / 4 : n;
  const int64_t blocks = (work + 255) / 256;
  adam_kernel<<<(unsigned)(blocks < 148 * 16 ? blocks : 148 * 16), 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, omb1, beta2, omb2, eps,
                                                                                                 step_size, bc2_sqrt, vec);
  LGB_LAUNCH_CHECK();
  return LGB_OK;
}
