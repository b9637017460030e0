// Fused Adam / AdamW step for the flat parameter buffers: one pass over (param, grad, exp_avg,
// exp_avg_sq) = 28 B/parameter of HBM traffic, against the seven elementwise passes of torch.optim.Adam's
// single-/multi-tensor path that train_hash2.py:141-142,227-228 runs over the 16.8 M-entry hash table every step.
// Arithmetic follows torch.optim.adam._single_tensor_adam (non-amsgrad, maximize=False):
//   [AdamW: p *= 1 - lr*wd]   m += (g - m)*(1-b1)   v = v*b2 + (1-b2)*g*g
//   p += -(lr / (1-b1^t)) * ( m / ( sqrt(v)/sqrt(1-b2^t) + eps ) )
// with the GradScaler protocol folded in (train_hash2.py:226-239): g is multiplied by inv_scale first and the whole
// step is skipped when *found_inf != 0 (torch's fused-optimizer contract).
#include "common.cuh"

namespace hbr {

struct AdamArgs {           // every scalar is formed in double on the host (torch uses Python floats) and rounded once
  float lr, beta2, eps, weight_decay, one_minus_b1, one_minus_b2, neg_step_size, bc2_sqrt, decay, inv_scale;
  int decoupled;
};

__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, const AdamArgs& a) {
  g = g * a.inv_scale;
  if (a.weight_decay != 0.f) {
    if (a.decoupled) p = p * a.decay;                           // AdamW: param.mul_(1 - lr*wd)
    else g = g + a.weight_decay * p;                            // Adam: grad.add(param, alpha=wd)
  }
  m = m + (g - m) * a.one_minus_b1;                             // exp_avg.lerp_(grad, 1-beta1)
  v = v * a.beta2 + a.one_minus_b2 * g * g;                     // exp_avg_sq.mul_(b2).addcmul_(g, g, 1-b2)
  const float denom = sqrtf(v) / a.bc2_sqrt + a.eps;
  p = p + a.neg_step_size * (m / denom);                        // param.addcdiv_(exp_avg, denom, value=-step_size)
}

__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, long long n,
            const float* __restrict__ found_inf, const AdamArgs a) {
  if (found_inf != nullptr && __ldg(found_inf) != 0.f) return;
  const long long n4 = n >> 2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 pp = reinterpret_cast<float4*>(p)[i], mm = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
    const float4 gg = __ldg(reinterpret_cast<const float4*>(g) + i);
    adam_one(pp.x, gg.x, mm.x, vv.x, a);
    adam_one(pp.y, gg.y, mm.y, vv.y, a);
    adam_one(pp.z, gg.z, mm.z, vv.z, a);
    adam_one(pp.w, gg.w, mm.w, vv.w, a);
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
  }
  for (long long i = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    adam_one(p[i], __ldg(g + i), m[i], v[i], a);
}

// ---- the same step with every per-step scalar on the DEVICE: step count, learning rate, gradient scale, found_inf --------
// (torch's "capturable" contract): nothing about a step is baked into the launch, so the optimiser can live inside the
// captured CUDA graph of the training step, and GradScaler's skipped steps do not advance the bias correction.
__global__ void adam_tick_kernel(long long* __restrict__ step, const float* __restrict__ found_inf) {
  if (found_inf == nullptr || *found_inf == 0.f) *step += 1;
}

__global__ void __launch_bounds__(256)
adam_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, long long n,
                const float* __restrict__ found_inf, AdamArgs a, double beta1, double beta2, double wd,
                const float* __restrict__ lr_dev, const long long* __restrict__ step_dev, const float* __restrict__ grad_scale) {
  if (found_inf != nullptr && __ldg(found_inf) != 0.f) return;
  __shared__ float sh[4];
  if (threadIdx.x == 0) {
    const double lr = (double)__ldg(lr_dev), t = (double)*step_dev;     // the tick kernel ran before: t is this step's number
    sh[0] = (float)(-(lr / (1.0 - pow(beta1, t))));
    sh[1] = (float)sqrt(1.0 - pow(beta2, t));
    sh[2] = (float)(1.0 - lr * wd);
    sh[3] = grad_scale != nullptr ? a.inv_scale / __ldg(grad_scale) : a.inv_scale;
  }
  __syncthreads();
  a.neg_step_size = sh[0]; a.bc2_sqrt = sh[1]; a.decay = sh[2]; a.inv_scale = sh[3];
  const long long n4 = n >> 2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 pp = reinterpret_cast<float4*>(p)[i], mm = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
    const float4 gg = __ldg(reinterpret_cast<const float4*>(g) + i);
    adam_one(pp.x, gg.x, mm.x, vv.x, a);
    adam_one(pp.y, gg.y, mm.y, vv.y, a);
    adam_one(pp.z, gg.z, mm.z, vv.z, a);
    adam_one(pp.w, gg.w, mm.w, vv.w, a);
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
  }
  for (long long i = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    adam_one(p[i], __ldg(g + i), m[i], v[i], a);
}

}  // namespace hbr

using namespace hbr;

extern "C" int hbr_adam_tick(long long* step_dev, const float* found_inf, void* stream) {
  HBR_REQUIRE(step_dev != nullptr, "step_dev is NULL");
  adam_tick_kernel<<<1, 1, 0, as_stream(stream)>>>(step_dev, found_inf);
  HBR_LAUNCH_CHECK();
  return HBR_OK;
}

extern "C" int hbr_adam_step_dev(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, const float* lr_dev,
                                 double beta1, double beta2, double eps, double weight_decay, int decoupled_weight_decay,
                                 const long long* step_dev, double inv_scale, const float* grad_scale_dev, const float* found_inf,
                                 void* stream) {
  HBR_REQUIRE(n >= 0, "n=%lld", (long long)n);
  if (n == 0) return HBR_OK;
  HBR_REQUIRE(param && grad && exp_avg && exp_avg_sq && lr_dev && step_dev, "NULL pointer");
  HBR_REQUIRE(((uintptr_t)param | (uintptr_t)grad | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) % 16 == 0,
              "buffers must be 16-byte aligned");
  AdamArgs a{};
  a.beta2 = (float)beta2; a.eps = (float)eps; a.weight_decay = (float)weight_decay;
  a.decoupled = decoupled_weight_decay; a.inv_scale = (float)inv_scale;
  a.one_minus_b1 = (float)(1.0 - beta1);
  a.one_minus_b2 = (float)(1.0 - beta2);
  const int64_t want = ceil_div(ceil_div(n, 4), 256);
  const int grid = (int)min64(want, (int64_t)sm_count() * 16);
  adam_dev_kernel<<<grid, 256, 0, as_stream(stream)>>>(param, grad, exp_avg, exp_avg_sq, n, found_inf, a, beta1, beta2,
                                                       weight_decay, lr_dev, step_dev, grad_scale_dev);
  HBR_LAUNCH_CHECK();
  return HBR_OK;
}

extern "C" int hbr_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, double lr,
                             double beta1, double beta2, double eps, double weight_decay, int decoupled_weight_decay,
                             int64_t step, double inv_scale, const float* found_inf, void* stream) {
  HBR_REQUIRE(n >= 0 && step >= 1, "n=%lld step=%lld", (long long)n, (long long)step);
  if (n == 0) return HBR_OK;
  HBR_REQUIRE(param && grad && exp_avg && exp_avg_sq, "NULL pointer");
  HBR_REQUIRE(((uintptr_t)param | (uintptr_t)grad | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) % 16 == 0,
              "buffers must be 16-byte aligned");
  AdamArgs a;
  a.lr = (float)lr; a.beta2 = (float)beta2; a.eps = (float)eps; a.weight_decay = (float)weight_decay;
  a.decoupled = decoupled_weight_decay; a.inv_scale = (float)inv_scale;
  a.one_minus_b1 = (float)(1.0 - beta1);
  a.one_minus_b2 = (float)(1.0 - beta2);
  a.decay = (float)(1.0 - lr * weight_decay);
  // bias corrections in double like torch's Python-float arithmetic (1 - beta**step), rounded once
  a.neg_step_size = (float)(-(lr / (1.0 - pow(beta1, (double)step))));
  a.bc2_sqrt = (float)sqrt(1.0 - pow(beta2, (double)step));
  const int64_t want = ceil_div(ceil_div(n, 4), 256);
  const int grid = (int)min64(want, (int64_t)sm_count() * 16);
  adam_kernel<<<grid, 256, 0, as_stream(stream)>>>(param, grad, exp_avg, exp_avg_sq, n, found_inf, a);
  HBR_LAUNCH_CHECK();
  return HBR_OK;
}
