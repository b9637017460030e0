// Small kernels of the training step that sit between the big ones (SURVEY 8a rows a8 and the loss of train_hash2.py:221):
// each replaces a handful of ATen elementwise / reduce launches that cost more in launch latency than in work at 4096 rays.
#include "common.cuh"

namespace hbr {

// strat_sampler (helper.py:231-232): t = lin + (u * span) / count, every operation rounded on its own exactly like the
// three ATen kernels it replaces (mul by a wrapped CPU scalar, div by a Python int, add); lin = torch.linspace(tn, tf, S)
// and u = torch.rand_like(lin) stay torch's (same RNG stream, same linspace arithmetic).
__global__ void strat_depths_kernel(const float* __restrict__ lin, const float* __restrict__ u, float span, float count,
                                    long long S, float* __restrict__ t) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < S) t[i] = __fadd_rn(__ldg(lin + i), __fdiv_rn(__fmul_rn(__ldg(u + i), span), count));
}

// loss = mean((a - gt)^2) [+ mean((b - gt)^2)]  (nn.MSELoss, train_hash2.py:177,221); partial sums per CTA in double,
// one atomicAdd per CTA into the zero-initialised result.
__global__ void __launch_bounds__(256)
mse_pair_fwd_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ gt, long long n,
                    double inv_n, float* __restrict__ loss) {
  // a thread sums at most a few dozen squares in fp32 (few-ulp error), everything above that in double
  float facc = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float g = __ldg(gt + i);
    const float da = __ldg(a + i) - g;
    facc = fmaf(da, da, facc);
    if (b != nullptr) {
      const float db = __ldg(b + i) - g;
      facc = fmaf(db, db, facc);
    }
  }
  double acc = (double)facc;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(kFull, acc, o);
  __shared__ double part[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) part[warp] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < 8; ++w) s += part[w];
    atomicAdd(loss, (float)(s * inv_n));
  }
}
// da = gout * 2 (a - gt) / n, db likewise (gout: the scalar upstream gradient, on the device)
__global__ void mse_pair_bwd_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ gt,
                                    long long n, float two_over_n, const float* __restrict__ gout, float* __restrict__ da,
                                    float* __restrict__ db) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float s = __ldg(gout) * two_over_n, g = __ldg(gt + i);
  da[i] = s * (__ldg(a + i) - g);
  if (b != nullptr) db[i] = s * (__ldg(b + i) - g);
}

}  // namespace hbr

using namespace hbr;

extern "C" int hbr_strat_depths(const float* lin, const float* u, float span, float count, int64_t S, float* t, void* stream) {
  HBR_REQUIRE(S >= 0 && count != 0.f, "S=%lld count=%g", (long long)S, (double)count);
  if (S == 0) return HBR_OK;
  HBR_REQUIRE(lin && u && t, "NULL pointer");
  strat_depths_kernel<<<(unsigned)ceil_div(S, 256), 256, 0, as_stream(stream)>>>(lin, u, span, count, S, t);
  HBR_LAUNCH_CHECK();
  return HBR_OK;
}

extern "C" int hbr_mse_pair_fwd(const float* a, const float* b, const float* gt, int64_t n, float scale, float* loss,
                                void* stream) {
  HBR_REQUIRE(n >= 1 && a && gt && loss, "bad argument");
  const int grid = (int)min64(ceil_div(n, 256 * 4), (long long)sm_count() * 4);
  mse_pair_fwd_kernel<<<grid, 256, 0, as_stream(stream)>>>(a, b, gt, n, (double)scale / (double)n, loss);
  HBR_LAUNCH_CHECK();
  return HBR_OK;
}

extern "C" int hbr_mse_pair_bwd(const float* a, const float* b, const float* gt, int64_t n, float scale, const float* gout,
                                float* da, float* db, void* stream) {
  HBR_REQUIRE(n >= 1 && a && gt && gout && da && (b == nullptr || db != nullptr), "bad argument");
  mse_pair_bwd_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, as_stream(stream)>>>(a, b, gt, n, 2.f * scale / (float)n, gout, da, db);
  HBR_LAUNCH_CHECK();
  return HBR_OK;
}
