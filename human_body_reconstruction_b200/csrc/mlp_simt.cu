// MLP_3D forward/backward in fp32 on CUDA cores (test_hash.py:52-77).  This is the full-precision path:
// nerf2mesh.py runs the field without autocast, and the 1e-5 gradient-parity tests run through it.
// The training path under autocast is the tcgen05 implementation in mlp_tc.cu.
//
// forward : thread = point, weights staged once per CTA in shared memory (padded to float4 rows and
//           read as warp-broadcast LDS.128), the 64 layer inputs in registers, outputs through a
//           thread-private shared-memory column.
// backward: kernel A (thread = point) walks the layers in reverse producing every pre-activation
//           gradient dz (kept in a (rows, n) scratch), dfeat and ddirs; kernel B turns (dz, activations)
//           into weight/bias gradients with 4x4 register tiles over 32-point shared-memory tiles.
#include "mlp_layout.cuh"

namespace hbr {

constexpr int kMlpThreads = 128;

struct SmemW {
  int W[6], b[6], KP[6], JP[6], total;
};
__host__ __device__ inline int round4(int v) { return (v + 3) & ~3; }
__host__ __device__ inline SmemW make_smem_layout(const MlpLayout& m, int in0p, int kcp) {
  SmemW s;
  int o = 0;
  for (int i = 0; i < 6; ++i) {
    s.KP[i] = i == 0 ? in0p : (i == 3 ? kcp : round4(m.K[i]));
    s.JP[i] = round4(m.J[i]);
    s.W[i] = o; o += s.JP[i] * s.KP[i];
    s.b[i] = o; o += s.JP[i];
  }
  s.total = o;
  return s;
}

__device__ __forceinline__ void stage_weights(const float* __restrict__ params, const MlpLayout& m, const SmemW& s,
                                              float* ws, int first, int last) {
  for (int i = first; i <= last; ++i) {
    const int KP = s.KP[i], K = m.K[i], J = m.J[i];
    for (int e = threadIdx.x; e < s.JP[i] * KP; e += blockDim.x) {
      const int j = e / KP, k = e - j * KP;
      ws[s.W[i] + e] = (j < J && k < K) ? __ldg(params + m.W[i] + j * K + k) : 0.f;
    }
    for (int j = threadIdx.x; j < s.JP[i]; j += blockDim.x) ws[s.b[i] + j] = j < J ? __ldg(params + m.b[i] + j) : 0.f;
  }
}

// out[j] = act(b[j] + sum_k W[j][k] in[k]); result to the thread's shared-memory column.
template <int KP>
__device__ __forceinline__ void dense_fwd(const float* __restrict__ Ws, const float* __restrict__ bs, int JP,
                                          const float (&in)[KP], float* col, bool relu) {
#pragma unroll 1
  for (int j = 0; j < JP; j += 4) {
    float a[4] = {bs[j], bs[j + 1], bs[j + 2], bs[j + 3]};
#pragma unroll
    for (int k = 0; k < KP; k += 4) {
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const float4 w = *reinterpret_cast<const float4*>(Ws + (j + r) * KP + k);
        a[r] = fmaf(w.x, in[k], a[r]);
        a[r] = fmaf(w.y, in[k + 1], a[r]);
        a[r] = fmaf(w.z, in[k + 2], a[r]);
        a[r] = fmaf(w.w, in[k + 3], a[r]);
      }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) col[(j + r) * kMlpThreads] = relu ? fmaxf(a[r], 0.f) : a[r];
  }
}

// din[k] = sum_j dz[j] W[j][k]   (dz read from the thread's column)
template <int KP>
__device__ __forceinline__ void dense_bwd(const float* __restrict__ Ws, int J, const float* col, float (&din)[KP]) {
#pragma unroll
  for (int k = 0; k < KP; ++k) din[k] = 0.f;
#pragma unroll 2
  for (int j = 0; j < J; ++j) {
    const float d = col[j * kMlpThreads];
#pragma unroll
    for (int k = 0; k < KP; k += 4) {
      const float4 w = *reinterpret_cast<const float4*>(Ws + j * KP + k);
      din[k] = fmaf(d, w.x, din[k]);
      din[k + 1] = fmaf(d, w.y, din[k + 1]);
      din[k + 2] = fmaf(d, w.z, din[k + 2]);
      din[k + 3] = fmaf(d, w.w, din[k + 3]);
    }
  }
}

template <int IN0P, int KCP>
__global__ void __launch_bounds__(kMlpThreads)
mlp_fwd_kernel(const float* __restrict__ feat, long long feat_stride, const float* __restrict__ dirs,
               long long dir_group, long long n, const float* __restrict__ params, int in0, int dv,
               float* __restrict__ out, float* __restrict__ act) {
  extern __shared__ __align__(16) float smem[];
  const MlpLayout m = make_layout(in0, dv);
  const SmemW s = make_smem_layout(m, IN0P, KCP);
  float* ws = smem;
  float* col = smem + s.total + threadIdx.x;         // [64][kMlpThreads] column of this thread
  stage_weights(params, m, s, ws, 0, dirs ? 5 : 2);
  __syncthreads();

  for (long long gp = (long long)blockIdx.x * kMlpThreads + threadIdx.x; gp < n; gp += (long long)gridDim.x * kMlpThreads) {
    {
      float in[IN0P];
#pragma unroll
      for (int k = 0; k < IN0P; ++k) in[k] = k < in0 ? __ldg(feat + gp * feat_stride + k) : 0.f;
      dense_fwd<IN0P>(ws + s.W[0], ws + s.b[0], kH, in, col, true);
    }
    float h[kH];
#pragma unroll
    for (int k = 0; k < kH; ++k) h[k] = col[k * kMlpThreads];
    if (act)
#pragma unroll
      for (int k = 0; k < kH; ++k) act[(size_t)(kRowH1 + k) * n + gp] = h[k];
    dense_fwd<kH>(ws + s.W[1], ws + s.b[1], kH, h, col, true);
#pragma unroll
    for (int k = 0; k < kH; ++k) h[k] = col[k * kMlpThreads];
    if (act)
#pragma unroll
      for (int k = 0; k < kH; ++k) act[(size_t)(kRowH2 + k) * n + gp] = h[k];
    dense_fwd<kH>(ws + s.W[2], ws + s.b[2], kSigOut, h, col, false);
    const float raw = col[0];
    const float density = raw > 0.f ? raw : 0.01f * raw;            // LeakyReLU, test_hash.py:62
    if (act)
#pragma unroll
      for (int k = 0; k < kSigOut; ++k) act[(size_t)(kRowO16 + k) * n + gp] = col[k * kMlpThreads];
    if (!dirs) {
      out[gp] = density;
      continue;
    }
    {
      float cin[KCP];
      const float* dr = dirs + (gp / dir_group) * dv;
#pragma unroll
      for (int k = 0; k < KCP; ++k) {
        float v = 0.f;
        if (k < kFeat) v = col[(1 + k) * kMlpThreads];                // feat_vec = dens_vec[:,1:] (:64)
        else if (k < kFeat + dv) v = __ldg(dr + (k - kFeat));         // concat viewdirs (:66)
        cin[k] = v;
      }
      dense_fwd<KCP>(ws + s.W[3], ws + s.b[3], kH, cin, col, true);
    }
#pragma unroll
    for (int k = 0; k < kH; ++k) h[k] = col[k * kMlpThreads];
    if (act)
#pragma unroll
      for (int k = 0; k < kH; ++k) act[(size_t)(kRowC1 + k) * n + gp] = h[k];
    dense_fwd<kH>(ws + s.W[4], ws + s.b[4], kH, h, col, true);
#pragma unroll
    for (int k = 0; k < kH; ++k) h[k] = col[k * kMlpThreads];
    if (act)
#pragma unroll
      for (int k = 0; k < kH; ++k) act[(size_t)(kRowC2 + k) * n + gp] = h[k];
    dense_fwd<kH>(ws + s.W[5], ws + s.b[5], 4, h, col, false);
    float4 o;
    float pre[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) pre[c] = col[c * kMlpThreads];
    o.x = pre[0] > 0.f ? pre[0] : expm1f(pre[0]);                    // ELU, test_hash.py:67
    o.y = pre[1] > 0.f ? pre[1] : expm1f(pre[1]);
    o.z = pre[2] > 0.f ? pre[2] : expm1f(pre[2]);
    o.w = density;
    *reinterpret_cast<float4*>(out + gp * 4) = o;                     // (rgb, sigma), :69
    if (act)
#pragma unroll
      for (int c = 0; c < 3; ++c) act[(size_t)(kRowRgb + c) * n + gp] = pre[c];
  }
}

// ---- backward A: per-point chain ---------------------------------------------------------------------------
template <int IN0P, int KCP>
__global__ void __launch_bounds__(kMlpThreads)
mlp_bwd_point_kernel(long long n, const float* __restrict__ params, int in0, int dv, long long dir_group,
                     const float* __restrict__ dout, const float* __restrict__ act, float* __restrict__ dz,
                     float* __restrict__ dfeat, long long dfeat_stride, float* __restrict__ ddirs, int dens_only) {
  // dens_only: the backward of the density-only forward (dirs == NULL there): dout is (n) -- the gradient of the
  // LeakyReLU density --, the colour net is not walked and the 15 feature outputs of the density head carry no gradient
  // (SDF mode's eikonal stencil, test_hash.py:78-84: six such evaluations per sample).
  extern __shared__ __align__(16) float smem[];
  const MlpLayout m = make_layout(in0, dv);
  const SmemW s = make_smem_layout(m, IN0P, KCP);
  float* ws = smem;
  float* col = smem + s.total + threadIdx.x;
  stage_weights(params, m, s, ws, 0, dens_only ? 2 : 5);
  __syncthreads();

  for (long long gp = (long long)blockIdx.x * kMlpThreads + threadIdx.x; gp < n; gp += (long long)gridDim.x * kMlpThreads) {
    float d[kH];
    float dsig16[kSigOut];
    float go_w;
    if (dens_only) {
      go_w = __ldg(dout + gp);
#pragma unroll
      for (int k = 1; k < kSigOut; ++k) dsig16[k] = 0.f;
    } else {
    const float4 go = *reinterpret_cast<const float4*>(dout + gp * 4);
    go_w = go.w;
    // colour head: ELU'
    {
      const float g[3] = {go.x, go.y, go.z};
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float pre = act[(size_t)(kRowRgb + c) * n + gp];
        const float dpre = g[c] * (pre > 0.f ? 1.f : expf(pre));
        col[c * kMlpThreads] = dpre;
        dz[(size_t)(kRowRgb + c) * n + gp] = dpre;
      }
    }
    dense_bwd<kH>(ws + s.W[5], 3, col, d);
#pragma unroll
    for (int k = 0; k < kH; ++k) {
      const float v = act[(size_t)(kRowC2 + k) * n + gp] > 0.f ? d[k] : 0.f;
      col[k * kMlpThreads] = v;
      dz[(size_t)(kRowC2 + k) * n + gp] = v;
    }
    dense_bwd<kH>(ws + s.W[4], kH, col, d);
#pragma unroll
    for (int k = 0; k < kH; ++k) {
      const float v = act[(size_t)(kRowC1 + k) * n + gp] > 0.f ? d[k] : 0.f;
      col[k * kMlpThreads] = v;
      dz[(size_t)(kRowC1 + k) * n + gp] = v;
    }
    {
      float dc[KCP];
      dense_bwd<KCP>(ws + s.W[3], kH, col, dc);
#pragma unroll
      for (int k = 0; k < kFeat; ++k) dsig16[1 + k] = dc[k];
      if (ddirs) {
        float* dd = ddirs + (gp / dir_group) * dv;
#pragma unroll
        for (int k = kFeat; k < KCP; ++k)
          if (k < kFeat + dv) atomicAdd(dd + (k - kFeat), dc[k]);
      }
    }
    }   // !dens_only
    dsig16[0] = go_w * (act[(size_t)kRowO16 * n + gp] > 0.f ? 1.f : 0.01f);   // LeakyReLU'
#pragma unroll
    for (int k = 0; k < kSigOut; ++k) {
      col[k * kMlpThreads] = dsig16[k];
      dz[(size_t)(kRowO16 + k) * n + gp] = dsig16[k];
    }
    dense_bwd<kH>(ws + s.W[2], kSigOut, col, d);
#pragma unroll
    for (int k = 0; k < kH; ++k) {
      const float v = act[(size_t)(kRowH2 + k) * n + gp] > 0.f ? d[k] : 0.f;
      col[k * kMlpThreads] = v;
      dz[(size_t)(kRowH2 + k) * n + gp] = v;
    }
    dense_bwd<kH>(ws + s.W[1], kH, col, d);
#pragma unroll
    for (int k = 0; k < kH; ++k) {
      const float v = act[(size_t)(kRowH1 + k) * n + gp] > 0.f ? d[k] : 0.f;
      col[k * kMlpThreads] = v;
      dz[(size_t)(kRowH1 + k) * n + gp] = v;
    }
    if (dfeat) {
      float df[IN0P];
      dense_bwd<IN0P>(ws + s.W[0], kH, col, df);
#pragma unroll
      for (int k = 0; k < IN0P; ++k)
        if (k < in0) dfeat[gp * dfeat_stride + k] = df[k];
    }
  }
}

// ---- backward B: weight / bias gradients --------------------------------------------------------------------
constexpr int kWgPts = 32;       // points per shared-memory tile
constexpr int kWgPitch = 68;     // floats; keeps float4 alignment

__global__ void __launch_bounds__(256)
mlp_wgrad_kernel(const float* __restrict__ feat, long long feat_stride, const float* __restrict__ dirs,
                 long long dir_group, long long n, int in0, int dv, const float* __restrict__ act,
                 const float* __restrict__ dz, float* __restrict__ dparams, long long chunk) {
  __shared__ __align__(16) float dzs[kWgPts * kWgPitch];
  __shared__ __align__(16) float as[kWgPts * kWgPitch];
  const MlpLayout m = make_layout(in0, dv);
  const int layer = blockIdx.y;
  const int J = m.J[layer], K = m.K[layer];
  const int dz_row = layer == 0 ? kRowH1 : layer == 1 ? kRowH2 : layer == 2 ? kRowO16 : layer == 3 ? kRowC1
                     : layer == 4 ? kRowC2 : kRowRgb;
  const int a_row = layer == 1 ? kRowH1 : layer == 2 ? kRowH2 : layer == 4 ? kRowC1 : kRowC2;   // layers 0,3 special
  const int tj = threadIdx.x >> 4, tk = threadIdx.x & 15;
  float acc[4][4] = {};
  float bacc[4] = {};
  const long long p_begin = (long long)blockIdx.x * chunk;
  const long long p_end = min(n, p_begin + chunk);
  for (long long p0 = p_begin; p0 < p_end; p0 += kWgPts) {
    // tile loads: element e -> (row c, point p); consecutive threads take consecutive points
    for (int e = threadIdx.x; e < 64 * kWgPts; e += 256) {
      const int c = e / kWgPts, p = e - c * kWgPts;
      const long long gp = p0 + p;
      const bool ok = gp < p_end;
      dzs[p * kWgPitch + c] = (ok && c < J) ? dz[(size_t)(dz_row + c) * n + gp] : 0.f;
      float a = 0.f;
      if (ok && c < K) {
        if (layer == 0) a = __ldg(feat + gp * feat_stride + c);
        else if (layer == 3) a = c < kFeat ? act[(size_t)(kRowO16 + 1 + c) * n + gp]
                                           : __ldg(dirs + (gp / dir_group) * dv + (c - kFeat));
        else a = act[(size_t)(a_row + c) * n + gp];
      }
      as[p * kWgPitch + c] = a;
    }
    __syncthreads();
    if (tj * 4 < J && tk * 4 < K) {
#pragma unroll 8
      for (int p = 0; p < kWgPts; ++p) {
        const float4 d4 = *reinterpret_cast<const float4*>(dzs + p * kWgPitch + tj * 4);
        const float4 a4 = *reinterpret_cast<const float4*>(as + p * kWgPitch + tk * 4);
        const float dd[4] = {d4.x, d4.y, d4.z, d4.w};
        const float aa[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
          for (int q = 0; q < 4; ++q) acc[r][q] = fmaf(dd[r], aa[q], acc[r][q]);
          if (tk == 0) bacc[r] += dd[r];
        }
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int j = tj * 4 + r;
    if (j >= J) continue;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int k = tk * 4 + q;
      if (k < K) atomicAdd(dparams + m.W[layer] + j * K + k, acc[r][q]);
    }
    if (tk == 0) atomicAdd(dparams + m.b[layer] + j, bacc[r]);
  }
}

template <typename Kern>
static int set_smem(Kern k, size_t bytes) {
  HBR_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return HBR_OK;
}

static size_t simt_smem_bytes(const hbr_mlp_dims* d, int in0p, int kcp) {
  const MlpLayout m = make_layout(d->in0, d->d_view);
  const SmemW s = make_smem_layout(m, in0p, kcp);
  return (size_t)(s.total + kH * kMlpThreads) * sizeof(float);
}

}  // namespace hbr

using namespace hbr;

extern "C" int64_t hbr_mlp_param_count(const hbr_mlp_dims* d) {
  if (check_dims(d)) return -1;
  return make_layout(d->in0, d->d_view).total;
}
extern "C" int64_t hbr_mlp_act_floats(void) { return kActRows; }

#define HBR_MLP_DISPATCH(KERN, ...)                                               \
  do {                                                                            \
    if (in0p == 32 && kcp == 40) { HBR_MLP_GO((KERN<32, 40>), __VA_ARGS__); }     \
    else if (in0p == 32) { HBR_MLP_GO((KERN<32, 64>), __VA_ARGS__); }             \
    else if (kcp == 40) { HBR_MLP_GO((KERN<64, 40>), __VA_ARGS__); }              \
    else { HBR_MLP_GO((KERN<64, 64>), __VA_ARGS__); }                             \
  } while (0)

extern "C" int hbr_mlp_fwd_f32(const float* feat, int64_t feat_stride, const float* dirs, int64_t dir_group, int64_t n,
                               const float* params, const hbr_mlp_dims* dims, float* out, float* act, void* stream) {
  if (int rc = check_dims(dims)) return rc;
  if (n == 0) return HBR_OK;
  HBR_REQUIRE(feat && params && out, "NULL pointer");
  HBR_REQUIRE(feat_stride >= dims->in0, "feat_stride %lld < in0", (long long)feat_stride);
  HBR_REQUIRE(!dirs || dir_group >= 1, "dir_group must be >= 1");
  HBR_REQUIRE(!dirs || (uintptr_t)out % 16 == 0, "out must be 16-byte aligned");
  const int in0p = dims->in0 <= 32 ? 32 : 64, kcp = dims->d_view + kFeat <= 40 ? 40 : 64;
  const size_t smem = simt_smem_bytes(dims, in0p, kcp);
  const int grid = (int)min64(ceil_div(n, kMlpThreads), (int64_t)sm_count() * 16);
  cudaStream_t st = as_stream(stream);
#define HBR_MLP_GO(K, ...)                               \
  if (int rc = set_smem(K, smem)) return rc;             \
  K<<<grid, kMlpThreads, smem, st>>>(__VA_ARGS__)
  HBR_MLP_DISPATCH(mlp_fwd_kernel, feat, feat_stride, dirs, dir_group, n, params, dims->in0, dims->d_view, out, act);
#undef HBR_MLP_GO
  HBR_LAUNCH_CHECK();
  return HBR_OK;
}

extern "C" int hbr_mlp_bwd_f32(const float* feat, int64_t feat_stride, const float* dirs, int64_t dir_group, int64_t n,
                               const float* params, const hbr_mlp_dims* dims, const float* dout, const float* act,
                               float* dz, float* dfeat, int64_t dfeat_stride, float* ddirs, float* dparams, void* stream) {
  if (int rc = check_dims(dims)) return rc;
  if (n == 0) return HBR_OK;
  HBR_REQUIRE(feat && params && dout && act && dz, "NULL pointer");
  const int dens_only = dirs == nullptr;             // backward of the density-only forward: dout is (n), no colour net
  HBR_REQUIRE(dens_only || (uintptr_t)dout % 16 == 0, "dout must be 16-byte aligned");
  HBR_REQUIRE(!dens_only || !ddirs, "ddirs requested without dirs");
  HBR_REQUIRE(!dfeat || dfeat_stride >= dims->in0, "dfeat_stride too small");
  const int in0p = dims->in0 <= 32 ? 32 : 64, kcp = dims->d_view + kFeat <= 40 ? 40 : 64;
  const size_t smem = simt_smem_bytes(dims, in0p, kcp);
  const int grid = (int)min64(ceil_div(n, kMlpThreads), (int64_t)sm_count() * 16);
  cudaStream_t st = as_stream(stream);
#define HBR_MLP_GO(K, ...)                               \
  if (int rc = set_smem(K, smem)) return rc;             \
  K<<<grid, kMlpThreads, smem, st>>>(__VA_ARGS__)
  HBR_MLP_DISPATCH(mlp_bwd_point_kernel, n, params, dims->in0, dims->d_view, dir_group, dout, act, dz, dfeat,
                   dfeat_stride, ddirs, dens_only);
#undef HBR_MLP_GO
  HBR_LAUNCH_CHECK();
  if (dparams) {
    const long long chunk = 2048;
    const dim3 g((unsigned)ceil_div(n, chunk), dens_only ? 3 : 6);   // density only: the three layers of the density head
    mlp_wgrad_kernel<<<g, 256, 0, st>>>(feat, feat_stride, dirs, dir_group, n, dims->in0, dims->d_view, act, dz,
                                        dparams, chunk);
    HBR_LAUNCH_CHECK();
  }
  return HBR_OK;
}
