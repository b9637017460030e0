// SDF mode (--use_sdf): the SDF branch of calc_color (helper.py:76-86,102-105) forward and backward, and the
// finite-difference eikonal term (test_hash.py:86-105 + helper.py:293-297) as a 6-point stencil around one encoder +
// density-head pass.  Replaces ~40 elementwise torch launches and 6 separate encoder/MLP passes per step.
//
// Compositing, one warp per ray (lanes take samples s = 32*i + lane, like composite.cu):
//   s_i   = sdf value (optionally formed here from the density head's LeakyReLU output d:  raw = d > 0 ? d : 100 d,
//           s = 2 sigmoid(raw) - 1, test_hash.py:59-60), clamped at -10 (helper.py:76)
//   phi_i = 1 / (1 + exp(-s_i b))                         VarModel, helper.py:18-21
//   a_i   = relu(1 - phi_{i+1} / phi_i), a_{S-1} = 0      helper.py:81-83
//   T_i   = prod_{j<i} (1 - a_j)                          cumprod_exclusive, helper.py:84
//   w_i   = T_i a_i,  C = sum_i w_i rgb_i                 helper.py:102-105
// Backward in closed form (c_i = gC . rgb_i + gw_i, suffix_i = sum_{k>i} w_k c_k):
//   e_i   = a_i > 0 ? suffix_i - T_i c_i (1 - a_i) : 0    (= dL/dr_i * r_i with r_i = phi_{i+1}/phi_i)
//   dL/d(s_i b) = (1 - phi_i)(e_{i-1} - e_i),  e_{-1} = e_{S-1} = 0
//   dL/ds_i = b * that (0 where the clamp hit), dL/db = sum_i s_i * that, d rgb_i = w_i gC.
#include "common.cuh"

namespace hbr {

constexpr int kSdfRaysPerCta = 8;

// sdf from the density head's LeakyReLU(0.01) output; gain = d sdf / d (that output)
__device__ __forceinline__ float sdf_from_density(float d, float& gain) {
  const bool pos = d > 0.f;
  const float raw = pos ? d : __fmul_rn(d, 100.f);
  const float sg = __fdiv_rn(1.f, __fadd_rn(1.f, expf(-raw)));
  gain = 2.f * sg * (1.f - sg) * (pos ? 1.f : 100.f);
  return __fsub_rn(__fmul_rn(2.f, sg), 1.f);
}

struct SdfIn {
  float s;       // clamped sdf
  float gain;    // d s / d (input value); 0 where the clamp hit
};

__device__ __forceinline__ SdfIn load_sdf(const float* __restrict__ sd, long long st, long long gs, int from_density) {
  SdfIn q;
  float v = __ldg(sd + gs * st);
  q.gain = 1.f;
  if (from_density) v = sdf_from_density(v, q.gain);
  if (v < -10.f) { v = -10.f; q.gain = 0.f; }                            // helper.py:76
  q.s = v;
  return q;
}

__device__ __forceinline__ float var_phi(float s, float b) {              // helper.py:18-21
  return __fdiv_rn(1.f, __fadd_rn(1.f, expf(-__fmul_rn(s, b))));
}

__device__ __forceinline__ float warp_incl_scan_mul(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float t = __shfl_up_sync(kFull, v, o);
    if (lane >= o) v *= t;
  }
  return v;
}

// One 32-sample chunk of a ray: phi of this lane's sample, alpha and the exclusive transmittance (carry = product of
// (1 - alpha) over all earlier chunks, updated).  Samples at or behind S get alpha = 0.
__device__ __forceinline__ void sdf_chunk(const float* __restrict__ sd, long long st, long long ray, int S, int s, int lane,
                                          int from_density, float b, float& carry, SdfIn& q, float& phi, float& alpha,
                                          float& T) {
  q.s = 0.f; q.gain = 0.f;
  phi = 1.f;
  if (s < S) {
    q = load_sdf(sd, st, ray * S + s, from_density);
    phi = var_phi(q.s, b);
  }
  float pn = __shfl_down_sync(kFull, phi, 1);
  if (lane == 31) pn = s + 1 < S ? var_phi(load_sdf(sd, st, ray * S + s + 1, from_density).s, b) : 1.f;
  alpha = 0.f;
  if (s + 1 < S) alpha = fmaxf(__fsub_rn(1.f, __fdiv_rn(pn, phi)), 0.f);  // helper.py:82-83; the last sample keeps 0
  const float incl = carry * warp_incl_scan_mul(1.f - alpha, lane);
  float prev = __shfl_up_sync(kFull, incl, 1);
  if (lane == 0) prev = carry;
  T = prev;                                                               // exclusive product (helper.py:84)
  carry = __shfl_sync(kFull, incl, 31);
}

__global__ void __launch_bounds__(kSdfRaysPerCta * 32)
composite_sdf_fwd_kernel(const float* __restrict__ rgb, long long rgb_st, const float* __restrict__ sd, long long sd_st,
                         int from_density, const float* __restrict__ b_ptr, long long R, int S, float* __restrict__ C,
                         float* __restrict__ w_out) {
  const int lane = threadIdx.x & 31;
  const long long ray = (long long)blockIdx.x * kSdfRaysPerCta + (threadIdx.x >> 5);
  if (ray >= R) return;
  const float b = __ldg(b_ptr);
  float carry = 1.f, c0 = 0.f, c1 = 0.f, c2 = 0.f;
  for (int s0 = 0; s0 < S; s0 += 32) {
    const int s = s0 + lane;
    SdfIn q;
    float phi, alpha, T;
    sdf_chunk(sd, sd_st, ray, S, s, lane, from_density, b, carry, q, phi, alpha, T);
    if (s < S) {
      const float w = __fmul_rn(T, alpha);
      if (w_out) w_out[ray * S + s] = w;
      const float* c = rgb + (ray * S + s) * rgb_st;
      c0 += w * __ldg(c + 0);
      c1 += w * __ldg(c + 1);
      c2 += w * __ldg(c + 2);
    }
  }
  c0 = warp_sum(c0); c1 = warp_sum(c1); c2 = warp_sum(c2);
  if (lane == 0) { C[ray * 3 + 0] = c0; C[ray * 3 + 1] = c1; C[ray * 3 + 2] = c2; }
}

template <int NCH>
__global__ void __launch_bounds__(kSdfRaysPerCta * 32)
composite_sdf_bwd_kernel(const float* __restrict__ rgb, long long rgb_st, const float* __restrict__ sd, long long sd_st,
                         int from_density, const float* __restrict__ b_ptr, long long R, int S, const float* __restrict__ gC,
                         const float* __restrict__ gw, float* __restrict__ drgb, long long drgb_st, float* __restrict__ dsd,
                         long long dsd_st, float* __restrict__ db_ray) {
  const int lane = threadIdx.x & 31;
  const long long ray = (long long)blockIdx.x * kSdfRaysPerCta + (threadIdx.x >> 5);
  if (ray >= R) return;
  const float b = __ldg(b_ptr);
  const float g0 = __ldg(gC + ray * 3 + 0), g1 = __ldg(gC + ray * 3 + 1), g2 = __ldg(gC + ray * 3 + 2);
  const bool packed = drgb_st == 4 && dsd_st == 4 && dsd == drgb + 3;

  float Tk[NCH], ak[NCH], ek[NCH], pk[NCH];      // transmittance, alpha, c (later e), (1 - phi) * gain-independent part
  float sk[NCH], gk[NCH];                        // clamped sdf, d sdf / d input
  float carry = 1.f;
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    Tk[i] = 0.f; ak[i] = 0.f; ek[i] = 0.f; pk[i] = 0.f; sk[i] = 0.f; gk[i] = 0.f;
    if (i * 32 < S) {
      const int s = i * 32 + lane;
      SdfIn q;
      float phi, alpha, T;
      sdf_chunk(sd, sd_st, ray, S, s, lane, from_density, b, carry, q, phi, alpha, T);
      if (s < S) {
        const float* c = rgb + (ray * S + s) * rgb_st;
        float cc = g0 * __ldg(c + 0) + g1 * __ldg(c + 1) + g2 * __ldg(c + 2);
        if (gw) cc += __ldg(gw + ray * S + s);
        Tk[i] = T; ak[i] = alpha; ek[i] = cc; pk[i] = 1.f - phi; sk[i] = q.s; gk[i] = q.gain;
      }
    }
  }
  float rcarry = 0.f;                            // sum of w*c over all later chunks
#pragma unroll
  for (int i = NCH - 1; i >= 0; --i) {
    if (i * 32 < S) {
      const float wc = Tk[i] * ak[i] * ek[i];
      const float rin = warp_rincl_scan(wc, lane);
      float after = __shfl_down_sync(kFull, rin, 1);
      if (lane == 31) after = 0.f;
      const float suffix = rcarry + after;       // sum_{k>i} w_k c_k
      rcarry += __shfl_sync(kFull, rin, 0);
      ek[i] = ak[i] > 0.f ? suffix - Tk[i] * ek[i] * (1.f - ak[i]) : 0.f;
    }
  }
  float db = 0.f;
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    if (i * 32 < S) {
      const int s = i * 32 + lane;
      float eprev = __shfl_up_sync(kFull, ek[i], 1);
      const float elast = i > 0 ? __shfl_sync(kFull, ek[i > 0 ? i - 1 : 0], 31) : 0.f;
      if (lane == 0) eprev = elast;
      if (s < S) {
        const float base = pk[i] * (eprev - ek[i]);                      // dL/d(s_i b)
        db += sk[i] * base;
        const float ds = b * base * gk[i];
        const float w = Tk[i] * ak[i];
        const long long gs = ray * S + s;
        if (packed) {
          *reinterpret_cast<float4*>(drgb + gs * 4) = make_float4(w * g0, w * g1, w * g2, ds);
        } else {
          float* o = drgb + gs * drgb_st;
          o[0] = w * g0; o[1] = w * g1; o[2] = w * g2;
          dsd[gs * dsd_st] = ds;
        }
      }
    }
  }
  db = warp_sum(db);
  if (lane == 0) db_ray[ray] = db;
}

// ---- eikonal stencil ------------------------------------------------------------------------------------------------
struct Box3 { float lo[3], hi[3]; };

// pts (6, n, 3): slab 2*axis = clamp(x + eps e_axis), slab 2*axis + 1 = clamp(x - eps e_axis)   (test_hash.py:91-102)
__global__ void sdf_stencil_points_kernel(const float* __restrict__ x, long long n, float eps, Box3 box, float* __restrict__ pts) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float p[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) p[k] = __ldg(x + i * 3 + k);
#pragma unroll
  for (int slab = 0; slab < 6; ++slab) {
    const int axis = slab >> 1;
    float* o = pts + ((long long)slab * n + i) * 3;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      float v = p[k];
      if (k == axis) v = (slab & 1) ? __fsub_rn(v, eps) : __fadd_rn(v, eps);
      o[k] = fminf(fmaxf(v, box.lo[k]), box.hi[k]);
    }
  }
}

// dens (6, n): LeakyReLU output of the density head at the stencil points -> |central-difference gradient of the sdf|
__global__ void sdf_eikonal_fwd_kernel(const float* __restrict__ dens, long long n, float eps, float* __restrict__ norm,
                                       float* __restrict__ grads) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float acc = 0.f, g[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    float gain;
    const float sp = sdf_from_density(__ldg(dens + (long long)(2 * a) * n + i), gain);
    const float sn = sdf_from_density(__ldg(dens + (long long)(2 * a + 1) * n + i), gain);
    g[a] = __fdiv_rn(__fmul_rn(0.5f, __fsub_rn(sp, sn)), eps);           // test_hash.py:104
    acc = __fadd_rn(acc, __fmul_rn(g[a], g[a]));                          // helper.py:296
  }
  norm[i] = sqrtf(acc);
  if (grads) { grads[i * 3 + 0] = g[0]; grads[i * 3 + 1] = g[1]; grads[i * 3 + 2] = g[2]; }
}

__global__ void sdf_eikonal_bwd_kernel(const float* __restrict__ dens, long long n, float eps, const float* __restrict__ gnorm,
                                       float* __restrict__ ddens) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float acc = 0.f, g[3], gp[3], gn[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float sp = sdf_from_density(__ldg(dens + (long long)(2 * a) * n + i), gp[a]);
    const float sn = sdf_from_density(__ldg(dens + (long long)(2 * a + 1) * n + i), gn[a]);
    g[a] = __fdiv_rn(__fmul_rn(0.5f, __fsub_rn(sp, sn)), eps);
    acc = __fadd_rn(acc, __fmul_rn(g[a], g[a]));
  }
  const float nrm = sqrtf(acc), go = __ldg(gnorm + i);
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float coef = go * (g[a] / nrm) * (0.5f / eps);                  // a zero norm gives NaN, as torch's sqrt backward
    ddens[(long long)(2 * a) * n + i] = coef * gp[a];
    ddens[(long long)(2 * a + 1) * n + i] = -coef * gn[a];
  }
}

}  // namespace hbr

using namespace hbr;

static int check_sdf(const float* rgb, const float* sd, const float* b, int64_t R, int64_t S) {
  HBR_REQUIRE(R >= 0 && S >= 1 && S <= 1024, "R=%lld S=%lld (S must be 1..1024)", (long long)R, (long long)S);
  HBR_REQUIRE(rgb && sd && b, "NULL pointer");
  return HBR_OK;
}

extern "C" int hbr_composite_sdf_fwd(const float* rgb, int64_t rgb_st, const float* sdf, int64_t sdf_st, int from_density,
                                     const float* b, int64_t R, int64_t S, float* C, float* w, void* stream) {
  if (R == 0) return HBR_OK;
  if (int rc = check_sdf(rgb, sdf, b, R, S)) return rc;
  HBR_REQUIRE(C != nullptr, "C is NULL");
  composite_sdf_fwd_kernel<<<(unsigned)ceil_div(R, kSdfRaysPerCta), kSdfRaysPerCta * 32, 0, as_stream(stream)>>>(
      rgb, rgb_st, sdf, sdf_st, from_density, b, R, (int)S, C, w);
  HBR_LAUNCH_CHECK();
  return HBR_OK;
}

extern "C" int hbr_composite_sdf_bwd(const float* rgb, int64_t rgb_st, const float* sdf, int64_t sdf_st, int from_density,
                                     const float* b, int64_t R, int64_t S, const float* gC, const float* gw, float* drgb,
                                     int64_t drgb_st, float* dsdf, int64_t dsdf_st, float* db_ray, void* stream) {
  if (R == 0) return HBR_OK;
  if (int rc = check_sdf(rgb, sdf, b, R, S)) return rc;
  HBR_REQUIRE(gC && drgb && dsdf && db_ray, "NULL pointer");
  if (drgb_st == 4 && dsdf_st == 4 && dsdf == drgb + 3)
    HBR_REQUIRE((uintptr_t)drgb % 16 == 0, "packed gradient buffer must be 16-byte aligned");
  const unsigned grid = (unsigned)ceil_div(R, kSdfRaysPerCta);
  cudaStream_t st = as_stream(stream);
#define HBR_SB(N)                                                                                                    \
  composite_sdf_bwd_kernel<N><<<grid, kSdfRaysPerCta * 32, 0, st>>>(rgb, rgb_st, sdf, sdf_st, from_density, b, R, (int)S, gC, \
                                                                    gw, drgb, drgb_st, dsdf, dsdf_st, db_ray)
  if (S <= 128) HBR_SB(4);
  else if (S <= 256) HBR_SB(8);
  else if (S <= 512) HBR_SB(16);
  else HBR_SB(32);
#undef HBR_SB
  HBR_LAUNCH_CHECK();
  return HBR_OK;
}

extern "C" int hbr_sdf_stencil_points(const float* x, int64_t n, float eps, const float* lo3_host, const float* hi3_host,
                                      float* pts, void* stream) {
  HBR_REQUIRE(n >= 0, "negative size");
  if (n == 0) return HBR_OK;
  HBR_REQUIRE(x && lo3_host && hi3_host && pts, "NULL pointer");
  Box3 box;
  for (int k = 0; k < 3; ++k) { box.lo[k] = lo3_host[k]; box.hi[k] = hi3_host[k]; }
  sdf_stencil_points_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, as_stream(stream)>>>(x, n, eps, box, pts);
  HBR_LAUNCH_CHECK();
  return HBR_OK;
}

extern "C" int hbr_sdf_eikonal_fwd(const float* dens6, int64_t n, float eps, float* norm, float* grads, void* stream) {
  HBR_REQUIRE(n >= 0 && eps > 0.f, "n=%lld eps=%g", (long long)n, (double)eps);
  if (n == 0) return HBR_OK;
  HBR_REQUIRE(dens6 && norm, "NULL pointer");
  sdf_eikonal_fwd_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, as_stream(stream)>>>(dens6, n, eps, norm, grads);
  HBR_LAUNCH_CHECK();
  return HBR_OK;
}

extern "C" int hbr_sdf_eikonal_bwd(const float* dens6, int64_t n, float eps, const float* gnorm, float* ddens6, void* stream) {
  HBR_REQUIRE(n >= 0 && eps > 0.f, "n=%lld eps=%g", (long long)n, (double)eps);
  if (n == 0) return HBR_OK;
  HBR_REQUIRE(dens6 && gnorm && ddens6, "NULL pointer");
  sdf_eikonal_bwd_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, as_stream(stream)>>>(dens6, n, eps, gnorm, ddens6);
  HBR_LAUNCH_CHECK();
  return HBR_OK;
}
