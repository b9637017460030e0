// C entry points of the tensor-core MLP (tcgen05 + TMEM).  The kernels live in mlp_tc_impl.cuh, instantiated once per
// operand format: bf16 (torch.autocast(bfloat16)) and fp16 (torch.autocast(float16) -- what the reference's trainer
// runs, train_hash2.py:218).  `operand` selects the instantiation (HBR_BF16 | HBR_F16).
#include "mlp_layout.cuh"
#include "tc_common.cuh"

#define HBR_OP tc::OpBf16
#define HBR_OPNS bf16
#include "mlp_tc_impl.cuh"
#undef HBR_OP
#undef HBR_OPNS
#define HBR_OP tc::OpF16
#define HBR_OPNS f16
#include "mlp_tc_impl.cuh"
#undef HBR_OP
#undef HBR_OPNS

using namespace hbr;
using bf16::narrow_shape;
using bf16::Scratch;
using tc::EncArgs;

static int check_operand(int operand, float grad_scale) {
  HBR_REQUIRE(operand == HBR_BF16 || operand == HBR_F16, "operand format %d (HBR_BF16 or HBR_F16)", operand);
  HBR_REQUIRE(grad_scale > 0.f && grad_scale < 3.0e38f, "grad_scale %g", (double)grad_scale);
  return HBR_OK;
}
#define HBR_BY_OPERAND(...) (operand == HBR_F16 ? f16::__VA_ARGS__ : bf16::__VA_ARGS__)

extern "C" int64_t hbr_mlp_tc_scratch_bytes(const hbr_mlp_dims* dims) {
  if (check_dims(dims)) return 0;
  return narrow_shape(dims) ? Scratch<32, 48>::total : Scratch<64, 64>::total;
}


// feat_dtype: HBR_F32, or the operand format itself (features already rounded by hbr_hash_encode_fwd_rays: contiguous rows)
static int check_feat(const void* feat, int feat_dtype, int64_t feat_stride, const hbr_mlp_dims* d, int operand) {
  HBR_REQUIRE(feat_dtype == HBR_F32 || feat_dtype == operand, "feat_dtype %d: HBR_F32 or the operand format %d", feat_dtype, operand);
  if (feat_dtype != HBR_F32) {
    const int kp = bf16::narrow_shape(d) ? 32 : 64;
    HBR_REQUIRE(d->in0 == kp && feat_stride == kp && (uintptr_t)feat % 16 == 0,
                "16-bit features need in0 == feat_stride == %d and a 16-byte aligned buffer", kp);
  }
  return HBR_OK;
}

extern "C" int hbr_mlp_fwd_tc(const void* feat_, int feat_dtype, int64_t feat_stride, const float* dirs, int64_t dir_group, int64_t n,
                              const float* params, const hbr_mlp_dims* dims, int operand, float* out, void* scratch,
                              int image_ready, const unsigned long long* n_dev, const int32_t* dir_rows, void* stream) {
  if (int rc = check_dims(dims)) return rc;
  if (int rc = check_operand(operand, 1.f)) return rc;
  if (n == 0) return HBR_OK;
  const float* feat = static_cast<const float*>(feat_);
  HBR_REQUIRE(feat && dirs && params && out, "NULL pointer");
  if (int rc = check_feat(feat_, feat_dtype, feat_stride, dims, operand)) return rc;
  const int f16 = feat_dtype != HBR_F32;
  HBR_REQUIRE(feat_stride >= dims->in0 && dir_group >= 1, "bad stride / dir_group");
  HBR_REQUIRE((uintptr_t)out % 16 == 0 && (uintptr_t)scratch % 256 == 0, "out / scratch alignment");
  cudaStream_t st = as_stream(stream);
  uint8_t* sc = static_cast<uint8_t*>(scratch);
  if (narrow_shape(dims))
    return HBR_BY_OPERAND(launch_fwd_tc<32, 48, 4, false>(feat, feat_stride, dirs, dir_group, n, params, dims->in0, dims->d_view,
                                                          out, sc, EncArgs{}, HashGeom{}, f16, image_ready, n_dev, dir_rows, st));
  return HBR_BY_OPERAND(launch_fwd_tc<64, 64, 4, false>(feat, feat_stride, dirs, dir_group, n, params, dims->in0, dims->d_view, out,
                                                        sc, EncArgs{}, HashGeom{}, f16, image_ready, n_dev, dir_rows, st));
}

extern "C" int hbr_mlp_bwd_tc(const void* feat_, int feat_dtype, int64_t feat_stride, const float* dirs, int64_t dir_group, int64_t n,
                              const float* params, const hbr_mlp_dims* dims, int operand, const float* out,
                              const float* dout, float* dfeat, int64_t dfeat_stride, float* ddirs, float* dparams,
                              float grad_scale, void* scratch, int image_ready, int defer_reduce,
                              const unsigned long long* n_dev, const int32_t* dir_rows, void* stream) {
  if (int rc = check_dims(dims)) return rc;
  HBR_REQUIRE(!defer_reduce || (scratch != nullptr && dparams != nullptr), "defer_reduce needs scratch and dparams");
  if (int rc = check_operand(operand, grad_scale)) return rc;
  if (n == 0) return HBR_OK;
  const float* feat = static_cast<const float*>(feat_);
  HBR_REQUIRE(feat && dirs && params && out && dout, "NULL pointer");
  if (int rc = check_feat(feat_, feat_dtype, feat_stride, dims, operand)) return rc;
  const int f16 = feat_dtype != HBR_F32;
  HBR_REQUIRE(feat_stride >= dims->in0 && dir_group >= 1, "bad stride / dir_group");
  HBR_REQUIRE((uintptr_t)dout % 16 == 0 && (uintptr_t)out % 16 == 0 && (uintptr_t)scratch % 256 == 0,
              "out / dout / scratch alignment");
  if (dfeat && dfeat_stride == HBR_DFEAT_LEVEL_MAJOR)
    HBR_REQUIRE(narrow_shape(dims) && dims->in0 == 32 && n % 2 == 0 && (uintptr_t)dfeat % 16 == 0 && n_dev == nullptr,
                "level-major dfeat: in0 = 32, even n, 16-byte aligned buffer, no compacted list");
  else
    HBR_REQUIRE(!dfeat || dfeat_stride >= dims->in0, "dfeat_stride too small");
  cudaStream_t st = as_stream(stream);
  uint8_t* sc = static_cast<uint8_t*>(scratch);
  if (narrow_shape(dims))
    return HBR_BY_OPERAND(launch_bwd_tc<32, 48, 2, false>(feat, feat_stride, dirs, dir_group, n, params, dims->in0, dims->d_view,
                                                          out, dout, dfeat, dfeat_stride, ddirs, dparams, sc, EncArgs{}, HashGeom{},
                                                          grad_scale, f16, image_ready, defer_reduce, n_dev, dir_rows, st));
  return HBR_BY_OPERAND(launch_bwd_tc<64, 64, 1, false>(feat, feat_stride, dirs, dir_group, n, params, dims->in0, dims->d_view, out,
                                                        dout, dfeat, dfeat_stride, ddirs, dparams, sc, EncArgs{}, HashGeom{}, grad_scale,
                                                        f16, image_ready, defer_reduce, n_dev, dir_rows, st));
}

extern "C" int hbr_mlp_tc_prepare(const float* params, const hbr_mlp_dims* dims, int operand, void* scratch, void* stream) {
  if (int rc = check_dims(dims)) return rc;
  if (int rc = check_operand(operand, 1.f)) return rc;
  HBR_REQUIRE(params && scratch && (uintptr_t)scratch % 256 == 0, "NULL / misaligned pointer");
  uint8_t* sc = static_cast<uint8_t*>(scratch);
  cudaStream_t st = as_stream(stream);
  if (narrow_shape(dims)) return HBR_BY_OPERAND(prepare_tc<32, 48>(params, dims->in0, dims->d_view, sc, st));
  return HBR_BY_OPERAND(prepare_tc<64, 64>(params, dims->in0, dims->d_view, sc, st));
}

extern "C" int hbr_mlp_tc_reduce_grads(const hbr_mlp_dims* dims, int64_t n, void* scratch, float* dparams, void* stream) {
  if (int rc = check_dims(dims)) return rc;
  if (n == 0) return HBR_OK;
  HBR_REQUIRE(scratch && dparams, "NULL pointer");
  uint8_t* sc = static_cast<uint8_t*>(scratch);
  cudaStream_t st = as_stream(stream);
  if (narrow_shape(dims)) return bf16::reduce_grads_tc<32, 48, 2>(n, dims->in0, dims->d_view, sc, dparams, st);
  return bf16::reduce_grads_tc<64, 64, 1>(n, dims->in0, dims->d_view, sc, dparams, st);
}

// ---- fused field evaluation: hash-grid encoder + MLP_3D in one kernel per direction ------------------------------
static int check_field(const hbr_hash_geom* geom, const hbr_mlp_dims* dims) {
  HBR_REQUIRE(geom != nullptr, "geom is NULL");
  if (int rc = check_dims(dims)) return rc;
  HBR_REQUIRE(geom->F == 2 && geom->L == 16 && geom->E == 0 && dims->in0 == 32 && is_pow2(geom->T) && geom->T >= 2 &&
                  geom->T <= (1u << 30) && narrow_shape(dims),
              "the fused field kernels cover F=2, L=16, E=0, power-of-two T, d_view <= 25 (use the separate kernels otherwise)");
  return HBR_OK;
}

extern "C" int hbr_field_fwd_tc(const float* x, int64_t n, const float* table, const hbr_hash_geom* geom, const float* dirs,
                                int64_t dir_group, const float* params, const hbr_mlp_dims* dims, int operand, float* out,
                                void* feat16, void* scratch, void* stream) {
  if (int rc = check_field(geom, dims)) return rc;
  if (int rc = check_operand(operand, 1.f)) return rc;
  if (n == 0) return HBR_OK;
  HBR_REQUIRE(x && table && dirs && params && out && feat16, "NULL pointer");
  HBR_REQUIRE(dir_group >= 1, "dir_group");
  HBR_REQUIRE((uintptr_t)out % 16 == 0 && (uintptr_t)table % 16 == 0 && (uintptr_t)feat16 % 16 == 0 &&
                  (uintptr_t)scratch % 256 == 0, "alignment");
  if (operand == HBR_F16) {
    EncArgs e{};
    e.x = x; e.table = table; e.feat16 = static_cast<uint16_t*>(feat16);
    return f16::launch_fwd_tc<32, 48, 4, true>(nullptr, 32, dirs, dir_group, n, params, 32, dims->d_view, out,
                                               static_cast<uint8_t*>(scratch), e, to_device_geom(*geom), 0, 0, nullptr, nullptr, as_stream(stream));
  }
  EncArgs e{};
  e.x = x; e.table = table; e.feat16 = static_cast<uint16_t*>(feat16);
  return bf16::launch_fwd_tc<32, 48, 4, true>(nullptr, 32, dirs, dir_group, n, params, 32, dims->d_view, out,
                                              static_cast<uint8_t*>(scratch), e, to_device_geom(*geom), 0, 0, nullptr, nullptr, as_stream(stream));
}

extern "C" int hbr_field_bwd_tc(const float* x, int64_t n, const hbr_hash_geom* geom, const float* dirs, int64_t dir_group,
                                const float* params, const hbr_mlp_dims* dims, int operand, const void* feat16,
                                const float* out, const float* dout, float* dtable, float* ddirs, float* dparams,
                                float grad_scale, void* scratch, void* stream) {
  if (int rc = check_field(geom, dims)) return rc;
  if (int rc = check_operand(operand, grad_scale)) return rc;
  if (n == 0) return HBR_OK;
  HBR_REQUIRE(x && dirs && params && feat16 && out && dout && dtable, "NULL pointer");
  HBR_REQUIRE(dir_group >= 1, "dir_group");
  HBR_REQUIRE((uintptr_t)out % 16 == 0 && (uintptr_t)dout % 16 == 0 && (uintptr_t)dtable % 16 == 0 &&
                  (uintptr_t)feat16 % 16 == 0 && (uintptr_t)scratch % 256 == 0, "alignment");
  uint16_t* f16p = const_cast<uint16_t*>(static_cast<const uint16_t*>(feat16));
  if (operand == HBR_F16) {
    EncArgs e{};
    e.x = x; e.dtable = dtable; e.feat16 = f16p;
    return f16::launch_bwd_tc<32, 48, 2, true>(nullptr, 32, dirs, dir_group, n, params, 32, dims->d_view, out, dout, nullptr, 32,
                                               ddirs, dparams, static_cast<uint8_t*>(scratch), e, to_device_geom(*geom),
                                               grad_scale, 0, 0, 0, nullptr, nullptr, as_stream(stream));
  }
  EncArgs e{};
  e.x = x; e.dtable = dtable; e.feat16 = f16p;
  return bf16::launch_bwd_tc<32, 48, 2, true>(nullptr, 32, dirs, dir_group, n, params, 32, dims->d_view, out, dout, nullptr, 32,
                                              ddirs, dparams, static_cast<uint8_t*>(scratch), e, to_device_geom(*geom),
                                              grad_scale, 0, 0, 0, nullptr, nullptr, as_stream(stream));
}

// Forward of the training step's field evaluation in ONE kernel: the hash-grid gather (hbr_hash_encode_fwd_rays) on dedicated
// warps of the tensor-core MLP kernel (mlp_fwd_tc_kernel<GATH>), feeding its tile groups tile by tile through shared
// memory.  out (R*S,4) = [rgb, sigma]; feat16 (R*S,32): the 16-bit features for the backward recompute.  image_ready != 0:
// the operand image in `scratch` was built for these parameters and format (hbr_mlp_tc_prepare); otherwise the kernel
// converts the parameters itself.
#ifndef HBR_FWD_GATHER_WARPS
#define HBR_FWD_GATHER_WARPS 8
#endif
extern "C" int hbr_field_fwd_rays_tc(const float* rays_o, const float* rays_d, const float* t, int64_t t_ray_stride, int64_t R,
                                     int64_t S, const float* table, const hbr_hash_geom* geom, const float* dirs,
                                     const float* params, const hbr_mlp_dims* dims, int operand, float* out, void* feat16,
                                     void* scratch, int image_ready, void* stream) {
  if (int rc = check_field(geom, dims)) return rc;
  if (int rc = check_operand(operand, 1.f)) return rc;
  HBR_REQUIRE(R >= 0 && S >= 1 && R < (1LL << 40) / S, "R=%lld S=%lld", (long long)R, (long long)S);
  HBR_REQUIRE(t_ray_stride == 0 || t_ray_stride >= S, "t_ray_stride %lld", (long long)t_ray_stride);
  HBR_REQUIRE(!image_ready || scratch != nullptr, "image_ready without scratch");
  const int64_t n = R * S;
  if (n == 0) return HBR_OK;
  HBR_REQUIRE(rays_o && rays_d && t && table && dirs && params && out && feat16, "NULL pointer");
  HBR_REQUIRE((uintptr_t)out % 16 == 0 && (uintptr_t)table % 16 == 0 && (uintptr_t)feat16 % 16 == 0 &&
                  (uintptr_t)scratch % 256 == 0, "alignment");
  EncArgs e{};
  e.table = table; e.feat16 = static_cast<uint16_t*>(feat16);
  e.ro = rays_o; e.rd = rays_d; e.rt = t; e.S = S; e.t_stride = t_ray_stride;
  const uint8_t* image = image_ready ? static_cast<const uint8_t*>(scratch) : nullptr;
  return HBR_BY_OPERAND(launch_fwd_gather_tc<HBR_FWD_GATHER_WARPS>(dirs, S, n, params, dims->d_view, out, image, e,
                                                                     to_device_geom(*geom), as_stream(stream)));
}

// MLP backward with the hash-grid scatter-add on dedicated warps of the same kernel (mlp_bwd_tc_kernel<SCAT = 7>): what
// hbr_mlp_bwd_tc followed by hbr_hash_encode_bwd_rays computes (same arithmetic, same run merging), without the fp32
// d(feature) tensor and with the two phases overlapped on every SM.  feat16: the (R*S,32) 16-bit features
// hbr_hash_encode_fwd_rays wrote in the operand format.  dtable is accumulated into (all 16 levels).
extern "C" int hbr_field_bwd_rays_tc(const void* feat16, const float* rays_o, const float* rays_d, const float* t,
                                     int64_t t_ray_stride, int64_t R, int64_t S, const hbr_hash_geom* geom, const float* dirs,
                                     const float* params, const hbr_mlp_dims* dims, int operand, const float* out,
                                     const float* dout, float* dtable, float* ddirs, float* dparams, float grad_scale,
                                     void* scratch, int image_ready, int defer_reduce, void* stream) {
  if (int rc = check_field(geom, dims)) return rc;
  if (int rc = check_operand(operand, grad_scale)) return rc;
  HBR_REQUIRE(!defer_reduce || (scratch != nullptr && dparams != nullptr), "defer_reduce needs scratch and dparams");
  HBR_REQUIRE(R >= 0 && S >= 1 && R < (1LL << 40) / S, "R=%lld S=%lld", (long long)R, (long long)S);
  HBR_REQUIRE(t_ray_stride == 0 || t_ray_stride >= S, "t_ray_stride %lld", (long long)t_ray_stride);
  const int64_t n = R * S;
  if (n == 0) return HBR_OK;
  HBR_REQUIRE(feat16 && rays_o && rays_d && t && dirs && params && out && dout && dtable, "NULL pointer");
  HBR_REQUIRE((uintptr_t)out % 16 == 0 && (uintptr_t)dout % 16 == 0 && (uintptr_t)dtable % 16 == 0 &&
                  (uintptr_t)feat16 % 16 == 0 && (uintptr_t)scratch % 256 == 0, "alignment");
  EncArgs e{};
  e.dtable = dtable; e.ro = rays_o; e.rd = rays_d; e.rt = t; e.S = S; e.t_stride = t_ray_stride;
  const float* feat = static_cast<const float*>(feat16);
  // one tile group + 11 scatter warps + the weight-gradient issuer per SM.  Measured at 524 288 points (T = 2^19, cold L2):
  // 312 us against 151 + 167 us for hbr_mlp_bwd_tc + hbr_hash_encode_bwd_rays; two tile groups + 7 scatter warps: 335 us
  // (the scatter warps alone take 222 us with 11 warps, 235 us with 7; the layer chains alone 208 us / 153 us).
  return HBR_BY_OPERAND(launch_bwd_tc<32, 48, 1, false, 11>(feat, 32, dirs, S, n, params, 32, dims->d_view, out, dout, nullptr, 32,
                                                            ddirs, dparams, static_cast<uint8_t*>(scratch), e,
                                                            to_device_geom(*geom), grad_scale, 1, image_ready, defer_reduce,
                                                            nullptr, nullptr, as_stream(stream)));
}
