// the step either side of the path.  The reference precomputes every pixel ray of every training view
// on the CPU (train_hash2.py:74-96: get_od per 50-view batch -> a 2.56 GB TensorDataset -> 8 DataLoader workers -> H2D
// of each 4096-ray batch).  Here the views (c2w + pixels) stay resident in HBM and a batch is generated where it is
// consumed: ray id -> (view, row, col) -> helper.py:176-208 arithmetic -> rays_o, unit rays_d, |d|, ground-truth rgb.
//
// Arithmetic of get_od, in the reference's order (separate roundings, IEEE division; no FMA contraction except the
// 3-term rotation, whose summation order inside the reference's BLAS call is unspecified anyway):
//   i = (col - K[0,2]) / K[0,0]      j = (row - K[1,2]) / K[1,1]      dirs = (i, -j, -1)
//   d = c2w[:3,:3] @ dirs            n = |d|                           rays_d = d / n      rays_o = c2w[:3,3]
// find_bounding_box (helper.py:109-141): min / max over all pixels of all views of  o + d*t,  t in {near, far + 1.5}.
#include "common.cuh"

namespace hbr {

struct Intrinsics { float fx, fy, cx, cy; };

__device__ __forceinline__ void pixel_ray(const float* __restrict__ c2w, long long view, int row, int col, const Intrinsics& k,
                                          float o[3], float d[3], float& norm) {
  const float* m = c2w + view * 16;                               // (4,4) row-major
  const float i = __fdiv_rn(__fsub_rn((float)col, k.cx), k.fx);
  const float j = __fdiv_rn(__fsub_rn((float)row, k.cy), k.fy);
  float v[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float r0 = __ldg(m + 4 * a), r1 = __ldg(m + 4 * a + 1), r2 = __ldg(m + 4 * a + 2);
    v[a] = fmaf(r2, -1.f, fmaf(r1, -j, __fmul_rn(r0, i)));
    o[a] = __ldg(m + 4 * a + 3);
  }
  norm = __fsqrt_rn(fmaf(v[2], v[2], fmaf(v[1], v[1], __fmul_rn(v[0], v[0]))));
#pragma unroll
  for (int a = 0; a < 3; ++a) d[a] = __fdiv_rn(v[a], norm);
}

template <typename PIX>
__global__ void __launch_bounds__(256)
ray_gen_kernel(const float* __restrict__ c2w, long long n_views, int H, int W, Intrinsics k, const long long* __restrict__ ids,
               long long first, long long n, const PIX* __restrict__ images, float* __restrict__ rays_o,
               float* __restrict__ rays_d, float* __restrict__ dir_norm, float* __restrict__ gt, int* __restrict__ bad) {
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  long long id = ids != nullptr ? ids[r] : first + r;
  const long long hw = (long long)H * W;
  if (id < 0 || id >= n_views * hw) {                              // out-of-range id: flag it, emit zeros
    if (bad != nullptr) atomicExch(bad, 1);
    for (int a = 0; a < 3; ++a) { rays_o[3 * r + a] = 0.f; rays_d[3 * r + a] = 0.f; if (gt) gt[3 * r + a] = 0.f; }
    dir_norm[r] = 0.f;
    return;
  }
  const long long view = id / hw;
  const int p = (int)(id - view * hw);
  const int row = p / W, col = p - row * W;
  float o[3], d[3], nrm;
  pixel_ray(c2w, view, row, col, k, o, d, nrm);
#pragma unroll
  for (int a = 0; a < 3; ++a) { rays_o[3 * r + a] = o[a]; rays_d[3 * r + a] = d[a]; }
  dir_norm[r] = nrm;
  if (gt != nullptr && images != nullptr) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      if constexpr (sizeof(PIX) == 1) gt[3 * r + a] = __fdiv_rn((float)images[3 * id + a], 255.f);   // torchvision ToTensor
      else gt[3 * r + a] = (float)images[3 * id + a];
    }
  }
}

__device__ __forceinline__ void atomic_min_f(float* addr, float v) {
  int old = __float_as_int(*addr);
  while (v < __int_as_float(old)) {
    const int seen = atomicCAS(reinterpret_cast<int*>(addr), old, __float_as_int(v));
    if (seen == old) break;
    old = seen;
  }
}
__device__ __forceinline__ void atomic_max_f(float* addr, float v) {
  int old = __float_as_int(*addr);
  while (v > __int_as_float(old)) {
    const int seen = atomicCAS(reinterpret_cast<int*>(addr), old, __float_as_int(v));
    if (seen == old) break;
    old = seen;
  }
}

__global__ void __launch_bounds__(256)
ray_bbox_kernel(const float* __restrict__ c2w, long long n_views, int H, int W, Intrinsics k, float t0, float t1,
                float* __restrict__ bounds) {
  const long long hw = (long long)H * W, total = n_views * hw;
  float mn[3] = {1e30f, 1e30f, 1e30f}, mx[3] = {-1e30f, -1e30f, -1e30f};
  for (long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x; id < total; id += (long long)gridDim.x * blockDim.x) {
    const long long view = id / hw;
    const int p = (int)(id - view * hw);
    const int row = p / W, col = p - row * W;
    float o[3], d[3], nrm;
    pixel_ray(c2w, view, row, col, k, o, d, nrm);
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const float p0 = __fadd_rn(o[a], __fmul_rn(d[a], t0)), p1 = __fadd_rn(o[a], __fmul_rn(d[a], t1));
      mn[a] = fminf(mn[a], fminf(p0, p1));
      mx[a] = fmaxf(mx[a], fmaxf(p0, p1));
    }
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
      mn[a] = fminf(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], s));
      mx[a] = fmaxf(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], s));
    }
  }
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      atomic_min_f(bounds + a, mn[a]);
      atomic_max_f(bounds + 3 + a, mx[a]);
    }
  }
}

}  // namespace hbr

using namespace hbr;

extern "C" int hbr_ray_gen(const float* c2w, int64_t n_views, int H, int W, float fx, float fy, float cx, float cy,
                           const int64_t* ray_ids, int64_t first, int64_t n_rays, const void* images, int image_dtype,
                           float* rays_o, float* rays_d, float* dir_norm, float* gt, int* bad_id, void* stream) {
  HBR_REQUIRE(n_rays >= 0 && n_views > 0 && H > 0 && W > 0, "n_rays=%lld n_views=%lld H=%d W=%d", (long long)n_rays,
              (long long)n_views, H, W);
  HBR_REQUIRE(fx != 0.f && fy != 0.f, "zero focal length");
  HBR_REQUIRE(image_dtype == HBR_F32 || image_dtype == HBR_U8, "images must be HBR_F32 or HBR_U8 (V,H,W,3)");
  HBR_REQUIRE((gt == nullptr) == (images == nullptr), "gt and images go together");
  if (n_rays == 0) return HBR_OK;
  HBR_REQUIRE(c2w && rays_o && rays_d && dir_norm, "NULL pointer");
  static_assert(sizeof(long long) == sizeof(int64_t), "id width");
  const Intrinsics k{fx, fy, cx, cy};
  const int grid = (int)ceil_div(n_rays, 256);
  const long long* ids = reinterpret_cast<const long long*>(ray_ids);
  if (image_dtype == HBR_U8)
    ray_gen_kernel<unsigned char><<<grid, 256, 0, as_stream(stream)>>>(c2w, n_views, H, W, k, ids, first, n_rays,
        static_cast<const unsigned char*>(images), rays_o, rays_d, dir_norm, gt, bad_id);
  else
    ray_gen_kernel<float><<<grid, 256, 0, as_stream(stream)>>>(c2w, n_views, H, W, k, ids, first, n_rays,
        static_cast<const float*>(images), rays_o, rays_d, dir_norm, gt, bad_id);
  HBR_LAUNCH_CHECK();
  return HBR_OK;
}

extern "C" int hbr_ray_bbox(const float* c2w, int64_t n_views, int H, int W, float fx, float fy, float cx, float cy,
                            float t0, float t1, float* bounds, void* stream) {
  HBR_REQUIRE(n_views > 0 && H > 0 && W > 0 && fx != 0.f && fy != 0.f, "n_views=%lld H=%d W=%d", (long long)n_views, H, W);
  HBR_REQUIRE(c2w && bounds, "NULL pointer");
  const int64_t total = n_views * (int64_t)H * W;
  const int grid = (int)min64(ceil_div(total, 256), (int64_t)sm_count() * 8);
  ray_bbox_kernel<<<grid, 256, 0, as_stream(stream)>>>(c2w, n_views, H, W, Intrinsics{fx, fy, cx, cy}, t0, t1, bounds);
  HBR_LAUNCH_CHECK();
  return HBR_OK;
}
