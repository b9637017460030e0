// the occupancy grid of vol_renderer.py:106-140 as a LIVE empty-space skipper.
// The reference keeps a 256^3 bool grid (max_dim / 4 per axis) but never lets it skip work: update_grid is commented out
// of vol_render (vol_renderer.py:205) and the masked samples are still formed, encoded and only then dropped from the MLP
// call.  Here:
//   * occupancy_update_kernel = update_grid (vol_renderer.py:116-131) as one pass: cells hit by a sample with alpha > 0
//     are set; "no sample hit anything" sets the whole grid (the reference's fallback), decided on the device;
//   * compact_samples_kernel forms the sample positions from the rays, looks each one up in the grid (get_mask,
//     vol_renderer.py:133-140, same arithmetic as hbr_occupancy_mask) and writes ONLY the live ones: their positions, the
//     ray they belong to, and a row map sample -> compacted row (-1 = skipped).  Each ray reserves one contiguous segment
//     of the compacted list (one atomicAdd per ray), so consecutive live samples of a ray stay adjacent -- the locality the
//     hash-grid kernels' coalescing and run merging rely on.  The live count stays on the device: the encoder / MLP /
//     compositor kernels that consume the list read it there (no host synchronisation, capturable in a CUDA graph).
#include "common.cuh"

namespace hbr {

struct Mu3o { float v[3]; };

__device__ __forceinline__ long long grid_cell(const float p[3], const Mu3o& mu, float sigma, int G) {
  long long q[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float u = __fmul_rn(__fdiv_rn(__fsub_rn(p[k], mu.v[k]), sigma), (float)G);
    long long c = __float2ll_rz(u);
    if (c < 0) c += G;                       // Python-style negative index
    c = c < 0 ? 0 : (c >= G ? G - 1 : c);    // the reference would raise here; clamp instead of faulting
    q[k] = c;
  }
  return (q[0] * G + q[1]) * G + q[2];
}

__global__ void occupancy_update_kernel(const float* __restrict__ pts, long long n, const float* __restrict__ alpha,
                                        long long alpha_stride, uint8_t* __restrict__ grid, int G, Mu3o mu, float sigma,
                                        unsigned* __restrict__ any) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  bool hit = false;
  if (i < n && __ldg(alpha + i * alpha_stride) > 0.f) {      // alpha <= 0 -> 0; ceil(alpha) > 0 marks the cell (:121-122)
    const float p[3] = {__ldg(pts + i * 3), __ldg(pts + i * 3 + 1), __ldg(pts + i * 3 + 2)};
    grid[grid_cell(p, mu, sigma, G)] = 1;
    hit = true;
  }
  if (__any_sync(kFull, hit) && (threadIdx.x & 31) == 0) atomicOr(any, 1u);
}
// "if torch.sum(tmp_arr > 0) == 0: bool_grid[...] = True" (:124-125), then the scratch is cleared (:129)
__global__ void occupancy_finish_kernel(uint8_t* __restrict__ grid, long long cells, unsigned* __restrict__ any, unsigned* __restrict__ done) {
  const bool none = *reinterpret_cast<volatile unsigned*>(any) == 0u;
  if (none)
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < cells; i += (long long)gridDim.x * blockDim.x) grid[i] = 1;
  __syncthreads();
  if (threadIdx.x == 0 && atomicAdd(done, 1u) == gridDim.x - 1) { *any = 0u; *done = 0u; }   // last CTA resets the flags
}

constexpr int kCompactWarps = 8;
constexpr int kMaxChunks = 32;                                // S <= 1024

__global__ void __launch_bounds__(kCompactWarps * 32)
compact_samples_kernel(const float* __restrict__ o, const float* __restrict__ d, const float* __restrict__ t, long long t_rs,
                       long long R, int S, const uint8_t* __restrict__ grid, int G, Mu3o mu, float sigma,
                       float* __restrict__ pts_c, int32_t* __restrict__ ray_c, int32_t* __restrict__ rowmap,
                       unsigned long long* __restrict__ count) {
  const int lane = threadIdx.x & 31;
  const long long ray = (long long)blockIdx.x * kCompactWarps + (threadIdx.x >> 5);
  if (ray >= R) return;
  const float ox = __ldg(o + ray * 3), oy = __ldg(o + ray * 3 + 1), oz = __ldg(o + ray * 3 + 2);
  const float dx = __ldg(d + ray * 3), dy = __ldg(d + ray * 3 + 1), dz = __ldg(d + ray * 3 + 2);
  const float* tr = t + ray * t_rs;
  unsigned bits[kMaxChunks];
  int total = 0;
#pragma unroll
  for (int c = 0; c < kMaxChunks; ++c) {
    bits[c] = 0u;
    if (c * 32 < S) {
      const int s = c * 32 + lane;
      bool live = false;
      if (s < S) {
        const float tt = __ldg(tr + s);
        const float p[3] = {__fadd_rn(ox, __fmul_rn(dx, tt)), __fadd_rn(oy, __fmul_rn(dy, tt)), __fadd_rn(oz, __fmul_rn(dz, tt))};
        live = grid[grid_cell(p, mu, sigma, G)] != 0;
      }
      bits[c] = __ballot_sync(kFull, live);
      total += __popc(bits[c]);
    }
  }
  unsigned long long base = 0;
  if (lane == 0 && total > 0) base = atomicAdd(count, (unsigned long long)total);
  base = __shfl_sync(kFull, base, 0);
  int before = 0;
#pragma unroll
  for (int c = 0; c < kMaxChunks; ++c) {
    if (c * 32 < S) {
      const int s = c * 32 + lane;
      if (s < S) {
        const bool live = (bits[c] >> lane) & 1u;
        long long row = -1;
        if (live) {
          row = (long long)base + before + __popc(bits[c] & ((1u << lane) - 1u));
          const float tt = __ldg(tr + s);
          pts_c[row * 3 + 0] = __fadd_rn(ox, __fmul_rn(dx, tt));
          pts_c[row * 3 + 1] = __fadd_rn(oy, __fmul_rn(dy, tt));
          pts_c[row * 3 + 2] = __fadd_rn(oz, __fmul_rn(dz, tt));
          ray_c[row] = (int32_t)ray;
        }
        rowmap[ray * S + s] = (int32_t)row;
      }
      before += __popc(bits[c]);
    }
  }
}

}  // namespace hbr

using namespace hbr;

extern "C" int hbr_occupancy_update(const float* pts, int64_t n, const float* alpha, int64_t alpha_stride, uint8_t* grid, int G,
                                    const float* mu3, float sigma, unsigned int* flags2, void* stream) {
  HBR_REQUIRE(n >= 0 && G > 0 && grid && mu3 && flags2, "bad argument");
  HBR_REQUIRE(n == 0 || (pts && alpha), "NULL pointer");
  Mu3o mu{{mu3[0], mu3[1], mu3[2]}};
  cudaStream_t st = as_stream(stream);
  if (n > 0) {
    occupancy_update_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(pts, n, alpha, alpha_stride, grid, G, mu, sigma, flags2);
    HBR_LAUNCH_CHECK();
  }
  const long long cells = (long long)G * G * G;
  occupancy_finish_kernel<<<(unsigned)min64(ceil_div(cells, 256 * 16), 1024), 256, 0, st>>>(grid, cells, flags2, flags2 + 1);
  HBR_LAUNCH_CHECK();
  return HBR_OK;
}

extern "C" int hbr_compact_samples(const float* rays_o, const float* rays_d, const float* t, int64_t t_ray_stride, int64_t R,
                                   int64_t S, const uint8_t* grid, int G, const float* mu3, float sigma, float* pts_c,
                                   int32_t* ray_c, int32_t* rowmap, unsigned long long* count, void* stream) {
  HBR_REQUIRE(R >= 0 && S >= 1 && S <= 32 * kMaxChunks, "R=%lld S=%lld (S must be 1..1024)", (long long)R, (long long)S);
  HBR_REQUIRE(R * S < (1LL << 31), "R*S=%lld does not fit the 32-bit row map", (long long)(R * S));
  HBR_REQUIRE(t_ray_stride == 0 || t_ray_stride >= S, "t_ray_stride %lld", (long long)t_ray_stride);
  if (R == 0) return HBR_OK;
  HBR_REQUIRE(rays_o && rays_d && t && grid && G > 0 && mu3 && pts_c && ray_c && rowmap && count, "NULL pointer");
  Mu3o mu{{mu3[0], mu3[1], mu3[2]}};
  compact_samples_kernel<<<(unsigned)ceil_div(R, kCompactWarps), kCompactWarps * 32, 0, as_stream(stream)>>>(
      rays_o, rays_d, t, t_ray_stride, R, (int)S, grid, G, mu, sigma, pts_c, ray_c, rowmap, count);
  HBR_LAUNCH_CHECK();
  return HBR_OK;
}
