// Shared device/host helpers for libhbr_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/hbr.h"

namespace hbr {

// ---- error plumbing -------------------------------------------------------------------------------
char* err_buf();                       // thread-local, 512 bytes (abi.cu)
int fail(int code, const char* fmt, ...);

#define HBR_REQUIRE(cond, ...)                                  \
  do {                                                          \
    if (!(cond)) return ::hbr::fail(HBR_ERR_ARG, __VA_ARGS__);  \
  } while (0)

#define HBR_CUDA(expr)                                                                         \
  do {                                                                                         \
    cudaError_t e__ = (expr);                                                                  \
    if (e__ != cudaSuccess)                                                                    \
      return ::hbr::fail(HBR_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), \
                         __FILE__, __LINE__);                                                  \
  } while (0)

#define HBR_LAUNCH_CHECK() HBR_CUDA(cudaGetLastError())

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }
static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int64_t min64(int64_t a, int64_t b) { return a < b ? a : b; }
int sm_count();                        // cached per device (abi.cu)

// ---- hash-grid math (hash_encoding.py:41-55,153-158) ------------------------------------------------
constexpr uint32_t kPrimeY = 2654435761u;   // == (uint32)(int32)-1640531535, hash_encoding.py:24
constexpr uint32_t kPrimeZ = 805459861u;
constexpr long long kPrimeY64 = -1640531535LL;
constexpr long long kPrimeZ64 = 805459861LL;

struct HashGeom {                      // device copy of hbr_hash_geom (+ derived)
  float mu[3];
  float sigma;
  int L, F, E;
  uint32_t T;
  float scale[HBR_MAX_LEVELS];
};

static inline HashGeom to_device_geom(const hbr_hash_geom& g) {
  HashGeom d;
  for (int i = 0; i < 3; ++i) d.mu[i] = g.mu[i];
  d.sigma = g.sigma;
  d.L = g.L; d.F = g.F; d.E = g.E; d.T = g.T;
  for (int i = 0; i < HBR_MAX_LEVELS; ++i) d.scale[i] = i < g.L ? g.scale[i] : 0.f;
  return d;
}

static inline bool is_pow2(uint32_t v) { return v && !(v & (v - 1)); }

template <typename XT> __device__ __forceinline__ float load_coord(const XT* p);
template <> __device__ __forceinline__ float load_coord<float>(const float* p) { return __ldg(p); }
template <> __device__ __forceinline__ float load_coord<__half>(const __half* p) { return __half2float(__ldg(p)); }

// un_x = ((x - mu) / sigma) * N_l, every op rounded on its own (no FMA contraction, IEEE division);
// cell = trunc toward zero as int64 (.long()); frac = un_x - float(cell).   hash_encoding.py:154-158
__device__ __forceinline__ void cell_of(float x, float mu, float sigma, float scale, long long& cell, float& frac) {
  const float u = __fmul_rn(__fdiv_rn(__fsub_rn(x, mu), sigma), scale);
  cell = __float2ll_rz(u);
  frac = __fsub_rn(u, __ll2float_rn(cell));
}

// Index of lattice corner (cx,cy,cz).  POW2: low 32 bits only, identical to the reference's int64
// floor-mod for power-of-two T.  Otherwise the int64 arithmetic is done literally.
template <bool POW2>
__device__ __forceinline__ uint32_t hash_corner(long long cx, long long cy, long long cz, uint32_t T) {
  if (POW2) {
    const uint32_t h = (uint32_t)cx ^ ((uint32_t)cy * kPrimeY) ^ ((uint32_t)cz * kPrimeZ);
    return h & (T - 1);
  } else {
    const unsigned long long a = (unsigned long long)cx;
    const unsigned long long b = (unsigned long long)cy * (unsigned long long)kPrimeY64;
    const unsigned long long c = (unsigned long long)cz * (unsigned long long)kPrimeZ64;
    long long v = (long long)(a ^ b ^ c);
    long long r = v % (long long)T;
    if (r < 0) r += (long long)T;
    return (uint32_t)r;
  }
}

// Corner weights in the reference's order: prod over x,y,z of (bit ? frac : 1-frac)  (:142-143)
__device__ __forceinline__ void corner_weights(float fx, float fy, float fz, float w[8]) {
  const float wx[2] = {__fsub_rn(1.f, fx), fx};
  const float wy[2] = {__fsub_rn(1.f, fy), fy};
  const float wz[2] = {__fsub_rn(1.f, fz), fz};
#pragma unroll
  for (int c = 0; c < 8; ++c) w[c] = __fmul_rn(__fmul_rn(wx[c & 1], wy[(c >> 1) & 1]), wz[(c >> 2) & 1]);
}

template <bool POW2>
__device__ __forceinline__ void corner_indices(long long ix, long long iy, long long iz, uint32_t T, uint32_t idx[8]) {
  if (POW2) {
    const uint32_t a[2] = {(uint32_t)ix, (uint32_t)ix + 1u};
    const uint32_t b0 = (uint32_t)iy * kPrimeY, c0 = (uint32_t)iz * kPrimeZ;
    const uint32_t b[2] = {b0, b0 + kPrimeY};
    const uint32_t c[2] = {c0, c0 + kPrimeZ};
#pragma unroll
    for (int k = 0; k < 8; ++k) idx[k] = (a[k & 1] ^ b[(k >> 1) & 1] ^ c[(k >> 2) & 1]) & (T - 1);
  } else {
#pragma unroll
    for (int k = 0; k < 8; ++k) idx[k] = hash_corner<false>(ix + (k & 1), iy + ((k >> 1) & 1), iz + ((k >> 2) & 1), T);
  }
}

// ---- warp helpers -----------------------------------------------------------------------------------
constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}
__device__ __forceinline__ float warp_incl_scan(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float t = __shfl_up_sync(kFull, v, o);
    if (lane >= o) v += t;
  }
  return v;
}
__device__ __forceinline__ float warp_rincl_scan(float v, int lane) {   // suffix-inclusive
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float t = __shfl_down_sync(kFull, v, o);
    if (lane + o < 32) v += t;
  }
  return v;
}

}  // namespace hbr
