// MLP_3D on the 5th-generation tensor cores (tcgen05 + TMEM): 16-bit operands (bf16 or fp16), fp32 accumulation.
// This file is the implementation body; mlp_tc.cu includes it once per operand format (HBR_OP = tc::OpBf16 in namespace
// hbr::bf16, tc::OpF16 in hbr::f16) and holds the C entry points.  "bf16" in the comments below stands for either format.
// This is the field evaluation of the training step (train_hash2.py:218-226 runs it under autocast); the fp32
// CUDA-core version in mlp_simt.cu serves nerf2mesh and the 1e-5 parity tests.
//
// Structure (both kernels): one persistent CTA per SM = G "tile groups" of 128 threads + MMA-issuing warp(s).
//   * A tile group owns one 128-point tile at a time (thread = point = TMEM lane) and walks the layer chain:
//     it writes the next layer's bf16 A tile into shared memory, signals `full[g]` (mbarrier, 128 arrivals),
//     waits on `done[g]`, pulls its accumulator row out of TMEM with tcgen05.ld and applies the activation.
//   * An MMA warp waits on `full`, one elected lane issues the tcgen05.mma sequence of that layer in straight-line
//     code (A = activation tile, B = weight tile, D = the group's 64 TMEM columns) and commits it to `done`.
//     The chain of one tile is strictly serial (measured: ~130 cycles from commit to the waiter waking up, the rest
//     is epilogue work and hand-off), so throughput comes from G tiles in flight per SM.
//   * The bias rides on the tensor core too: the first MMA of every layer is ones[128x16] x biasB[Nx16]^T with
//     biasB = (bf16(b), bf16(b - bf16(b)), 0...), i.e. the accumulator starts at b (to 2^-17 relative), which removes
//     the bias loads/adds from the epilogues (the epilogue is tcgen05.ld -> cvt.relu.bf16x2 -> st.shared).
// Forward: one MMA warp per tile group (independent accumulators).
// Backward: recomputes the forward activations in shared memory (nothing but the features is re-read from HBM), then
// walks the layers in reverse: per layer one dgrad GEMM (dA = dZ W, B = the SAME weight tile read MN-major) and one
// weight-gradient GEMM (reduction over the 128 points, both operands read MN-major from tiles already in smem);
// the gradient accumulators stay resident in TMEM across all tiles of the CTA and are flushed once with atomics.
// Because all tiles of a CTA accumulate into the same TMEM columns, ONE warp issues every MMA of the CTA, visiting
// the groups in a fixed order.  Bias gradients: a GEMM against a ones column (64-wide layers), a 1.0 planted in a
// padding column of the colour-net input tile, or -- for the two 16-wide layers, whose gradient is accumulated
// transposed with M = 128 -- a ones column group placed right behind the activation tile.  dZ of a layer is written
// IN PLACE over the activation tile whose consumer GEMMs have completed.  ELU' and LeakyReLU' come from the saved
// forward output (elu'(x) = x > 0 ? 1 : elu(x) + 1), so the last forward layer is not recomputed.
// Layout conventions: tc_common.cuh.
#ifndef HBR_OP
#error "define HBR_OP (tc::OpBf16 | tc::OpF16) and HBR_OPNS before including mlp_tc_impl.cuh"
#endif

// fused backward (one tile group + scatter warps): d(feature) tiles in flight between the tile group and the scatter warps
#ifndef HBR_RING_SLOTS
#define HBR_RING_SLOTS 2
#endif

namespace hbr {
namespace HBR_OPNS {
using namespace tc;
using OP = HBR_OP;


constexpr int kTile = 128;            // points per tile == threads per tile group
constexpr int kLBO128 = kTile * 16;   // column-group stride of a 128-row tile (bytes)

template <int K0P, int KCP>
struct WOfs {                         // byte offsets of the six bf16 weight tiles [JP rows x KP cols]
  static constexpr int w0 = 0;
  static constexpr int w1 = w0 + 64 * K0P * 2;
  static constexpr int w2 = w1 + 64 * 64 * 2;
  static constexpr int w3 = w2 + 16 * 64 * 2;
  static constexpr int w4 = w3 + 64 * KCP * 2;
  static constexpr int w5 = w4 + 64 * 64 * 2;
  static constexpr int total = w5 + 16 * 64 * 2;
};

struct BOfs {                         // byte offsets of the six [JP x 16] bias tiles
  __host__ __device__ static constexpr int ofs(int i) {
    return i == 0 ? 0 : i == 1 ? 2048 : i == 2 ? 4096 : i == 3 ? 4608 : i == 4 ? 6656 : 8704;
  }
  static constexpr int total = 9216;
};

// fp32 (J,K) row-major weights -> bf16 canonical tile [JP rows x KP cols], zero padded; one 16-byte chunk per step
template <int K0P, int KCP>
__device__ __forceinline__ void stage_weights_bf16(const float* __restrict__ params, const MlpLayout& m, uint8_t* wsm,
                                                   uint8_t* bias_sm, uint8_t* ones16, float* bias_f32 = nullptr,
                                                   bool skip_weights = false, int tid = -1, int nthr = 0) {
  if (tid < 0) { tid = threadIdx.x; nthr = blockDim.x; }
  const int JP[6] = {64, 64, 16, 64, 64, 16};
  const int KP[6] = {K0P, 64, 64, KCP, 64, 64};
  const int wofs[6] = {WOfs<K0P, KCP>::w0, WOfs<K0P, KCP>::w1, WOfs<K0P, KCP>::w2, WOfs<K0P, KCP>::w3,
                       WOfs<K0P, KCP>::w4, WOfs<K0P, KCP>::w5};
  // weight tiles: one flat loop over all 16-byte chunks of the six tiles so that every thread has several chunks'
  // worth of global loads in flight (the whole image is staged in ~1 us instead of six dependent passes)
  if (!skip_weights) {
    constexpr int kChunks = WOfs<K0P, KCP>::total / 16;
    const int cbeg[7] = {0, wofs[1] / 16, wofs[2] / 16, wofs[3] / 16, wofs[4] / 16, wofs[5] / 16, kChunks};
#pragma unroll 3
    for (int c = tid; c < kChunks; c += nthr) {
      int i = 0;
#pragma unroll
      for (int t = 1; t < 6; ++t) i += c >= cbeg[t] ? 1 : 0;
      const int J = m.J[i], K = m.K[i];
      const int jp = (i == 2 || i == 5) ? 16 : 64;
      const int e = c - cbeg[i];
      const int cg = e / jp, j = e - cg * jp;                // consecutive threads -> consecutive rows: conflict-free STS.128
      const float* src = params + m.W[i] + j * K + cg * 8;
      float v[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) v[q] = (j < J && cg * 8 + q < K) ? __ldg(src + q) : 0.f;
      store_chunk<OP>(wsm + wofs[i], j, cg, jp, v);
    }
  }
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    const int J = m.J[i];
    if (bias_f32 == nullptr && bias_sm == nullptr) continue;          // weight tiles only
    if (bias_f32 != nullptr) {               // plain fp32 biases for kernels that add them in the epilogue
      for (int j = tid; j < 64; j += nthr) bias_f32[i * 64 + j] = j < J ? __ldg(params + m.b[i] + j) : 0.f;
      continue;
    }
    // bias tile [JP x 16] (K-major B operand of the bias MMA): col 0 = bf16(b), col 1 = bf16(b - bf16(b))
    uint8_t* bt = bias_sm + BOfs::ofs(i);
    for (int e = tid; e < JP[i] * 2; e += nthr) {
      const int cg = e / JP[i], j = e - cg * JP[i];
      float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (cg == 0 && j < J) {
        const float b = __ldg(params + m.b[i] + j);
        v[0] = OP::round(b);
        v[1] = b - v[0];
      }
      store_chunk<OP>(bt, j, cg, JP[i], v);
    }
  }
  if (bias_f32 != nullptr || ones16 == nullptr) return;
  // ones16 [128 x 16] (A operand of the bias MMA): cols 0,1 = 1
  for (int e = tid; e < kTile * 2; e += nthr) {
    const int cg = e / kTile, rr = e - cg * kTile;
    float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (cg == 0) v[0] = v[1] = 1.f;
    store_chunk<OP>(ones16, rr, cg, kTile, v);
  }
}

// Shared-memory operand addresses are carried in 16-byte units (a4 = byte address >> 4) so that a descriptor is one
// integer add on the low word (address + leading-byte-offset bits) next to a constant high word.
__device__ __forceinline__ uint64_t desc64(uint32_t a4, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  const uint32_t lo = a4 + ((lbo_bytes >> 4) << 16);
  const uint32_t hi = (sbo_bytes >> 4) | (1u << 14);                   // bit 46: tcgen05 descriptor version
  return ((uint64_t)hi << 32) | lo;
}
__device__ __forceinline__ uint32_t a4_of(const void* p) { return smem_u32(p) >> 4; }

// D[128 x N] (+)= A[128 x K] * W^T : A K-major activation tile, B K-major weight tile (forward)
__device__ __forceinline__ void issue_fwd(uint32_t tmem_d, uint32_t a4, uint32_t w4, int JP, int KP,
                                          bool accumulate = false) {
  const uint32_t idesc = make_idesc<OP>(128, JP, false, false);
#pragma unroll
  for (int kk = 0; kk < KP / 16; ++kk) {
    const uint64_t a = desc64(a4 + kk * (2 * kLBO128 / 16), kLBO128, 128);
    const uint64_t b = desc64(w4 + kk * 2 * JP, JP * 16, 128);
    mma_f16(tmem_d, a, b, idesc, accumulate || kk > 0);
  }
}
// same GEMM with the A operand in tensor memory: K packed two bf16 per column, 8 columns per K = 16 step
__device__ __forceinline__ void issue_fwd_ts(uint32_t tmem_d, uint32_t tmem_a, uint32_t w4, int JP, int KP, bool accumulate) {
  const uint32_t idesc = make_idesc<OP>(128, JP, false, false);
#pragma unroll
  for (int kk = 0; kk < KP / 16; ++kk) {
    const uint64_t b = desc64(w4 + kk * 2 * JP, JP * 16, 128);
    mma_f16_ts(tmem_d, tmem_a + kk * 8, b, idesc, accumulate || kk > 0);
  }
}
// D[128 x JP] = b (broadcast over rows) + A W^T: the bias enters as ones16 x biasB^T
__device__ __forceinline__ void issue_layer(uint32_t tmem_d, uint32_t a4, uint32_t w4, uint32_t ones16_4,
                                            uint32_t bias4, int JP, int KP) {
  issue_fwd(tmem_d, ones16_4, bias4, JP, 16, false);
  issue_fwd(tmem_d, a4, w4, JP, KP, true);
}
// D[128 x KP] = dZ[128 x JP] * W : A K-major dZ tile, B = weight tile [JP x KP] read MN-major (N = k, K = j)
__device__ __forceinline__ void issue_dgrad(uint32_t tmem_d, uint32_t dz4, uint32_t w4, int JP, int KP) {
  const uint32_t idesc = make_idesc<OP>(128, KP, false, true);
#pragma unroll
  for (int kk = 0; kk < JP / 16; ++kk) {
    const uint64_t a = desc64(dz4 + kk * (2 * kLBO128 / 16), kLBO128, 128);
    const uint64_t b = desc64(w4 + kk * 16, 128, JP * 16);
    mma_f16(tmem_d, a, b, idesc, kk > 0);
  }
}
// G[M x N] += At^T[M x 128] * Bt[128 x N]: both tiles [128 points x cols] read MN-major, reduction over points.
// (weight gradient: At = dZ, Bt = activation tile + its ones column group, M = 64; for the 16-wide layers the roles
//  swap and M = 128 so that the ones column group behind the activation tile adds the bias-gradient row: transposed)
__device__ __forceinline__ void issue_wgrad(uint32_t tmem_g, uint32_t a4, uint32_t b4, int N, bool accumulate,
                                            int M = 64) {
  const uint32_t idesc = make_idesc<OP>(M, N, true, true);
#pragma unroll
  for (int kk = 0; kk < kTile / 16; ++kk) {
    const uint64_t a = desc64(a4 + kk * 16, 128, kLBO128);
    const uint64_t b = desc64(b4 + kk * 16, 128, kLBO128);
    mma_f16(tmem_g, a, b, idesc, accumulate || kk > 0);
  }
}

__device__ __forceinline__ float elu1(float x) { return x > 0.f ? x : expm1f(x); }

#ifdef HBR_DEBUG_ENTRY
// ---------------------------------------------------------------------------------------------------------------
// debug / unit-test GEMM: exercises exactly the operand modes used below (tests/test_gpu_tc.py)
//   mode 0: D[128 x N] = A[128 x K] * B[N x K]^T      (K-major A, K-major B, M = 128)
//   mode 1: D[128 x N] = A[128 x K] * Bt[K x N]       (K-major A, MN-major B)
//   mode 2: D[64 x N]  = At[K x 64]^T * Bt[K x N]     (MN-major A, MN-major B, M = 64), K = 128
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) umma_debug_kernel(int mode, const float* __restrict__ A, const float* __restrict__ B,
                                                          float* __restrict__ D, int N, int K) {
  extern __shared__ __align__(128) uint8_t sm[];
  uint8_t* a_t = sm;                       // up to 128 x 128 bf16
  uint8_t* b_t = sm + 32768;
  __shared__ uint64_t mbar;
  __shared__ uint32_t tslot;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tmem_alloc<64>(&tslot);
  if (tid == 0) { mbar_init(&mbar, 1); fence_mbar_init(); }
  // fill tiles from row-major fp32
  const int a_rows = mode == 2 ? K : 128, a_cols = mode == 2 ? 64 : K;
  for (int e = tid; e < a_rows * a_cols; e += 128) {
    const int r = e / a_cols, c = e - r * a_cols;
    *reinterpret_cast<uint16_t*>(a_t + chunk_off(r, c >> 3, a_rows) + (c & 7) * 2) = OP::bits(A[e]);
  }
  const int b_rows = mode == 0 ? N : K, b_cols = mode == 0 ? K : N;
  for (int e = tid; e < b_rows * b_cols; e += 128) {
    const int r = e / b_cols, c = e - r * b_cols;
    *reinterpret_cast<uint16_t*>(b_t + chunk_off(r, c >> 3, b_rows) + (c & 7) * 2) = OP::bits(B[e]);
  }
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tbase = tslot;
  if (tid == 0) {
    const uint32_t at = a4_of(a_t), bt = a4_of(b_t);
    if (mode == 0) {
      issue_fwd(tbase, at, bt, N, K);
    } else if (mode == 1) {
      // here the "weight tile" is Bt [K rows(j) x N cols(k)]: dgrad convention JP = K, KP = N
      issue_dgrad(tbase, at, bt, K, N);
    } else {
      issue_wgrad(tbase, at, bt, N, false);
    }
    commit(&mbar);
  }
  mbar_wait(&mbar, 0);
  fence_after_sync();
  float v[64];
  const uint32_t taddr = tbase + ((uint32_t)(warp * 32) << 16);
  if (N > 48) tmem_ld<64>(taddr, v);
  else if (N > 32) tmem_ld<48>(taddr, v);
  else if (N > 16) tmem_ld<32>(taddr, v);
  else tmem_ld<16>(taddr, v);
  const int lane = tid & 31;
  if (mode == 2) {
    if (lane < 16) {
      const int row = warp * 16 + lane;
      for (int c = 0; c < N; ++c) D[row * N + c] = v[c];
    }
  } else {
    for (int c = 0; c < N; ++c) D[tid * N + c] = v[c];
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<64>(tbase);
}


// ---------------------------------------------------------------------------------------------------------------
// tensor-pipe micro-benchmark (debug): `reps` tcgen05.mma of shape M x N x 16 issued back to back by one thread,
// round-robin over `nacc` accumulators (nacc = 1: one dependent accumulation chain), then one commit.
// cycles[0] = issue start -> completion observed, cycles[1] = issue start -> last MMA issued.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) umma_bench_kernel(int M, int N, int reps, int nacc, int mn_major,
                                                          long long* __restrict__ cycles) {
  extern __shared__ __align__(128) uint8_t sm[];
  __shared__ uint64_t mbar;
  __shared__ uint32_t tslot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int e = tid; e < 65536 / 16; e += 128) reinterpret_cast<uint4*>(sm)[e] = make_uint4(0, 0, 0, 0);
  if (warp == 0) tmem_alloc<512>(&tslot);
  if (tid == 0) { mbar_init(&mbar, 1); fence_mbar_init(); }
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tbase = tslot;
  if (warp == 0) {
    long long t0 = 0, t1 = 0;
    if (elect_one()) {
      const uint32_t a4 = a4_of(sm), b4 = a4_of(sm + 32768);
      const uint32_t idesc = make_idesc<OP>(M, N, mn_major != 0, mn_major != 0);
      const uint64_t a = mn_major ? desc64(a4, 128, kLBO128) : desc64(a4, kLBO128, 128);
      const uint64_t b = mn_major ? desc64(b4, 128, kLBO128) : desc64(b4, N * 16, 128);
      t0 = clock64();
      for (int i = 0; i < reps; ++i) mma_f16(tbase + (i % nacc) * N, a, b, idesc, i >= nacc);
      t1 = clock64();
      commit(&mbar);
    }
    __syncwarp();
    mbar_wait(&mbar, 0);
    const long long t2 = clock64();
    t0 = __shfl_sync(kFull, t0, 0);   // elected lane is lane 0 in practice; good enough for a probe
    t1 = __shfl_sync(kFull, t1, 0);
    if (tid == 0) { cycles[0] = t2 - t0; cycles[1] = t1 - t0; }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tbase);
}


// Steady-state tensor-pipe cost of the three GEMM chains the MLP kernels issue (debug): `reps` back-to-back chains,
// straight-line issue code as in the kernels.  kind 0: forward layer (64x64, 4 K-steps, K-major x K-major);
// 1: dgrad (4 K-steps, B MN-major); 2: weight gradient M=64 N=72 (8 K-steps, both MN-major); 3: transposed weight
// gradient M=128 N=16 (8 K-steps).  cycles[0] = total, cycles[1] = issue only.
__global__ void __launch_bounds__(128) umma_chain_bench_kernel(int kind, int reps, int nacc, long long* __restrict__ cycles) {
  extern __shared__ __align__(128) uint8_t sm[];
  __shared__ uint64_t mbar;
  __shared__ uint32_t tslot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int e = tid; e < 98304 / 16; e += 128) reinterpret_cast<uint4*>(sm)[e] = make_uint4(0, 0, 0, 0);
  if (warp == 0) tmem_alloc<512>(&tslot);
  if (tid == 0) { mbar_init(&mbar, 1); fence_mbar_init(); }
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tbase = tslot;
  if (warp == 0) {
    long long t0 = 0, t1 = 0;
    const uint32_t a4 = a4_of(sm), b4 = a4_of(sm + 49152);
    if (elect_one()) {
      t0 = clock64();
      for (int i = 0; i < reps; ++i) {
        const uint32_t d = tbase + (i % nacc) * 80;
        if (kind == 0) issue_fwd(d, a4, b4, 64, 64);
        else if (kind == 1) issue_dgrad(d, a4, b4, 64, 64);
        else if (kind == 2) issue_wgrad(d, a4, b4, 72, true);
        else issue_wgrad(d, a4, b4, 16, true, 128);
      }
      t1 = clock64();
      commit(&mbar);
    }
    __syncwarp();
    mbar_wait(&mbar, 0);
    const long long t2 = clock64();
    t0 = __shfl_sync(kFull, t0, 0);
    t1 = __shfl_sync(kFull, t1, 0);
    if (tid == 0) { cycles[0] = t2 - t0; cycles[1] = t1 - t0; }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tbase);
}

#endif  // HBR_DEBUG_ENTRY

// ---------------------------------------------------------------------------------------------------------------
// tile-group side helpers (r = this thread's row in the tile, 0..127)
// ---------------------------------------------------------------------------------------------------------------
// One tile of features -> bf16 A tile.  Fast path (contiguous fp32 rows of exactly K0P floats): the tile is a contiguous
// 128*K0P*4-byte block, read with lane-contiguous float4 loads (4 lines per warp instruction instead of 32) and
// scattered into the canonical layout with 8-byte shared-memory stores.
template <int K0P, int NT = kTile>
__device__ __forceinline__ void load_features(const float* __restrict__ feat, long long stride, long long tile0, long long n,
                                              int in0, bool vec_ok, int gt, int r, int h, uint8_t* x0) {
  // NT threads share the tile: gt = thread index in the group; (r, h) = row and column half (NT == 2 * kTile) or h = 0
  if (vec_ok) {                                   // in0 == stride == K0P, 16-byte aligned base
    constexpr int kQ = K0P / 4;                   // float4 per row
    constexpr int kIt = kTile * kQ / NT;
    const float4* src = reinterpret_cast<const float4*>(feat + tile0 * K0P);
    float4 q[kIt];
#pragma unroll
    for (int it = 0; it < kIt; ++it) {
      const int idx = it * NT + gt;
      q[it] = tile0 + idx / kQ < n ? __ldg(src + idx) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int it = 0; it < kIt; ++it) {
      const int idx = it * NT + gt, row = idx / kQ, c4 = idx % kQ;
      uint2 o;
      o.x = OP::pack(q[it].x, q[it].y);
      o.y = OP::pack(q[it].z, q[it].w);
      *reinterpret_cast<uint2*>(x0 + chunk_off(row, c4 >> 1, kTile) + (c4 & 1) * 8) = o;
    }
  } else {
    const long long gp = tile0 + r;
    constexpr int kCgs = K0P / 8 * kTile / NT;    // column groups per thread
#pragma unroll
    for (int c = 0; c < kCgs; ++c) {
      const int cg = h * kCgs + c;
      float v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int k = cg * 8 + i;
        v[i] = (gp < n && k < in0) ? __ldg(feat + gp * stride + k) : 0.f;
      }
      store_chunk<OP>(x0, r, cg, kTile, v);
    }
  }
}

// ReLU on 64 accumulator columns (bias already inside) -> bf16 activations packed into 32 TMEM columns (the next
// layer's A operand): no shared-memory round trip, no proxy fence
__device__ __forceinline__ void relu_epilogue64_tmem(uint32_t taddr_acc, uint32_t taddr_a) {
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    float v[32];
    tmem_ld<32>(taddr_acc + half * 32, v);
    uint32_t p[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) p[q] = OP::pack_relu(v[2 * q], v[2 * q + 1]);
    tmem_st16(taddr_a + half * 16, p);
  }
  tmem_st_wait();
}

// ReLU on 64 accumulator columns (bias already inside) -> bf16 activation tile
__device__ __forceinline__ void relu_epilogue64(uint32_t taddr, int r, uint8_t* tile) {
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    float v[32];
    tmem_ld<32>(taddr + half * 32, v);
#pragma unroll
    for (int cg = 0; cg < 4; ++cg) {
      const float* p = v + cg * 8;
      uint4 o;
      o.x = OP::pack_relu(p[0], p[1]); o.y = OP::pack_relu(p[2], p[3]);
      o.z = OP::pack_relu(p[4], p[5]); o.w = OP::pack_relu(p[6], p[7]);
      *reinterpret_cast<uint4*>(tile + chunk_off(r, half * 4 + cg, kTile)) = o;
    }
  }
}

// dA (64 accumulator columns) * [activation > 0] -> bf16 dZ, in two steps for the backward chain: (1) dZ -> registers + tensor memory (the A operand of the next dgrad GEMM,
// all the chain needs); (2) later, once the weight-gradient GEMM that still reads the activation tile has finished, the
// in-place store of the dZ tile (the operand of the NEXT weight-gradient GEMM), off the critical chain.
__device__ __forceinline__ void masked_dz_to_tmem64(uint32_t taddr, int r, const uint8_t* tile, uint32_t taddr_a, uint4* o) {
  float v[64];
  tmem_ld<64>(taddr, v);                         // all four loads in flight, one wait
#pragma unroll
  for (int cg = 0; cg < 8; ++cg) {
    const uint4 h = *reinterpret_cast<const uint4*>(tile + chunk_off(r, cg, kTile));
    const float* p = v + cg * 8;
    uint4& q = o[cg];
    q.x = OP::mask_pos(OP::pack(p[0], p[1]), h.x); q.y = OP::mask_pos(OP::pack(p[2], p[3]), h.y);
    q.z = OP::mask_pos(OP::pack(p[4], p[5]), h.z); q.w = OP::mask_pos(OP::pack(p[6], p[7]), h.w);
  }
  tmem_st16(taddr_a, reinterpret_cast<const uint32_t*>(o));
  tmem_st16(taddr_a + 16, reinterpret_cast<const uint32_t*>(o) + 16);
  tmem_st_wait();
}
__device__ __forceinline__ void store_tile64(int r, uint8_t* tile, const uint4* o) {
#pragma unroll
  for (int cg = 0; cg < 8; ++cg) *reinterpret_cast<uint4*>(tile + chunk_off(r, cg, kTile)) = o[cg];
}

// colour-net input tile: [15 features | direction encoding | (optionally a 1.0 at column 15+dv) | 0 ...]
// (all rows of a ray read the same direction row: broadcast loads that hit L1 after the first touch)
template <int KCP, bool PLANT_ONE>
__device__ __forceinline__ void build_cin(const float* o16, const float* __restrict__ dirs, long long dir_row, int dv,
                                          bool valid, int r, uint8_t* cin) {
#pragma unroll
  for (int cg = 0; cg < KCP / 8; ++cg) {
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int k = cg * 8 + i;
      float x = 0.f;
      if (k < kFeat) x = o16[1 + k];                                    // feat_vec = dens_vec[:,1:]  (test_hash.py:64)
      else if (k < kFeat + dv) x = valid ? __ldg(dirs + dir_row * dv + (k - kFeat)) : 0.f;   // concat(viewdirs) (:66)
      else if (PLANT_ONE && k == kFeat + dv) x = 1.f;                   // meets a zero weight column; feeds the bias gradient
      v[i] = x;
    }
    store_chunk<OP>(cin, r, cg, kTile, v);
  }
}

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// group -> MMA warp: "my A tile is written (and my TMEM reads are finished)";  MMA warp -> group: commit on done
#define HBR_SIGNAL()          \
  do {                        \
    fence_async_smem();       \
    fence_before_sync();      \
    mbar_arrive(full);        \
  } while (0)
#define HBR_WAIT()            \
  do {                        \
    mbar_wait(done, dphase);  \
    dphase ^= 1;              \
    fence_after_sync();       \
  } while (0)

__device__ __forceinline__ long long tiles_of_slot(long long ntiles, long long slot, long long nslots) {
  return slot < ntiles ? (ntiles - slot + nslots - 1) / nslots : 0;
}

// ---------------------------------------------------------------------------------------------------------------
// optional caller-provided scratch (hbr_mlp_tc_scratch_bytes): [bf16 weight tiles | bias tiles | ones16 | fp32 biases |
// per-CTA gradient rows].  A small prep kernel builds the operand image once per call, the persistent CTAs then copy
// it with 16-byte loads instead of each converting all 14 227 parameters; the backward CTAs write their gradient image
// as one row each and a reduce kernel sums the rows (instead of 148-way contended atomics).
// ---------------------------------------------------------------------------------------------------------------
template <int K0P, int KCP>
struct Scratch {
  static constexpr int off_bias = WOfs<K0P, KCP>::total;
  static constexpr int off_ones16 = off_bias + BOfs::total;
  static constexpr int off_bias_f32 = off_ones16 + kTile * 16 * 2;
  static constexpr int off_grad = (off_bias_f32 + 6 * 64 * 4 + 255) & ~255;
  static constexpr int kMaxRows = 256;                                   // >= persistent grid size
  static constexpr int kRowFloats = 18432;                               // >= parameter count (<= 17 875 at in0 = 64, d_view = 49)
  static constexpr long long total = (long long)off_grad + (long long)kMaxRows * kRowFloats * 4;
};

constexpr int kPrepCtas = 11;                                           // 9 x 256 threads cover the <= 2304 weight chunks
template <int K0P, int KCP>
__global__ void __launch_bounds__(256) mlp_prep_kernel(const float* __restrict__ params, int in0, int dv, uint8_t* img) {
  using SC = Scratch<K0P, KCP>;
  const MlpLayout m = make_layout(in0, dv);
  if (blockIdx.x < kPrepCtas - 2) {            // weight tiles: one 16-byte chunk per thread
    constexpr int kChunks = WOfs<K0P, KCP>::total / 16;
    const int c = blockIdx.x * 256 + threadIdx.x;
    if (c < kChunks) stage_weights_bf16<K0P, KCP>(params, m, img, nullptr, nullptr, nullptr, false, c, 1 << 30);
  } else if (blockIdx.x == kPrepCtas - 2) {    // bias tiles + ones16 (weights skipped)
    stage_weights_bf16<K0P, KCP>(params, m, img, img + SC::off_bias, img + SC::off_ones16, nullptr, true);
  } else {                                      // fp32 biases
    stage_weights_bf16<K0P, KCP>(params, m, img, nullptr, nullptr, reinterpret_cast<float*>(img + SC::off_bias_f32), true);
  }
}

__device__ __forceinline__ void cp_async16(void* dst, const void* src, bool valid) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(smem_u32(dst)), "l"(src), "r"(valid ? 16 : 0) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

// Operand image (built once per launch by mlp_prep_kernel) -> shared memory.  Every 16-byte chunk is its own cp.async, so
// all of a thread's ~13 L2 reads are in flight together (a load/store loop pays the L2 latency once per iteration: measured
// ~10 us of set-up per CTA).  The caller waits (cp_async_wait_all) before the CTA-wide barrier.
__device__ __forceinline__ void copy_image_async(uint8_t* dst, const uint8_t* __restrict__ src, int bytes) {
  for (int e = threadIdx.x; e < bytes / 16; e += blockDim.x) cp_async16(dst + e * 16, src + e * 16, true);
  cp_async_commit();
}

constexpr int kReduceSlices = 8;
__global__ void __launch_bounds__(256) mlp_grad_reduce_kernel(const float* __restrict__ rows, int nrows, int row_floats,
                                                               int total, float* __restrict__ dparams) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  const int per = (nrows + kReduceSlices - 1) / kReduceSlices;
  const int r0 = blockIdx.y * per, r1 = min(nrows, r0 + per);
  float acc = 0.f;
#pragma unroll 4
  for (int r = r0; r < r1; ++r) acc += __ldg(rows + (size_t)r * row_floats + e);
  if (r1 > r0) atomicAdd(dparams + e, acc);
}

// ---------------------------------------------------------------------------------------------------------------
// fused hash-grid encoder (F = 2, L*F = 32, power-of-two T): the tile group gathers its own 128 x 32 feature tile
// straight into the A operand (forward) and scatters d(features) straight out of the accumulator (backward); the fp32
// feature / d(feature) tensors of the unfused path never exist.  Same arithmetic as hash_grid.cu (hash_encoding.py:146-170).
// ---------------------------------------------------------------------------------------------------------------

// this thread's point -> 32 bf16 features into row r of the canonical A tile (+ its row of feat16)
__device__ __forceinline__ void encode_row(const EncArgs& e, const HashGeom& g, long long gp, long long n, int r, uint8_t* x0) {
  float pt[3] = {0.f, 0.f, 0.f};
  if (gp < n) {
    pt[0] = __ldg(e.x + gp * 3 + 0); pt[1] = __ldg(e.x + gp * 3 + 1); pt[2] = __ldg(e.x + gp * 3 + 2);
  }
  uint32_t packed[16];
#pragma unroll 2
  for (int l = 0; l < 16; ++l) {
    const float s = g.scale[l];
    long long ix, iy, iz;
    float fx, fy, fz;
    cell_of(pt[0], g.mu[0], g.sigma, s, ix, fx);
    cell_of(pt[1], g.mu[1], g.sigma, s, iy, fy);
    cell_of(pt[2], g.mu[2], g.sigma, s, iz, fz);
    uint32_t idx[8];
    corner_indices<true>(ix, iy, iz, g.T, idx);
    const float* lvl = e.table + (size_t)l * g.T * 2;
    float v[8][2];
    if (!(ix & 1)) {             // even x: corners (x, x+1) share one aligned 16-byte slot
#pragma unroll
      for (int c = 0; c < 8; c += 2) {
        const float4 q = __ldg(reinterpret_cast<const float4*>(lvl) + (idx[c] >> 1));
        const bool odd = idx[c] & 1;
        v[c][0] = odd ? q.z : q.x;     v[c][1] = odd ? q.w : q.y;
        v[c + 1][0] = odd ? q.x : q.z; v[c + 1][1] = odd ? q.y : q.w;
      }
    } else {
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float2 q = __ldg(reinterpret_cast<const float2*>(lvl) + idx[c]);
        v[c][0] = q.x; v[c][1] = q.y;
      }
    }
    float w[8];
    corner_weights(fx, fy, fz, w);
    float a0 = __fmul_rn(v[0][0], w[0]), a1 = __fmul_rn(v[0][1], w[0]);
#pragma unroll
    for (int c = 1; c < 8; ++c) {
      a0 = __fadd_rn(a0, __fmul_rn(v[c][0], w[c]));
      a1 = __fadd_rn(a1, __fmul_rn(v[c][1], w[c]));
    }
    packed[l] = OP::pack(a0, a1);
  }
#pragma unroll
  for (int cg = 0; cg < 4; ++cg) {
    const uint4 q = make_uint4(packed[4 * cg], packed[4 * cg + 1], packed[4 * cg + 2], packed[4 * cg + 3]);
    *reinterpret_cast<uint4*>(x0 + chunk_off(r, cg, kTile)) = q;
    if (gp < n) reinterpret_cast<uint4*>(e.feat16 + gp * 32)[cg] = q;
  }
}

// backward recompute: this thread's saved bf16 feature row -> row r of the canonical A tile
__device__ __forceinline__ void load_feat16_row(const EncArgs& e, long long gp, long long n, int r, uint8_t* x0) {
#pragma unroll
  for (int cg = 0; cg < 4; ++cg) {
    const uint4 q = gp < n ? __ldg(reinterpret_cast<const uint4*>(e.feat16 + gp * 32) + cg) : make_uint4(0, 0, 0, 0);
    *reinterpret_cast<uint4*>(x0 + chunk_off(r, cg, kTile)) = q;
  }
}

// d(features) of this thread's point (32 fp32 values in registers) -> scatter-add into the table gradient.  A warp is 32
// consecutive points (consecutive samples of a ray): runs of lanes in the same cell are merged with shuffles and the
// run head issues the reductions, paired into red.global.add.v4.f32 where the two corners share a 16-byte slot.
__device__ __forceinline__ void scatter_row(const EncArgs& e, const HashGeom& g, const float pt[3], bool valid, int lane,
                                            const float* df) {
#pragma unroll 1
  for (int l = 0; l < 16; ++l) {
    const float s = g.scale[l];
    long long ix, iy, iz;
    float fx, fy, fz;
    cell_of(pt[0], g.mu[0], g.sigma, s, ix, fx);
    cell_of(pt[1], g.mu[1], g.sigma, s, iy, fy);
    cell_of(pt[2], g.mu[2], g.sigma, s, iz, fz);
    float w[8];
    corner_weights(fx, fy, fz, w);
    float g0 = 0.f, g1 = 0.f;
#pragma unroll
    for (int q = 0; q < 16; ++q)
      if (q == l) { g0 = df[2 * q]; g1 = df[2 * q + 1]; }
    float val[8][2];
#pragma unroll
    for (int c = 0; c < 8; ++c) { val[c][0] = w[c] * g0; val[c][1] = w[c] * g1; }
    const long long pix = __shfl_up_sync(kFull, ix, 1);
    const long long piy = __shfl_up_sync(kFull, iy, 1);
    const long long piz = __shfl_up_sync(kFull, iz, 1);
    const int pvalid = __shfl_up_sync(kFull, (int)valid, 1);
    const bool head = lane == 0 || !valid || !pvalid || pix != ix || piy != iy || piz != iz;
    const unsigned heads = __ballot_sync(kFull, head);
    if (heads != kFull) {
      const unsigned above = lane == 31 ? 0u : (heads & (0xfffffffeu << lane));
      const int end = above ? (__ffs(above) - 1) : 32;                    // first lane of the next run
      const int maxrun = __reduce_max_sync(kFull, head ? end - lane : 0);
      for (int d = 1; d < maxrun; d <<= 1) {
#pragma unroll
        for (int c = 0; c < 8; ++c)
#pragma unroll
          for (int f = 0; f < 2; ++f) {
            const float t = __shfl_down_sync(kFull, val[c][f], d);
            if (lane + d < end) val[c][f] += t;
          }
      }
    }
    if (head && valid) {
      uint32_t idx[8];
      corner_indices<true>(ix, iy, iz, g.T, idx);
      float* lvl = e.dtable + (size_t)l * g.T * 2;
      if (!(ix & 1)) {
#pragma unroll
        for (int c = 0; c < 8; c += 2) {
          const bool odd = idx[c] & 1;
          const float4 q = odd ? make_float4(val[c + 1][0], val[c + 1][1], val[c][0], val[c][1])
                               : make_float4(val[c][0], val[c][1], val[c + 1][0], val[c + 1][1]);
          atomicAdd(reinterpret_cast<float4*>(lvl) + (idx[c] >> 1), q);
        }
      } else {
#pragma unroll
        for (int c = 0; c < 8; ++c) atomicAdd(reinterpret_cast<float2*>(lvl) + idx[c], make_float2(val[c][0], val[c][1]));
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------------------------
template <int K0P, int KCP, int G>
struct FwdSmem {
  static constexpr int off_bias = WOfs<K0P, KCP>::total;
  static constexpr int off_ones16 = off_bias + BOfs::total;
  static constexpr int off_buf = off_ones16 + kTile * 16 * 2;
  static constexpr int buf_bytes = kTile * 64 * 2;                      // bf16 feature tile (layer 0's A operand)
  static constexpr int off_stage = off_buf + G * buf_bytes;             // fp32 staging of the NEXT tile's features (cp.async)
  static constexpr int stage_bytes = K0P == 32 ? kTile * K0P * 4 : 0;   // the wide variant loads features directly
  static constexpr int off_bar = off_stage + G * stage_bytes;
  // full[G], done[G], tslot (16 bytes), then the gather variant's xfull[G][2], xempty[G][2]
  static constexpr int off_xbar = off_bar + 2 * G * 8 + 16;
  static constexpr int total = off_xbar + 4 * G * 8;
  static_assert(total <= 232448, "shared memory budget exceeded");
};

// (x - mu) / sigma is level-independent: the scatter warps keep it per slice and finish cell_of per level
__device__ __forceinline__ void cell_of_norm(float u0, float scale, long long& cell, float& frac) {
  const float u = __fmul_rn(u0, scale);
  cell = __float2ll_rz(u);
  frac = __fsub_rn(u, __ll2float_rn(cell));
}

// ---- gather warps of the fused forward (GATH > 0): one (32-point slice, 4-level group) unit -----------------------------
// Lane = point: 4 levels x 8 corner gathers = 32 independent 8-byte loads in flight, interpolated exactly like
// hash_fwd_kernel (hash_grid.cu; products and sums rounded separately, hash_encoding.py:144), rounded to the operand
// format in pairs and stored as ONE 16-byte chunk of the canonical A tile (row r, column group lg) -- and as the same
// 16 bytes of the point's feat16 row for the backward recompute.
template <int NLV>   // levels per unit: 4 (one 16-byte chunk, 32 loads in flight) or 2 (half a chunk, 16 loads in flight)
__device__ __forceinline__ void gather_unit(const EncArgs& e, const HashGeom& g, long long n, long long tile, int slice, int lg,
                                            int lane, uint8_t* x0) {
  const int r = slice * 32 + lane;
  const long long gp = tile * kTile + r;
  float un[3] = {0.f, 0.f, 0.f};
  if (gp < n) {
    long long ray, smp;
    if (n <= 0xffffffffLL) {
      const unsigned q = (unsigned)gp / (unsigned)e.S;
      ray = q;
      smp = (long long)((unsigned)gp - q * (unsigned)e.S);
    } else {
      ray = gp / e.S;
      smp = gp - ray * e.S;
    }
    const float tt = __ldg(e.rt + ray * e.t_stride + smp);
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const float x = __fadd_rn(__ldg(e.ro + ray * 3 + a), __fmul_rn(__ldg(e.rd + ray * 3 + a), tt));   // vol_renderer.py:165
      un[a] = __fdiv_rn(__fsub_rn(x, g.mu[a]), g.sigma);
    }
  }
  float2 v[NLV][8];
  float fr[NLV][3];
#pragma unroll
  for (int q = 0; q < NLV; ++q) {
    const int l = lg * NLV + q;
    const float s = g.scale[l];
    long long ix, iy, iz;
    float fx, fy, fz;
    cell_of_norm(un[0], s, ix, fx);
    cell_of_norm(un[1], s, iy, fy);
    cell_of_norm(un[2], s, iz, fz);
    uint32_t idx[8];
    corner_indices<true>(ix, iy, iz, g.T, idx);
    const float2* lvl = reinterpret_cast<const float2*>(e.table) + (size_t)l * g.T;
#pragma unroll
    for (int c = 0; c < 8; ++c) v[q][c] = __ldg(lvl + idx[c]);
    fr[q][0] = fx; fr[q][1] = fy; fr[q][2] = fz;
  }
  uint32_t packed[NLV];
#pragma unroll
  for (int q = 0; q < NLV; ++q) {
    float w[8];
    corner_weights(fr[q][0], fr[q][1], fr[q][2], w);
    float a0 = __fmul_rn(v[q][0].x, w[0]), a1 = __fmul_rn(v[q][0].y, w[0]);
#pragma unroll
    for (int c = 1; c < 8; ++c) {
      a0 = __fadd_rn(a0, __fmul_rn(v[q][c].x, w[c]));
      a1 = __fadd_rn(a1, __fmul_rn(v[q][c].y, w[c]));
    }
    packed[q] = OP::pack(a0, a1);
  }
  if constexpr (NLV == 4) {
    const uint4 out4 = make_uint4(packed[0], packed[1], packed[2], packed[3]);
    *reinterpret_cast<uint4*>(x0 + chunk_off(r, lg, kTile)) = out4;
    if (gp < n) reinterpret_cast<uint4*>(e.feat16 + gp * 32)[lg] = out4;
  } else {
    const uint2 out2 = make_uint2(packed[0], packed[1]);
    *reinterpret_cast<uint2*>(x0 + chunk_off(r, lg >> 1, kTile) + (lg & 1) * 8) = out2;
    if (gp < n) reinterpret_cast<uint2*>(e.feat16 + gp * 32)[lg] = out2;
  }
}

// 16-byte asynchronous global -> shared copy; valid == false zero-fills the destination

// Feature pipeline of the forward kernel (contiguous fp32 rows of K0P floats): every thread copies the float4 elements
// idx = it*128 + r of a tile into the staging buffer with cp.async while the previous tile runs its layer chain, and at
// the start of the tile converts exactly the elements it copied itself (no cross-thread hazard) into the bf16 A tile.
template <int K0P, int NT = kTile>
__device__ __forceinline__ void stage_features_async(const float* __restrict__ feat, long long tile0, long long n, int gt,
                                                     uint8_t* stage) {
  constexpr int kQ = K0P / 4;
  const float4* src = reinterpret_cast<const float4*>(feat + tile0 * K0P);
#pragma unroll
  for (int it = 0; it < kTile * kQ / NT; ++it) {
    const int idx = it * NT + gt;
    const bool ok = tile0 + idx / kQ < n;
    cp_async16(stage + idx * 16, ok ? (const void*)(src + idx) : (const void*)feat, ok);
  }
  cp_async_commit();
}
template <int K0P, int NT = kTile>
__device__ __forceinline__ void convert_staged_features(const uint8_t* stage, int gt, uint8_t* x0) {
  constexpr int kQ = K0P / 4;
#pragma unroll
  for (int it = 0; it < kTile * kQ / NT; ++it) {
    const int idx = it * NT + gt, row = idx / kQ, c4 = idx % kQ;
    const float4 q = *reinterpret_cast<const float4*>(stage + idx * 16);
    uint2 o;
    o.x = OP::pack(q.x, q.y);
    o.y = OP::pack(q.z, q.w);
    *reinterpret_cast<uint2*>(x0 + chunk_off(row, c4 >> 1, kTile) + (c4 & 1) * 8) = o;
  }
}

// Features that arrive in the 16-bit operand format already (hbr_hash_encode_fwd_rays writes them so): rows of K0P
// 16-bit values, contiguous.  Every 16-byte chunk (8 columns of one row) is its own cp.async from the row-major global
// row straight to its place in the canonical no-swizzle tile: no registers, no conversion, no staging buffer.
// Thread (r, part) of NPART copies the chunks [part * K0P/8/NPART, ...) of row r.
template <int K0P, int NPART>
__device__ __forceinline__ void copy_feat16_async(const uint16_t* __restrict__ feat16, long long tile0, long long n, int r, int part,
                                                  uint8_t* dst_tile) {
  constexpr int kPer = K0P / 8 / NPART;
  const long long gp = tile0 + r;
  const bool ok = gp < n;
  const uint16_t* src = feat16 + (ok ? gp : 0) * K0P;
#pragma unroll
  for (int c = 0; c < kPer; ++c) {
    const int cg = part * kPer + c;
    cp_async16(dst_tile + chunk_off(r, cg, kTile), src + cg * 8, ok);
  }
  cp_async_commit();
}

// TRACE: clock64 stamps of group 0 / its MMA warp in CTA 0 (debug entry point hbr_debug_mlp_trace; compiled out otherwise)
// GATH > 0 (= hbr_field_fwd_rays_tc, the training step's forward on one GPU): the hash-grid gather runs INSIDE this kernel
// on GATH dedicated warps beside G = 2 tile groups.  A tile's 16 units (32-point slice x 4-level group, gather_unit) are
// dealt to the gather warps, which fill the group's feature tile -- two 8 KB buffers per group, xfull / xempty mbarriers --
// one tile ahead of the layer chain; the chain (latency-bound, tensor pipe ~40 %, no LSU traffic beyond layer 0's
// operand) hides behind the gather (bound by the L1TEX wavefront rate of the scattered 8-byte loads).  The sample
// positions come from the rays, the fp32 feature tensor never exists, the 16-bit features are written once for the
// backward recompute, and the weights are staged by the tile groups while the gather warps are already at work (no prep
// kernel in front).  Same arithmetic as hbr_hash_encode_fwd_rays + hbr_mlp_fwd_tc.
template <int K0P, int KCP, int G, bool TRACE = false, bool ENC = false, int GATH = 0>
__global__ void __launch_bounds__(G * kTile + GATH * 32, 1)
mlp_fwd_tc_kernel(const float* __restrict__ feat, long long feat_stride, const float* __restrict__ dirs, long long dir_group,
                  long long n_arg, const float* __restrict__ params, int in0, int dv, float* __restrict__ out,
                  const uint8_t* __restrict__ image, long long* __restrict__ trace, const EncArgs enc,
                  const __grid_constant__ HashGeom geom, int feat16, const unsigned long long* __restrict__ n_dev,
                  const int32_t* __restrict__ dir_rows) {
  // feat16 != 0: `feat` points at 16-bit features in the operand format (rows of K0P values, contiguous)
  // n_dev / dir_rows: compacted sample lists: the live point count is read from the device, and point
  // p takes the direction row dir_rows[p] (the ray it belongs to) instead of p / dir_group
  const long long n = n_dev != nullptr ? min(n_arg, (long long)__ldg(n_dev)) : n_arg;
  using SM = FwdSmem<K0P, KCP, G>;
  using WO = WOfs<K0P, KCP>;
  // TMEM per tile group: 64 accumulator columns + 32 columns holding the current layer input (bf16 pairs)
  constexpr int kGrpCols = 96;
  constexpr int kCols = G * kGrpCols <= 128 ? 128 : (G * kGrpCols <= 256 ? 256 : 512);
  static_assert(G * kGrpCols <= 512, "TMEM budget exceeded");
  extern __shared__ __align__(128) uint8_t sm[];
  const MlpLayout m = make_layout(in0, dv);
  uint8_t* wsm = sm;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + SM::off_bar);      // full[0..G), done[0..G)
  uint32_t* tslot = reinterpret_cast<uint32_t*>(bars + 2 * G);
  const int warp = __shfl_sync(kFull, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;

  uint64_t* xfull = reinterpret_cast<uint64_t*>(sm + SM::off_xbar);     // [G][2] feature buffer written by the gather warps
  uint64_t* xempty = xfull + 2 * G;                                     // [G][2] layer 0 has consumed the buffer
  constexpr int kXBuf = kTile * K0P * 2;                                // one feature buffer (the group's buf holds two)
  static_assert(GATH == 0 || (K0P == 32 && !ENC && 2 * kXBuf <= SM::buf_bytes), "gather variant: 32 features, two buffers");
  if (warp == 0) tmem_alloc<kCols>(tslot);
  if (threadIdx.x == 32) {
    for (int g = 0; g < G; ++g) mbar_init(bars + G + g, 1);             // done[g]: the group's MMAs have completed
    if (GATH > 0)
      for (int q = 0; q < 2 * G; ++q) {
        mbar_init(xfull + q, GATH * 32);
        mbar_init(xempty + q, 1);
      }
    fence_mbar_init();
  }
  if (GATH == 0) {
    if (image != nullptr) {
      copy_image_async(sm, image, SM::off_buf);                         // [weights | bias tiles | ones16], same layout
      cp_async_wait_all();
    } else {
      stage_weights_bf16<K0P, KCP>(params, m, wsm, sm + SM::off_bias, sm + SM::off_ones16);
    }
    fence_async_smem();
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tbase = *tslot;
  const long long ntiles = (n + kTile - 1) / kTile;
  const long long nslots = (long long)gridDim.x * G;

  // GATH <= 8: 16 units of 4 levels per tile; more gather warps: 32 units of 2 levels, and the registers are rebalanced
  // with setmaxnreg (launch at 80: the tile groups rise to 128, the gather warps drop to 56)
  constexpr int kNlv = GATH > 8 ? 2 : 4;
  constexpr int kUnits = 64 / kNlv;
  if (GATH > 0 && warp >= 4 * G) {
    // ===== gather warps =====
    if (GATH > 8) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(56));
    const int gw = warp - 4 * G;
    uint32_t epar = 0;                                                   // bit (2 g + slot): parity of xempty[g][slot]
#pragma unroll 1
    for (long long k = 0;; ++k) {
      bool any = false;
#pragma unroll 1
      for (int g = 0; g < G; ++g) {
        const long long tile = (long long)g * gridDim.x + blockIdx.x + k * nslots;
        if (tile >= ntiles) continue;
        any = true;
        const int slot = (int)(k & 1), q = 2 * g + slot;
        if (k >= 2) {                                                    // the buffer's previous tile has been consumed
          mbar_wait(xempty + q, (epar >> q) & 1u);
          epar ^= 1u << q;
        }
        uint8_t* x0 = sm + SM::off_buf + g * SM::buf_bytes + slot * kXBuf;
#pragma unroll 1
        for (int u = gw; u < kUnits; u += GATH) gather_unit<kNlv>(enc, geom, n, tile, u & 3, u >> 2, lane, x0);
        fence_async_smem();                                              // generic-proxy stores -> visible to the tensor core
        mbar_arrive(xfull + q);
      }
      if (!any) break;
    }
  } else {
    if (GATH > 8) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(128));
    if (GATH > 0) {
      // the tile groups stage the weights themselves (threads 0 .. G*128-1) while the gather warps already work
      if (image != nullptr) {
        for (int e = threadIdx.x; e < SM::off_buf / 16; e += G * kTile) cp_async16(sm + e * 16, image + e * 16, true);
        cp_async_commit();
        cp_async_wait_all();
      } else {
        stage_weights_bf16<K0P, KCP>(params, m, wsm, sm + SM::off_bias, sm + SM::off_ones16, nullptr, false, threadIdx.x,
                                     G * kTile);
      }
      fence_async_smem();
      asm volatile("bar.sync 7, %0;" ::"n"(G * kTile) : "memory");
    }
    // ===== tile group g (warps 4g .. 4g+3): 128 threads = the 128 points of a tile; the group's first warp also issues
    // the group's MMAs (no separate issuer warps: 512 threads keep 128 registers each, and the hand-off is one named
    // barrier instead of an mbarrier round trip through another warp) =====
    const int g = warp >> 2;
    const int r = threadIdx.x & (kTile - 1);
    const bool issuer = (warp & 3) == 0;
    uint8_t* buf = sm + SM::off_buf + g * SM::buf_bytes;
    uint64_t* done = bars + G + g;
    const uint32_t tgrp = tbase + g * kGrpCols;                          // D: [tgrp, tgrp+64), A: [tgrp+64, tgrp+96)
    const uint32_t taddr = tgrp + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t taddr_a = taddr + 64;
    const uint32_t wa = a4_of(wsm), ba = a4_of(sm + SM::off_bias), o16a = a4_of(sm + SM::off_ones16), xa = a4_of(buf);
    uint32_t dphase = 0;
    const bool vec_ok = in0 == K0P && feat_stride == K0P && ((uintptr_t)feat & 15) == 0;
    int tgi = 0, tmi = 0;
    (void)tgi; (void)tmi;
    int tgk = 0;                                 // gather variant: tiles done by this group; xpar: parity of xfull[g][slot]
    uint32_t xpar = 0;
    (void)tgk; (void)xpar;
    uint8_t* stage = sm + SM::off_stage + g * SM::stage_bytes;
    const bool f16in = !ENC && GATH == 0 && feat16 != 0;
    const uint16_t* featq = reinterpret_cast<const uint16_t*>(feat);
    const bool staged = !ENC && GATH == 0 && !f16in && vec_ok && SM::stage_bytes > 0;
    if (staged && (long long)g * gridDim.x + blockIdx.x < ntiles)
      stage_features_async<K0P>(feat, ((long long)g * gridDim.x + blockIdx.x) * kTile, n, r, stage);
    if (f16in && (long long)g * gridDim.x + blockIdx.x < ntiles)
      copy_feat16_async<K0P, 1>(featq, ((long long)g * gridDim.x + blockIdx.x) * kTile, n, r, 0, buf);
#define TR()                                                                                  \
  do {                                                                                        \
    if (TRACE && blockIdx.x == 0 && threadIdx.x == 0 && tgi < 1000) trace[tgi++] = clock64(); \
  } while (0)
    // operands written (SMEM: shared-memory tile, needs the generic->async proxy fence; otherwise tensor memory only):
    // group barrier, the first warp's elected lane issues the layer's MMAs and commits, everybody waits for completion
#define HBR_LAYER(SMEM, BODY)                                                                           \
  do {                                                                                                  \
    TR();                                                                                               \
    if (SMEM) fence_async_smem();                                                                       \
    fence_before_sync();                                                                                \
    asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory");                                          \
    TR();                                                                                               \
    if (issuer) {                                                                                       \
      if (TRACE && g == 0 && blockIdx.x == 0 && lane == 0 && tmi < 500) trace[1024 + tmi++] = clock64(); \
      fence_after_sync();                                                                               \
      if (elect_one()) {                                                                                \
        BODY;                                                                                           \
        commit(done);                                                                                   \
      }                                                                                                 \
      __syncwarp();                                                                                     \
      if (TRACE && g == 0 && blockIdx.x == 0 && lane == 0 && tmi < 500) trace[1024 + tmi++] = clock64(); \
    }                                                                                                   \
    mbar_wait(done, dphase);                                                                            \
    dphase ^= 1;                                                                                        \
    fence_after_sync();                                                                                 \
    TR();                                                                                               \
  } while (0)
#define HBR_TS_LAYER(I, W, JP, KP)                                      \
  HBR_LAYER(false, {                                                    \
    issue_fwd(tgrp, o16a, ba + BOfs::ofs(I) / 16, JP, 16, false);       \
    issue_fwd_ts(tgrp, tgrp + 64, wa + WO::W / 16, JP, KP, true);       \
  })
    for (long long tile = (long long)g * gridDim.x + blockIdx.x; tile < ntiles; tile += nslots) {
      const long long gp = tile * kTile + r;
      const bool valid = gp < n;
      const long long dir_row = valid ? (dir_rows != nullptr ? (long long)__ldg(dir_rows + gp) : gp / dir_group) : 0;
      if (valid && lane == 0) prefetch_l1(dirs + dir_row * dv);
      TR();
      uint32_t xofs = 0;                         // gather variant: which of the group's two feature buffers holds this tile
      if (GATH > 0) {
        const int slot = tgk & 1;
        mbar_wait(xfull + 2 * g + slot, (xpar >> slot) & 1u);
        xpar ^= 1u << slot;
        xofs = slot * (kXBuf / 16);
      } else if (ENC) {
        encode_row(enc, geom, gp, n, r, buf);
      } else if (f16in) {
        cp_async_wait_all();                     // this tile's rows: copied while the previous tile ran layers 1..5
      } else if (staged) {
        cp_async_wait_all();
        TR();
        convert_staged_features<K0P>(stage, r, buf);
        TR();
        if (tile + nslots < ntiles) stage_features_async<K0P>(feat, (tile + nslots) * kTile, n, r, stage);
      } else {
        load_features<K0P>(feat, feat_stride, tile * kTile, n, in0, vec_ok, r, r, 0, buf);
      }
      // layer 0 reads the feature tile from shared memory; the later layers read their input from tensor memory
      HBR_LAYER(true, issue_layer(tgrp, xa + xofs, wa + WO::w0 / 16, o16a, ba + BOfs::ofs(0) / 16, 64, K0P));
      // layer 0 has consumed the tile: the next tile's rows travel into it during the rest of this chain
      if (GATH > 0) {
        if (r == 0) mbar_arrive(xempty + 2 * g + (tgk & 1));             // hand the buffer back to the gather warps
        ++tgk;
      }
      if (f16in && tile + nslots < ntiles) copy_feat16_async<K0P, 1>(featq, (tile + nslots) * kTile, n, r, 0, buf);
      relu_epilogue64_tmem(taddr, taddr_a);
      HBR_TS_LAYER(1, w1, 64, 64);
      relu_epilogue64_tmem(taddr, taddr_a);
      HBR_TS_LAYER(2, w2, 16, 64);
      float o16[16];
      tmem_ld<16>(taddr, o16);
      const float density = o16[0] > 0.f ? o16[0] : 0.01f * o16[0];     // LeakyReLU (test_hash.py:62)
      {
        // colour-net input [15 features | direction encoding | 0...] packed straight into tensor memory
        uint32_t p[KCP / 2];
#pragma unroll
        for (int q = 0; q < KCP / 2; ++q) {
          float e[2];
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const int k = 2 * q + j;
            float x = 0.f;
            if (k < kFeat) x = o16[1 + k];                              // feat_vec = dens_vec[:,1:]  (test_hash.py:64)
            else if (k < kFeat + dv) x = valid ? __ldg(dirs + dir_row * dv + (k - kFeat)) : 0.f;   // concat(viewdirs) (:66)
            e[j] = x;
          }
          p[q] = OP::pack(e[0], e[1]);
        }
        tmem_st16(taddr_a, p);
        if (KCP == 48) tmem_st8(taddr_a + 16, p + 16);
        else tmem_st16(taddr_a + 16, p + 16);
        tmem_st_wait();
      }
      HBR_TS_LAYER(3, w3, 64, KCP);
      relu_epilogue64_tmem(taddr, taddr_a);
      HBR_TS_LAYER(4, w4, 64, 64);
      relu_epilogue64_tmem(taddr, taddr_a);
      HBR_TS_LAYER(5, w5, 16, 64);
      float c16[16];
      tmem_ld<16>(taddr, c16);
      if (valid) {
        float4 o;
        o.x = elu1(c16[0]);                                             // ELU (test_hash.py:67)
        o.y = elu1(c16[1]);
        o.z = elu1(c16[2]);
        o.w = density;
        *reinterpret_cast<float4*>(out + gp * 4) = o;                   // (rgb, sigma), test_hash.py:69
      }
    }
#undef HBR_TS_LAYER
#undef HBR_LAYER
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<kCols>(tbase);
}

// ---------------------------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------------------------
constexpr int kCg = kTile * 16;                                        // bytes of one 8-column group of a 128-row tile

// bias + ReLU on 64 accumulator columns -> bf16 activations in registers (o[8], one 16-byte chunk per column group) and,
// with a valid taddr_a, in tensor memory as the next layer's A operand -- all the chain needs.  The caller stores o[] to
// the activation tile in shared memory (the weight-gradient operand) behind the next GEMM's issue (store_tile64).
// (The backward kernel adds the bias here: its two tile groups leave the CUDA cores mostly idle, while every MMA saved
// shortens the issue-bound critical path.)
__device__ __forceinline__ void relu_bias_to_tmem64(uint32_t taddr, const float* bias, uint32_t taddr_a, uint4* o) {
  float v[64];
  tmem_ld<64>(taddr, v);                         // all four loads in flight, one wait
#pragma unroll
  for (int cg = 0; cg < 8; ++cg) {
    const float4 b0 = *reinterpret_cast<const float4*>(bias + cg * 8);
    const float4 b1 = *reinterpret_cast<const float4*>(bias + cg * 8 + 4);
    const float* p = v + cg * 8;
    o[cg].x = OP::pack_relu(p[0] + b0.x, p[1] + b0.y); o[cg].y = OP::pack_relu(p[2] + b0.z, p[3] + b0.w);
    o[cg].z = OP::pack_relu(p[4] + b1.x, p[5] + b1.y); o[cg].w = OP::pack_relu(p[6] + b1.z, p[7] + b1.w);
  }
  if (taddr_a != 0xffffffffu) {
    tmem_st16(taddr_a, reinterpret_cast<const uint32_t*>(o));
    tmem_st16(taddr_a + 16, reinterpret_cast<const uint32_t*>(o) + 16);
    tmem_st_wait();
  }
}

// 16-wide dZ (the two 16-output layers): bf16 into the shared-memory tile (weight-gradient operand) and into tensor
// memory (A operand of the dgrad GEMM)
__device__ __forceinline__ void store_dz16_both(const float* dz16, int r, uint8_t* dzs, uint32_t taddr_a) {
  uint32_t p[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) p[q] = OP::pack(dz16[2 * q], dz16[2 * q + 1]);
  *reinterpret_cast<uint4*>(dzs + chunk_off(r, 0, kTile)) = make_uint4(p[0], p[1], p[2], p[3]);
  *reinterpret_cast<uint4*>(dzs + chunk_off(r, 1, kTile)) = make_uint4(p[4], p[5], p[6], p[7]);
  tmem_st8(taddr_a, p);
  tmem_st_wait();
}

// dgrad with the dZ operand in tensor memory: D[128 x KP] = dZ[128 x JP] * W, B = weight tile [JP x KP] read MN-major
__device__ __forceinline__ void issue_dgrad_ts(uint32_t tmem_d, uint32_t tmem_a, uint32_t w4, int JP, int KP) {
  const uint32_t idesc = make_idesc<OP>(128, KP, false, true);
#pragma unroll
  for (int kk = 0; kk < JP / 16; ++kk) {
    const uint64_t b = desc64(w4 + kk * 16, 128, JP * 16);
    mma_f16_ts(tmem_d, tmem_a + kk * 8, b, idesc, kk > 0);
  }
}

template <int K0P, int KCP, int G>
struct BwdSmem {
  // KCP == 48 (15 + d_view <= 40): column group 5 of the colour-net input tile is pure padding and doubles as the
  // first half of the 16-wide dZ tile (its partner column group sits right behind the tile)
  static constexpr bool kAliasDzs = KCP == 48;
  static constexpr int off_bias = WOfs<K0P, KCP>::total;                // 6 x 64 fp32
  static constexpr int off_grp = off_bias + 6 * 64 * 4;
  // per group: every activation tile is followed by a column group of 1.0 (bias gradient through the weight-gradient
  // GEMM); h2 and c2 additionally by >= 7 further column groups of finite data (their M = 128 operand reads them)
  static constexpr int h2 = 0;
  static constexpr int c2 = h2 + 9 * kCg;
  static constexpr int x0 = c2 + 9 * kCg;
  static constexpr int h1 = x0 + kTile * K0P * 2 + kCg;
  static constexpr int c1 = h1 + 9 * kCg;
  static constexpr int cin = c1 + 9 * kCg;
  static constexpr int dzs = kAliasDzs ? cin + 5 * kCg : cin + kTile * KCP * 2;
  static constexpr int grp_bytes = dzs + 2 * kCg;
  // scatter variant with one tile group (SCAT = 11): two 16 KB slots for the d(feature) tiles on their way to the scatter
  // warps (with two groups the slot is the group's own dead h2 region)
  static constexpr int ring_slot = kTile * 32 * 4;
  static constexpr int ring_slots = HBR_RING_SLOTS;                     // one group: d(feature) tiles in flight towards the scatter warps
  static constexpr int off_ring = off_grp + G * grp_bytes;
  static constexpr int off_bar = off_ring + (G == 1 && K0P == 32 ? ring_slots * ring_slot : 0);   // full[G], doneA[G], doneB[G], startB[G], dfull[slots], dempty[slots]
  // normalised sample positions of the scatter warps' current / next tile: [2][128 points][3] fp32
  static constexpr int off_pos = off_bar + (4 * G + 2 * ring_slots) * 8 + 16;
  static constexpr int total = off_pos + 2 * kTile * 3 * 4;
  static_assert(total <= 232448, "shared memory budget exceeded");
};

// TMEM columns: [0, 128) work accumulators of the (up to two) groups; then the gradient accumulators:
//   layers 0,1,3,4 (M = 64): G[j][k], rows = output neuron; KP columns + 8 bias-gradient columns from the ones group
//     (layer 3 with KCP == 48 has its bias gradient in column 15 + d_view: the 1.0 planted in the input tile);
//   layers 2,5 (M = 128): transposed G^T[k][j], 16 columns, rows 0..63 = input index, row 64 = bias gradient.
template <int K0P, int KCP, int WORK = 128>
struct BwdTmem {
  static constexpr bool kCinOne = KCP == 48;
  static constexpr int n0 = K0P + 8, n1 = 72, n3 = KCP + (kCinOne ? 0 : 8), n4 = 72;
  static constexpr int g0 = WORK;                                       // columns below: the groups' work accumulators
  static constexpr int g1 = g0 + n0;
  static constexpr int g2 = g1 + n1;
  static constexpr int g3 = g2 + 16;
  static constexpr int g4 = g3 + n3;
  static constexpr int g5 = g4 + n4;
  static constexpr int end = g5 + 16;
  static_assert(end <= 512 && g4 + 80 <= 512, "TMEM budget exceeded");
};

// ---- flush the gradient accumulators: TMEM -> registers -> the CTA's row of the scratch (summed over CTAs by
//      mlp_grad_reduce_kernel), or atomics into dparams when there is no scratch.  Straight from registers: accumulator
//      row q (one lane) owns the K contiguous floats of dW[q][:], written as 16-byte stores.
//      Written as ROLLED loops over 8-column TMEM loads on purpose: this code runs once per CTA, so its cost is its
//      instruction fetch -- the fully unrolled version (~20 KB of SASS, cold in the instruction cache) measured ~31 k
//      cycles per CTA, an earlier one staging through shared memory ~50 k (plus 16-way bank conflicts).
//      (Also measured: the same rolled loop staging through a padded, conflict-free shared-memory image with
//      lane-contiguous stores behind a named barrier: ~21 k cycles against ~16.5 k for the direct stores.)
//      The accumulators are split over the CTA's tile-group warps (a warp reads the TMEM lane quarter warp % 4).
template <int K0P, int KCP, int WORK>
__device__ __noinline__ void flush_gradients(uint32_t tbase, int warp, int lane, int ngroups, const MlpLayout& m,
                                             bool has_tiles, float* __restrict__ dparams, float* __restrict__ grad_rows, float ginv,
                                             int nthreads) {
  using TM = BwdTmem<K0P, KCP, WORK>;
  constexpr bool kCinOne = TM::kCinOne;
  if (dparams == nullptr) return;
  float* row = grad_rows != nullptr ? grad_rows + (size_t)blockIdx.x * Scratch<K0P, KCP>::kRowFloats : nullptr;
  if (!has_tiles) {
    if (row != nullptr)
      for (int e = threadIdx.x; e < m.total; e += nthreads) row[e] = 0.f;      // called by the first `nthreads` threads
    return;
  }
  if (warp >= 4 * ngroups) return;
  const int wq = warp & 3, unit = warp >> 2, units = ngroups;     // the six accumulators are dealt to the groups' warp quartets
  const uint32_t trow = tbase + ((uint32_t)(wq * 32) << 16);
  auto put = [&](int idx, float v) {
    if (row != nullptr) row[idx] = v * ginv;
    else atomicAdd(dparams + idx, v * ginv);
  };
  // layers 0,1,3,4 -- M = 64 layout: accumulator row q (= output neuron) lives in lane (q % 16) + 32 * (q / 16)
  const int q = wq * 16 + lane;                  // meaningful for lane < 16
#pragma unroll 1
  for (int t = 0; t < 4; ++t) {
    if ((t % units) != unit) continue;           // warp-uniform
    const int i = t < 2 ? t : t + 1;
    const int gcol = t == 0 ? TM::g0 : t == 1 ? TM::g1 : t == 2 ? TM::g3 : TM::g4;
    const int ncol = t == 0 ? TM::n0 : t == 1 ? TM::n1 : t == 2 ? TM::n3 : TM::n4;      // multiples of 8
    const int Ki = m.K[i], Ji = m.J[i], Wi = m.W[i], bi = m.b[i];
    // bias gradient: first column of the ones group, or the planted 1.0 column of the colour-net input
    const int bias_col = (i == 3 && kCinOne) ? Ki : (i == 0 ? K0P : (i == 3 ? KCP : 64));
    const bool mine = lane < 16 && q < Ji;
    const bool vec = row != nullptr && (Ki & 3) == 0 && (Wi & 3) == 0;
    const int base = Wi + q * Ki;
#pragma unroll 1
    for (int c = 0; c < ncol; c += 8) {
      float v[8];
      tmem_ld8_wait(trow + gcol + c, v);         // executed by the whole warp (.sync.aligned)
      if (mine) {
        if (vec) {
          if (c < Ki) *reinterpret_cast<float4*>(row + base + c) = make_float4(v[0] * ginv, v[1] * ginv, v[2] * ginv, v[3] * ginv);
          if (c + 4 < Ki) *reinterpret_cast<float4*>(row + base + c + 4) = make_float4(v[4] * ginv, v[5] * ginv, v[6] * ginv, v[7] * ginv);
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (c + j < Ki) put(base + c + j, v[j]);
        }
        if ((bias_col & ~7) == c) {
          float bv = v[0];
#pragma unroll
          for (int j = 1; j < 8; ++j)
            if ((bias_col & 7) == j) bv = v[j];
          put(bi + q, bv);
        }
      }
    }
  }
  // layers 2,5 -- M = 128 layout: accumulator row = lane; rows 0..63 = input index k, row 64 = bias gradient
  const int r = wq * 32 + lane;
#pragma unroll 1
  for (int t = 0; t < 2; ++t) {
    if (((t + 2) % units) != unit) continue;
    const int i = t == 0 ? 2 : 5;
    const int Ji = m.J[i], Wi = m.W[i], bi = m.b[i];
    float gacc[16];
    tmem_ld<16>(trow + (t == 0 ? TM::g2 : TM::g5), gacc);
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      if (j < Ji) {
        if (r < 64) put(Wi + j * 64 + r, gacc[j]);                       // lanes write consecutive floats: coalesced
        else if (r == 64) put(bi + j, gacc[j]);
      }
    }
  }
}

// SCAT > 0 (= 7): the hash-grid scatter-add of the tile's d(features) (hash_grid.cu: hash_bwd_kernel) runs INSIDE this
// kernel on SCAT dedicated warps: a tile group leaves its 128 x 32 fp32 d(feature) tile in shared memory (the dead h2
// region, [level][point] float2), the scatter warps pull their values into registers, hand the buffer back and issue the
// run-merged red.global.add while the groups walk the next tiles' layer chains.  The chain is latency-bound (tensor pipe
// ~25 % busy, LSU idle), the scatter is bound by the red.global path: side by side on one SM they overlap instead of
// adding up, and the (N,32) fp32 d(feature) tensor is neither written nor re-read.
// Work split of the scatter warps: a tile is 64 units (32-point slice s, level l); warp w of 7 takes the units
// u = w, w + 7, ... (u = 4 l + s) -- a unit is exactly what one warp of hash_bwd_kernel does for one level.
// Registers: the register file is per SM sub-partition (4 warps of this CTA on each: 2 tile-group warps + 2 others), so
// the CTA is 16 warps = 512 threads launched at 128 registers, and setmaxnreg moves them: warps 8..15 (7 scatter warps
// + the weight-gradient issuer: setmaxnreg works on whole warpgroups) drop to kScatRegs, the tile groups rise to
// kGroupRegs; per sub-partition 2 x 168 + 2 x 88 = 512 = what the launch allocated.
constexpr int kGroupRegs = 168;
// G = 2, SCAT = 7: 2 x 168 + 2 x 88 = 512 per sub-partition;  G = 1, SCAT = 11: 168 + 3 x 112 = 504
__host__ __device__ constexpr int scat_regs(int G) { return G == 2 ? 88 : 112; }


template <int K0P, int KCP, int G, bool TRACE = false, bool ENC = false, int SCAT = 0>
__global__ void __launch_bounds__(G * kTile + SCAT * 32 + 32, 1)
mlp_bwd_tc_kernel(const float* __restrict__ feat, long long feat_stride, const float* __restrict__ dirs, long long dir_group,
                  long long n_arg, const float* __restrict__ params, int in0, int dv, const float* __restrict__ out,
                  const float* __restrict__ dout, float* __restrict__ dfeat, long long dfeat_stride,
                  float* __restrict__ ddirs, float* __restrict__ dparams, const uint8_t* __restrict__ image,
                  float* __restrict__ grad_rows, long long* __restrict__ trace, const EncArgs enc,
                  const __grid_constant__ HashGeom geom, float gscale, int feat16, const unsigned long long* __restrict__ n_dev,
                  const int32_t* __restrict__ dir_rows) {
  // feat16 != 0: `feat` points at 16-bit features in the operand format (rows of K0P values, contiguous)
  // n_dev / dir_rows: compacted sample lists, as in the forward kernel
  const long long n = n_dev != nullptr ? min(n_arg, (long long)__ldg(n_dev)) : n_arg;
  // gscale: power of two applied to the upstream gradient before it is rounded to the 16-bit operand format and divided
  // out of every result (fp16 has 5 exponent bits: unscaled gradients of a mean-reduced loss underflow; the reference
  // relies on GradScaler for the same reason, train_hash2.py:156,226).  1.0 = off.
  const float ginv = 1.f / gscale;
  using SM = BwdSmem<K0P, KCP, G>;
  using WO = WOfs<K0P, KCP>;
  // TMEM per tile group: 64 work-accumulator columns + 32 columns holding the A operand (bf16 pairs) of the group's next
  // forward / dgrad GEMM; the gradient accumulators start behind the groups
  constexpr int kGrpCols = 96;
  using TM = BwdTmem<K0P, KCP, 2 * kGrpCols>;
  constexpr bool kCinOne = TM::kCinOne;
  static_assert(G >= 1 && G <= 2, "two work accumulators");
  extern __shared__ __align__(128) uint8_t sm[];
  const MlpLayout m = make_layout(in0, dv);
  uint8_t* wsm = sm;
  float* bias = reinterpret_cast<float*>(sm + SM::off_bias);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + SM::off_bar);
  constexpr int kSlots = SM::ring_slots;
  uint32_t* tslot = reinterpret_cast<uint32_t*>(bars + 4 * G + 2 * kSlots);
  uint64_t* dfull = bars + 4 * G;                                        // [kSlots] d(feature) slot written (128 arrivals)
  uint64_t* dempty = bars + 4 * G + kSlots;                              // [kSlots] slot read by every scatter lane
  const int warp = __shfl_sync(kFull, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  constexpr int kIssuerWarp = 4 * G + SCAT;
  if (TRACE && blockIdx.x == 0 && threadIdx.x == 0) trace[2000] = clock64();

  if (warp == 0) tmem_alloc<512>(tslot);
  if (threadIdx.x == 32) {
    for (int g = 0; g < G; ++g) {                                       // done[g], doneB[g], startB[g] (slot 0 unused)
      mbar_init(bars + G + g, 1);
      mbar_init(bars + 2 * G + g, 1);
      mbar_init(bars + 3 * G + g, 1);
    }
    for (int q = 0; q < kSlots; ++q) {
      mbar_init(dfull + q, kTile);
      mbar_init(dempty + q, SCAT > 0 ? SCAT * 32 : 1);
    }
    fence_mbar_init();
  }
  if (image != nullptr) {
    copy_image_async(wsm, image, WO::total);                            // lands while the group regions are cleared below
    copy_image_async(sm + SM::off_bias, image + Scratch<K0P, KCP>::off_bias_f32, 6 * 64 * 4);
  } else {
    stage_weights_bf16<K0P, KCP>(params, m, wsm, nullptr, nullptr, bias);
  }
  {
    // zero the group regions (the M = 128 operands read column groups they do not own: keep them finite), then the ones
    for (int e = threadIdx.x; e < G * SM::grp_bytes / 16; e += blockDim.x)
      reinterpret_cast<uint4*>(sm + SM::off_grp)[e] = make_uint4(0, 0, 0, 0);
    __syncthreads();
    const uint32_t one2 = OP::pack(1.f, 1.f);
    const uint4 ones4 = make_uint4(one2, one2, one2, one2);
    for (int e = threadIdx.x; e < kTile; e += blockDim.x) {
      for (int g = 0; g < G; ++g) {
        uint8_t* gb = sm + SM::off_grp + g * SM::grp_bytes;
        reinterpret_cast<uint4*>(gb + SM::h2 + 8 * kCg)[e] = ones4;
        reinterpret_cast<uint4*>(gb + SM::c2 + 8 * kCg)[e] = ones4;
        reinterpret_cast<uint4*>(gb + SM::x0 + kTile * K0P * 2)[e] = ones4;
        reinterpret_cast<uint4*>(gb + SM::h1 + 8 * kCg)[e] = ones4;
        reinterpret_cast<uint4*>(gb + SM::c1 + 8 * kCg)[e] = ones4;
      }
    }
  }
  cp_async_wait_all();
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tbase = *tslot;
  if (TRACE && blockIdx.x == 0 && threadIdx.x == 0) trace[2001] = clock64();
  const long long ntiles = (n + kTile - 1) / kTile;
  const long long nslots = (long long)gridDim.x * G;
  long long nt[G];
  long long cta_tiles = 0;
#pragma unroll
  for (int g = 0; g < G; ++g) {
    nt[g] = tiles_of_slot(ntiles, (long long)g * gridDim.x + blockIdx.x, nslots);
    cta_tiles += nt[g];
  }
  const long long kmax = nt[0];                  // the slot of group 0 never has fewer tiles than a later group's

  auto wgrad_issuer = [&]() {
    // ===== weight-gradient issuer (one converged warp): every weight/bias-gradient GEMM of the CTA -- the accumulators
    // all tiles share -- comes from this one thread sequence, visiting the groups in a fixed order.  It waits on
    // startB[g], committed by the group right behind the stage's dgrad (so the chain-critical dgrad never queues behind a
    // weight gradient in the in-order tensor pipe), and commits to doneB[g].  startB advances only on backward stages and
    // a group cannot pass one without doneB, so the phase tracking cannot fall behind.
    // The operand descriptors are rebuilt from two laundered base values in every stage: left to itself the compiler
    // hoists the loop-invariant descriptors out of the tile loop and spills them. =====
    const uint32_t tb0 = __shfl_sync(kFull, tbase, 0);
    const uint32_t sm0 = a4_of(sm);
    uint32_t par[G];
#pragma unroll
    for (int g = 0; g < G; ++g) par[g] = 0;
    bool first = true;                           // gradient accumulators not yet written
#define HBR_WGRAD_STAGE(BODY)                                              \
  _Pragma("unroll") for (int g = 0; g < G; ++g) {                          \
    if (k < nt[g]) {                                                       \
      mbar_wait(bars + 3 * G + g, par[g]);                                 \
      par[g] ^= 1;                                                         \
      fence_after_sync();                                                  \
      if (elect_one()) {                                                   \
        uint32_t sb = sm0, tb = tb0;                                       \
        asm volatile("" : "+r"(sb), "+r"(tb));                             \
        const uint32_t base = sb + (SM::off_grp + g * SM::grp_bytes) / 16; \
        const uint32_t x0a = base + SM::x0 / 16, h1a = base + SM::h1 / 16, h2a = base + SM::h2 / 16, \
                       cina = base + SM::cin / 16, c1a = base + SM::c1 / 16, c2a = base + SM::c2 / 16, \
                       dzsa = base + SM::dzs / 16;                         \
        (void)x0a; (void)h1a; (void)h2a; (void)cina; (void)c1a; (void)c2a; (void)dzsa; \
        const bool acc = !(first && g == 0);                               \
        BODY;                                                              \
        commit(bars + 2 * G + g);                                          \
      }                                                                    \
      __syncwarp();                                                        \
    }                                                                      \
  }
    for (long long k = 0; k < kmax; ++k) {
      // ---- weight + bias gradients: reduction over the tile's 128 points ----
      HBR_WGRAD_STAGE(issue_wgrad(tb + TM::g5, c2a, dzsa, 16, acc, 128));   // transposed; input c2 | ones
      HBR_WGRAD_STAGE(issue_wgrad(tb + TM::g4, c2a, c1a, TM::n4, acc));     // dZ = c2 tile, input c1 | ones
      HBR_WGRAD_STAGE(issue_wgrad(tb + TM::g3, c1a, cina, TM::n3, acc));    // dZ = c1 tile, input cin (planted 1.0)
      HBR_WGRAD_STAGE(issue_wgrad(tb + TM::g2, h2a, dzsa, 16, acc, 128));   // transposed; input h2 | ones
      HBR_WGRAD_STAGE(issue_wgrad(tb + TM::g1, h2a, h1a, TM::n1, acc));     // dZ = h2 tile, input h1 | ones
      HBR_WGRAD_STAGE(issue_wgrad(tb + TM::g0, h1a, x0a, TM::n0, acc));     // dZ = h1 tile, input x0 | ones
      first = false;
    }
  };
  if (SCAT > 0 && warp >= 4 * G) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(scat_regs(G)));
    if (warp == kIssuerWarp) {
      wgrad_issuer();
    } else {
      // ===== scatter warps: the hash-grid backward of every tile of this CTA (same arithmetic as hash_bwd_kernel) =====
      // ROLLED loops on purpose: the tile groups' layer chain is ~115 KB of straight-line code; with the unit body unrolled
      // (10 units x 2 groups = another ~90 KB) both instruction streams missed the instruction cache all the time
      // (measured: 435 us for the kernel against 153 + 169 us for the two separate kernels).
      const int sw = warp - 4 * G;                                        // 0 .. SCAT-1
      constexpr int kUnits = 64;
      uint32_t fpar = 0;                                                  // bit g: parity of dfull[g]
      int pbuf = 0;
#pragma unroll 1
      for (long long kg = 0; kg < kmax * G; ++kg) {
        const long long k = kg / G;
        const int g = (int)(kg - k * G);
        if (k >= (g == 0 ? nt[0] : nt[G - 1])) continue;
        const long long tile = (long long)g * gridDim.x + blockIdx.x + k * nslots;
        // normalised positions ((o + d t) - mu) / sigma of the tile's 128 points: warp w < 4 forms those of slice w and
        // leaves them in shared memory for everybody (every warp forming all four slices itself cost ~3 k cycles per tile,
        // on every warp at the same time: the scatter warps alone ran 250 us whether they were 7 or 11)
        float* pos = reinterpret_cast<float*>(sm + SM::off_pos) + pbuf * (kTile * 3);
        pbuf ^= 1;
        if (sw < 4) {
          const long long gp = tile * kTile + sw * 32 + lane;
          float x3[3] = {0.f, 0.f, 0.f};
          if (gp < n) {
            long long ray, smp;
            if (gp < (1LL << 31) && enc.S < (1LL << 31)) {
              const unsigned q = (unsigned)gp / (unsigned)enc.S;
              ray = q; smp = (long long)((unsigned)gp - q * (unsigned)enc.S);
            } else {
              ray = gp / enc.S; smp = gp - ray * enc.S;
            }
            const float tt = __ldg(enc.rt + ray * enc.t_stride + smp);
#pragma unroll
            for (int a = 0; a < 3; ++a) {
              const float x = __fadd_rn(__ldg(enc.ro + ray * 3 + a), __fmul_rn(__ldg(enc.rd + ray * 3 + a), tt));
              x3[a] = __fdiv_rn(__fsub_rn(x, geom.mu[a]), geom.sigma);
            }
          }
#pragma unroll
          for (int a = 0; a < 3; ++a) pos[(sw * 32 + lane) * 3 + a] = x3[a];
        }
        asm volatile("bar.sync 3, %0;" ::"n"(SCAT * 32) : "memory");
        // slot of this tile: the group's h2 region (two groups), or the ring slot k & 1 (one group)
        const int slot = G == 2 ? g : (int)(k % kSlots);
        const float2* stg = reinterpret_cast<const float2*>(G == 2 ? sm + SM::off_grp + g * SM::grp_bytes + SM::h2
                                                                   : sm + SM::off_ring + slot * SM::ring_slot);
        mbar_wait(dfull + slot, (fpar >> slot) & 1u);
        fpar ^= 1u << slot;
        // two units per iteration (u and u + SCAT), written side by side: a unit is one long dependent chain (cells ->
        // run test -> log-step shuffle merge -> reductions), and seven warps of such chains leave the SM's issue slots
        // mostly empty (measured: 319 us for the kernel with the reductions compiled out) -- two independent chains per
        // warp give the scheduler twice the work to pick from
#pragma unroll 1
        for (int ua = sw; ua < kUnits; ua += 2 * SCAT) {
          int cx[2][3];
          float val[2][8][2];
          bool valid[2], head[2];
          int end[2], lvl_of[2];
          int maxrun = 0;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int u = ua + h * SCAT;
            const bool live = u < kUnits;                                 // warp-uniform; the last pair of a warp may be half
            const int uc = live ? u : ua;
            const int l = uc >> 2, sl = uc & 3;
            lvl_of[h] = l;
            valid[h] = live && tile * kTile + sl * 32 + lane < n;
            const float2 gy = stg[l * kTile + sl * 32 + lane];
            float un[3];
#pragma unroll
            for (int a = 0; a < 3; ++a) un[a] = pos[(sl * 32 + lane) * 3 + a];
            const float s = geom.scale[l];
            // cells as 32-bit integers: the power-of-two hash and the run test below only look at the low 32 bits of the
            // reference's int64 cell, and trunc / frac agree with the int64 path whenever |u| < 2^31 (else: that path)
            float fr[3];
#pragma unroll
            for (int a = 0; a < 3; ++a) {
              const float uu = __fmul_rn(un[a], s);
              if (fabsf(uu) < 2147483648.f) {
                cx[h][a] = __float2int_rz(uu);
                fr[a] = __fsub_rn(uu, __int2float_rn(cx[h][a]));
              } else {
                const long long c64 = __float2ll_rz(uu);
                cx[h][a] = (int)(uint32_t)(unsigned long long)c64;
                fr[a] = __fsub_rn(uu, __ll2float_rn(c64));
              }
            }
            float w[8];
            corner_weights(fr[0], fr[1], fr[2], w);
#pragma unroll
            for (int c = 0; c < 8; ++c) { val[h][c][0] = w[c] * gy.x; val[h][c][1] = w[c] * gy.y; }
            // runs of consecutive lanes in the same cell (a ray crosses a cell in one contiguous stretch)
            const int pcx = __shfl_up_sync(kFull, cx[h][0], 1);
            const int pcy = __shfl_up_sync(kFull, cx[h][1], 1);
            const int pcz = __shfl_up_sync(kFull, cx[h][2], 1);
            const int pvalid = __shfl_up_sync(kFull, (int)valid[h], 1);
            head[h] = lane == 0 || !valid[h] || !pvalid || pcx != cx[h][0] || pcy != cx[h][1] || pcz != cx[h][2];
            const unsigned heads = __ballot_sync(kFull, head[h]);
            const unsigned above = lane == 31 ? 0u : (heads & (0xfffffffeu << lane));
            end[h] = above ? (__ffs(above) - 1) : 32;                     // first lane of the next run
            maxrun = max(maxrun, head[h] ? end[h] - lane : 0);
          }
          maxrun = __reduce_max_sync(kFull, maxrun);
          for (int d = 1; d < maxrun; d <<= 1) {
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
              for (int c = 0; c < 8; ++c)
#pragma unroll
                for (int f = 0; f < 2; ++f) {
                  const float t = __shfl_down_sync(kFull, val[h][c][f], d);
                  if (lane + d < end[h]) val[h][c][f] += t;
                }
          }
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            if (head[h] && valid[h]) {
              uint32_t idx[8];
              {
                const uint32_t a0 = (uint32_t)cx[h][0], b0 = (uint32_t)cx[h][1] * kPrimeY, c0 = (uint32_t)cx[h][2] * kPrimeZ;
                const uint32_t aa[2] = {a0, a0 + 1u}, bb[2] = {b0, b0 + kPrimeY}, cc[2] = {c0, c0 + kPrimeZ};
#pragma unroll
                for (int q = 0; q < 8; ++q) idx[q] = (aa[q & 1] ^ bb[(q >> 1) & 1] ^ cc[(q >> 2) & 1]) & (geom.T - 1);
              }
              float* lvl = enc.dtable + (size_t)lvl_of[h] * geom.T * 2;
              if (!(cx[h][0] & 1)) {   // even x: corners (x, x+1) share one aligned 16-byte slot, one red.global.add.v4.f32
#pragma unroll
                for (int c = 0; c < 8; c += 2) {
                  const bool odd = idx[c] & 1;
                  const float4 q = odd ? make_float4(val[h][c + 1][0], val[h][c + 1][1], val[h][c][0], val[h][c][1])
                                       : make_float4(val[h][c][0], val[h][c][1], val[h][c + 1][0], val[h][c + 1][1]);
                  atomicAdd(reinterpret_cast<float4*>(lvl) + (idx[c] >> 1), q);
                }
              } else {
#pragma unroll
                for (int c = 0; c < 8; ++c)
                  atomicAdd(reinterpret_cast<float2*>(lvl) + idx[c], make_float2(val[h][c][0], val[h][c][1]));
              }
            }
          }
        }
        mbar_arrive(dempty + slot);                                      // every unit's values have been read: the slot may be rewritten
      }
    }
  } else if (SCAT == 0 && warp >= 4 * G) {
    wgrad_issuer();
  } else {
    // ===== tile group =====
    if (SCAT > 0) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(kGroupRegs));
    const int g = warp >> 2;
    const int r = threadIdx.x & (kTile - 1);
    uint8_t* gb = sm + SM::off_grp + g * SM::grp_bytes;
    uint8_t *x0 = gb + SM::x0, *h1 = gb + SM::h1, *h2 = gb + SM::h2, *cin = gb + SM::cin, *c1 = gb + SM::c1,
            *c2 = gb + SM::c2, *dzs = gb + SM::dzs;
    // The group's first warp issues the group's own forward-recompute / dgrad GEMMs (independent work accumulator): the
    // hand-off is one named barrier over the 128 threads instead of an mbarrier round trip through an issuer warp.
    uint64_t* done = bars + G + g;
    uint64_t* doneb = bars + 2 * G + g;
    uint64_t* startb = bars + 3 * G + g;
    const bool issuer = (warp & 3) == 0;
    const uint32_t tgrp = tbase + g * kGrpCols, tgrp_a = tgrp + 64;
    const uint32_t taddr = tgrp + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t taddr_a = taddr + 64;
    const uint32_t wa = a4_of(wsm), x0a = a4_of(x0);
    uint32_t dphase = 0, bphase = 0, ephase = 0;
    (void)ephase;
#define HBR_BSTAGE(BWD, BODY) HBR_BSTAGE_T(BODY, {})
    // forward-recompute stage; TRAIL runs between the issue and the wait (work the chain does not need)
#define HBR_BSTAGE_T(BODY, ...)                                            \
  do {                                                                     \
    HBR_STAMP(0);                                                          \
    fence_async_smem();                                                    \
    fence_before_sync();                                                   \
    asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory");             \
    if (issuer) {                                                          \
      fence_after_sync();                                                  \
      if (elect_one()) {                                                   \
        BODY;                                                              \
        commit(done);                                                      \
      }                                                                    \
      __syncwarp();                                                        \
    }                                                                      \
    __VA_ARGS__;                                                           \
    mbar_wait(done, dphase);                                               \
    dphase ^= 1;                                                           \
    fence_after_sync();                                                    \
    HBR_STAMP(1);                                                          \
  } while (0)
    // Backward stage.  The A operand of its dgrad GEMM is already in tensor memory (written by the epilogue before), so
    // the chain-critical dgrad is issued first; TRAIL is what the chain does not need -- wait for the previous
    // weight-gradient GEMM, then the in-place store of the dZ tile this stage's weight-gradient GEMM reads -- and runs in
    // the shadow of the dgrad; a plain arrive on startB then lets the weight-gradient warp issue behind the dgrad.
    // (Waiting for the weight-gradient GEMM BEFORE issuing the next dgrad, as the first version did, put ~500 cycles of
    // weight-gradient latency on every backward stage of the chain.)
#define HBR_STAMP(j)                                                       \
  do {                                                                     \
    if (TRACE && blockIdx.x == 0 && threadIdx.x == 0 && tgi < 12) trace[tgi * 80 + stamp_i++] = clock64(); \
  } while (0)
#define HBR_BSTAGE_BWD(BODY, ...)                                          \
  do {                                                                     \
    HBR_STAMP(0);                                                          \
    fence_before_sync();                                                   \
    asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory");             \
    HBR_STAMP(1);                                                          \
    if (issuer) {                                                          \
      fence_after_sync();                                                  \
      if (elect_one()) {                                                   \
        BODY;                                                              \
        commit(done);                                                      \
      }                                                                    \
      __syncwarp();                                                        \
    }                                                                      \
    HBR_STAMP(2);                                                          \
    __VA_ARGS__;                                                           \
    HBR_STAMP(3);                                                          \
    fence_async_smem();                                                    \
    asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory");             \
    if (r == 0) mbar_arrive(startb);                                       \
    HBR_STAMP(4);                                                          \
    mbar_wait(done, dphase);                                               \
    dphase ^= 1;                                                           \
    fence_after_sync();                                                    \
    HBR_STAMP(5);                                                          \
  } while (0)
    // weight-gradient GEMM of the stage has finished reading its tiles (they are about to be overwritten in place)
#define HBR_WAIT_B()           \
  do {                         \
    mbar_wait(doneb, bphase);  \
    bphase ^= 1;               \
  } while (0)
    const bool vec_ok = in0 == K0P && feat_stride == K0P && ((uintptr_t)feat & 15) == 0;
    const bool dvec_ok = dfeat != nullptr && in0 == K0P && dfeat_stride == K0P && ((uintptr_t)dfeat & 15) == 0 && K0P <= 32;
    // dfeat_stride == HBR_DFEAT_LEVEL_MAJOR: d(features) leave as (K0P/2 levels, n, 2) -- what the level-major scatter-add
    // (hash_grid.cu: hash_bwd_lm_kernel) reads with coalesced 8-byte loads, level by level (checked by the entry point)
    const bool dlm = dfeat != nullptr && dfeat_stride == HBR_DFEAT_LEVEL_MAJOR && K0P <= 32;
    int tgi = 0;
    (void)tgi;
    // The next tile's fp32 feature rows travel into the (by then dead) c2 tile with cp.async during the last backward
    // stages and are converted from shared memory at the start of the tile: a tile that begins by waiting for its own
    // global loads spent ~2 300 cycles (15 % of its chain) there.  Contiguous K0P == 32 layout only (16 KB <= the tile).
    constexpr bool kStage = K0P == 32 && !ENC;
    const bool f16in = !ENC && feat16 != 0;
    const uint16_t* featq = reinterpret_cast<const uint16_t*>(feat);
    const bool stage_ok = kStage && (vec_ok || f16in);
    bool staged = false;
    for (long long tile = (long long)g * gridDim.x + blockIdx.x; tile < ntiles; tile += nslots, ++tgi) {
      int stamp_i = 0;
      (void)stamp_i;
      HBR_STAMP(0);                              // tile start
      const long long gp = tile * kTile + r;
      const bool valid = gp < n;
      const long long dir_row = valid ? (dir_rows != nullptr ? (long long)__ldg(dir_rows + gp) : gp / dir_group) : 0;
      if (valid && lane == 0) prefetch_l1(dirs + dir_row * dv);
      float4 fo = make_float4(0.f, 0.f, 0.f, 0.f), go = fo;           // saved forward output, upstream gradient
      if (valid) {
        fo = __ldg(reinterpret_cast<const float4*>(out + gp * 4));
        go = __ldg(reinterpret_cast<const float4*>(dout + gp * 4));
        go.x *= gscale; go.y *= gscale; go.z *= gscale; go.w *= gscale;
      }
      // ---- recompute the forward activations ----
      float pt[3] = {0.f, 0.f, 0.f};
      if (ENC) {
        if (valid) { pt[0] = __ldg(enc.x + gp * 3 + 0); pt[1] = __ldg(enc.x + gp * 3 + 1); pt[2] = __ldg(enc.x + gp * 3 + 2); }
        load_feat16_row(enc, gp, n, r, x0);
      } else if (staged && f16in) {
        // the rows were copied into the (dead) c2 tile in canonical layout already: move this thread's own row over
        cp_async_wait_all();
#pragma unroll
        for (int cg = 0; cg < K0P / 8; ++cg) {
          const uint32_t o = chunk_off(r, cg, kTile);
          *reinterpret_cast<uint4*>(x0 + o) = *reinterpret_cast<const uint4*>(c2 + o);
        }
      } else if (f16in) {
        copy_feat16_async<K0P, 1>(featq, tile * kTile, n, r, 0, x0);
        cp_async_wait_all();
      } else if (staged) {
        cp_async_wait_all();
        convert_staged_features<K0P>(c2, r, x0);   // every thread converts exactly the elements it copied itself
      } else {
        load_features<K0P>(feat, feat_stride, tile * kTile, n, in0, vec_ok, r, r, 0, x0);
      }
      if ((tile + nslots) * kTile + r < n) {
        if (ENC) prefetch_l2(enc.feat16 + ((tile + nslots) * kTile + r) * 32);
        else if (!stage_ok && !f16in) prefetch_l2(feat + ((tile + nslots) * kTile + r) * feat_stride);
        if ((r & 7) == 0) {
          prefetch_l2(out + ((tile + nslots) * kTile + r) * 4);
          prefetch_l2(dout + ((tile + nslots) * kTile + r) * 4);
        }
      }
      uint4 dzt[8];                              // activation / dZ tile formed by the last epilogue, stored one stage later
      HBR_BSTAGE(false, issue_fwd(tgrp, x0a, wa + WO::w0 / 16, 64, K0P));                       // F0 (x0 from shared memory)
      relu_bias_to_tmem64(taddr, bias + 0, taddr_a, dzt);
      HBR_BSTAGE_T(issue_fwd_ts(tgrp, tgrp_a, wa + WO::w1 / 16, 64, 64, false),                 // F1
                   { store_tile64(r, h1, dzt); });
      relu_bias_to_tmem64(taddr, bias + 64, taddr_a, dzt);
      HBR_BSTAGE_T(issue_fwd_ts(tgrp, tgrp_a, wa + WO::w2 / 16, 16, 64, false),                 // F2
                   {
                     if (SCAT > 0 && G == 2 && tgi > 0) {   // the scatter warps have taken the previous tile's d(features) out of h2
                       mbar_wait(dempty + g, ephase & 1u);
                       ephase ^= 1u;
                     }
                     store_tile64(r, h2, dzt);
                   });
      {
        float o16[16];
        tmem_ld<16>(taddr, o16);
#pragma unroll
        for (int k = 0; k < 16; ++k) o16[k] += bias[128 + k];
        build_cin<KCP, kCinOne>(o16, dirs, dir_row, dv, valid, r, cin);
        // the same row as the A operand of the colour net's first GEMM: copy this thread's chunks smem -> TMEM
        uint32_t p[KCP / 2];
#pragma unroll
        for (int cg = 0; cg < KCP / 8; ++cg) {
          const uint4 q = *reinterpret_cast<const uint4*>(cin + chunk_off(r, cg, kTile));
          p[4 * cg] = q.x; p[4 * cg + 1] = q.y; p[4 * cg + 2] = q.z; p[4 * cg + 3] = q.w;
        }
        tmem_st16(taddr_a, p);
        if (KCP == 48) tmem_st8(taddr_a + 16, p + 16);
        else tmem_st16(taddr_a + 16, p + 16);
        tmem_st_wait();
      }
      HBR_BSTAGE(false, issue_fwd_ts(tgrp, tgrp_a, wa + WO::w3 / 16, 64, KCP, false));          // F3
      relu_bias_to_tmem64(taddr, bias + 192, taddr_a, dzt);
      HBR_BSTAGE_T(issue_fwd_ts(tgrp, tgrp_a, wa + WO::w4 / 16, 64, 64, false),                 // F4
                   { store_tile64(r, c1, dzt); });
      relu_bias_to_tmem64(taddr, bias + 256, 0xffffffffu, dzt);         // c2 feeds no forward GEMM here
      {
        // d(rgb_pre) = g * ELU'(pre), with ELU'(pre) = pre > 0 ? 1 : exp(pre) = elu(pre) + 1 from the saved output
        float dz16[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) dz16[c] = 0.f;
        dz16[0] = go.x * (fo.x > 0.f ? 1.f : fo.x + 1.f);
        dz16[1] = go.y * (fo.y > 0.f ? 1.f : fo.y + 1.f);
        dz16[2] = go.z * (fo.z > 0.f ? 1.f : fo.z + 1.f);
        store_dz16_both(dz16, r, dzs, taddr_a);
      }
      HBR_BSTAGE_BWD(issue_dgrad_ts(tgrp, tgrp_a, wa + WO::w5 / 16, 16, 64),                    // col_model.4: work = dA(c2)
                     { store_tile64(r, c2, dzt); });   // the c2 activations, before this stage's weight-gradient GEMM starts
      masked_dz_to_tmem64(taddr, r, c2, taddr_a, dzt);
      HBR_BSTAGE_BWD(issue_dgrad_ts(tgrp, tgrp_a, wa + WO::w4 / 16, 64, 64),                    // col_model.2: work = dA(c1)
                     { HBR_WAIT_B(); store_tile64(r, c2, dzt); });
      masked_dz_to_tmem64(taddr, r, c1, taddr_a, dzt);
      staged = stage_ok && tile + nslots < ntiles;
      HBR_BSTAGE_BWD(issue_dgrad_ts(tgrp, tgrp_a, wa + WO::w3 / 16, 64, KCP), {                 // col_model.0: work[0,KCP) = d(cin)
        HBR_WAIT_B();                            // the weight-gradient GEMM reading c2 (dZ) and c1 has finished: c2 is dead
        store_tile64(r, c1, dzt);
        if (staged && f16in) copy_feat16_async<K0P, 1>(featq, (tile + nslots) * kTile, n, r, 0, c2);
        else if (staged) stage_features_async<K0P>(feat, (tile + nslots) * kTile, n, r, c2);
      });
      float dd[KCP - kFeat];                     // d(direction encoding), reduced into ddirs off the chain
      uint32_t dzp[8];
      {
        float dc[KCP], dz16[16];
        tmem_ld<KCP>(taddr, dc);
        dz16[0] = go.w * (fo.w > 0.f ? 1.f : 0.01f);                    // LeakyReLU' from the saved density
#pragma unroll
        for (int k = 0; k < kFeat; ++k) dz16[1 + k] = dc[k];
#pragma unroll
        for (int k = kFeat; k < KCP; ++k) dd[k - kFeat] = dc[k] * ginv;
#pragma unroll
        for (int q = 0; q < 8; ++q) dzp[q] = OP::pack(dz16[2 * q], dz16[2 * q + 1]);
        tmem_st8(taddr_a, dzp);
        tmem_st_wait();
      }
      HBR_BSTAGE_BWD(issue_dgrad_ts(tgrp, tgrp_a, wa + WO::w2 / 16, 16, 64), {                  // sig_model.4: work = dA(h2)
        HBR_WAIT_B();                            // the 16-wide dZ tile may alias the padding of the colour-net input tile
        *reinterpret_cast<uint4*>(dzs + chunk_off(r, 0, kTile)) = make_uint4(dzp[0], dzp[1], dzp[2], dzp[3]);
        *reinterpret_cast<uint4*>(dzs + chunk_off(r, 1, kTile)) = make_uint4(dzp[4], dzp[5], dzp[6], dzp[7]);
        if (ddirs != nullptr) {
          // rows of one warp usually belong to one ray: reduce over the warp first, one atomic per column
          const long long row0 = __shfl_sync(kFull, dir_row, 0);
          const bool uniform = __all_sync(kFull, dir_row == row0 && valid);
          _Pragma("unroll") for (int k = kFeat; k < KCP; ++k) {
            if (k < kFeat + dv) {
              if (uniform) {
                const float sdd = warp_sum(dd[k - kFeat]);
                if (lane == 0) atomicAdd(ddirs + row0 * dv + (k - kFeat), sdd);
              } else if (valid) {
                atomicAdd(ddirs + dir_row * dv + (k - kFeat), dd[k - kFeat]);
              }
            }
          }
        }
      });
      masked_dz_to_tmem64(taddr, r, h2, taddr_a, dzt);
      HBR_BSTAGE_BWD(issue_dgrad_ts(tgrp, tgrp_a, wa + WO::w1 / 16, 64, 64),                    // sig_model.2: work = dA(h1)
                     { HBR_WAIT_B(); store_tile64(r, h2, dzt); });
      masked_dz_to_tmem64(taddr, r, h1, taddr_a, dzt);
      HBR_BSTAGE_BWD(issue_dgrad_ts(tgrp, tgrp_a, wa + WO::w0 / 16, 64, K0P),                   // sig_model.0: work[0,K0P) = d(feat)
                     { HBR_WAIT_B(); store_tile64(r, h1, dzt); });
      if (SCAT > 0) {
        // d(features) of this row -> the (dead) h2 region as [level][point] float2 (conflict-free both ways), then the
        // scatter warps take over: 128 arrivals (release) on dfull
        float df[K0P];
        tmem_ld<K0P>(taddr, df);
        const int slot = G == 2 ? g : (tgi % kSlots);
        if (G == 1 && tgi >= kSlots) {             // ring slot: its previous tile (kSlots tiles back) has been read
          mbar_wait(dempty + slot, (ephase >> slot) & 1u);
          ephase ^= 1u << slot;
        }
        float2* stg = reinterpret_cast<float2*>(G == 2 ? h2 : sm + SM::off_ring + slot * SM::ring_slot);
#pragma unroll
        for (int l = 0; l < K0P / 2; ++l) stg[l * kTile + r] = make_float2(df[2 * l] * ginv, df[2 * l + 1] * ginv);
        mbar_arrive(dfull + slot);
      } else if (ENC) {
        float df[K0P];
        tmem_ld<K0P>(taddr, df);
#pragma unroll
        for (int k = 0; k < K0P; ++k) df[k] *= ginv;
        scatter_row(enc, geom, pt, valid, lane, df);
      } else if (dfeat != nullptr) {
        float df[K0P];
        tmem_ld<K0P>(taddr, df);
#pragma unroll
        for (int k = 0; k < K0P; ++k) df[k] *= ginv;
        if (dlm) {
          // rows -> the (dead) h2 tile as [level][point] float2 (conflict-free), then the tile's 1 KB run of every level
          // goes out with lane-contiguous 16-byte stores
          float2* stg = reinterpret_cast<float2*>(h2);
#pragma unroll
          for (int l = 0; l < K0P / 2; ++l) stg[l * kTile + r] = make_float2(df[2 * l], df[2 * l + 1]);
          asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory");
          const float4* s4 = reinterpret_cast<const float4*>(h2);
#pragma unroll
          for (int it = 0; it < K0P / 4; ++it) {
            const int idx = it * kTile + r, l = idx >> 6, j = idx & 63;            // 64 float4 = 128 points per level
            if (tile * kTile + 2 * j < n)                                          // n is even: a pair is valid or not
              reinterpret_cast<float4*>(dfeat + ((size_t)l * n + tile * kTile) * 2)[j] = s4[idx];
          }
        } else if (dvec_ok) {
          // rows -> the (dead) h2 tile as fp32 with an XOR swizzle on the 16-byte chunk index, then lane-contiguous
          // float4 stores of the tile's contiguous 128*K0P*4-byte block of dfeat (measured: eight 16-byte stores per
          // thread straight from registers, 32 partial sectors per instruction, are ~400 cycles slower per tile)
          constexpr int kQ = K0P / 4;
          float4* stg = reinterpret_cast<float4*>(h2);
#pragma unroll
          for (int c = 0; c < kQ; ++c)
            stg[r * kQ + (c ^ (r & 7))] = make_float4(df[4 * c], df[4 * c + 1], df[4 * c + 2], df[4 * c + 3]);
          asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory");
          float4* dst = reinterpret_cast<float4*>(dfeat + tile * kTile * K0P);
#pragma unroll
          for (int it = 0; it < kQ; ++it) {
            const int idx = it * kTile + r, row = idx / kQ, c = idx % kQ;
            if (tile * kTile + row < n) dst[idx] = stg[row * kQ + (c ^ (row & 7))];
          }
        } else if (valid) {
#pragma unroll
          for (int k = 0; k < K0P; ++k)
            if (k < in0) dfeat[gp * dfeat_stride + k] = df[k];
        }
      }
      HBR_STAMP(0);                              // d(feat) written
      HBR_WAIT_B();                              // x0 / h1 are rewritten by the next tile
      HBR_STAMP(0);                              // tile end
    }
    if (TRACE && blockIdx.x == 0 && threadIdx.x == 0) trace[2004] = clock64();
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  if (TRACE && blockIdx.x == 0 && threadIdx.x == 0) trace[2002] = clock64();

  if (warp < 4 * G) flush_gradients<K0P, KCP, 2 * kGrpCols>(tbase, warp, lane, G, m, cta_tiles > 0, dparams, grad_rows, ginv, 4 * G * 32);
  fence_before_sync();
  __syncthreads();
  if (TRACE && blockIdx.x == 0 && threadIdx.x == 0) trace[2003] = clock64();
  if (warp == 0) tmem_dealloc<512>(tbase);
}

// The reference's configuration (in0 = 32, d_view = 24) gets the widest pipelines (4 forward / 2 backward tile groups
// per SM); other widths (in0 <= 64, 15 + d_view <= 64) run the same kernels with padded K and fewer groups.
static inline bool narrow_shape(const hbr_mlp_dims* d) { return d->in0 <= 32 && d->d_view + kFeat <= 40; }

template <int K0P, int KCP, int G, bool ENC>
static int launch_fwd_tc(const float* feat, int64_t feat_stride, const float* dirs, int64_t dir_group, int64_t n,
                         const float* params, int in0, int dv, float* out, uint8_t* scratch, const EncArgs& enc,
                         const HashGeom& geom, int feat16, int image_ready, const unsigned long long* n_dev, const int32_t* dir_rows,
                         cudaStream_t st) {
  constexpr int smem = FwdSmem<K0P, KCP, G>::total;
  const int grid = (int)min64(ceil_div(ceil_div(n, kTile), G), sm_count());
  // the full image (incl. the fp32 biases only the backward reads): the backward call of the step can then skip its prep
  if (scratch != nullptr && !image_ready) mlp_prep_kernel<K0P, KCP><<<kPrepCtas, 256, 0, st>>>(params, in0, dv, scratch);
  auto kern = mlp_fwd_tc_kernel<K0P, KCP, G, false, ENC>;
  HBR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  kern<<<grid, G * kTile, smem, st>>>(feat, feat_stride, dirs, dir_group, n, params, in0, dv, out, scratch, nullptr, enc,
                                             geom, feat16, n_dev, dir_rows);
  HBR_LAUNCH_CHECK();
  return HBR_OK;
}

// forward with the hash-grid gather on dedicated warps (mlp_fwd_tc_kernel<GATH>): 2 tile groups + GATH gather warps per SM.
// image != 0: the operand image in `scratch` is current (copied by the tile groups); otherwise they convert the fp32
// parameters themselves while the gather warps work -- no prep kernel in front of this one either way.
template <int GATH>
static int launch_fwd_gather_tc(const float* dirs, int64_t dir_group, int64_t n, const float* params, int dv, float* out,
                                const uint8_t* image, const EncArgs& enc, const HashGeom& geom, cudaStream_t st) {
  constexpr int G = 2;
  constexpr int smem = FwdSmem<32, 48, G>::total;
  const int grid = (int)min64(ceil_div(ceil_div(n, kTile), G), sm_count());
  auto kern = mlp_fwd_tc_kernel<32, 48, G, false, false, GATH>;
  HBR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  kern<<<grid, G * kTile + GATH * 32, smem, st>>>(nullptr, 32, dirs, dir_group, n, params, 32, dv, out, image, nullptr, enc, geom,
                                                  1, nullptr, nullptr);
  HBR_LAUNCH_CHECK();
  return HBR_OK;
}

template <int K0P, int KCP, int G, bool ENC, int SCAT = 0>
static int launch_bwd_tc(const float* feat, int64_t feat_stride, const float* dirs, int64_t dir_group, int64_t n,
                         const float* params, int in0, int dv, const float* out, const float* dout, float* dfeat,
                         int64_t dfeat_stride, float* ddirs, float* dparams, uint8_t* scratch, const EncArgs& enc,
                         const HashGeom& geom, float gscale, int feat16, int image_ready, int defer_reduce,
                         const unsigned long long* n_dev, const int32_t* dir_rows, cudaStream_t st) {
  using SC = Scratch<K0P, KCP>;
  constexpr int smem = BwdSmem<K0P, KCP, G>::total;
  const int grid = (int)min64(ceil_div(ceil_div(n, kTile), G), sm_count());
  const bool rows = scratch != nullptr && dparams != nullptr && grid <= SC::kMaxRows;
  if (scratch != nullptr && !image_ready) mlp_prep_kernel<K0P, KCP><<<kPrepCtas, 256, 0, st>>>(params, in0, dv, scratch);
  auto kern = mlp_bwd_tc_kernel<K0P, KCP, G, false, ENC, SCAT>;
  HBR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  if constexpr (SCAT > 0) {
    // the setmaxnreg budget (see the kernel): 16 warps, 4 per SM sub-partition, launched at `regs` registers per thread;
    // afterwards every sub-partition holds 2 tile-group warps + 2 scatter / issuer warps
    static_assert((G == 2 && SCAT == 7) || (G == 1 && SCAT == 11), "register budget: 16 warps, 4 per sub-partition");
    constexpr int kScatRegs = scat_regs(G);
    static int regs = 0;
    if (regs == 0) {
      cudaFuncAttributes fa;
      HBR_CUDA(cudaFuncGetAttributes(&fa, kern));
      regs = fa.numRegs;
    }
    HBR_REQUIRE(regs >= kScatRegs && regs <= kGroupRegs && 4 * regs <= 512 && G * kGroupRegs + (4 - G) * kScatRegs <= 4 * regs,
                "fused backward kernel compiled with %d registers per thread: the setmaxnreg budget does not hold", regs);
  }
  kern<<<grid, G * kTile + SCAT * 32 + 32, smem, st>>>(feat, feat_stride, dirs, dir_group, n, params, in0, dv, out, dout, dfeat,
                                           dfeat_stride, ddirs, dparams, scratch,
                                           rows ? reinterpret_cast<float*>(scratch + SC::off_grad) : nullptr, nullptr, enc, geom, gscale, feat16, n_dev, dir_rows);
  if (rows && !defer_reduce) {
    const int total = make_layout(in0, dv).total;
    mlp_grad_reduce_kernel<<<dim3((total + 255) / 256, kReduceSlices), 256, 0, st>>>(
        reinterpret_cast<const float*>(scratch + SC::off_grad), grid, SC::kRowFloats, total, dparams);
  }
  HBR_LAUNCH_CHECK();
  return HBR_OK;
}

// grid size / row use of launch_bwd_tc for n points (the deferred reduce must see the same numbers)
template <int K0P, int KCP, int G>
static int reduce_grads_tc(int64_t n, int in0, int dv, uint8_t* scratch, float* dparams, cudaStream_t st) {
  using SC = Scratch<K0P, KCP>;
  const int grid = (int)min64(ceil_div(ceil_div(n, kTile), G), sm_count());
  HBR_REQUIRE(grid <= SC::kMaxRows, "grid %d exceeds the gradient rows", grid);
  const int total = make_layout(in0, dv).total;
  mlp_grad_reduce_kernel<<<dim3((total + 255) / 256, kReduceSlices), 256, 0, st>>>(
      reinterpret_cast<const float*>(scratch + SC::off_grad), grid, SC::kRowFloats, total, dparams);
  HBR_LAUNCH_CHECK();
  return HBR_OK;
}
template <int K0P, int KCP>
static int prepare_tc(const float* params, int in0, int dv, uint8_t* scratch, cudaStream_t st) {
  mlp_prep_kernel<K0P, KCP><<<kPrepCtas, 256, 0, st>>>(params, in0, dv, scratch);
  HBR_LAUNCH_CHECK();
  return HBR_OK;
}


}  // namespace HBR_OPNS
}  // namespace hbr
