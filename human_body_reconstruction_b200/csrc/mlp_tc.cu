// MLP_3D on the 5th-generation tensor cores (tcgen05 + TMEM): bf16 operands, fp32 accumulation.
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
#include "mlp_layout.cuh"
#include "tc_common.cuh"

namespace hbr {
using namespace tc;

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
                                                   uint8_t* bias_sm, uint8_t* ones16) {
  const int JP[6] = {64, 64, 16, 64, 64, 16};
  const int KP[6] = {K0P, 64, 64, KCP, 64, 64};
  const int wofs[6] = {WOfs<K0P, KCP>::w0, WOfs<K0P, KCP>::w1, WOfs<K0P, KCP>::w2, WOfs<K0P, KCP>::w3,
                       WOfs<K0P, KCP>::w4, WOfs<K0P, KCP>::w5};
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    const int J = m.J[i], K = m.K[i];
    uint8_t* w = wsm + wofs[i];
    const int ncg = KP[i] / 8;
    for (int e = threadIdx.x; e < JP[i] * ncg; e += blockDim.x) {
      const int cg = e / JP[i], j = e - cg * JP[i];          // consecutive threads -> consecutive rows: conflict-free STS.128
      float v[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int k = cg * 8 + q;
        v[q] = (j < J && k < K) ? __ldg(params + m.W[i] + j * K + k) : 0.f;
      }
      store_chunk(w, j, cg, JP[i], v);
    }
    // bias tile [JP x 16] (K-major B operand of the bias MMA): col 0 = bf16(b), col 1 = bf16(b - bf16(b))
    uint8_t* bt = bias_sm + BOfs::ofs(i);
    for (int e = threadIdx.x; e < JP[i] * 2; e += blockDim.x) {
      const int cg = e / JP[i], j = e - cg * JP[i];
      float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (cg == 0 && j < J) {
        const float b = __ldg(params + m.b[i] + j);
        v[0] = bf16_round(b);
        v[1] = b - v[0];
      }
      store_chunk(bt, j, cg, JP[i], v);
    }
  }
  // ones16 [128 x 16] (A operand of the bias MMA): cols 0,1 = 1
  for (int e = threadIdx.x; e < kTile * 2; e += blockDim.x) {
    const int cg = e / kTile, rr = e - cg * kTile;
    float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (cg == 0) v[0] = v[1] = 1.f;
    store_chunk(ones16, rr, cg, kTile, v);
  }
}

// D[128 x N] (+)= A[128 x K] * W^T : A K-major activation tile, B K-major weight tile (forward)
__device__ __forceinline__ void issue_fwd(uint32_t tmem_d, uint32_t a_tile, uint32_t w_tile, int JP, int KP,
                                          bool accumulate = false) {
  const uint32_t idesc = make_idesc(128, JP, false, false);
#pragma unroll
  for (int kk = 0; kk < KP / 16; ++kk) {
    const uint64_t a = make_desc(a_tile + kk * 2 * kLBO128, kLBO128, 128);
    const uint64_t b = make_desc(w_tile + kk * 2 * JP * 16, JP * 16, 128);
    mma_f16(tmem_d, a, b, idesc, accumulate || kk > 0);
  }
}
// D[128 x JP] = b (broadcast over rows) + A W^T: the bias enters as ones16 x biasB^T
__device__ __forceinline__ void issue_layer(uint32_t tmem_d, uint32_t a_tile, uint32_t w_tile, uint32_t ones16,
                                            uint32_t bias_tile, int JP, int KP) {
  issue_fwd(tmem_d, ones16, bias_tile, JP, 16, false);
  issue_fwd(tmem_d, a_tile, w_tile, JP, KP, true);
}
// D[128 x KP] = dZ[128 x JP] * W : A K-major dZ tile, B = weight tile [JP x KP] read MN-major (N = k, K = j)
__device__ __forceinline__ void issue_dgrad(uint32_t tmem_d, uint32_t dz_tile, uint32_t w_tile, int JP, int KP) {
  const uint32_t idesc = make_idesc(128, KP, false, true);
#pragma unroll
  for (int kk = 0; kk < JP / 16; ++kk) {
    const uint64_t a = make_desc(dz_tile + kk * 2 * kLBO128, kLBO128, 128);
    const uint64_t b = make_desc(w_tile + kk * 256, 128, JP * 16);
    mma_f16(tmem_d, a, b, idesc, kk > 0);
  }
}
// G[M x N] += At^T[M x 128] * Bt[128 x N]: both tiles [128 points x cols] read MN-major, reduction over points.
// (weight gradient: At = dZ, Bt = activation, M = 64; for the 16-wide layers the roles swap and M = 128 so that the
//  ones column group behind the activation tile adds the bias-gradient row: transposed gradient)
__device__ __forceinline__ void issue_wgrad(uint32_t tmem_g, uint32_t a_tile, uint32_t b_tile, int N, bool accumulate,
                                            int M = 64) {
  const uint32_t idesc = make_idesc(M, N, true, true);
#pragma unroll
  for (int kk = 0; kk < kTile / 16; ++kk) {
    const uint64_t a = make_desc(a_tile + kk * 256, 128, kLBO128);
    const uint64_t b = make_desc(b_tile + kk * 256, 128, kLBO128);
    mma_f16(tmem_g, a, b, idesc, accumulate || kk > 0);
  }
}

__device__ __forceinline__ float elu1(float x) { return x > 0.f ? x : expm1f(x); }

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
    *reinterpret_cast<__nv_bfloat16*>(a_t + chunk_off(r, c >> 3, a_rows) + (c & 7) * 2) = __float2bfloat16_rn(A[e]);
  }
  const int b_rows = mode == 0 ? N : K, b_cols = mode == 0 ? K : N;
  for (int e = tid; e < b_rows * b_cols; e += 128) {
    const int r = e / b_cols, c = e - r * b_cols;
    *reinterpret_cast<__nv_bfloat16*>(b_t + chunk_off(r, c >> 3, b_rows) + (c & 7) * 2) = __float2bfloat16_rn(B[e]);
  }
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tbase = tslot;
  if (tid == 0) {
    const uint32_t at = smem_u32(a_t), bt = smem_u32(b_t);
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
// tile-group side helpers (r = this thread's row in the tile, 0..127)
// ---------------------------------------------------------------------------------------------------------------
template <int K0P>
__device__ __forceinline__ void load_features(const float* __restrict__ feat, long long stride, long long gp, long long n,
                                              int in0, bool vec_ok, int r, uint8_t* x0) {
  if (vec_ok) {                                   // in0 == K0P, 16-byte aligned rows
    const float4* src = reinterpret_cast<const float4*>(feat + gp * stride);
    float4 q[K0P / 4];
#pragma unroll
    for (int i = 0; i < K0P / 4; ++i) q[i] = gp < n ? __ldg(src + i) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int cg = 0; cg < K0P / 8; ++cg) {
      uint4 o;
      o.x = pack_bf16(q[2 * cg].x, q[2 * cg].y); o.y = pack_bf16(q[2 * cg].z, q[2 * cg].w);
      o.z = pack_bf16(q[2 * cg + 1].x, q[2 * cg + 1].y); o.w = pack_bf16(q[2 * cg + 1].z, q[2 * cg + 1].w);
      *reinterpret_cast<uint4*>(x0 + chunk_off(r, cg, kTile)) = o;
    }
  } else {
#pragma unroll
    for (int cg = 0; cg < K0P / 8; ++cg) {
      float v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int k = cg * 8 + i;
        v[i] = (gp < n && k < in0) ? __ldg(feat + gp * stride + k) : 0.f;
      }
      store_chunk(x0, r, cg, kTile, v);
    }
  }
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
      o.x = pack_bf16_relu(p[0], p[1]); o.y = pack_bf16_relu(p[2], p[3]);
      o.z = pack_bf16_relu(p[4], p[5]); o.w = pack_bf16_relu(p[6], p[7]);
      *reinterpret_cast<uint4*>(tile + chunk_off(r, half * 4 + cg, kTile)) = o;
    }
  }
}

// dA (64 accumulator columns) * [activation > 0] -> bf16 dZ written over the activation tile itself
__device__ __forceinline__ void masked_dz_inplace64(uint32_t taddr, int r, uint8_t* tile) {
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    float v[32];
    tmem_ld<32>(taddr + half * 32, v);
#pragma unroll
    for (int cg = 0; cg < 4; ++cg) {
      uint4* q = reinterpret_cast<uint4*>(tile + chunk_off(r, half * 4 + cg, kTile));
      const uint4 h = *q;
      const float* p = v + cg * 8;
      uint4 o;
      o.x = mask_pos_bf16x2(pack_bf16(p[0], p[1]), h.x); o.y = mask_pos_bf16x2(pack_bf16(p[2], p[3]), h.y);
      o.z = mask_pos_bf16x2(pack_bf16(p[4], p[5]), h.z); o.w = mask_pos_bf16x2(pack_bf16(p[6], p[7]), h.w);
      *q = o;
    }
  }
}

// colour-net input tile: [15 features | direction encoding | (optionally a 1.0 at column 15+dv) | 0 ...]
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
    store_chunk(cin, r, cg, kTile, v);
  }
}

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
// forward
// ---------------------------------------------------------------------------------------------------------------
template <int K0P, int KCP, int G>
struct FwdSmem {
  static constexpr int off_bias = WOfs<K0P, KCP>::total;
  static constexpr int off_ones16 = off_bias + BOfs::total;
  static constexpr int off_buf = off_ones16 + kTile * 16 * 2;
  static constexpr int buf_bytes = kTile * 64 * 2;
  static constexpr int off_bar = off_buf + G * buf_bytes;
  static constexpr int total = off_bar + 2 * G * 8 + 16;
};

// TRACE: clock64 stamps of group 0 / its MMA warp in CTA 0 (debug entry point hbr_debug_mlp_trace; compiled out otherwise)
template <int K0P, int KCP, int G, bool TRACE = false>
__global__ void __launch_bounds__(G * (kTile + 32), 1)
mlp_fwd_tc_kernel(const float* __restrict__ feat, long long feat_stride, const float* __restrict__ dirs, long long dir_group,
                  long long n, const float* __restrict__ params, int in0, int dv, float* __restrict__ out,
                  long long* __restrict__ trace = nullptr) {
  using SM = FwdSmem<K0P, KCP, G>;
  using WO = WOfs<K0P, KCP>;
  constexpr int kCols = G * 64 <= 64 ? 64 : (G * 64 <= 128 ? 128 : (G * 64 <= 256 ? 256 : 512));
  extern __shared__ __align__(128) uint8_t sm[];
  const MlpLayout m = make_layout(in0, dv);
  uint8_t* wsm = sm;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + SM::off_bar);      // full[0..G), done[0..G)
  uint32_t* tslot = reinterpret_cast<uint32_t*>(bars + 2 * G);
  const int warp = __shfl_sync(kFull, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;

  if (warp == 0) tmem_alloc<kCols>(tslot);
  if (threadIdx.x == 32) {
    for (int g = 0; g < G; ++g) { mbar_init(bars + g, kTile); mbar_init(bars + G + g, 1); }
    fence_mbar_init();
  }
  stage_weights_bf16<K0P, KCP>(params, m, wsm, sm + SM::off_bias, sm + SM::off_ones16);
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tbase = *tslot;
  const long long ntiles = (n + kTile - 1) / kTile;
  const long long nslots = (long long)gridDim.x * G;

  if (warp >= 4 * G) {
    // ===== MMA issuer of group g: converged warp, straight-line layer sequence, one elected lane issues =====
    const int g = warp - 4 * G;
    const uint32_t tb = __shfl_sync(kFull, tbase, 0);
    const uint32_t wa = smem_u32(wsm), ba = smem_u32(sm + SM::off_bias), o16a = smem_u32(sm + SM::off_ones16);
    const uint32_t d = tb + g * 64;
    const uint32_t a = smem_u32(sm + SM::off_buf + g * SM::buf_bytes);
    uint64_t* full = bars + g;
    uint64_t* done = bars + G + g;
    uint32_t par = 0;
    int tmi = 0;
    (void)tmi;
#define HBR_MMA_STAGE(BODY)                                                                             \
  do {                                                                                                  \
    mbar_wait(full, par);                                                                               \
    par ^= 1;                                                                                           \
    if (TRACE && g == 0 && blockIdx.x == 0 && lane == 0 && tmi < 500) trace[1024 + tmi++] = clock64();  \
    fence_after_sync();                                                                                 \
    if (elect_one()) {                                                                                  \
      BODY;                                                                                             \
      commit(done);                                                                                     \
    }                                                                                                   \
    __syncwarp();                                                                                       \
    if (TRACE && g == 0 && blockIdx.x == 0 && lane == 0 && tmi < 500) trace[1024 + tmi++] = clock64();  \
  } while (0)
    for (long long tile = (long long)g * gridDim.x + blockIdx.x; tile < ntiles; tile += nslots) {
      HBR_MMA_STAGE(issue_layer(d, a, wa + WO::w0, o16a, ba + BOfs::ofs(0), 64, K0P));
      HBR_MMA_STAGE(issue_layer(d, a, wa + WO::w1, o16a, ba + BOfs::ofs(1), 64, 64));
      HBR_MMA_STAGE(issue_layer(d, a, wa + WO::w2, o16a, ba + BOfs::ofs(2), 16, 64));
      HBR_MMA_STAGE(issue_layer(d, a, wa + WO::w3, o16a, ba + BOfs::ofs(3), 64, KCP));
      HBR_MMA_STAGE(issue_layer(d, a, wa + WO::w4, o16a, ba + BOfs::ofs(4), 64, 64));
      HBR_MMA_STAGE(issue_layer(d, a, wa + WO::w5, o16a, ba + BOfs::ofs(5), 16, 64));
    }
  } else {
    // ===== tile group =====
    const int g = warp >> 2;
    const int r = threadIdx.x & (kTile - 1);
    uint8_t* buf = sm + SM::off_buf + g * SM::buf_bytes;
    uint64_t* full = bars + g;
    uint64_t* done = bars + G + g;
    const uint32_t taddr = tbase + ((uint32_t)((warp & 3) * 32) << 16) + g * 64;
    uint32_t dphase = 0;
    const bool vec_ok = in0 == K0P && (feat_stride & 3) == 0 && ((uintptr_t)feat & 15) == 0;
    int tgi = 0;
    (void)tgi;
#define TR()                                                                                  \
  do {                                                                                        \
    if (TRACE && blockIdx.x == 0 && threadIdx.x == 0 && tgi < 1000) trace[tgi++] = clock64(); \
  } while (0)
    for (long long tile = (long long)g * gridDim.x + blockIdx.x; tile < ntiles; tile += nslots) {
      const long long gp = tile * kTile + r;
      const bool valid = gp < n;
      TR();
      load_features<K0P>(feat, feat_stride, gp, n, in0, vec_ok, r, buf);
      TR(); HBR_SIGNAL(); TR(); HBR_WAIT(); TR();
      relu_epilogue64(taddr, r, buf);
      TR(); HBR_SIGNAL(); TR(); HBR_WAIT(); TR();
      relu_epilogue64(taddr, r, buf);
      TR(); HBR_SIGNAL(); TR(); HBR_WAIT(); TR();
      float o16[16];
      tmem_ld<16>(taddr, o16);
      const float density = o16[0] > 0.f ? o16[0] : 0.01f * o16[0];     // LeakyReLU (test_hash.py:62)
      build_cin<KCP, false>(o16, dirs, valid ? gp / dir_group : 0, dv, valid, r, buf);
      TR(); HBR_SIGNAL(); TR(); HBR_WAIT(); TR();
      relu_epilogue64(taddr, r, buf);
      TR(); HBR_SIGNAL(); TR(); HBR_WAIT(); TR();
      relu_epilogue64(taddr, r, buf);
      TR(); HBR_SIGNAL(); TR(); HBR_WAIT(); TR();
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
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<kCols>(tbase);
}

// ---------------------------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------------------------
constexpr int kCg = kTile * 16;                                        // bytes of one 8-column group of a 128-row tile

template <int K0P, int KCP, int G>
struct BwdSmem {
  // KCP == 48 (15 + d_view <= 40): column group 5 of the colour-net input tile is pure padding and doubles as the
  // first half of the 16-wide dZ tile (its partner column group sits right behind the tile)
  static constexpr bool kAliasDzs = KCP == 48;
  static constexpr int off_bias = WOfs<K0P, KCP>::total;
  static constexpr int off_ones16 = off_bias + BOfs::total;
  static constexpr int off_onesb = off_ones16 + kTile * 16 * 2;         // one column group of 1.0 (B operand, N = 8)
  static constexpr int off_grp = off_onesb + kCg;
  // per group; h2 and c2 are each followed by a ones column group and >= 14 KB of further tiles (M = 128 operand)
  static constexpr int h2 = 0;
  static constexpr int c2 = h2 + 9 * kCg;
  static constexpr int x0 = c2 + 9 * kCg;
  static constexpr int h1 = x0 + kTile * K0P * 2;
  static constexpr int c1 = h1 + 8 * kCg;
  static constexpr int cin = c1 + 8 * kCg;
  static constexpr int dzs = kAliasDzs ? cin + 5 * kCg : cin + kTile * KCP * 2;
  static constexpr int grp_bytes = dzs + 2 * kCg;
  static constexpr int off_bar = off_grp + G * grp_bytes;
  static constexpr int total = off_bar + 2 * G * 8 + 16;
  static_assert(total <= 232448, "shared memory budget exceeded");
};

// TMEM columns: [0, 128) work accumulators of the (up to two) groups; then the gradient accumulators:
//   layers 0,1,3,4 (M = 64): G[j][k], rows = output neuron, KP columns, followed by 8 bias-gradient columns
//     (layer 3 with KCP == 48 has its bias gradient in column 15 + d_view instead);
//   layers 2,5 (M = 128): transposed G^T[k][j], 16 columns, rows 0..63 = input index, row 64 = bias gradient.
template <int K0P, int KCP>
struct BwdTmem {
  static constexpr bool kCinOne = KCP == 48;
  static constexpr int g0 = 128, b0 = g0 + K0P;
  static constexpr int g1 = b0 + 8, b1 = g1 + 64;
  static constexpr int g2 = b1 + 8;
  static constexpr int g3 = g2 + 16, b3 = g3 + KCP;
  static constexpr int g4 = b3 + (kCinOne ? 0 : 8), b4 = g4 + 64;
  static constexpr int g5 = b4 + 8;
  static constexpr int end = g5 + 16;
  static_assert(end <= 512, "TMEM budget exceeded");
};

template <int K0P, int KCP, int G>
__global__ void __launch_bounds__(G * kTile + 32, 1)
mlp_bwd_tc_kernel(const float* __restrict__ feat, long long feat_stride, const float* __restrict__ dirs, long long dir_group,
                  long long n, const float* __restrict__ params, int in0, int dv, const float* __restrict__ out,
                  const float* __restrict__ dout, float* __restrict__ dfeat, long long dfeat_stride,
                  float* __restrict__ ddirs, float* __restrict__ dparams) {
  using SM = BwdSmem<K0P, KCP, G>;
  using WO = WOfs<K0P, KCP>;
  using TM = BwdTmem<K0P, KCP>;
  constexpr bool kCinOne = TM::kCinOne;
  static_assert(G >= 1 && G <= 2, "two work accumulators");
  extern __shared__ __align__(128) uint8_t sm[];
  const MlpLayout m = make_layout(in0, dv);
  uint8_t* wsm = sm;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + SM::off_bar);
  uint32_t* tslot = reinterpret_cast<uint32_t*>(bars + 2 * G);
  const int warp = __shfl_sync(kFull, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;

  if (warp == 0) tmem_alloc<512>(tslot);
  if (threadIdx.x == 32) {
    for (int g = 0; g < G; ++g) { mbar_init(bars + g, kTile); mbar_init(bars + G + g, 1); }
    fence_mbar_init();
  }
  stage_weights_bf16<K0P, KCP>(params, m, wsm, sm + SM::off_bias, sm + SM::off_ones16);
  {
    const uint32_t one2 = pack_bf16(1.f, 1.f);
    const uint4 ones4 = make_uint4(one2, one2, one2, one2);
    for (int e = threadIdx.x; e < kTile; e += blockDim.x) {
      reinterpret_cast<uint4*>(sm + SM::off_onesb)[e] = ones4;
      for (int g = 0; g < G; ++g) {
        uint8_t* gb = sm + SM::off_grp + g * SM::grp_bytes;
        reinterpret_cast<uint4*>(gb + SM::h2 + 8 * kCg)[e] = ones4;
        reinterpret_cast<uint4*>(gb + SM::c2 + 8 * kCg)[e] = ones4;
      }
    }
    // the M = 128 operands read 7 column groups past the ones group: keep that memory finite from the start
    for (int e = threadIdx.x; e < G * SM::grp_bytes / 16; e += blockDim.x) {
      const int g = e / (SM::grp_bytes / 16), o = (e - g * (SM::grp_bytes / 16)) * 16;
      if (o >= SM::x0) *reinterpret_cast<uint4*>(sm + SM::off_grp + g * SM::grp_bytes + o) = make_uint4(0, 0, 0, 0);
    }
  }
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tbase = *tslot;
  const long long ntiles = (n + kTile - 1) / kTile;
  const long long nslots = (long long)gridDim.x * G;
  long long nt[G];
  long long cta_tiles = 0;
#pragma unroll
  for (int g = 0; g < G; ++g) {
    nt[g] = tiles_of_slot(ntiles, (long long)g * gridDim.x + blockIdx.x, nslots);
    cta_tiles += nt[g];
  }

  if (warp == 4 * G) {
    // ===== the CTA's only MMA issuer: converged warp, fixed visiting order (stage-major, group-minor), straight-line
    //       issue code; every gradient-accumulating MMA of the CTA comes from this one thread sequence =====
    const uint32_t tb0 = __shfl_sync(kFull, tbase, 0);
    const uint32_t sm0 = smem_u32(sm);
    uint32_t par[G];
#pragma unroll
    for (int g = 0; g < G; ++g) par[g] = 0;
    bool first = true;                           // gradient accumulators not yet written
    // The operand descriptors are rebuilt from two laundered base values in every stage: left to itself the compiler
    // hoists all ~400 loop-invariant descriptors out of the tile loop and spills them.
#define HBR_BWD_STAGE(BODY)                                                \
  _Pragma("unroll") for (int g = 0; g < G; ++g) {                          \
    if (k < nt[g]) {                                                       \
      mbar_wait(bars + g, par[g]);                                         \
      par[g] ^= 1;                                                         \
      fence_after_sync();                                                  \
      if (elect_one()) {                                                   \
        uint32_t sb = sm0, tb = tb0;                                       \
        asm volatile("" : "+r"(sb), "+r"(tb));                             \
        const uint32_t wa = sb, ba = sb + SM::off_bias, o16a = sb + SM::off_ones16, onesb = sb + SM::off_onesb; \
        const uint32_t d = tb + g * 64;                                    \
        const uint32_t base = sb + SM::off_grp + g * SM::grp_bytes;        \
        const uint32_t x0a = base + SM::x0, h1a = base + SM::h1, h2a = base + SM::h2, cina = base + SM::cin, \
                       c1a = base + SM::c1, c2a = base + SM::c2, dzsa = base + SM::dzs; \
        (void)wa; (void)ba; (void)o16a; (void)onesb;                       \
        (void)x0a; (void)h1a; (void)h2a; (void)cina; (void)c1a; (void)c2a; (void)dzsa; \
        const bool acc = !(first && g == 0);                               \
        (void)acc;                                                         \
        BODY;                                                              \
        commit(bars + G + g);                                              \
      }                                                                    \
      __syncwarp();                                                        \
    }                                                                      \
  }
    const long long kmax = nt[0];                // slot of group 0 never has fewer tiles than a later group's
    for (long long k = 0; k < kmax; ++k) {
      // ---- forward recompute (layers 0..4) ----
      HBR_BWD_STAGE(issue_layer(d, x0a, wa + WO::w0, o16a, ba + BOfs::ofs(0), 64, K0P));
      HBR_BWD_STAGE(issue_layer(d, h1a, wa + WO::w1, o16a, ba + BOfs::ofs(1), 64, 64));
      HBR_BWD_STAGE(issue_layer(d, h2a, wa + WO::w2, o16a, ba + BOfs::ofs(2), 16, 64));
      HBR_BWD_STAGE(issue_layer(d, cina, wa + WO::w3, o16a, ba + BOfs::ofs(3), 64, KCP));
      HBR_BWD_STAGE(issue_layer(d, c1a, wa + WO::w4, o16a, ba + BOfs::ofs(4), 64, 64));
      // ---- backward: dgrad into the work accumulator, weight/bias gradients into the resident accumulators ----
      HBR_BWD_STAGE({   // col_model.4: dZ = dzs (16 wide), input c2 (+ ones group): transposed gradient, M = 128
        issue_dgrad(d, dzsa, wa + WO::w5, 16, 64);
        issue_wgrad(tb + TM::g5, c2a, dzsa, 16, acc, 128);
      });
      HBR_BWD_STAGE({   // col_model.2: dZ in the c2 tile, input c1
        issue_dgrad(d, c2a, wa + WO::w4, 64, 64);
        issue_wgrad(tb + TM::g4, c2a, c1a, 64, acc);
        issue_wgrad(tb + TM::b4, c2a, onesb, 8, acc);
      });
      HBR_BWD_STAGE({   // col_model.0: dZ in the c1 tile, input cin (bias gradient through the planted 1.0 column)
        issue_dgrad(d, c1a, wa + WO::w3, 64, KCP);
        issue_wgrad(tb + TM::g3, c1a, cina, KCP, acc);
        if (!kCinOne) issue_wgrad(tb + TM::b3, c1a, onesb, 8, acc);
      });
      HBR_BWD_STAGE({   // sig_model.4: dZ = dzs (16 wide), input h2 (+ ones group): transposed
        issue_dgrad(d, dzsa, wa + WO::w2, 16, 64);
        issue_wgrad(tb + TM::g2, h2a, dzsa, 16, acc, 128);
      });
      HBR_BWD_STAGE({   // sig_model.2: dZ in the h2 tile, input h1
        issue_dgrad(d, h2a, wa + WO::w1, 64, 64);
        issue_wgrad(tb + TM::g1, h2a, h1a, 64, acc);
        issue_wgrad(tb + TM::b1, h2a, onesb, 8, acc);
      });
      HBR_BWD_STAGE({   // sig_model.0: dZ in the h1 tile, input x0
        issue_dgrad(d, h1a, wa + WO::w0, 64, K0P);
        issue_wgrad(tb + TM::g0, h1a, x0a, K0P, acc);
        issue_wgrad(tb + TM::b0, h1a, onesb, 8, acc);
      });
      first = false;
    }
  } else {
    // ===== tile group =====
    const int g = warp >> 2;
    const int r = threadIdx.x & (kTile - 1);
    uint8_t* gb = sm + SM::off_grp + g * SM::grp_bytes;
    uint8_t *x0 = gb + SM::x0, *h1 = gb + SM::h1, *h2 = gb + SM::h2, *cin = gb + SM::cin, *c1 = gb + SM::c1,
            *c2 = gb + SM::c2, *dzs = gb + SM::dzs;
    uint64_t* full = bars + g;
    uint64_t* done = bars + G + g;
    const uint32_t taddr = tbase + ((uint32_t)((warp & 3) * 32) << 16) + g * 64;
    uint32_t dphase = 0;
    const bool vec_ok = in0 == K0P && (feat_stride & 3) == 0 && ((uintptr_t)feat & 15) == 0;
    const bool dvec_ok = dfeat != nullptr && in0 == K0P && (dfeat_stride & 3) == 0 && ((uintptr_t)dfeat & 15) == 0;
    for (long long tile = (long long)g * gridDim.x + blockIdx.x; tile < ntiles; tile += nslots) {
      const long long gp = tile * kTile + r;
      const bool valid = gp < n;
      float4 fo = make_float4(0.f, 0.f, 0.f, 0.f), go = fo;           // saved forward output, upstream gradient
      if (valid) {
        fo = __ldg(reinterpret_cast<const float4*>(out + gp * 4));
        go = __ldg(reinterpret_cast<const float4*>(dout + gp * 4));
      }
      // ---- recompute the forward activations ----
      load_features<K0P>(feat, feat_stride, gp, n, in0, vec_ok, r, x0);
      HBR_SIGNAL(); HBR_WAIT();
      relu_epilogue64(taddr, r, h1);
      HBR_SIGNAL(); HBR_WAIT();
      relu_epilogue64(taddr, r, h2);
      HBR_SIGNAL(); HBR_WAIT();
      const long long dir_row = valid ? gp / dir_group : 0;
      {
        float o16[16];
        tmem_ld<16>(taddr, o16);
        build_cin<KCP, kCinOne>(o16, dirs, dir_row, dv, valid, r, cin);
      }
      HBR_SIGNAL(); HBR_WAIT();
      relu_epilogue64(taddr, r, c1);
      HBR_SIGNAL(); HBR_WAIT();
      relu_epilogue64(taddr, r, c2);
      {
        // d(rgb_pre) = g * ELU'(pre), with ELU'(pre) = pre > 0 ? 1 : exp(pre) = elu(pre) + 1 from the saved output
        float dz16[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) dz16[c] = 0.f;
        dz16[0] = go.x * (fo.x > 0.f ? 1.f : fo.x + 1.f);
        dz16[1] = go.y * (fo.y > 0.f ? 1.f : fo.y + 1.f);
        dz16[2] = go.z * (fo.z > 0.f ? 1.f : fo.z + 1.f);
        store_chunk(dzs, r, 0, kTile, dz16);
        store_chunk(dzs, r, 1, kTile, dz16 + 8);
      }
      HBR_SIGNAL(); HBR_WAIT();                  // col_model.4 backward done: work = dA(c2)
      masked_dz_inplace64(taddr, r, c2);
      HBR_SIGNAL(); HBR_WAIT();                  // col_model.2 done: work = dA(c1)
      masked_dz_inplace64(taddr, r, c1);
      HBR_SIGNAL(); HBR_WAIT();                  // col_model.0 done: work[0,KCP) = d(cin)
      {
        float dc[KCP], dz16[16];
        tmem_ld<KCP>(taddr, dc);
        dz16[0] = go.w * (fo.w > 0.f ? 1.f : 0.01f);                    // LeakyReLU' from the saved density
#pragma unroll
        for (int k = 0; k < kFeat; ++k) dz16[1 + k] = dc[k];
        store_chunk(dzs, r, 0, kTile, dz16);
        store_chunk(dzs, r, 1, kTile, dz16 + 8);
        if (ddirs != nullptr) {
          // rows of one warp usually belong to one ray: reduce over the warp first, one atomic per column
          const long long row0 = __shfl_sync(kFull, dir_row, 0);
          const bool uniform = __all_sync(kFull, dir_row == row0 && valid);
#pragma unroll
          for (int k = kFeat; k < KCP; ++k) {
            if (k < kFeat + dv) {
              if (uniform) {
                const float s = warp_sum(dc[k]);
                if (lane == 0) atomicAdd(ddirs + row0 * dv + (k - kFeat), s);
              } else if (valid) {
                atomicAdd(ddirs + dir_row * dv + (k - kFeat), dc[k]);
              }
            }
          }
        }
      }
      HBR_SIGNAL(); HBR_WAIT();                  // sig_model.4 done: work = dA(h2)
      masked_dz_inplace64(taddr, r, h2);
      HBR_SIGNAL(); HBR_WAIT();                  // sig_model.2 done: work = dA(h1)
      masked_dz_inplace64(taddr, r, h1);
      HBR_SIGNAL(); HBR_WAIT();                  // sig_model.0 done: work[0,K0P) = d(feat)
      if (dfeat != nullptr) {
        float df[K0P];
        tmem_ld<K0P>(taddr, df);
        if (valid) {
          if (dvec_ok) {
            float4* dst = reinterpret_cast<float4*>(dfeat + gp * dfeat_stride);
#pragma unroll
            for (int i = 0; i < K0P / 4; ++i) dst[i] = make_float4(df[4 * i], df[4 * i + 1], df[4 * i + 2], df[4 * i + 3]);
          } else {
#pragma unroll
            for (int k = 0; k < K0P; ++k)
              if (k < in0) dfeat[gp * dfeat_stride + k] = df[k];
          }
        }
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();

  // ---- flush the gradient accumulators ----
  if (cta_tiles > 0 && dparams != nullptr && warp < 4) {
    const uint32_t trow = tbase + ((uint32_t)(warp * 32) << 16);
    // layers 0,1,3,4 -- M = 64 layout: accumulator row q (= output neuron) lives in lane (q % 16) + 32 * (q / 16)
    {
      const int q = warp * 16 + lane;            // meaningful for lane < 16
      const int gcol[4] = {TM::g0, TM::g1, TM::g3, TM::g4};
      const int bcol[4] = {TM::b0, TM::b1, TM::b3, TM::b4};
      const int li[4] = {0, 1, 3, 4};
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int i = li[t];
        float gacc[64], gbias[16];
        const int KP = i == 0 ? K0P : (i == 3 ? KCP : 64);
        if (KP > 48) tmem_ld<64>(trow + gcol[t], gacc);
        else if (KP > 32) tmem_ld<48>(trow + gcol[t], gacc);
        else tmem_ld<32>(trow + gcol[t], gacc);
        if (i == 3 && kCinOne) gbias[0] = 0.f;
        else tmem_ld<16>(trow + bcol[t], gbias);    // 8 valid columns; the excess belongs to the next accumulator
        if (lane < 16 && q < m.J[i]) {
          float bsum = gbias[0];
          for (int k = 0; k < KP; ++k) {
            if (k < m.K[i]) atomicAdd(dparams + m.W[i] + q * m.K[i] + k, gacc[k]);
            if (i == 3 && kCinOne && k == m.K[i]) bsum = gacc[k];
          }
          atomicAdd(dparams + m.b[i] + q, bsum);
        }
      }
    }
    // layers 2,5 -- M = 128 layout: accumulator row = lane; rows 0..63 = input index k, row 64 = bias gradient
    {
      const int row = warp * 32 + lane;
      const int gcol[2] = {TM::g2, TM::g5};
      const int li[2] = {2, 5};
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const int i = li[t];
        float gacc[16];
        tmem_ld<16>(trow + gcol[t], gacc);
        if (row < 64) {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (j < m.J[i]) atomicAdd(dparams + m.W[i] + j * 64 + row, gacc[j]);
        } else if (row == 64) {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (j < m.J[i]) atomicAdd(dparams + m.b[i] + j, gacc[j]);
        }
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tbase);
}

}  // namespace hbr

using namespace hbr;

extern "C" int hbr_debug_umma(int mode, const float* A, const float* B, float* D, int N, int K, void* stream) {
  HBR_REQUIRE(mode >= 0 && mode <= 2, "mode %d", mode);
  HBR_REQUIRE(N % 16 == 0 && N >= 16 && N <= 64 && K % 16 == 0 && K >= 16 && K <= 128, "N=%d K=%d", N, K);
  HBR_REQUIRE(mode != 2 || K == 128, "mode 2 needs K=128");
  HBR_CUDA(cudaFuncSetAttribute(umma_debug_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  umma_debug_kernel<<<1, 128, 65536, as_stream(stream)>>>(mode, A, B, D, N, K);
  HBR_LAUNCH_CHECK();
  return HBR_OK;
}

extern "C" int hbr_debug_mlp_trace(const float* feat, const float* dirs, int64_t dir_group, int64_t n, const float* params,
                                   float* out, long long* trace, void* stream) {
  constexpr int smem = FwdSmem<32, 48, 4>::total;
  HBR_CUDA(cudaFuncSetAttribute(mlp_fwd_tc_kernel<32, 48, 4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int grid = (int)min64(ceil_div(ceil_div(n, kTile), 4), sm_count());
  mlp_fwd_tc_kernel<32, 48, 4, true><<<grid, 4 * (kTile + 32), smem, as_stream(stream)>>>(feat, 32, dirs, dir_group, n,
                                                                                         params, 32, 24, out, trace);
  HBR_LAUNCH_CHECK();
  return HBR_OK;
}

// The reference's configuration (in0 = 32, d_view = 24) gets the widest pipelines (4 forward / 2 backward tile groups
// per SM); other widths (in0 <= 64, 15 + d_view <= 64) run the same kernels with padded K and fewer groups.
static inline bool narrow_shape(const hbr_mlp_dims* d) { return d->in0 <= 32 && d->d_view + kFeat <= 40; }

extern "C" int hbr_mlp_fwd_tc(const float* feat, int64_t feat_stride, const float* dirs, int64_t dir_group, int64_t n,
                              const float* params, const hbr_mlp_dims* dims, float* out, void* stream) {
  if (int rc = check_dims(dims)) return rc;
  if (n == 0) return HBR_OK;
  HBR_REQUIRE(feat && dirs && params && out, "NULL pointer");
  HBR_REQUIRE(feat_stride >= dims->in0 && dir_group >= 1, "bad stride / dir_group");
  HBR_REQUIRE((uintptr_t)out % 16 == 0, "out must be 16-byte aligned");
  cudaStream_t st = as_stream(stream);
  const int64_t ntiles = ceil_div(n, kTile);
  const int in0 = dims->in0, dv = dims->d_view;
  const int grid = (int)min64(ceil_div(ntiles, 4), sm_count());
  if (narrow_shape(dims)) {
    constexpr int smem = FwdSmem<32, 48, 4>::total;
    HBR_CUDA(cudaFuncSetAttribute(mlp_fwd_tc_kernel<32, 48, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    mlp_fwd_tc_kernel<32, 48, 4><<<grid, 4 * (kTile + 32), smem, st>>>(feat, feat_stride, dirs, dir_group, n, params, in0, dv,
                                                                     out);
  } else {
    constexpr int smem = FwdSmem<64, 64, 4>::total;
    HBR_CUDA(cudaFuncSetAttribute(mlp_fwd_tc_kernel<64, 64, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    mlp_fwd_tc_kernel<64, 64, 4><<<grid, 4 * (kTile + 32), smem, st>>>(feat, feat_stride, dirs, dir_group, n, params, in0, dv,
                                                                     out);
  }
  HBR_LAUNCH_CHECK();
  return HBR_OK;
}

extern "C" int hbr_mlp_bwd_tc(const float* feat, int64_t feat_stride, const float* dirs, int64_t dir_group, int64_t n,
                              const float* params, const hbr_mlp_dims* dims, const float* out, const float* dout,
                              float* dfeat, int64_t dfeat_stride, float* ddirs, float* dparams, void* stream) {
  if (int rc = check_dims(dims)) return rc;
  if (n == 0) return HBR_OK;
  HBR_REQUIRE(feat && dirs && params && out && dout, "NULL pointer");
  HBR_REQUIRE(feat_stride >= dims->in0 && dir_group >= 1, "bad stride / dir_group");
  HBR_REQUIRE((uintptr_t)dout % 16 == 0 && (uintptr_t)out % 16 == 0, "out / dout must be 16-byte aligned");
  HBR_REQUIRE(!dfeat || dfeat_stride >= dims->in0, "dfeat_stride too small");
  cudaStream_t st = as_stream(stream);
  const int64_t ntiles = ceil_div(n, kTile);
  const int in0 = dims->in0, dv = dims->d_view;
  if (narrow_shape(dims)) {
    constexpr int smem = BwdSmem<32, 48, 2>::total;
    const int grid = (int)min64(ceil_div(ntiles, 2), sm_count());
    HBR_CUDA(cudaFuncSetAttribute(mlp_bwd_tc_kernel<32, 48, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    mlp_bwd_tc_kernel<32, 48, 2><<<grid, 2 * kTile + 32, smem, st>>>(feat, feat_stride, dirs, dir_group, n, params, in0, dv,
                                                                   out, dout, dfeat, dfeat_stride, ddirs, dparams);
  } else {
    constexpr int smem = BwdSmem<64, 64, 1>::total;
    const int grid = (int)min64(ntiles, sm_count());
    HBR_CUDA(cudaFuncSetAttribute(mlp_bwd_tc_kernel<64, 64, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    mlp_bwd_tc_kernel<64, 64, 1><<<grid, kTile + 32, smem, st>>>(feat, feat_stride, dirs, dir_group, n, params, in0, dv, out,
                                                               dout, dfeat, dfeat_stride, ddirs, dparams);
  }
  HBR_LAUNCH_CHECK();
  return HBR_OK;
}
