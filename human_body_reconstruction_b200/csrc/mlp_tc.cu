// MLP_3D on the 5th-generation tensor cores (tcgen05 + TMEM): bf16 operands, fp32 accumulation.
// This is the field evaluation of the training step (train_hash2.py:218-226 runs it under autocast); the fp32
// CUDA-core version in mlp_simt.cu serves nerf2mesh and the 1e-5 parity tests.
//
// Structure (both kernels): one persistent CTA per SM = G "tile groups" of 128 threads + one MMA-issuing warp.
//   * A tile group owns one 128-point tile at a time (thread = point = TMEM lane) and walks the layer chain:
//     it writes the next layer's bf16 A tile into shared memory, signals `full[g]` (mbarrier, 128 arrivals),
//     waits on `done[g]`, pulls its accumulator row out of TMEM with tcgen05.ld and applies bias/activation.
//   * Lane 0 of the MMA warp polls the G `full` barriers, issues the tcgen05.mma sequence of whichever group
//     is ready (A = activation tile, B = weight tile, D = that group's 64 TMEM columns) and commits it to
//     `done[g]`.  The chain of one tile is strictly serial, so throughput comes from G tiles in flight per SM:
//     while one group waits for the tensor pipe the others run their epilogues.
//   * Every MMA is issued by the same thread, so the weight-gradient accumulators that all tiles of a CTA share
//     (backward) are updated in issue order.
// Backward recomputes the forward activations in shared memory (nothing but the features is re-read from HBM),
// then walks the layers in reverse: per layer one dgrad GEMM (dA = dZ W, B = the SAME weight tile read MN-major),
// one weight-gradient GEMM (reduction over the 128 points, both operands read MN-major from tiles already in
// smem) and one bias-gradient GEMM against a ones tile; the gradient accumulators stay resident in TMEM across
// all tiles of the CTA and are flushed once with atomics.  dZ of a layer is written IN PLACE over the
// activation tile whose consumer GEMMs have completed, so one group needs 88 KB and two groups fit an SM.
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

// fp32 (J,K) row-major weights -> bf16 canonical tile [JP rows x KP cols], zero padded; one 16-byte chunk per step
template <int K0P, int KCP>
__device__ __forceinline__ void stage_weights_bf16(const float* __restrict__ params, const MlpLayout& m, uint8_t* wsm,
                                                   float* bias_sm) {
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
    for (int j = threadIdx.x; j < 64; j += blockDim.x) bias_sm[i * 64 + j] = j < J ? __ldg(params + m.b[i] + j) : 0.f;
  }
}

// D[128 x N] (+)= A[128 x K] * W^T : A K-major activation tile, B K-major weight tile (forward)
__device__ __forceinline__ void issue_fwd(uint32_t tmem_d, uint32_t a_tile, uint32_t w_tile, int JP, int KP) {
  const uint32_t idesc = make_idesc(128, JP, false, false);
  for (int kk = 0; kk < KP / 16; ++kk) {
    const uint64_t a = make_desc(a_tile + kk * 2 * kLBO128, kLBO128, 128);
    const uint64_t b = make_desc(w_tile + kk * 2 * JP * 16, JP * 16, 128);
    mma_f16(tmem_d, a, b, idesc, kk > 0);
  }
}
// D[128 x KP] = dZ[128 x JP] * W : A K-major dZ tile, B = weight tile [JP x KP] read MN-major (N = k, K = j)
__device__ __forceinline__ void issue_dgrad(uint32_t tmem_d, uint32_t dz_tile, uint32_t w_tile, int JP, int KP) {
  const uint32_t idesc = make_idesc(128, KP, false, true);
  for (int kk = 0; kk < JP / 16; ++kk) {
    const uint64_t a = make_desc(dz_tile + kk * 2 * kLBO128, kLBO128, 128);
    const uint64_t b = make_desc(w_tile + kk * 256, 128, JP * 16);
    mma_f16(tmem_d, a, b, idesc, kk > 0);
  }
}
// G[64 x N] += At^T[64 x 128] * Bt[128 x N]: both tiles [128 points x cols] read MN-major, reduction over points.
// (weight gradient: At = dZ, Bt = activation; for the 16-wide layers the roles swap, giving the transposed gradient)
__device__ __forceinline__ void issue_wgrad(uint32_t tmem_g, uint32_t a_tile, uint32_t b_tile, int N, bool accumulate) {
  const uint32_t idesc = make_idesc(64, N, true, true);
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

// bias + ReLU on 64 accumulator columns -> bf16 activation tile (two 32-column halves to bound registers)
__device__ __forceinline__ void relu_epilogue64(uint32_t taddr, const float* bias, int r, uint8_t* tile) {
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    float v[32];
    tmem_ld<32>(taddr + half * 32, v);
#pragma unroll
    for (int cg = 0; cg < 4; ++cg) {
      const float4 b0 = *reinterpret_cast<const float4*>(bias + half * 32 + cg * 8);
      const float4 b1 = *reinterpret_cast<const float4*>(bias + half * 32 + cg * 8 + 4);
      const float* p = v + cg * 8;
      uint4 o;
      o.x = pack_bf16_relu(p[0] + b0.x, p[1] + b0.y); o.y = pack_bf16_relu(p[2] + b0.z, p[3] + b0.w);
      o.z = pack_bf16_relu(p[4] + b1.x, p[5] + b1.y); o.w = pack_bf16_relu(p[6] + b1.z, p[7] + b1.w);
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

template <int KCP>
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
      else if (k < kFeat + dv && valid) x = __ldg(dirs + dir_row * dv + (k - kFeat));   // concat(viewdirs) (:66)
      v[i] = x;
    }
    store_chunk(cin, r, cg, kTile, v);
  }
}

// group -> MMA thread: "my A tile is written (and my TMEM reads are finished)";  MMA thread -> group: commit on done
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
  static constexpr int off_buf = off_bias + 6 * 64 * 4;
  static constexpr int buf_bytes = kTile * 64 * 2;
  static constexpr int off_bar = off_buf + G * buf_bytes;
  static constexpr int total = off_bar + 2 * G * 8 + 16;
};

template <int K0P, int KCP, int G>
__global__ void __launch_bounds__(G * kTile + 32, 1)
mlp_fwd_tc_kernel(const float* __restrict__ feat, long long feat_stride, const float* __restrict__ dirs, long long dir_group,
                  long long n, const float* __restrict__ params, int in0, int dv, float* __restrict__ out) {
  using SM = FwdSmem<K0P, KCP, G>;
  using WO = WOfs<K0P, KCP>;
  constexpr int kCols = G * 64 <= 64 ? 64 : (G * 64 <= 128 ? 128 : (G * 64 <= 256 ? 256 : 512));
  extern __shared__ __align__(128) uint8_t sm[];
  const MlpLayout m = make_layout(in0, dv);
  uint8_t* wsm = sm;
  float* bias = reinterpret_cast<float*>(sm + SM::off_bias);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + SM::off_bar);      // full[0..G), done[0..G)
  uint32_t* tslot = reinterpret_cast<uint32_t*>(bars + 2 * G);
  const int warp = __shfl_sync(kFull, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;

  if (warp == 0) tmem_alloc<kCols>(tslot);
  if (threadIdx.x == 32) {
    for (int g = 0; g < G; ++g) { mbar_init(bars + g, kTile); mbar_init(bars + G + g, 1); }
    fence_mbar_init();
  }
  stage_weights_bf16<K0P, KCP>(params, m, wsm, bias);
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tbase = *tslot;
  const long long ntiles = (n + kTile - 1) / kTile;
  const long long nslots = (long long)gridDim.x * G;

  if (warp == 4 * G) {
    // ===== MMA issuer: the warp runs the loop converged (uniform control flow, operands on the uniform datapath),
    //       one elected lane issues =====
    const uint32_t tb = __shfl_sync(kFull, tbase, 0);
    const uint32_t wa = smem_u32(wsm);
    int left[G], layer[G];
    uint32_t par[G];
    int remaining = 0;
#pragma unroll
    for (int g = 0; g < G; ++g) {
      left[g] = 6 * (int)tiles_of_slot(ntiles, (long long)g * gridDim.x + blockIdx.x, nslots);
      remaining += left[g];
      layer[g] = 0;
      par[g] = 0;
    }
    while (remaining > 0) {
#pragma unroll
      for (int g = 0; g < G; ++g) {
        if (left[g] > 0 && __all_sync(kFull, mbar_test_wait(bars + g, par[g]))) {
          fence_after_sync();
          if (elect_one()) {
            const uint32_t d = tb + g * 64;
            const uint32_t a = smem_u32(sm + SM::off_buf + g * SM::buf_bytes);
            if (layer[g] == 0) issue_fwd(d, a, wa + WO::w0, 64, K0P);
            else if (layer[g] == 1) issue_fwd(d, a, wa + WO::w1, 64, 64);
            else if (layer[g] == 2) issue_fwd(d, a, wa + WO::w2, 16, 64);
            else if (layer[g] == 3) issue_fwd(d, a, wa + WO::w3, 64, KCP);
            else if (layer[g] == 4) issue_fwd(d, a, wa + WO::w4, 64, 64);
            else issue_fwd(d, a, wa + WO::w5, 16, 64);
            commit(bars + G + g);
          }
          __syncwarp();
          layer[g] = layer[g] == 5 ? 0 : layer[g] + 1;
          par[g] ^= 1;
          --left[g];
          --remaining;
        }
      }
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
    for (long long tile = (long long)g * gridDim.x + blockIdx.x; tile < ntiles; tile += nslots) {
      const long long gp = tile * kTile + r;
      const bool valid = gp < n;
      load_features<K0P>(feat, feat_stride, gp, n, in0, vec_ok, r, buf);
      HBR_SIGNAL(); HBR_WAIT();
      relu_epilogue64(taddr, bias + 0, r, buf);
      HBR_SIGNAL(); HBR_WAIT();
      relu_epilogue64(taddr, bias + 64, r, buf);
      HBR_SIGNAL(); HBR_WAIT();
      float o16[16];
      tmem_ld<16>(taddr, o16);
#pragma unroll
      for (int k = 0; k < 16; ++k) o16[k] += bias[128 + k];
      const float density = o16[0] > 0.f ? o16[0] : 0.01f * o16[0];     // LeakyReLU (test_hash.py:62)
      build_cin<KCP>(o16, dirs, valid ? gp / dir_group : 0, dv, valid, r, buf);
      HBR_SIGNAL(); HBR_WAIT();
      relu_epilogue64(taddr, bias + 192, r, buf);
      HBR_SIGNAL(); HBR_WAIT();
      relu_epilogue64(taddr, bias + 256, r, buf);
      HBR_SIGNAL(); HBR_WAIT();
      float c16[16];
      tmem_ld<16>(taddr, c16);
      if (valid) {
        float4 o;
        o.x = elu1(c16[0] + bias[320]);                                 // ELU (test_hash.py:67)
        o.y = elu1(c16[1] + bias[321]);
        o.z = elu1(c16[2] + bias[322]);
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
template <int K0P, int KCP, int G>
struct BwdSmem {
  static constexpr int off_bias = WOfs<K0P, KCP>::total;
  static constexpr int off_ones = off_bias + 6 * 64 * 4;               // [128 x 64] bf16 of 1.0
  static constexpr int off_grp = off_ones + kTile * 64 * 2;
  // per group
  static constexpr int x0 = 0;
  static constexpr int h1 = x0 + kTile * K0P * 2;
  static constexpr int h2 = h1 + kTile * 64 * 2;
  static constexpr int cin = h2 + kTile * 64 * 2;
  static constexpr int c1 = cin + kTile * KCP * 2;
  static constexpr int c2 = c1 + kTile * 64 * 2;
  static constexpr int dzs = c2 + kTile * 64 * 2;                      // [128 x 16] dZ of the two 16-wide layers
  static constexpr int grp_bytes = dzs + kTile * 16 * 2;
  static constexpr int off_bar = off_grp + G * grp_bytes;
  static constexpr int total = off_bar + 2 * G * 8 + 16;
};

// TMEM columns: [0, 128) two work accumulators; then the weight-gradient accumulators (64 lanes each):
//   layers 0,1,3,4: G[j][k] (rows = output neuron, KP columns); layers 2,5 (16 outputs): transposed G^T[k][j], 16 columns
// then the bias-gradient accumulators: 8 columns (layers 0,1,3,4), 16 columns (layers 2,5, every row equal).
template <int K0P, int KCP>
struct BwdTmem {
  static constexpr int g0 = 128, g1 = g0 + K0P, g2 = g1 + 64, g3 = g2 + 16, g4 = g3 + KCP, g5 = g4 + 64;
  static constexpr int b0 = g5 + 16, b1 = b0 + 8, b2 = b1 + 8, b3 = b2 + 16, b4 = b3 + 8, b5 = b4 + 8;
  static constexpr int end = b5 + 16;
  static_assert(end <= 512, "TMEM budget exceeded");
};

template <int K0P, int KCP, int G>
__global__ void __launch_bounds__(G * kTile + 32, 1)
mlp_bwd_tc_kernel(const float* __restrict__ feat, long long feat_stride, const float* __restrict__ dirs, long long dir_group,
                  long long n, const float* __restrict__ params, int in0, int dv, const float* __restrict__ dout,
                  float* __restrict__ dfeat, long long dfeat_stride, float* __restrict__ ddirs, float* __restrict__ dparams) {
  using SM = BwdSmem<K0P, KCP, G>;
  using WO = WOfs<K0P, KCP>;
  using TM = BwdTmem<K0P, KCP>;
  extern __shared__ __align__(128) uint8_t sm[];
  const MlpLayout m = make_layout(in0, dv);
  uint8_t* wsm = sm;
  float* bias = reinterpret_cast<float*>(sm + SM::off_bias);
  uint8_t* ones = sm + SM::off_ones;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + SM::off_bar);
  uint32_t* tslot = reinterpret_cast<uint32_t*>(bars + 2 * G);
  const int warp = __shfl_sync(kFull, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;

  if (warp == 0) tmem_alloc<512>(tslot);
  if (threadIdx.x == 32) {
    for (int g = 0; g < G; ++g) { mbar_init(bars + g, kTile); mbar_init(bars + G + g, 1); }
    fence_mbar_init();
  }
  stage_weights_bf16<K0P, KCP>(params, m, wsm, bias);
  {
    const uint32_t one2 = pack_bf16(1.f, 1.f);
    for (int e = threadIdx.x; e < kTile * 64 * 2 / 16; e += blockDim.x)
      reinterpret_cast<uint4*>(ones)[e] = make_uint4(one2, one2, one2, one2);
  }
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tbase = *tslot;
  const long long ntiles = (n + kTile - 1) / kTile;
  const long long nslots = (long long)gridDim.x * G;
  long long cta_tiles = 0;
#pragma unroll
  for (int g = 0; g < G; ++g) cta_tiles += tiles_of_slot(ntiles, (long long)g * gridDim.x + blockIdx.x, nslots);

  if (warp == 4 * G) {
    // ===== MMA issuer (converged warp, one elected lane issues; see the forward kernel) =====
    const uint32_t tb = __shfl_sync(kFull, tbase, 0);
    const uint32_t wa = smem_u32(wsm), onesa = smem_u32(ones);
    int left[G], stage[G];
    uint32_t par[G];
    int remaining = 0;
#pragma unroll
    for (int g = 0; g < G; ++g) {
      left[g] = 12 * (int)tiles_of_slot(ntiles, (long long)g * gridDim.x + blockIdx.x, nslots);
      remaining += left[g];
      stage[g] = 0;
      par[g] = 0;
    }
    uint32_t inited = 0;                        // bit i: layer i's gradient accumulators hold a first contribution
    while (remaining > 0) {
#pragma unroll
      for (int g = 0; g < G; ++g) {
        if (left[g] > 0 && __all_sync(kFull, mbar_test_wait(bars + g, par[g]))) {
          fence_after_sync();
          const int stg = stage[g];
          if (elect_one()) {
            const uint32_t d = tb + g * 64;
            const uint32_t base = smem_u32(sm + SM::off_grp + g * SM::grp_bytes);
            const uint32_t x0a = base + SM::x0, h1a = base + SM::h1, h2a = base + SM::h2, cina = base + SM::cin,
                           c1a = base + SM::c1, c2a = base + SM::c2, dzsa = base + SM::dzs;
            // ---- forward recompute ----
            if (stg == 0) issue_fwd(d, x0a, wa + WO::w0, 64, K0P);
            else if (stg == 1) issue_fwd(d, h1a, wa + WO::w1, 64, 64);
            else if (stg == 2) issue_fwd(d, h2a, wa + WO::w2, 16, 64);
            else if (stg == 3) issue_fwd(d, cina, wa + WO::w3, 64, KCP);
            else if (stg == 4) issue_fwd(d, c1a, wa + WO::w4, 64, 64);
            else if (stg == 5) issue_fwd(d, c2a, wa + WO::w5, 16, 64);
            // ---- backward: dgrad into the work accumulator, weight/bias gradients into the resident ones ----
            else if (stg == 6) {    // col_model.4: dZ = dzs (16 wide), input c2; transposed gradient [k][j]
              const bool acc = (inited >> 5) & 1;
              issue_dgrad(d, dzsa, wa + WO::w5, 16, 64);
              issue_wgrad(tb + TM::g5, c2a, dzsa, 16, acc);
              issue_wgrad(tb + TM::b5, onesa, dzsa, 16, acc);
            } else if (stg == 7) {  // col_model.2: dZ in the c2 tile, input c1
              const bool acc = (inited >> 4) & 1;
              issue_dgrad(d, c2a, wa + WO::w4, 64, 64);
              issue_wgrad(tb + TM::g4, c2a, c1a, 64, acc);
              issue_wgrad(tb + TM::b4, c2a, onesa, 8, acc);
            } else if (stg == 8) {  // col_model.0: dZ in the c1 tile, input cin
              const bool acc = (inited >> 3) & 1;
              issue_dgrad(d, c1a, wa + WO::w3, 64, KCP);
              issue_wgrad(tb + TM::g3, c1a, cina, KCP, acc);
              issue_wgrad(tb + TM::b3, c1a, onesa, 8, acc);
            } else if (stg == 9) {  // sig_model.4: dZ = dzs (16 wide), input h2; transposed
              const bool acc = (inited >> 2) & 1;
              issue_dgrad(d, dzsa, wa + WO::w2, 16, 64);
              issue_wgrad(tb + TM::g2, h2a, dzsa, 16, acc);
              issue_wgrad(tb + TM::b2, onesa, dzsa, 16, acc);
            } else if (stg == 10) { // sig_model.2: dZ in the h2 tile, input h1
              const bool acc = (inited >> 1) & 1;
              issue_dgrad(d, h2a, wa + WO::w1, 64, 64);
              issue_wgrad(tb + TM::g1, h2a, h1a, 64, acc);
              issue_wgrad(tb + TM::b1, h2a, onesa, 8, acc);
            } else {                // sig_model.0: dZ in the h1 tile, input x0
              const bool acc = inited & 1;
              issue_dgrad(d, h1a, wa + WO::w0, 64, K0P);
              issue_wgrad(tb + TM::g0, h1a, x0a, K0P, acc);
              issue_wgrad(tb + TM::b0, h1a, onesa, 8, acc);
            }
            commit(bars + G + g);
          }
          __syncwarp();
          if (stg >= 6) inited |= 1u << (11 - stg);
          stage[g] = stg == 11 ? 0 : stg + 1;
          par[g] ^= 1;
          --left[g];
          --remaining;
        }
      }
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
      // ---- recompute the forward activations ----
      load_features<K0P>(feat, feat_stride, gp, n, in0, vec_ok, r, x0);
      HBR_SIGNAL(); HBR_WAIT();
      relu_epilogue64(taddr, bias + 0, r, h1);
      HBR_SIGNAL(); HBR_WAIT();
      relu_epilogue64(taddr, bias + 64, r, h2);
      HBR_SIGNAL(); HBR_WAIT();
      float lrelu_slope;
      const long long dir_row = valid ? gp / dir_group : 0;
      {
        float o16[16];
        tmem_ld<16>(taddr, o16);
#pragma unroll
        for (int k = 0; k < 16; ++k) o16[k] += bias[128 + k];
        lrelu_slope = o16[0] > 0.f ? 1.f : 0.01f;
        build_cin<KCP>(o16, dirs, dir_row, dv, valid, r, cin);
      }
      HBR_SIGNAL(); HBR_WAIT();
      relu_epilogue64(taddr, bias + 192, r, c1);
      HBR_SIGNAL(); HBR_WAIT();
      relu_epilogue64(taddr, bias + 256, r, c2);
      HBR_SIGNAL(); HBR_WAIT();
      float g_density;
      {
        float c16[16], dz16[16];
        tmem_ld<16>(taddr, c16);
        float4 go = make_float4(0.f, 0.f, 0.f, 0.f);
        if (valid) go = __ldg(reinterpret_cast<const float4*>(dout + gp * 4));
        const float gg[3] = {go.x, go.y, go.z};
#pragma unroll
        for (int c = 0; c < 16; ++c) dz16[c] = 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float pre = c16[c] + bias[320 + c];
          dz16[c] = gg[c] * (pre > 0.f ? 1.f : expf(pre));               // ELU'
        }
        g_density = go.w * lrelu_slope;                                  // LeakyReLU'
        store_chunk(dzs, r, 0, kTile, dz16);
        store_chunk(dzs, r, 1, kTile, dz16 + 8);
      }
      HBR_SIGNAL(); HBR_WAIT();                  // stage 6 done: work = dA(c2)
      masked_dz_inplace64(taddr, r, c2);
      HBR_SIGNAL(); HBR_WAIT();                  // stage 7 done: work = dA(c1)
      masked_dz_inplace64(taddr, r, c1);
      HBR_SIGNAL(); HBR_WAIT();                  // stage 8 done: work[0,KCP) = d(cin)
      {
        float dc[KCP], dz16[16];
        tmem_ld<KCP>(taddr, dc);
        dz16[0] = g_density;
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
      HBR_SIGNAL(); HBR_WAIT();                  // stage 9 done: work = dA(h2)
      masked_dz_inplace64(taddr, r, h2);
      HBR_SIGNAL(); HBR_WAIT();                  // stage 10 done: work = dA(h1)
      masked_dz_inplace64(taddr, r, h1);
      HBR_SIGNAL(); HBR_WAIT();                  // stage 11 done: work[0,K0P) = d(feat)
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

  // ---- flush the gradient accumulators (M = 64 layout: row q lives in lane (q%16) + 32*(q/16)) ----
  if (cta_tiles > 0 && dparams != nullptr && warp < 4) {
    const int q = warp * 16 + lane;              // accumulator row, meaningful for lane < 16
    const uint32_t trow = tbase + ((uint32_t)(warp * 32) << 16);
    const int gcol[6] = {TM::g0, TM::g1, TM::g2, TM::g3, TM::g4, TM::g5};
    const int bcol[6] = {TM::b0, TM::b1, TM::b2, TM::b3, TM::b4, TM::b5};
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      const bool transposed = i == 2 || i == 5;
      float gacc[64];
      const int KP = i == 0 ? K0P : (i == 3 ? KCP : (transposed ? 16 : 64));
      if (KP > 48) tmem_ld<64>(trow + gcol[i], gacc);
      else if (KP > 32) tmem_ld<48>(trow + gcol[i], gacc);
      else if (KP > 16) tmem_ld<32>(trow + gcol[i], gacc);
      else tmem_ld<16>(trow + gcol[i], gacc);
      float gbias[16];
      tmem_ld<16>(trow + bcol[i], gbias);        // 8 or 16 valid columns; any excess belongs to the next accumulator
      if (lane < 16) {
        if (!transposed) {
          if (q < m.J[i]) {
            for (int k = 0; k < m.K[i]; ++k) atomicAdd(dparams + m.W[i] + q * m.K[i] + k, gacc[k]);
            atomicAdd(dparams + m.b[i] + q, gbias[0]);
          }
        } else {
          // row q = input index k (K = 64), column = output neuron j
          for (int j = 0; j < m.J[i]; ++j) atomicAdd(dparams + m.W[i] + j * m.K[i] + q, gacc[j]);
          if (q == 0)
            for (int j = 0; j < m.J[i]; ++j) atomicAdd(dparams + m.b[i] + j, gbias[j]);
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

// The reference's configuration (in0 = 32, d_view = 24) gets the widest pipelines (4 forward / 2 backward tile groups
// per SM); other widths (in0 <= 64, 15 + d_view <= 64) run the same kernels with padded K and fewer groups.
#define HBR_TC_LAUNCH(KERN, K0P_, KCP_, G_, SMEM_T, ...)                                                            \
  do {                                                                                                              \
    constexpr int smem = SMEM_T<K0P_, KCP_, G_>::total;                                                             \
    HBR_CUDA(cudaFuncSetAttribute(KERN<K0P_, KCP_, G_>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));        \
    KERN<K0P_, KCP_, G_><<<grid, G_ * kTile + 32, smem, st>>>(__VA_ARGS__);                                         \
  } while (0)

extern "C" int hbr_mlp_fwd_tc(const float* feat, int64_t feat_stride, const float* dirs, int64_t dir_group, int64_t n,
                              const float* params, const hbr_mlp_dims* dims, float* out, void* stream) {
  if (int rc = check_dims(dims)) return rc;
  if (n == 0) return HBR_OK;
  HBR_REQUIRE(feat && dirs && params && out, "NULL pointer");
  HBR_REQUIRE(feat_stride >= dims->in0 && dir_group >= 1, "bad stride / dir_group");
  HBR_REQUIRE((uintptr_t)out % 16 == 0, "out must be 16-byte aligned");
  const int k0p = dims->in0 <= 32 ? 32 : 64, kcp = dims->d_view + kFeat <= 48 ? 48 : 64;
  cudaStream_t st = as_stream(stream);
  const int64_t ntiles = ceil_div(n, kTile);
  const int in0 = dims->in0, dv = dims->d_view;
  if (k0p == 32 && kcp == 48) {
    const int grid = (int)min64(ceil_div(ntiles, 4), sm_count());
    HBR_TC_LAUNCH(mlp_fwd_tc_kernel, 32, 48, 4, FwdSmem, feat, feat_stride, dirs, dir_group, n, params, in0, dv, out);
  } else {
    const int grid = (int)min64(ceil_div(ntiles, 4), sm_count());
    HBR_TC_LAUNCH(mlp_fwd_tc_kernel, 64, 64, 4, FwdSmem, feat, feat_stride, dirs, dir_group, n, params, in0, dv, out);
  }
  HBR_LAUNCH_CHECK();
  return HBR_OK;
}

extern "C" int hbr_mlp_bwd_tc(const float* feat, int64_t feat_stride, const float* dirs, int64_t dir_group, int64_t n,
                              const float* params, const hbr_mlp_dims* dims, const float* dout, float* dfeat,
                              int64_t dfeat_stride, float* ddirs, float* dparams, void* stream) {
  if (int rc = check_dims(dims)) return rc;
  if (n == 0) return HBR_OK;
  HBR_REQUIRE(feat && dirs && params && dout, "NULL pointer");
  HBR_REQUIRE(feat_stride >= dims->in0 && dir_group >= 1, "bad stride / dir_group");
  HBR_REQUIRE((uintptr_t)dout % 16 == 0, "dout must be 16-byte aligned");
  HBR_REQUIRE(!dfeat || dfeat_stride >= dims->in0, "dfeat_stride too small");
  const int k0p = dims->in0 <= 32 ? 32 : 64, kcp = dims->d_view + kFeat <= 48 ? 48 : 64;
  cudaStream_t st = as_stream(stream);
  const int64_t ntiles = ceil_div(n, kTile);
  const int in0 = dims->in0, dv = dims->d_view;
  if (k0p == 32 && kcp == 48) {
    const int grid = (int)min64(ceil_div(ntiles, 2), sm_count());
    HBR_TC_LAUNCH(mlp_bwd_tc_kernel, 32, 48, 2, BwdSmem, feat, feat_stride, dirs, dir_group, n, params, in0, dv, dout,
                  dfeat, dfeat_stride, ddirs, dparams);
  } else {
    const int grid = (int)min64(ntiles, sm_count());
    HBR_TC_LAUNCH(mlp_bwd_tc_kernel, 64, 64, 1, BwdSmem, feat, feat_stride, dirs, dir_group, n, params, in0, dv, dout,
                  dfeat, dfeat_stride, ddirs, dparams);
  }
  HBR_LAUNCH_CHECK();
  return HBR_OK;
}
