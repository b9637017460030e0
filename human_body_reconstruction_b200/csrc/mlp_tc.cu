// MLP_3D on the 5th-generation tensor cores (tcgen05 + TMEM): bf16 operands, fp32 accumulation.
// This is the field evaluation of the training step (train_hash2.py:218-226 runs it under autocast); the fp32
// CUDA-core version in mlp_simt.cu serves nerf2mesh and the 1e-5 parity tests.
//
// One CTA = 128 threads = one 128-point tile at a time (persistent over tiles).  Per layer:
//   thread 0 issues tcgen05.mma (A = activation tile in smem, B = weight tile in smem, D = 128 x N fp32 in TMEM)
//   and commits to an mbarrier; all 4 warps wait, pull their 32 TMEM lanes (one point per thread) with
//   tcgen05.ld, apply bias + activation in registers, repack to bf16 and write the next layer's A tile.
// Backward recomputes the forward activations in shared memory (nothing but the features is re-read from HBM),
// then walks the layers in reverse: per layer one dgrad GEMM (dA = dZ W, B = the SAME weight tile read MN-major)
// and one weight-gradient GEMM (dW = dZ^T A, both operands read MN-major from the tiles already in smem, M = 64)
// whose fp32 accumulators stay resident in TMEM across all tiles of the persistent CTA; bias gradients ride
// along as a GEMM against a ones tile.  They are flushed once per CTA with atomics.
// Layout conventions: tc_common.cuh.
#include "mlp_layout.cuh"
#include "tc_common.cuh"

namespace hbr {
using namespace tc;

constexpr int kTile = 128;            // points per tile == threads per CTA
constexpr int kLBO128 = kTile * 16;   // column-group stride of a 128-row tile (bytes)

struct TcShape {                      // padded GEMM shapes of the six layers
  int JP[6], KP[6], wofs[6], wbytes;
};
__host__ __device__ inline TcShape make_tc_shape(int k0p, int kcp) {
  TcShape s;
  const int JP[6] = {64, 64, 16, 64, 64, 16};
  const int KP[6] = {k0p, 64, 64, kcp, 64, 64};
  int o = 0;
  for (int i = 0; i < 6; ++i) {
    s.JP[i] = JP[i]; s.KP[i] = KP[i];
    s.wofs[i] = o; o += JP[i] * KP[i] * 2;
  }
  s.wbytes = o;
  return s;
}

// fp32 (J,K) row-major weights -> bf16 canonical tile [JP rows x KP cols], zero padded
__device__ __forceinline__ void stage_weights_bf16(const float* __restrict__ params, const MlpLayout& m, const TcShape& s,
                                                   uint8_t* wsm, float* bias_sm, int nlayers) {
  for (int i = 0; i < nlayers; ++i) {
    const int JP = s.JP[i], KP = s.KP[i], J = m.J[i], K = m.K[i];
    uint8_t* w = wsm + s.wofs[i];
    for (int e = threadIdx.x; e < JP * KP; e += blockDim.x) {
      const int j = e / KP, k = e - j * KP;
      const float v = (j < J && k < K) ? __ldg(params + m.W[i] + j * K + k) : 0.f;
      *reinterpret_cast<__nv_bfloat16*>(w + chunk_off(j, k >> 3, JP) + (k & 7) * 2) = __float2bfloat16_rn(v);
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
// G[64 x N] += dZ^T[64 x 128] * Act[128 x N]: both tiles [128 points x cols] read MN-major, reduction over points
__device__ __forceinline__ void issue_wgrad(uint32_t tmem_g, uint32_t dz_tile, uint32_t act_tile, int N, bool accumulate) {
  const uint32_t idesc = make_idesc(64, N, true, true);
  for (int kk = 0; kk < kTile / 16; ++kk) {
    const uint64_t a = make_desc(dz_tile + kk * 256, 128, kLBO128);
    const uint64_t b = make_desc(act_tile + kk * 256, 128, kLBO128);
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
// shared pieces of the forward pass (also the recompute phase of the backward kernel)
// ---------------------------------------------------------------------------------------------------------------
template <int K0P>
__device__ __forceinline__ void load_features(const float* __restrict__ feat, long long stride, long long gp, long long n,
                                              int in0, uint8_t* x0) {
  const int r = threadIdx.x;
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

// bias + ReLU on N accumulator columns, write the bf16 activation tile, return the sign mask
template <int N>
__device__ __forceinline__ unsigned long long relu_epilogue(uint32_t taddr, const float* bias, uint8_t* tile) {
  float v[N];
  tmem_ld<N>(taddr, v);
  unsigned long long mask = 0;
#pragma unroll
  for (int k = 0; k < N; ++k) {
    const float a = v[k] + bias[k];
    if (a > 0.f) mask |= 1ull << k;
    v[k] = fmaxf(a, 0.f);
  }
#pragma unroll
  for (int cg = 0; cg < N / 8; ++cg) store_chunk(tile, threadIdx.x, cg, kTile, v + cg * 8);
  return mask;
}

template <int KCP>
__device__ __forceinline__ void build_cin(const float* o16, const float* __restrict__ dirs, long long dir_row, int dv,
                                          bool valid, uint8_t* cin) {
  float v[KCP];
#pragma unroll
  for (int k = 0; k < KCP; ++k) {
    float x = 0.f;
    if (k < kFeat) x = o16[1 + k];                                    // feat_vec = dens_vec[:,1:]  (test_hash.py:64)
    else if (k < kFeat + dv && valid) x = __ldg(dirs + dir_row * dv + (k - kFeat));   // concat(viewdirs) (:66)
    v[k] = x;
  }
#pragma unroll
  for (int cg = 0; cg < KCP / 8; ++cg) store_chunk(cin, threadIdx.x, cg, kTile, v + cg * 8);
}

#define HBR_TC_SYNC_ISSUE(BODY)            \
  do {                                     \
    fence_async_smem();                    \
    fence_before_sync();                   \
    __syncthreads();                       \
    if (threadIdx.x == 0) {                \
      fence_after_sync();                  \
      BODY;                                \
      commit(mbar);                        \
    }                                      \
    mbar_wait(mbar, phase);                \
    phase ^= 1;                            \
    fence_after_sync();                    \
  } while (0)

// ---------------------------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------------------------
template <int K0P, int KCP>
__global__ void __launch_bounds__(kTile)
mlp_fwd_tc_kernel(const float* __restrict__ feat, long long feat_stride, const float* __restrict__ dirs, long long dir_group,
                  long long n, const float* __restrict__ params, int in0, int dv, float* __restrict__ out) {
  extern __shared__ __align__(128) uint8_t sm[];
  const MlpLayout m = make_layout(in0, dv);
  const TcShape s = make_tc_shape(K0P, KCP);
  uint8_t* wsm = sm;
  float* bias = reinterpret_cast<float*>(sm + s.wbytes);
  uint8_t* x0 = sm + s.wbytes + 6 * 64 * 4;
  uint8_t* h = x0 + kTile * K0P * 2;
  uint8_t* cin = h + kTile * 64 * 2;
  uint64_t* mbar = reinterpret_cast<uint64_t*>(cin + kTile * KCP * 2);
  uint32_t* tslot = reinterpret_cast<uint32_t*>(mbar + 1);
  const int warp = threadIdx.x >> 5;

  if (warp == 0) tmem_alloc<64>(tslot);
  if (threadIdx.x == 0) { mbar_init(mbar, 1); fence_mbar_init(); }
  stage_weights_bf16(params, m, s, wsm, bias, 6);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tbase = *tslot;
  const uint32_t taddr = tbase + ((uint32_t)(warp * 32) << 16);
  const uint32_t wa = smem_u32(wsm), x0a = smem_u32(x0), ha = smem_u32(h), cina = smem_u32(cin);
  uint32_t phase = 0;
  const long long ntiles = (n + kTile - 1) / kTile;

  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long gp = tile * kTile + threadIdx.x;
    const bool valid = gp < n;
    load_features<K0P>(feat, feat_stride, gp, n, in0, x0);
    HBR_TC_SYNC_ISSUE(issue_fwd(tbase, x0a, wa + s.wofs[0], 64, K0P));
    relu_epilogue<64>(taddr, bias + 0, h);
    HBR_TC_SYNC_ISSUE(issue_fwd(tbase, ha, wa + s.wofs[1], 64, 64));
    relu_epilogue<64>(taddr, bias + 64, h);
    HBR_TC_SYNC_ISSUE(issue_fwd(tbase, ha, wa + s.wofs[2], 16, 64));
    float o16[16];
    tmem_ld<16>(taddr, o16);
#pragma unroll
    for (int k = 0; k < 16; ++k) o16[k] += bias[128 + k];
    const float density = o16[0] > 0.f ? o16[0] : 0.01f * o16[0];     // LeakyReLU (test_hash.py:62)
    build_cin<KCP>(o16, dirs, valid ? gp / dir_group : 0, dv, valid, cin);
    HBR_TC_SYNC_ISSUE(issue_fwd(tbase, cina, wa + s.wofs[3], 64, KCP));
    relu_epilogue<64>(taddr, bias + 192, h);
    HBR_TC_SYNC_ISSUE(issue_fwd(tbase, ha, wa + s.wofs[4], 64, 64));
    relu_epilogue<64>(taddr, bias + 256, h);
    HBR_TC_SYNC_ISSUE(issue_fwd(tbase, ha, wa + s.wofs[5], 16, 64));
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
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<64>(tbase);
}

// ---------------------------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------------------------
template <int N>
__device__ __forceinline__ void masked_dz_epilogue(uint32_t taddr, unsigned long long mask, uint8_t* dz) {
  float v[N];
  tmem_ld<N>(taddr, v);
#pragma unroll
  for (int k = 0; k < N; ++k) v[k] = ((mask >> k) & 1ull) ? v[k] : 0.f;
#pragma unroll
  for (int cg = 0; cg < N / 8; ++cg) store_chunk(dz, threadIdx.x, cg, kTile, v + cg * 8);
}

// write a 16-wide dZ (layers with 16 / 3 outputs) into the 64-column dZ tile, zeroing the other columns
__device__ __forceinline__ void store_dz16(const float* v16, uint8_t* dz) {
  store_chunk(dz, threadIdx.x, 0, kTile, v16);
  store_chunk(dz, threadIdx.x, 1, kTile, v16 + 8);
  const uint4 z = make_uint4(0, 0, 0, 0);
#pragma unroll
  for (int cg = 2; cg < 8; ++cg) *reinterpret_cast<uint4*>(dz + chunk_off(threadIdx.x, cg, kTile)) = z;
}

template <int K0P, int KCP>
__global__ void __launch_bounds__(kTile, 1)
mlp_bwd_tc_kernel(const float* __restrict__ feat, long long feat_stride, const float* __restrict__ dirs, long long dir_group,
                  long long n, const float* __restrict__ params, int in0, int dv, const float* __restrict__ dout,
                  float* __restrict__ dfeat, long long dfeat_stride, float* __restrict__ ddirs, float* __restrict__ dparams) {
  extern __shared__ __align__(128) uint8_t sm[];
  const MlpLayout m = make_layout(in0, dv);
  const TcShape s = make_tc_shape(K0P, KCP);
  uint8_t* wsm = sm;
  float* bias = reinterpret_cast<float*>(sm + s.wbytes);
  uint8_t* x0 = sm + s.wbytes + 6 * 64 * 4;
  uint8_t* h1 = x0 + kTile * K0P * 2;
  uint8_t* h2 = h1 + kTile * 64 * 2;
  uint8_t* cin = h2 + kTile * 64 * 2;
  uint8_t* c1 = cin + kTile * KCP * 2;
  uint8_t* c2 = c1 + kTile * 64 * 2;
  uint8_t* dz = c2 + kTile * 64 * 2;
  uint8_t* ones = dz + kTile * 64 * 2;                       // [128 x 8] bf16 of 1.0
  uint64_t* mbar = reinterpret_cast<uint64_t*>(ones + kTile * 8 * 2);
  uint32_t* tslot = reinterpret_cast<uint32_t*>(mbar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0) tmem_alloc<512>(tslot);
  if (threadIdx.x == 0) { mbar_init(mbar, 1); fence_mbar_init(); }
  stage_weights_bf16(params, m, s, wsm, bias, 6);
  {
    const uint32_t one2 = pack_bf16(1.f, 1.f);
    *reinterpret_cast<uint4*>(ones + chunk_off(threadIdx.x, 0, kTile)) = make_uint4(one2, one2, one2, one2);
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tbase = *tslot;
  const uint32_t taddr = tbase + ((uint32_t)(warp * 32) << 16);
  // TMEM columns: [0,64) work accumulator; then one weight-gradient accumulator per layer (64 rows x KP cols);
  // then one 8-column bias-gradient accumulator per layer
  int gcol[6], bcol[6];
  {
    int c = 64;
    for (int i = 0; i < 6; ++i) { gcol[i] = c; c += s.KP[i]; }
    for (int i = 0; i < 6; ++i) { bcol[i] = c; c += 8; }
  }
  const uint32_t wa = smem_u32(wsm), x0a = smem_u32(x0), h1a = smem_u32(h1), h2a = smem_u32(h2), cina = smem_u32(cin),
                 c1a = smem_u32(c1), c2a = smem_u32(c2), dza = smem_u32(dz), onesa = smem_u32(ones);
  uint32_t phase = 0;
  const long long ntiles = (n + kTile - 1) / kTile;
  bool acc = false;

  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long gp = tile * kTile + threadIdx.x;
    const bool valid = gp < n;
    // ---- recompute the forward activations ----
    load_features<K0P>(feat, feat_stride, gp, n, in0, x0);
    HBR_TC_SYNC_ISSUE(issue_fwd(tbase, x0a, wa + s.wofs[0], 64, K0P));
    const unsigned long long m_h1 = relu_epilogue<64>(taddr, bias + 0, h1);
    HBR_TC_SYNC_ISSUE(issue_fwd(tbase, h1a, wa + s.wofs[1], 64, 64));
    const unsigned long long m_h2 = relu_epilogue<64>(taddr, bias + 64, h2);
    HBR_TC_SYNC_ISSUE(issue_fwd(tbase, h2a, wa + s.wofs[2], 16, 64));
    float o16[16];
    tmem_ld<16>(taddr, o16);
#pragma unroll
    for (int k = 0; k < 16; ++k) o16[k] += bias[128 + k];
    const float lrelu_slope = o16[0] > 0.f ? 1.f : 0.01f;
    const long long dir_row = valid ? gp / dir_group : 0;
    build_cin<KCP>(o16, dirs, dir_row, dv, valid, cin);
    HBR_TC_SYNC_ISSUE(issue_fwd(tbase, cina, wa + s.wofs[3], 64, KCP));
    const unsigned long long m_c1 = relu_epilogue<64>(taddr, bias + 192, c1);
    HBR_TC_SYNC_ISSUE(issue_fwd(tbase, c1a, wa + s.wofs[4], 64, 64));
    const unsigned long long m_c2 = relu_epilogue<64>(taddr, bias + 256, c2);
    HBR_TC_SYNC_ISSUE(issue_fwd(tbase, c2a, wa + s.wofs[5], 16, 64));
    float dz16[16];
    float g_density = 0.f;
    {
      float c16[16];
      tmem_ld<16>(taddr, c16);
      float4 go = make_float4(0.f, 0.f, 0.f, 0.f);
      if (valid) go = *reinterpret_cast<const float4*>(dout + gp * 4);
      const float g[3] = {go.x, go.y, go.z};
#pragma unroll
      for (int c = 0; c < 16; ++c) dz16[c] = 0.f;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float pre = c16[c] + bias[320 + c];
        dz16[c] = g[c] * (pre > 0.f ? 1.f : expf(pre));                // ELU'
      }
      g_density = go.w * lrelu_slope;                                   // LeakyReLU'
    }
    // ---- layer 5 (col_model.4): dz = d(rgb_pre) ----
    store_dz16(dz16, dz);
    HBR_TC_SYNC_ISSUE({
      issue_dgrad(tbase, dza, wa + s.wofs[5], 16, 64);
      issue_wgrad(tbase + gcol[5], dza, c2a, 64, acc);
      issue_wgrad(tbase + bcol[5], dza, onesa, 8, acc);
    });
    masked_dz_epilogue<64>(taddr, m_c2, dz);
    // ---- layer 4 (col_model.2) ----
    HBR_TC_SYNC_ISSUE({
      issue_dgrad(tbase, dza, wa + s.wofs[4], 64, 64);
      issue_wgrad(tbase + gcol[4], dza, c1a, 64, acc);
      issue_wgrad(tbase + bcol[4], dza, onesa, 8, acc);
    });
    masked_dz_epilogue<64>(taddr, m_c1, dz);
    // ---- layer 3 (col_model.0): input = cat(feat15, dirs) ----
    HBR_TC_SYNC_ISSUE({
      issue_dgrad(tbase, dza, wa + s.wofs[3], 64, KCP);
      issue_wgrad(tbase + gcol[3], dza, cina, KCP, acc);
      issue_wgrad(tbase + bcol[3], dza, onesa, 8, acc);
    });
    {
      float dc[KCP];
      tmem_ld<KCP>(taddr, dc);
      dz16[0] = g_density;
#pragma unroll
      for (int k = 0; k < kFeat; ++k) dz16[1 + k] = dc[k];
      if (ddirs != nullptr && valid) {
#pragma unroll
        for (int k = kFeat; k < KCP; ++k)
          if (k < kFeat + dv) atomicAdd(ddirs + dir_row * dv + (k - kFeat), dc[k]);
      }
    }
    // ---- layer 2 (sig_model.4): 16 outputs, no activation on the vector itself ----
    store_dz16(dz16, dz);
    HBR_TC_SYNC_ISSUE({
      issue_dgrad(tbase, dza, wa + s.wofs[2], 16, 64);
      issue_wgrad(tbase + gcol[2], dza, h2a, 64, acc);
      issue_wgrad(tbase + bcol[2], dza, onesa, 8, acc);
    });
    masked_dz_epilogue<64>(taddr, m_h2, dz);
    // ---- layer 1 (sig_model.2) ----
    HBR_TC_SYNC_ISSUE({
      issue_dgrad(tbase, dza, wa + s.wofs[1], 64, 64);
      issue_wgrad(tbase + gcol[1], dza, h1a, 64, acc);
      issue_wgrad(tbase + bcol[1], dza, onesa, 8, acc);
    });
    masked_dz_epilogue<64>(taddr, m_h1, dz);
    // ---- layer 0 (sig_model.0) ----
    HBR_TC_SYNC_ISSUE({
      issue_dgrad(tbase, dza, wa + s.wofs[0], 64, K0P);
      issue_wgrad(tbase + gcol[0], dza, x0a, K0P, acc);
      issue_wgrad(tbase + bcol[0], dza, onesa, 8, acc);
    });
    if (dfeat != nullptr) {
      float df[K0P];
      tmem_ld<K0P>(taddr, df);
      if (valid) {
#pragma unroll
        for (int k = 0; k < K0P; ++k)
          if (k < in0) dfeat[gp * dfeat_stride + k] = df[k];
      }
    }
    acc = true;
  }

  // ---- flush the weight / bias gradient accumulators (M = 64 layout: row j lives in lane (j%16) + 32*(j/16)) ----
  if (acc && dparams != nullptr) {
    const int j = warp * 16 + lane;            // meaningful for lane < 16
    for (int i = 0; i < 6; ++i) {
      float g[64];
      const int KP = s.KP[i];
      if (KP > 48) tmem_ld<64>(taddr + gcol[i], g);
      else if (KP > 32) tmem_ld<48>(taddr + gcol[i], g);
      else tmem_ld<32>(taddr + gcol[i], g);
      float gb[16];
      tmem_ld<16>(taddr + bcol[i], gb);        // 8 valid columns (all equal); the rest belongs to the next accumulator
      if (lane < 16 && j < m.J[i]) {
        for (int k = 0; k < m.K[i]; ++k) atomicAdd(dparams + m.W[i] + j * m.K[i] + k, g[k]);
        atomicAdd(dparams + m.b[i] + j, gb[0]);
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tbase);
}

static size_t fwd_smem_bytes(int k0p, int kcp) {
  const TcShape s = make_tc_shape(k0p, kcp);
  return (size_t)s.wbytes + 6 * 64 * 4 + kTile * k0p * 2 + kTile * 64 * 2 + kTile * kcp * 2 + 16 + 128;
}
static size_t bwd_smem_bytes(int k0p, int kcp) {
  const TcShape s = make_tc_shape(k0p, kcp);
  return (size_t)s.wbytes + 6 * 64 * 4 + kTile * k0p * 2 + 5 * kTile * 64 * 2 + kTile * kcp * 2 + kTile * 8 * 2 + 16 + 128;
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

#define HBR_TC_DISPATCH(KERN, SMEMFN, GRID, ...)                                              \
  do {                                                                                         \
    const size_t smem = SMEMFN(k0p, kcp);                                                      \
    if (k0p == 32 && kcp == 48) {                                                              \
      HBR_CUDA(cudaFuncSetAttribute(KERN<32, 48>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      KERN<32, 48><<<GRID, kTile, smem, st>>>(__VA_ARGS__);                                    \
    } else if (k0p == 32) {                                                                    \
      HBR_CUDA(cudaFuncSetAttribute(KERN<32, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      KERN<32, 64><<<GRID, kTile, smem, st>>>(__VA_ARGS__);                                    \
    } else if (kcp == 48) {                                                                    \
      HBR_CUDA(cudaFuncSetAttribute(KERN<64, 48>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      KERN<64, 48><<<GRID, kTile, smem, st>>>(__VA_ARGS__);                                    \
    } else {                                                                                   \
      HBR_CUDA(cudaFuncSetAttribute(KERN<64, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      KERN<64, 64><<<GRID, kTile, smem, st>>>(__VA_ARGS__);                                    \
    }                                                                                          \
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
  const int grid = (int)min64(ceil_div(n, kTile), (int64_t)sm_count() * 3);
  HBR_TC_DISPATCH(mlp_fwd_tc_kernel, fwd_smem_bytes, grid, feat, feat_stride, dirs, dir_group, n, params, dims->in0,
                  dims->d_view, out);
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
  const int grid = (int)min64(ceil_div(n, kTile), (int64_t)sm_count());
  HBR_TC_DISPATCH(mlp_bwd_tc_kernel, bwd_smem_bytes, grid, feat, feat_stride, dirs, dir_group, n, params, dims->in0,
                  dims->d_view, dout, dfeat, dfeat_stride, ddirs, dparams);
  HBR_LAUNCH_CHECK();
  return HBR_OK;
}
