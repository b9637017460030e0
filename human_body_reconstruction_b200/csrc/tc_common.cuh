// tcgen05 / TMEM / mbarrier primitives for sm_100a, written as inline PTX.
// Shared-memory operand layout used everywhere in this library: the UMMA "no-swizzle" canonical layout.
//   A [rows x cols] bf16 tile is stored as 8x8 "core matrices" (8 rows x 16 bytes, contiguous 128 B):
//     byte_offset(r, c) = (c % 8) * 2 + (r % 8) * 16 + (r / 8) * 128 + (c / 8) * (rows / 8) * 128
//   i.e. core matrices are contiguous along the row dimension, column groups of 8 are (rows*16) bytes apart.
//   The same bytes can be handed to the tensor core in two ways:
//     K-major  operand (rows = M|N index, cols = K index):  SBO = 128, LBO = rows*16, one K=16 step = 2*LBO
//     MN-major operand (cols = M|N index, rows = K index):  SBO = rows*16, LBO = 128, one K=16 step = 256 B
//   so an activation tile written once serves as the K-major A operand of the forward / dgrad GEMMs and as the
//   MN-major operand of the weight-gradient GEMM (reduction over the point dimension) without a transpose.
#pragma once
#include "common.cuh"

namespace hbr {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- shared memory matrix descriptor (SM100 "version 1", no swizzle) -------------------------------------------
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;                      // descriptor version for tcgen05
  return d;                                    // base_offset = 0, lbo_mode = 0, layout_type = SWIZZLE_NONE (0)
}

// ---- operand formats of kind::f16: bf16 (torch.autocast(bfloat16)) or fp16 (torch.autocast(float16), what the
//      reference's trainer runs: train_hash2.py:218).  Everything format-specific -- conversions, the ReLU-mask
//      product of the backward, the descriptor's A/B format field -- sits behind these two types.
struct OpBf16 {
  static constexpr uint32_t kFmt = 1;          // instruction-descriptor A/B format: 1 = bf16
  static __device__ __forceinline__ uint32_t pack(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
  // max(x,0) fused into the conversion: low half <- lo, high half <- hi
  static __device__ __forceinline__ uint32_t pack_relu(float lo, float hi) {
    uint32_t d;
    asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;\n" : "=r"(d) : "f"(hi), "f"(lo));
    return d;
  }
  // dz * [h > 0] on packed pairs (h is a stored post-ReLU activation: h > 0 <=> pre-activation > 0)
  static __device__ __forceinline__ uint32_t mask_pos(uint32_t dz, uint32_t h) {
    const __nv_bfloat162 zero = __floats2bfloat162_rn(0.f, 0.f);
    const __nv_bfloat162 m = __hgt2(*reinterpret_cast<const __nv_bfloat162*>(&h), zero);      // 1.0 / 0.0 per half
    const __nv_bfloat162 r = __hmul2(*reinterpret_cast<const __nv_bfloat162*>(&dz), m);
    return *reinterpret_cast<const uint32_t*>(&r);
  }
  static __device__ __forceinline__ float round(float a) { return __bfloat162float(__float2bfloat16_rn(a)); }
  static __device__ __forceinline__ uint16_t bits(float a) {
    const __nv_bfloat16 h = __float2bfloat16_rn(a);
    return *reinterpret_cast<const uint16_t*>(&h);
  }
};
struct OpF16 {
  static constexpr uint32_t kFmt = 0;          // 0 = fp16
  static __device__ __forceinline__ uint32_t pack(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
  static __device__ __forceinline__ uint32_t pack_relu(float lo, float hi) {
    uint32_t d;
    asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;\n" : "=r"(d) : "f"(hi), "f"(lo));
    return d;
  }
  static __device__ __forceinline__ uint32_t mask_pos(uint32_t dz, uint32_t h) {
    const __half2 zero = __floats2half2_rn(0.f, 0.f);
    const __half2 m = __hgt2(*reinterpret_cast<const __half2*>(&h), zero);
    const __half2 r = __hmul2(*reinterpret_cast<const __half2*>(&dz), m);
    return *reinterpret_cast<const uint32_t*>(&r);
  }
  static __device__ __forceinline__ float round(float a) { return __half2float(__float2half_rn(a)); }
  static __device__ __forceinline__ uint16_t bits(float a) {
    const __half h = __float2half_rn(a);
    return *reinterpret_cast<const uint16_t*>(&h);
  }
};

// ---- instruction descriptor: kind::f16, (bf16 | fp16) x same -> fp32 ------------------------------------------------
template <class OP>
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, bool a_mn_major, bool b_mn_major) {
  return (1u << 4)                             // D format: f32
         | (OP::kFmt << 7)                     // A format
         | (OP::kFmt << 10)                    // B format
         | ((a_mn_major ? 1u : 0u) << 15) | ((b_mn_major ? 1u : 0u) << 16)
         | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"((uint32_t)accumulate) : "memory");
}

// same with the A operand read from tensor memory (rows = TMEM lanes, K packed two bf16 per 32-bit column)
__device__ __forceinline__ void mma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
      :: "r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"((uint32_t)accumulate) : "memory");
}

// all previously issued tcgen05.mma of this thread arrive on the mbarrier when complete
__device__ __forceinline__ void commit(uint64_t* mbar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n"
               :: "r"(smem_u32(mbar)) : "memory");
}

__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
// generic-proxy shared-memory writes -> visible to the async proxy (tensor core operand reads)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

// exactly one lane of a converged warp gets true (lets ptxas keep the MMA operands on the uniform datapath)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}\n"
      : "=r"(pred));
  return pred != 0;
}

// ---- TMEM allocation (one full warp executes these) --------------------------------------------------------------
template <int NCOLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" :: "r"(smem_u32(smem_slot)), "n"(NCOLS) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" :: "r"(taddr), "n"(NCOLS) : "memory");
}

// ---- mbarrier ----------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* mbar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"(smem_u32(mbar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
__device__ __forceinline__ bool mbar_try_wait(uint64_t* mbar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}\n"
      : "=r"(ok) : "r"(smem_u32(mbar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* mbar, uint32_t parity) {
  while (!mbar_try_wait(mbar, parity)) {
  }
}
// non-blocking probe (the MMA-issuing thread polls several barriers round-robin)
__device__ __forceinline__ bool mbar_test_wait(uint64_t* mbar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}\n"
      : "=r"(ok) : "r"(smem_u32(mbar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_arrive(uint64_t* mbar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" :: "r"(smem_u32(mbar)) : "memory");
}

// ---- TMEM -> registers: this thread's lane, N consecutive 32-bit columns ------------------------------------------
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// 8 columns, waits for the data (rolled loops of code that runs once per CTA)
__device__ __forceinline__ void tmem_ld8_wait(uint32_t taddr, float* v) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr) : "memory");
  tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

// registers -> TMEM: this thread's lane, 16 consecutive 32-bit columns
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};\n"
      :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
         "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};\n"
      :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory"); }

// N columns (multiple of 16) starting at taddr; issues all loads, then one wait
template <int N>
__device__ __forceinline__ void tmem_ld(uint32_t taddr, float* v) {
  static_assert(N % 16 == 0, "N must be a multiple of 16");
#pragma unroll
  for (int i = 0; i < N; i += 16) tmem_ld16(taddr + i, v + i);
  tmem_ld_wait();
}

// ---- canonical no-swizzle tile addressing -------------------------------------------------------------------------
// byte offset of the 16-byte chunk holding columns [8*cg, 8*cg+8) of row r in a tile with `rows` rows
__device__ __forceinline__ uint32_t chunk_off(int r, int cg, int rows) {
  return (uint32_t)((r & 7) * 16 + (r >> 3) * 128 + cg * rows * 16);
}

// store 8 consecutive columns (one chunk) of this thread's row
template <class OP>
__device__ __forceinline__ void store_chunk(uint8_t* tile, int r, int cg, int rows, const float* v8) {
  uint4 q;
  q.x = OP::pack(v8[0], v8[1]); q.y = OP::pack(v8[2], v8[3]);
  q.z = OP::pack(v8[4], v8[5]); q.w = OP::pack(v8[6], v8[7]);
  *reinterpret_cast<uint4*>(tile + chunk_off(r, cg, rows)) = q;
}

// arguments of the fused encoder + MLP variants (hbr_field_*_tc)
struct EncArgs {
  const float* x;                // (n,3) fp32 sample positions
  const float* table;            // (L,T,2) fp32                      (forward)
  float* dtable;                 // (L,T,2) fp32, accumulated into    (backward)
  uint16_t* feat16;               // (n,32) bf16 features: written by the forward, re-read by the backward recompute
  // scatter warps of the backward kernel (SCAT): sample positions formed from the rays, p = o + d * t (vol_renderer.py:165)
  const float* ro;               // (R,3)
  const float* rd;               // (R,3)
  const float* rt;               // (S) shared (t_stride = 0) or (R,S) per ray
  long long S, t_stride;
};

}  // namespace tc
}  // namespace hbr
