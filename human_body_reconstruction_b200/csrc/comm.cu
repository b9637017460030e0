// Gradient all-reduce of the data-parallel training step over NVLink peer memory.
//
// The step's only exchange is the average of the flat gradient buffer [hash table (L,T,F) | MLP parameters] over the
// ranks (67 MB at T = 2^19).  The scatter-add of the hash-grid backward writes that buffer; every rank maps every other
// rank's buffer (CUDA IPC, one process per GPU), so the collective is ONE kernel that runs right behind the scatter-add:
//
//   barrier (flags in peer memory)  ->  rank r owns slice r of the buffer: reads the slice from every rank over NVLink
//   (fixed rank order, so every rank ends with bit-identical values), scales, writes the result into every rank's
//   buffer  ->  barrier.
//
// Per GPU that is S(W-1)/W bytes of loads in and S(W-1)/W bytes of stores out (+ the peers' stores in) against the
// 2S(W-1)/W of a ring in 2(W-1) latency-bound steps; with an NVLS multicast mapping of the buffers (`multicast` != NULL)
// the reduction happens in the NVSwitch (multimem.ld_reduce / multimem.st) and each direction carries S bytes.
// The barriers are self-resetting CAS flags (put: 0 -> 1 on the peer, take: 1 -> 0 on my side), one word per
// (CTA, sender): CTA b of every rank only talks to CTA b of the others, so the grid must be co-resident (<= SM count).
// A barrier that does not complete within kBarrierTimeoutNs sets *status (sticky; also checked on entry) and the kernel
// returns WITHOUT touching the gradients: the failure is contained -- this rank keeps its own unreduced gradient -- and
// the host side raises when it sees the word (peer.PeerRegion.raise_if_failed, called by dist.PeerGradAllReduce every
// step), so no optimiser step runs on gradients the ranks disagree on.  The kernel never hangs a GPU.
#include "common.cuh"
#include <string.h>

namespace hbr {

constexpr int kMaxPeers = HBR_MAX_PEERS;
constexpr int kMaxCtas = HBR_PEER_MAX_CTAS;
constexpr unsigned long long kBarrierTimeoutNs = 10000000000ull;   // 10 s (first-step skew between ranks can be seconds)

struct PeerPtrs {
  float* buf[kMaxPeers];
  unsigned* flag[kMaxPeers];
};

__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ float4 ld_peer(const float* p) {      // system-coherent 16-byte load (peer memory: never L1)
  float4 v;
  asm volatile("ld.volatile.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_peer(float* p, const float4& v) {
  asm volatile("st.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 mc_ld_reduce(const float* p) {   // sum over every rank's copy, formed in the switch
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void mc_st(float* p, const float4& v) {  // one store, delivered to every rank's copy
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

// All threads of CTA b on every rank have arrived (and their earlier writes are visible system-wide).
__device__ __forceinline__ bool peer_barrier(const PeerPtrs& p, int rank, int world, unsigned* status) {
  __syncthreads();
  const int q = threadIdx.x;
  bool ok = true;
  if (q < world && q != rank) {
    __threadfence_system();
    unsigned* put = p.flag[q] + (size_t)blockIdx.x * kMaxPeers + rank;
    unsigned* take = p.flag[rank] + (size_t)blockIdx.x * kMaxPeers + q;
    const unsigned long long t0 = global_ns();
    while (atomicCAS_system(put, 0u, 1u) != 0u)
      if (global_ns() - t0 > kBarrierTimeoutNs) { ok = false; break; }
    while (ok && atomicCAS_system(take, 1u, 0u) != 1u)
      if (global_ns() - t0 > kBarrierTimeoutNs) { ok = false; break; }
    if (!ok && status != nullptr) atomicExch(status, 1u);
    __threadfence_system();
  }
  return __syncthreads_and(ok) != 0;
}

__device__ __forceinline__ void add4(float4& a, const float4& b) { a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }
__device__ __forceinline__ void mul4(float4& a, float s) { a.x *= s; a.y *= s; a.z *= s; a.w *= s; }

template <int WORLD, bool MC>
__global__ void __launch_bounds__(512)
allreduce_peer_kernel(const PeerPtrs p, float* __restrict__ mc, int rank, long long n4, float scale, unsigned* status) {
  if (status != nullptr && *reinterpret_cast<volatile unsigned*>(status) != 0u) return;   // an earlier exchange failed: the flags are not trustworthy
  if (!peer_barrier(p, rank, WORLD, status)) return;              // every rank's scatter-add (previous kernel) is complete
  const long long per = (n4 + WORLD - 1) / WORLD;
  const long long lo = (long long)rank * per;
  const long long hi = lo + per < n4 ? lo + per : n4;
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long i = lo + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (MC) {
    for (; i + stride < hi; i += 2 * stride) {                     // two independent switch reductions in flight
      float4 a = mc_ld_reduce(mc + 4 * i), b = mc_ld_reduce(mc + 4 * (i + stride));
      mul4(a, scale); mul4(b, scale);
      mc_st(mc + 4 * i, a); mc_st(mc + 4 * (i + stride), b);
    }
    for (; i < hi; i += stride) {
      float4 a = mc_ld_reduce(mc + 4 * i);
      mul4(a, scale);
      mc_st(mc + 4 * i, a);
    }
  } else {
    for (; i + stride < hi; i += 2 * stride) {                     // 2 x WORLD independent 16-byte loads in flight
      float4 a[WORLD], b[WORLD];
#pragma unroll
      for (int q = 0; q < WORLD; ++q) {
        a[q] = ld_peer(p.buf[q] + 4 * i);
        b[q] = ld_peer(p.buf[q] + 4 * (i + stride));
      }
#pragma unroll
      for (int q = 1; q < WORLD; ++q) { add4(a[0], a[q]); add4(b[0], b[q]); }   // fixed rank order: identical everywhere
      mul4(a[0], scale); mul4(b[0], scale);
#pragma unroll
      for (int q = 0; q < WORLD; ++q) {
        st_peer(p.buf[q] + 4 * i, a[0]);
        st_peer(p.buf[q] + 4 * (i + stride), b[0]);
      }
    }
    for (; i < hi; i += stride) {
      float4 a[WORLD];
#pragma unroll
      for (int q = 0; q < WORLD; ++q) a[q] = ld_peer(p.buf[q] + 4 * i);
#pragma unroll
      for (int q = 1; q < WORLD; ++q) add4(a[0], a[q]);
      mul4(a[0], scale);
#pragma unroll
      for (int q = 0; q < WORLD; ++q) st_peer(p.buf[q] + 4 * i, a[0]);
    }
  }
  peer_barrier(p, rank, WORLD, status);                           // every slice has landed in my buffer; nobody reads it any more
}

// ---- streamed exchange: the all-reduce runs BESIDE the scatter-add that produces the table gradient -------------------
// One launch on a side stream walks a list of pieces of the region (the MLP gradient, then the level chunks of the table
// gradient in the order hash_bwd_kernel<STREAM> completes them).  Per piece: CTA-leader waits until the local producer's
// counter done[i] has reached need[i] (every CTA of that chunk has issued its reductions and fenced), flag barrier with
// the same CTA of every other rank (their chunk is complete too), then the slice this rank owns is summed over the ranks
// and the result written into every rank's buffer -- peer loads/stores or multimem through the NVSwitch as above.
// Pieces are disjoint, and a slice is read only by its owner, so no barrier is needed between pieces; one closing barrier
// makes sure every slice has landed everywhere before the kernel (and with it the backward pass) ends.
constexpr int kMaxPieces = HBR_MAX_LEVELS + 2;
struct StreamPlan {
  int npieces;
  long long off4[kMaxPieces];      // float4 offset of the piece in the region
  long long n4[kMaxPieces];        // float4 count
  unsigned need[kMaxPieces];       // wait for done[done_idx] >= need (0: complete when the kernel starts, by stream order)
  int done_idx[kMaxPieces];
};

__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

template <int WORLD, bool MC>
__device__ __forceinline__ void reduce_slice(const PeerPtrs& p, float* __restrict__ mc, long long lo, long long hi, float scale) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long i = lo + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (MC) {
    constexpr int U = 4;                                           // independent switch reductions in flight per thread
    for (; i + (U - 1) * stride < hi; i += U * stride) {
      float4 a[U];
#pragma unroll
      for (int u = 0; u < U; ++u) a[u] = mc_ld_reduce(mc + 4 * (i + u * stride));
#pragma unroll
      for (int u = 0; u < U; ++u) { mul4(a[u], scale); mc_st(mc + 4 * (i + u * stride), a[u]); }
    }
    for (; i < hi; i += stride) {
      float4 a = mc_ld_reduce(mc + 4 * i);
      mul4(a, scale);
      mc_st(mc + 4 * i, a);
    }
  } else {
    for (; i + stride < hi; i += 2 * stride) {
      float4 a[WORLD], b[WORLD];
#pragma unroll
      for (int q = 0; q < WORLD; ++q) {
        a[q] = ld_peer(p.buf[q] + 4 * i);
        b[q] = ld_peer(p.buf[q] + 4 * (i + stride));
      }
#pragma unroll
      for (int q = 1; q < WORLD; ++q) { add4(a[0], a[q]); add4(b[0], b[q]); }   // fixed rank order: identical everywhere
      mul4(a[0], scale); mul4(b[0], scale);
#pragma unroll
      for (int q = 0; q < WORLD; ++q) {
        st_peer(p.buf[q] + 4 * i, a[0]);
        st_peer(p.buf[q] + 4 * (i + stride), b[0]);
      }
    }
    for (; i < hi; i += stride) {
      float4 a[WORLD];
#pragma unroll
      for (int q = 0; q < WORLD; ++q) a[q] = ld_peer(p.buf[q] + 4 * i);
#pragma unroll
      for (int q = 1; q < WORLD; ++q) add4(a[0], a[q]);
      mul4(a[0], scale);
#pragma unroll
      for (int q = 0; q < WORLD; ++q) st_peer(p.buf[q] + 4 * i, a[0]);
    }
  }
}

template <int WORLD, bool MC>
__global__ void __launch_bounds__(512)
allreduce_stream_kernel(const PeerPtrs p, float* __restrict__ mc, int rank, const __grid_constant__ StreamPlan plan,
                        const unsigned* __restrict__ done, float scale, unsigned* status) {
  if (status != nullptr && *reinterpret_cast<volatile unsigned*>(status) != 0u) return;
  __shared__ int s_ok;
  for (int c = 0; c < plan.npieces; ++c) {
    if (plan.need[c] != 0u) {
      if (threadIdx.x == 0) {
        const unsigned long long t0 = global_ns();
        bool ok = true;
        while (ld_acquire_gpu(done + plan.done_idx[c]) < plan.need[c]) {
          if (global_ns() - t0 > kBarrierTimeoutNs) { ok = false; break; }
          __nanosleep(100);
        }
        if (!ok && status != nullptr) atomicExch(status, 1u);
        s_ok = ok ? 1 : 0;
      }
      __syncthreads();
      const bool ok = s_ok != 0;
      __syncthreads();                                             // s_ok is rewritten for the next piece
      if (!ok) return;                                             // the producer never finished: contained like a barrier timeout
    }
    if (!peer_barrier(p, rank, WORLD, status)) return;             // the piece is complete on every rank
    const long long per = (plan.n4[c] + WORLD - 1) / WORLD;
    const long long lo = plan.off4[c] + (long long)rank * per;
    const long long end = plan.off4[c] + plan.n4[c];
    const long long hi = lo + per < end ? lo + per : end;
    reduce_slice<WORLD, MC>(p, mc, lo, hi, scale);
  }
  peer_barrier(p, rank, WORLD, status);                            // every slice has landed in my buffer
}

template <int WORLD>
static int launch_allreduce_stream(const PeerPtrs& p, float* mc, int rank, const StreamPlan& plan, const unsigned* done,
                                   float scale, int ctas, unsigned* status, cudaStream_t s) {
  if (mc != nullptr)
    allreduce_stream_kernel<WORLD, true><<<ctas, 512, 0, s>>>(p, mc, rank, plan, done, scale, status);
  else
    allreduce_stream_kernel<WORLD, false><<<ctas, 512, 0, s>>>(p, mc, rank, plan, done, scale, status);
  HBR_LAUNCH_CHECK();
  return HBR_OK;
}

template <int WORLD>
static int launch_allreduce(const PeerPtrs& p, float* mc, int rank, long long n4, float scale, int ctas, unsigned* status,
                            cudaStream_t s) {
  if (mc != nullptr)
    allreduce_peer_kernel<WORLD, true><<<ctas, 512, 0, s>>>(p, mc, rank, n4, scale, status);
  else
    allreduce_peer_kernel<WORLD, false><<<ctas, 512, 0, s>>>(p, mc, rank, n4, scale, status);
  HBR_LAUNCH_CHECK();
  return HBR_OK;
}

}  // namespace hbr

using namespace hbr;

extern "C" int hbr_peer_alloc(void** ptr, int64_t bytes) {
  HBR_REQUIRE(ptr != nullptr && bytes > 0, "ptr=%p bytes=%lld", (void*)ptr, (long long)bytes);
  HBR_CUDA(cudaMalloc(ptr, (size_t)bytes));
  HBR_CUDA(cudaMemset(*ptr, 0, (size_t)bytes));
  HBR_CUDA(cudaDeviceSynchronize());
  return HBR_OK;
}

extern "C" int hbr_peer_free(void* ptr) {
  if (ptr != nullptr) HBR_CUDA(cudaFree(ptr));
  return HBR_OK;
}

extern "C" int hbr_peer_export(void* ptr, unsigned char handle[HBR_PEER_HANDLE_BYTES]) {
  static_assert(sizeof(cudaIpcMemHandle_t) == HBR_PEER_HANDLE_BYTES, "handle size");
  HBR_REQUIRE(ptr != nullptr && handle != nullptr, "NULL pointer");
  cudaIpcMemHandle_t h;
  HBR_CUDA(cudaIpcGetMemHandle(&h, ptr));
  memcpy(handle, &h, sizeof(h));
  return HBR_OK;
}

extern "C" int hbr_peer_import(const unsigned char handle[HBR_PEER_HANDLE_BYTES], void** ptr) {
  HBR_REQUIRE(ptr != nullptr && handle != nullptr, "NULL pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  HBR_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return HBR_OK;
}

extern "C" int hbr_peer_release(void* ptr) {
  if (ptr != nullptr) HBR_CUDA(cudaIpcCloseMemHandle(ptr));
  return HBR_OK;
}

extern "C" int hbr_allreduce_peer(void* const* bufs, void* const* flags, void* multicast, int rank, int world, int64_t n,
                                  float scale, int ctas, unsigned int* status, void* stream) {
  HBR_REQUIRE(world >= 1 && world <= kMaxPeers && rank >= 0 && rank < world, "rank=%d world=%d", rank, world);
  HBR_REQUIRE(n >= 0 && n % 4 == 0, "n=%lld must be a multiple of 4 floats", (long long)n);
  HBR_REQUIRE(bufs != nullptr && flags != nullptr, "NULL pointer table");
  if (n == 0) return HBR_OK;
  PeerPtrs p{};
  for (int q = 0; q < world; ++q) {
    HBR_REQUIRE(bufs[q] != nullptr && flags[q] != nullptr, "rank %d: NULL buffer/flags", q);
    HBR_REQUIRE((uintptr_t)bufs[q] % 16 == 0 && (uintptr_t)flags[q] % 4 == 0, "rank %d: misaligned", q);
    p.buf[q] = static_cast<float*>(bufs[q]);
    p.flag[q] = static_cast<unsigned*>(flags[q]);
  }
  if (ctas <= 0) ctas = 64;
  if (ctas > kMaxCtas) ctas = kMaxCtas;
  if (ctas > sm_count()) ctas = sm_count();                      // the flag barriers need a co-resident grid
  cudaStream_t s = as_stream(stream);
  float* mc = static_cast<float*>(multicast);
  const long long n4 = n / 4;
  switch (world) {
    case 1: return launch_allreduce<1>(p, mc, rank, n4, scale, ctas, status, s);
    case 2: return launch_allreduce<2>(p, mc, rank, n4, scale, ctas, status, s);
    case 3: return launch_allreduce<3>(p, mc, rank, n4, scale, ctas, status, s);
    case 4: return launch_allreduce<4>(p, mc, rank, n4, scale, ctas, status, s);
    case 5: return launch_allreduce<5>(p, mc, rank, n4, scale, ctas, status, s);
    case 6: return launch_allreduce<6>(p, mc, rank, n4, scale, ctas, status, s);
    case 7: return launch_allreduce<7>(p, mc, rank, n4, scale, ctas, status, s);
    case 8: return launch_allreduce<8>(p, mc, rank, n4, scale, ctas, status, s);
    default: return fail(HBR_ERR_ARG, "world=%d: the peer all-reduce is built for 1..8 ranks of one NVSwitch domain", world);
  }
}

extern "C" int hbr_allreduce_peer_stream(void* const* bufs, void* const* flags, void* multicast, int rank, int world,
                                         int npieces, const int64_t* piece_off, const int64_t* piece_n,
                                         const unsigned int* piece_need, const int* piece_done_idx, const unsigned int* done,
                                         float scale, int ctas, unsigned int* status, void* stream) {
  HBR_REQUIRE(world >= 1 && world <= kMaxPeers && rank >= 0 && rank < world, "rank=%d world=%d", rank, world);
  HBR_REQUIRE(bufs != nullptr && flags != nullptr, "NULL pointer table");
  HBR_REQUIRE(npieces >= 0 && npieces <= kMaxPieces, "npieces=%d out of range [0,%d]", npieces, kMaxPieces);
  if (npieces == 0) return HBR_OK;
  HBR_REQUIRE(piece_off && piece_n && piece_need && piece_done_idx, "NULL piece table");
  StreamPlan plan{};
  plan.npieces = npieces;
  for (int c = 0; c < npieces; ++c) {
    HBR_REQUIRE(piece_off[c] >= 0 && piece_off[c] % 4 == 0 && piece_n[c] > 0 && piece_n[c] % 4 == 0,
                "piece %d: offset %lld / n %lld must be multiples of 4 floats", c, (long long)piece_off[c], (long long)piece_n[c]);
    HBR_REQUIRE(piece_need[c] == 0 || (done != nullptr && piece_done_idx[c] >= 0), "piece %d waits on a counter but done is NULL", c);
    plan.off4[c] = piece_off[c] / 4;
    plan.n4[c] = piece_n[c] / 4;
    plan.need[c] = piece_need[c];
    plan.done_idx[c] = piece_done_idx[c];
  }
  PeerPtrs p{};
  for (int q = 0; q < world; ++q) {
    HBR_REQUIRE(bufs[q] != nullptr && flags[q] != nullptr, "rank %d: NULL buffer/flags", q);
    HBR_REQUIRE((uintptr_t)bufs[q] % 16 == 0 && (uintptr_t)flags[q] % 4 == 0, "rank %d: misaligned", q);
    p.buf[q] = static_cast<float*>(bufs[q]);
    p.flag[q] = static_cast<unsigned*>(flags[q]);
  }
  if (ctas <= 0) ctas = 32;
  if (ctas > kMaxCtas) ctas = kMaxCtas;
  if (ctas > sm_count()) ctas = sm_count();
  cudaStream_t s = as_stream(stream);
  float* mc = static_cast<float*>(multicast);
  switch (world) {
    case 1: return launch_allreduce_stream<1>(p, mc, rank, plan, done, scale, ctas, status, s);
    case 2: return launch_allreduce_stream<2>(p, mc, rank, plan, done, scale, ctas, status, s);
    case 3: return launch_allreduce_stream<3>(p, mc, rank, plan, done, scale, ctas, status, s);
    case 4: return launch_allreduce_stream<4>(p, mc, rank, plan, done, scale, ctas, status, s);
    case 5: return launch_allreduce_stream<5>(p, mc, rank, plan, done, scale, ctas, status, s);
    case 6: return launch_allreduce_stream<6>(p, mc, rank, plan, done, scale, ctas, status, s);
    case 7: return launch_allreduce_stream<7>(p, mc, rank, plan, done, scale, ctas, status, s);
    case 8: return launch_allreduce_stream<8>(p, mc, rank, plan, done, scale, ctas, status, s);
    default: return fail(HBR_ERR_ARG, "world=%d: the peer all-reduce is built for 1..8 ranks of one NVSwitch domain", world);
  }
}
