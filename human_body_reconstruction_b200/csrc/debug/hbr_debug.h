/* Debug / probe entry points (libhbr_b200_debug.so).  Not part of the product ABI: nothing in the package's product
 * path calls these; tests/test_gpu_tc.py uses the operand-mode self test, scripts/dbg_*.py the probes. */
#ifndef HBR_B200_DEBUG_H_
#define HBR_B200_DEBUG_H_
#include "../../../include/hbr.h"
#ifdef __cplusplus
extern "C" {
#endif
/* Self-test of the three UMMA operand modes the MLP kernels rely on (one 128-thread CTA, 16-bit inputs rounded from
 * fp32 in the given operand format, fp32 result): mode 0: D[128,N] = A[128,K] B[N,K]^T; mode 1: D[128,N] = A[128,K] Bt[K,N];
 * mode 2: D[64,N] = At[128,64]^T Bt[128,N]. */
int hbr_debug_umma(int mode, int operand, const float* A, const float* B, float* D, int N, int K, void* stream);
/* Tensor-pipe probe: `reps` tcgen05.mma (M x N x 16, bf16) issued back to back by one thread round-robin over `nacc`
 * accumulators; cycles[0] = first issue -> completion observed, cycles[1] = first issue -> last issue (SM clocks). */
int hbr_debug_umma_bench(int M, int N, int reps, int nacc, int mn_major, long long* cycles, void* stream);
/* Steady-state cost of the GEMM chains the MLP kernels issue: kind 0 forward layer (4 MMAs), 1 dgrad (4), 2 weight
 * gradient M=64 N=72 (8), 3 transposed weight gradient M=128 N=16 (8); `reps` chains round-robin over `nacc`
 * accumulators.  cycles[0] = total, cycles[1] = issue only (SM clocks). */
int hbr_debug_umma_chain_bench(int kind, int reps, int nacc, long long* cycles, void* stream);
/* Latency probe of the forward kernel (in0 = 32, d_view = 24): trace[0..1000) = clock64 stamps of tile group 0 of CTA 0,
 * trace[1024..1524) = the issuing warp's (ready-seen, committed) pairs for that group.  trace holds 2048 int64. */
int hbr_debug_mlp_trace(const float* feat, const float* dirs, int64_t dir_group, int64_t n, const float* params,
                        float* out, long long* trace, void* stream);
/* Same probe for the backward kernel. */
int hbr_debug_mlp_trace_bwd(const float* feat, const float* dirs, int64_t dir_group, int64_t n, const float* params,
                            const float* out, const float* dout, float* dfeat, float* dparams, void* scratch,
                            long long* trace, void* stream);
#ifdef __cplusplus
}
#endif
#endif
