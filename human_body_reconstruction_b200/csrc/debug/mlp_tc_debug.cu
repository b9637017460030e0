// Debug / probe entry points of the tensor-core MLP kernels (NOT part of the product ABI in include/hbr.h): the UMMA
// operand-mode self test, tensor-pipe micro-benchmarks and the in-kernel clock64 traces.  Built separately into
// libhbr_b200_debug.so (human_body_reconstruction_b200/_lib.py: build_debug()) for tests/test_gpu_tc.py and scripts/dbg_*.py.
#define HBR_DEBUG_ENTRY 1
#include "../mlp_layout.cuh"
#include "../tc_common.cuh"
#include "hbr_debug.h"

#define HBR_OP tc::OpBf16
#define HBR_OPNS bf16
#include "../mlp_tc_impl.cuh"
#undef HBR_OP
#undef HBR_OPNS
#define HBR_OP tc::OpF16
#define HBR_OPNS f16
#include "../mlp_tc_impl.cuh"
#undef HBR_OP
#undef HBR_OPNS

using namespace hbr;
using namespace hbr::bf16;

extern "C" int hbr_debug_umma(int mode, int operand, const float* A, const float* B, float* D, int N, int K, void* stream) {
  HBR_REQUIRE(mode >= 0 && mode <= 2, "mode %d", mode);
  HBR_REQUIRE(operand == HBR_BF16 || operand == HBR_F16, "operand %d", operand);
  HBR_REQUIRE(N % 16 == 0 && N >= 16 && N <= 64 && K % 16 == 0 && K >= 16 && K <= 128, "N=%d K=%d", N, K);
  HBR_REQUIRE(mode != 2 || K == 128, "mode 2 needs K=128");
  auto kern = operand == HBR_F16 ? f16::umma_debug_kernel : bf16::umma_debug_kernel;
  HBR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  kern<<<1, 128, 65536, as_stream(stream)>>>(mode, A, B, D, N, K);
  HBR_LAUNCH_CHECK();
  return HBR_OK;
}

extern "C" int hbr_debug_umma_bench(int M, int N, int reps, int nacc, int mn_major, long long* cycles, void* stream) {
  HBR_REQUIRE((M == 64 || M == 128) && N >= 8 && N <= 256 && N % 8 == 0 && nacc >= 1 && nacc * N <= 512 && reps >= 1,
              "bad shape");
  HBR_CUDA(cudaFuncSetAttribute(umma_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  umma_bench_kernel<<<1, 128, 65536, as_stream(stream)>>>(M, N, reps, nacc, mn_major, cycles);
  HBR_LAUNCH_CHECK();
  return HBR_OK;
}

extern "C" int hbr_debug_umma_chain_bench(int kind, int reps, int nacc, long long* cycles, void* stream) {
  HBR_REQUIRE(kind >= 0 && kind <= 3 && reps >= 1 && nacc >= 1 && nacc <= 6, "bad arguments");
  HBR_CUDA(cudaFuncSetAttribute(umma_chain_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 98304));
  umma_chain_bench_kernel<<<1, 128, 98304, as_stream(stream)>>>(kind, reps, nacc, cycles);
  HBR_LAUNCH_CHECK();
  return HBR_OK;
}

extern "C" int hbr_debug_mlp_trace_bwd(const float* feat, const float* dirs, int64_t dir_group, int64_t n,
                                       const float* params, const float* out, const float* dout, float* dfeat,
                                       float* dparams, void* scratch, long long* trace, void* stream) {
  using SC = Scratch<32, 48>;
  constexpr int smem = BwdSmem<32, 48, 2>::total;
  HBR_CUDA(cudaFuncSetAttribute(mlp_bwd_tc_kernel<32, 48, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int grid = (int)min64(ceil_div(ceil_div(n, kTile), 2), sm_count());
  uint8_t* sc = static_cast<uint8_t*>(scratch);                        // as hbr_mlp_bwd_tc: operand image + gradient rows
  if (sc != nullptr) mlp_prep_kernel<32, 48><<<kPrepCtas, 256, 0, as_stream(stream)>>>(params, 32, 24, sc);
  mlp_bwd_tc_kernel<32, 48, 2, true><<<grid, 2 * kTile + 32, smem, as_stream(stream)>>>(
      feat, 32, dirs, dir_group, n, params, 32, 24, out, dout, dfeat, 32, nullptr, dparams, sc,
      sc != nullptr ? reinterpret_cast<float*>(sc + SC::off_grad) : nullptr, trace, EncArgs{}, HashGeom{}, 1.f, 0, nullptr, nullptr);
  HBR_LAUNCH_CHECK();
  return HBR_OK;
}

extern "C" int hbr_debug_mlp_trace(const float* feat, const float* dirs, int64_t dir_group, int64_t n, const float* params,
                                   float* out, long long* trace, void* stream) {
  constexpr int smem = FwdSmem<32, 48, 4>::total;
  HBR_CUDA(cudaFuncSetAttribute(mlp_fwd_tc_kernel<32, 48, 4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int grid = (int)min64(ceil_div(ceil_div(n, kTile), 4), sm_count());
  mlp_fwd_tc_kernel<32, 48, 4, true><<<grid, 4 * kTile, smem, as_stream(stream)>>>(
      feat, 32, dirs, dir_group, n, params, 32, 24, out, nullptr, trace, EncArgs{}, HashGeom{}, 0, nullptr, nullptr);
  HBR_LAUNCH_CHECK();
  return HBR_OK;
}

