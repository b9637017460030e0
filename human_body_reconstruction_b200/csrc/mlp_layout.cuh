// Flat parameter layout of MLP_3D(num_sig=2, num_col=2, h_size=64) -- test_hash.py:21-51 -- shared by the
// fp32 (CUDA-core) and bf16 (tcgen05) implementations.  Order == state_dict order.
#pragma once
#include "common.cuh"

namespace hbr {

constexpr int kH = 64;          // hidden width
constexpr int kSigOut = 16;     // 1 density + 15 features (test_hash.py:31)
constexpr int kFeat = 15;

struct MlpLayout {
  int in0, dv, kc;              // kc = 15 + d_view
  int W[6], b[6];               // offsets (floats) of weight / bias of layer i in the flat buffer
  int J[6], K[6];               // (out, in) of layer i
  int total;
};

__host__ __device__ inline MlpLayout make_layout(int in0, int dv) {
  MlpLayout m;
  m.in0 = in0; m.dv = dv; m.kc = kFeat + dv;
  const int J[6] = {kH, kH, kSigOut, kH, kH, 3};
  const int K[6] = {in0, kH, kH, kFeat + dv, kH, kH};
  int o = 0;
  for (int i = 0; i < 6; ++i) {
    m.J[i] = J[i]; m.K[i] = K[i];
    m.W[i] = o; o += J[i] * K[i];
    m.b[i] = o; o += J[i];
  }
  m.total = o;
  return m;
}

// rows of the fp32 activation / pre-activation-gradient scratch, each row holds n floats
constexpr int kRowH1 = 0, kRowH2 = 64, kRowO16 = 128, kRowC1 = 144, kRowC2 = 208, kRowRgb = 272, kActRows = 276;

static inline int check_dims(const hbr_mlp_dims* d) {
  HBR_REQUIRE(d != nullptr, "dims is NULL");
  HBR_REQUIRE(d->in0 >= 1 && d->in0 <= 64, "in0=%d out of range [1,64]", d->in0);
  HBR_REQUIRE(d->d_view >= 0 && d->d_view + kFeat <= 64, "d_view=%d out of range [0,49]", d->d_view);
  return HBR_OK;
}

}  // namespace hbr
