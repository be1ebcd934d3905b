// Gram / covariance SYRK and the StyleLoss reductions (gram.cu).
#pragma once
#include "common.cuh"
#include "pointwise.cuh"

namespace maua {

size_t gram_workspace_bytes(int C);
// gram[c][d] = (sum_p f[p][c] f[p][d] - [cov] P mu_c mu_d) / (Cn * P);   f is NHWC-flattened [P][C].
// Cn (0 = C) is the channel count the normalisations use: the real count of a layer whose channels are zero-padded to C.
// Optional fusion of the StyleLoss value into the finalize kernel: diff = gram - target, *loss_out = scale * mse.
struct GramLossFuse {
    const float* target = nullptr;
    float* diff = nullptr;
    float* loss_out = nullptr;
    float value_scale = 1.f;
    ReduceScratch rs;
};
int gram_launch(const float* f, long P, int C, int use_cov, float* gram, float* mean_out, void* workspace, int impl,
                cudaStream_t st, const GramLossFuse* fuse = nullptr, int Cn = 0);
int style_loss_fwd_launch(const float* gram, const float* target, int C, float value_scale, float* loss_out,
                          float* diff, ReduceScratch rs, cudaStream_t st);
int style_loss_bwd_prep_launch(const float* diff, const float* mean, int C, long P, const float* coef, float* aux_d,
                               float* aux_bias, cudaStream_t st);
int style_loss_bwd_bias_launch(const float* aux_d, const float* mean, int C, float* aux_bias, cudaStream_t st);
// target = (accumulate ? target : 0) + weight * gram     (StyleLoss capture, loss.py:146-151)
int axpby_launch(const float* x, float* y, long n, float a, int accumulate, cudaStream_t st);

// shared with conv_tc.cu
// atom32 != 0 selects CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B (needed for MN-major TF32 UMMA operands)
int make_tmap_2d(CUtensorMap* m, const float* ptr, long rows, long cols, int box_rows, int atom32);

}  // namespace maua
