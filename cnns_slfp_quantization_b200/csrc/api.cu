// api.cu -- C-ABI dispatch for the convolution entry points (include/slfp_b200.h).
#include "slfp_common.cuh"

namespace slfp {
int conv2d_fwd_dense(const SlfpConvDesc* d, const uint8_t* x_codes, const void* w_f16, const SlfpEpilogue* epi, cudaStream_t st);
bool conv2d_fwd_dense_v2_supported(const SlfpConvDesc* d);
int conv2d_fwd_dense_v2_impl(const SlfpConvDesc* d, const uint8_t* x_codes, const SlfpConvDesc* d2, const uint8_t* x2_codes,
                             const void* w_f16, const SlfpEpilogue* epi, cudaStream_t st);
int conv2d_fwd_grouped(const SlfpConvDesc* d, const uint8_t* x_codes, const void* w_codes, const SlfpEpilogue* epi, cudaStream_t st);
int conv2d_bwd_direct(const SlfpConvDesc* d, const float* gy, const uint8_t* x_codes, const uint8_t* w_codes, int wfmt,
                      float ka, float kw, float* dx, float* dwt, long long so, long long sc, long long sr, long long ss,
                      float* db, cudaStream_t st);
size_t conv2d_bwd_tc_workspace(const SlfpConvDesc* d, int need_dx, int need_dw);
int conv2d_bwd_tc(const SlfpConvDesc* d, const float* gy, const uint8_t* x_codes, const uint8_t* w_codes, int wfmt,
                  float ka, float kw, float* dx, float* dwt, long long so, long long sc, long long sr, long long ss,
                  float* db, void* workspace, size_t ws_bytes, cudaStream_t st, const float* gy_absmax);
}  // namespace slfp

using namespace slfp;

static int check_desc(const SlfpConvDesc* d, const char* who) {
    if (!d) return set_error(SLFP_ERR_BAD_ARG, "%s: null descriptor", who);
    if (d->n <= 0 || d->h <= 0 || d->w <= 0 || d->c <= 0 || d->k <= 0 || d->r <= 0 || d->s <= 0 || d->stride_h <= 0 ||
        d->stride_w <= 0 || d->dil_h <= 0 || d->dil_w <= 0 || d->pad_h < 0 || d->pad_w < 0 || d->groups <= 0 ||
        d->c_phys < d->c)
        return set_error(SLFP_ERR_BAD_ARG, "%s: invalid convolution descriptor", who);
    return 0;
}

extern "C" int slfp_conv2d_fwd(const SlfpConvDesc* desc, const uint8_t* x_codes, const void* w_prepared,
                               const SlfpEpilogue* epi, slfp_stream_t stream) {
    int rc = check_desc(desc, "slfp_conv2d_fwd");
    if (rc) return rc;
    if (!x_codes || !w_prepared || !epi) return set_error(SLFP_ERR_BAD_ARG, "slfp_conv2d_fwd: null pointer");
    if (!epi->y_f32 && !epi->y_f16 && !epi->y_codes) return set_error(SLFP_ERR_BAD_ARG, "slfp_conv2d_fwd: no output");
    // SLFP_FMT_F16Q in / store_f16 out exist in the warp-specialised dense kernel only (conv_igemm_v2.cu)
    if ((desc->fmt == SLFP_FMT_F16Q || epi->store_f16) && (desc->groups != 1 || !conv2d_fwd_dense_v2_supported(desc)))
        return set_error(SLFP_ERR_UNSUPPORTED, "slfp_conv2d_fwd: float16-image activations need a dense layer with c_phys %% 16 == 0");
    if (desc->groups == 1) return conv2d_fwd_dense(desc, x_codes, w_prepared, epi, (cudaStream_t)stream);
    return conv2d_fwd_grouped(desc, x_codes, w_prepared, epi, (cudaStream_t)stream);
}

extern "C" int slfp_conv2d_fwd_dual(const SlfpConvDesc* desc1, const uint8_t* x1_codes, const SlfpConvDesc* desc2,
                                    const uint8_t* x2_codes, const void* w_cat, const SlfpEpilogue* epi, slfp_stream_t stream) {
    int rc = check_desc(desc1, "slfp_conv2d_fwd_dual");
    if (!rc) rc = check_desc(desc2, "slfp_conv2d_fwd_dual");
    if (rc) return rc;
    if (!x1_codes || !x2_codes || !w_cat || !epi) return set_error(SLFP_ERR_BAD_ARG, "slfp_conv2d_fwd_dual: null pointer");
    if (!epi->y_f32 && !epi->y_f16 && !epi->y_codes) return set_error(SLFP_ERR_BAD_ARG, "slfp_conv2d_fwd_dual: no output");
    if (desc1->groups != 1 || !conv2d_fwd_dense_v2_supported(desc1))
        return set_error(SLFP_ERR_UNSUPPORTED, "slfp_conv2d_fwd_dual: dense convolutions with c_phys %% 64 == 0 only");
    return conv2d_fwd_dense_v2_impl(desc1, x1_codes, desc2, x2_codes, w_cat, epi, (cudaStream_t)stream);
}

extern "C" int slfp_conv2d_bwd(const SlfpConvDesc* desc, const float* gy, const uint8_t* x_codes, const uint8_t* w_codes,
                               int wfmt, float ka, float kw, float* dx, float* dw, long long so, long long sc,
                               long long sr, long long ss, float* db, slfp_stream_t stream) {
    int rc = check_desc(desc, "slfp_conv2d_bwd");
    if (rc) return rc;
    if (!gy) return set_error(SLFP_ERR_BAD_ARG, "slfp_conv2d_bwd: null gy");
    return conv2d_bwd_direct(desc, gy, x_codes, w_codes, wfmt, ka, kw, dx, dw, so, sc, sr, ss, db, (cudaStream_t)stream);
}

extern "C" size_t slfp_conv2d_bwd_workspace_size(const SlfpConvDesc* desc, int need_dx, int need_dw) {
    if (check_desc(desc, "slfp_conv2d_bwd_workspace_size")) return 0;
    return conv2d_bwd_tc_workspace(desc, need_dx, need_dw);
}

extern "C" int slfp_conv2d_bwd_ws(const SlfpConvDesc* desc, const float* gy, const uint8_t* x_codes, const uint8_t* w_codes,
                                  int wfmt, float ka, float kw, float* dx, float* dw, long long so, long long sc,
                                  long long sr, long long ss, float* db, void* workspace, size_t workspace_bytes,
                                  slfp_stream_t stream) {
    int rc = check_desc(desc, "slfp_conv2d_bwd_ws");
    if (rc) return rc;
    if (!gy) return set_error(SLFP_ERR_BAD_ARG, "slfp_conv2d_bwd_ws: null gy");
    return conv2d_bwd_tc(desc, gy, x_codes, w_codes, wfmt, ka, kw, dx, dw, so, sc, sr, ss, db, workspace, workspace_bytes,
                         (cudaStream_t)stream, nullptr);
}

extern "C" int slfp_conv2d_bwd_ws_absmax(const SlfpConvDesc* desc, const float* gy, const float* gy_absmax, const uint8_t* x_codes,
                                         const uint8_t* w_codes, int wfmt, float ka, float kw, float* dx, float* dw, long long so,
                                         long long sc, long long sr, long long ss, float* db, void* workspace, size_t workspace_bytes,
                                         slfp_stream_t stream) {
    int rc = check_desc(desc, "slfp_conv2d_bwd_ws_absmax");
    if (rc) return rc;
    if (!gy) return set_error(SLFP_ERR_BAD_ARG, "slfp_conv2d_bwd_ws_absmax: null gy");
    return conv2d_bwd_tc(desc, gy, x_codes, w_codes, wfmt, ka, kw, dx, dw, so, sc, sr, ss, db, workspace, workspace_bytes,
                         (cudaStream_t)stream, gy_absmax);
}
