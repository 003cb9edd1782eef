// internal launchers of the meta-training path (train_kernels.cu, train_graph.cu)
#pragma once
#include "common.cuh"

// C[b] (M x N, row stride ldc) = alpha * A[b] * B[b] + beta * C[b] in FP32 on the CUDA cores, with
// arbitrary element strides:  A(m, k) = A[b * bsA + m * sAm + k * sAk],  B(k, n) = B[b * bsB +
// k * sBk + n * sBn].  splits > 1 cuts K into `splits` ranges whose partial products are summed in
// range order by a second kernel (deterministic); `partial` then needs batch * splits * M * N floats.
int launch_sgemm(const float* A, int64_t sAm, int64_t sAk, int64_t bsA, const float* B, int64_t sBk,
                 int64_t sBn, int64_t bsB, float* C, int64_t ldc, int64_t bsC, int M, int N, int K,
                 int batch, float alpha, float beta, int splits, float* partial, cudaStream_t st);
// C = alpha * sum over splits of partial[split] (M x N each, dense) + beta * C, fixed order
int launch_sgemm_reduce(const float* partial, int M, int N, int splits, float alpha, float beta,
                        float* C, int64_t ldc, cudaStream_t st);
// tensor-core C[m][n] = sum_r A[r][m] B[r][n] into split partials (tc_gemm.cu)
int launch_gemm_tn_tc(const float* A, int64_t lda, const float* B, int64_t ldb, int M, int N,
                      int64_t R, int max_splits, float* partial, int* splits_out, cudaStream_t st);
// number of K ranges launch_sgemm should use for a (M x N) output reduced over K
int sgemm_splits(int M, int N, int64_t K, int batch);

// BatchNorm with batch statistics over `rows` rows of C channels (x rows ldx floats apart).
// stats[0..C) = mean, stats[C..2C) = 1 / sqrt(var_biased + eps).  running (mean | var, 2C floats,
// may be NULL) is updated with momentum (unbiased variance), as nn.BatchNorm does in train mode.
// scratch: 2 * C * BN_MAX_BLOCKS doubles.
#define BN_MAX_BLOCKS 592
int launch_bn_stats(const float* x, int64_t ldx, int64_t rows, int C, float eps, float momentum,
                    float* running, float* stats, double* scratch, cudaStream_t st);
// y = act(gamma * (x - mean) * invstd + beta)
int launch_bn_act(const float* x, int64_t ldx, int64_t rows, int C, const float* stats,
                  const float* gamma, const float* beta, int act, float* y, int64_t ldy,
                  cudaStream_t st);
// backward of y = act(BN(x)): dx (may alias dy) and dgamma += / dbeta += (accumulated).
int launch_bn_act_bwd(const float* dy, int64_t ld_dy, const float* x, int64_t ldx, int64_t rows,
                      int C, const float* stats, const float* gamma, const float* beta, int act,
                      float* dx, int64_t ld_dx, float* dgamma, float* dbeta, double* scratch,
                      cudaStream_t st);
// out[c] += sum over rows of x[r][c]
int launch_col_sum_acc(const float* x, int64_t ldx, int64_t rows, int C, float* out, double* scratch,
                       cudaStream_t st);

// EdgeConv pieces with the (rows = B*N*k edges, 64 channels) activations materialised
int launch_edge_pre(const float* PQ, const int32_t* idx, int64_t B, int N, int k, float* h1pre,
                    cudaStream_t st);
int launch_edge_max(const float* h2pre, const float* stats, const float* gamma, const float* beta,
                    int64_t M, int k, float* y, int64_t ldy, uint8_t* arg, cudaStream_t st);
int launch_edge_max_bwd(const float* dy, int64_t ld_dy, const uint8_t* arg, int64_t M, int k,
                        float* dA2, cudaStream_t st);
int launch_edge_pre_bwd(const float* dh1, const int32_t* idx, int64_t B, int N, int k, float* dPQ,
                        cudaStream_t st);
int launch_unfold_w1_grad(const float* dWf, int C, float* dW1, cudaStream_t st);

// attention map: softmax over each row of S (rows x n) in place; with a keep mask (may be NULL)
// Pd = P * mask / (1 - p) is written to Pd (else Pd is not touched)
int launch_softmax_rows(float* S, int64_t rows, int n, const uint8_t* mask, float p, float* Pd,
                        cudaStream_t st);
// dS (in place over dPd) = P * (dP - sum_j dP * P),  dP = dPd * mask / (1 - p)
int launch_softmax_rows_bwd(const float* P, float* dPd, int64_t rows, int n, const uint8_t* mask,
                            float p, cudaStream_t st);
int launch_dropout_mask(uint64_t seed, int64_t n, float p, uint8_t* mask, cudaStream_t st);

int launch_add_cols(const float* src, int64_t lds, int64_t rows, int ncols, float* dst, int64_t ldd,
                    cudaStream_t st);  // dst[:, 0:ncols] += src[:, 0:ncols]
int launch_copy_cols_plain(const float* src, int64_t lds, int64_t rows, int ncols, float* dst,
                           int64_t ldd, cudaStream_t st);
int launch_transpose(const float* src, int rows, int cols, float* dst, cudaStream_t st);
int launch_fill_f32(float* p, int64_t n, float v, cudaStream_t st);

int launch_adam(float* p, const float* g, float* m, float* v, int64_t n, int64_t n_group0,
                float lr0, float lr1, float beta1, float beta2, float eps, float bc1, float bc2,
                float grad_scale, cudaStream_t st);

// ---- graph half (train_graph.cu) -------------------------------------------------------------
// dZ of the mean cross-entropy over the query rows (scaled by w), zero elsewhere; (nn, nc) rows
int launch_ce_grad(const float* Z, int nn, int q_off, int nq, int nc, const int64_t* qy, float w,
                   float* dZ, cudaStream_t st);
// adjoint of label propagation on the merged rows: per-edge dL/dsim (times the Gaussian's own
// derivative factor -sim / sigma^2) -> gE (nn, k)
int launch_lp_adjoint_edges(const int32_t* rowptr, const int32_t* rowlen, const uint16_t* mcol,
                            const float* mval, const float* dinv, const uint8_t* valid,
                            const int32_t* nbr, const float* sim, int nn, int k, int nc,
                            const float* Z, const float* Gm, float alpha, float sigma, float* dD,
                            float* gE, cudaStream_t st);
// dF (node rows, D floats) += edge terms of gE: d/df_i and d/df_j of exp(-0.5 |f_i - f_j + 1e-6|^2 / s^2)
int launch_sim_bwd(const float* F, int D, const uint8_t* valid, const int32_t* nbr, const float* gE,
                   int nn, int k, float* dF, cudaStream_t st);
// member counts of every prototype from the chunk partial counts of proto_partial_kernel
int launch_proto_counts(const int32_t* pcount, const int32_t* set_n, const int32_t* proto_cnt,
                        int n_sets, int m_max, int n_chunks, int32_t* count_out, cudaStream_t st);
// gradient of the support rows: every support point receives dproto[its prototype] / members
// (graph prototypes, rows of dFnode) plus, for foreground points, the same from its shot's
// contrast prototypes (dcproto, may be NULL)
int launch_support_grad(const float* dFnode, int slot, const int32_t* assign,
                        const int32_t* pcnt_members, const float* dcproto, int cslot,
                        const int32_t* cassign, const int32_t* ccnt_members, int n_way, int k_shot,
                        int N, int D, const int32_t* sy, const int32_t* cloud_bg_off,
                        const int32_t* cloud_fg_off, float* dFsup, cudaStream_t st);
// way-contrast loss (models/mpti.py:226-313) of one way from the per-shot prototypes;
// backward != 0 also accumulates dproj_w / dproj_b / dcproto (scaled by w)
int launch_contrast(const float* cproto, const int32_t* cproto_cnt, int cslot, int D,
                    const int32_t* support_flag, int n_way, int k_shot, int way,
                    const float* proj_w, const float* proj_b, float temp, float* loss_way,
                    int backward, float w, float* dproj_w, float* dproj_b, float* dcproto,
                    cudaStream_t st);
