// Backward of the graph half of the training forward (reference models/mpti.py:488-571 under
// autograd) and the way-contrast loss (models/mpti.py:226-313).
//
// Label propagation Z = (I - alpha S)^-1 Y with S = r W r, r = (rowsum(W) + eps)^-1/2, W = A + A^T:
// with G = (I - alpha S)^-1 dZ (S is symmetric, so the adjoint solve is the forward solver again),
//   dS_ij = alpha G_i . Z_j,
//   dr_i  = sum_j (dS_ij + dS_ji) W_ij r_j,       dD_i = -1/2 r_i^3 dr_i,
//   dA_ij = r_i r_j (dS_ij + dS_ji) + dD_i + dD_j         (a_ij feeds W_ij and W_ji).
// Only the k-sparse pattern is ever touched: kNN indices carry no gradient.
#include "train.cuh"

// ---------------------------------------------------------------------------------------------
__global__ void ce_grad_kernel(const float* __restrict__ Z, int nn, int q_off, int nq, int nc,
                               const int64_t* __restrict__ qy, float w, float* __restrict__ dZ) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nn) return;
  float* d = dZ + (int64_t)i * nc;
  if (i < q_off || i >= q_off + nq) {
    for (int c = 0; c < nc; ++c) d[c] = 0.f;
    return;
  }
  const float* z = Z + (int64_t)i * nc;
  float mx = -INFINITY;
  for (int c = 0; c < nc; ++c) mx = fmaxf(mx, z[c]);
  float se = 0.f;
  for (int c = 0; c < nc; ++c) se += expf(z[c] - mx);
  const int y = (int)qy[i - q_off];
  const float s = w / (float)nq;
  for (int c = 0; c < nc; ++c) d[c] = s * (expf(z[c] - mx) / se - (c == y ? 1.f : 0.f));
}

int launch_ce_grad(const float* Z, int nn, int q_off, int nq, int nc, const int64_t* qy, float w,
                   float* dZ, cudaStream_t st) {
  ce_grad_kernel<<<(nn + 255) / 256, 256, 0, st>>>(Z, nn, q_off, nq, nc, qy, w, dZ);
  R3DFS_CHECK_LAUNCH();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// q_ij = alpha (G_i . Z_j + G_j . Z_i)
__device__ __forceinline__ float pair_q(const float* __restrict__ Z, const float* __restrict__ G,
                                        int i, int j, int nc, float alpha) {
  float s = 0.f;
  for (int c = 0; c < nc; ++c)
    s += G[(int64_t)i * nc + c] * Z[(int64_t)j * nc + c] + G[(int64_t)j * nc + c] * Z[(int64_t)i * nc + c];
  return alpha * s;
}

// dD_i = -1/2 r_i^3 dr_i with dr_i = (1 / r_i) sum_{j in row i} q_ij S_ij   (W_ij r_j = S_ij / r_i)
__global__ __launch_bounds__(256) void lp_adjoint_degree_kernel(
    const int32_t* __restrict__ rowptr, const int32_t* __restrict__ rowlen,
    const uint16_t* __restrict__ mcol, const float* __restrict__ mval,
    const float* __restrict__ dinv, int nn, int nc, const float* __restrict__ Z,
    const float* __restrict__ G, float alpha, float* __restrict__ dD) {
  const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (i >= nn) return;
  const int lane = threadIdx.x & 31;
  const int L = rowlen[i];
  float acc = 0.f;
  if (L > 0) {
    const int64_t mb = rowptr[i];
    for (int t = lane; t < L; t += 32) {
      const int j = mcol[mb + t];
      acc += pair_q(Z, G, i, j, nc, alpha) * mval[mb + t];
    }
  }
  acc = warp_sum(acc);
  if (lane == 0) {
    const float r = dinv[i];
    dD[i] = L > 0 ? -0.5f * r * r * acc : 0.f;
  }
}

// gE[i][s] = dL/dsim_is * d sim / d(dist^2 / 2 sigma^2 term) = (r_i r_j q_ij + dD_i + dD_j) * (-sim / sigma^2)
__global__ void lp_adjoint_edge_kernel(const int32_t* __restrict__ nbr, const float* __restrict__ sim,
                                       const uint8_t* __restrict__ valid,
                                       const float* __restrict__ dinv, const float* __restrict__ dD,
                                       int nn, int k, int nc, const float* __restrict__ Z,
                                       const float* __restrict__ G, float alpha, float inv_s2,
                                       float* __restrict__ gE) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (int64_t)nn * k) return;
  const int i = (int)(e / k);
  if (!valid[i]) {
    gE[e] = 0.f;
    return;
  }
  const int j = nbr[e];
  const float dA = dinv[i] * dinv[j] * pair_q(Z, G, i, j, nc, alpha) + dD[i] + dD[j];
  gE[e] = -dA * sim[e] * inv_s2;
}

int launch_lp_adjoint_edges(const int32_t* rowptr, const int32_t* rowlen, const uint16_t* mcol,
                            const float* mval, const float* dinv, const uint8_t* valid,
                            const int32_t* nbr, const float* sim, int nn, int k, int nc,
                            const float* Z, const float* Gm, float alpha, float sigma, float* dD,
                            float* gE, cudaStream_t st) {
  lp_adjoint_degree_kernel<<<(nn + 7) / 8, 256, 0, st>>>(rowptr, rowlen, mcol, mval, dinv, nn, nc, Z,
                                                        Gm, alpha, dD);
  R3DFS_CHECK_LAUNCH();
  const int64_t edges = (int64_t)nn * k;
  lp_adjoint_edge_kernel<<<(unsigned)((edges + 255) / 256), 256, 0, st>>>(
      nbr, sim, valid, dinv, dD, nn, k, nc, Z, Gm, alpha, 1.f / (sigma * sigma), gE);
  R3DFS_CHECK_LAUNCH();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// dF_i += g (f_i - f_j + 1e-6),  dF_j -= g (f_i - f_j + 1e-6)  for every kept edge (i -> j).
// One warp per node; the node's own row is accumulated in registers, the neighbours' rows with
// float atomics (red.global.add.f32).
// ---------------------------------------------------------------------------------------------
#define SB_MAX_F4 2  // D <= 256

__global__ __launch_bounds__(256) void sim_bwd_kernel(const float* __restrict__ F, int D,
                                                      const uint8_t* __restrict__ valid,
                                                      const int32_t* __restrict__ nbr,
                                                      const float* __restrict__ gE, int nn, int k,
                                                      float* __restrict__ dF) {
  const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (i >= nn || !valid[i]) return;
  const int lane = threadIdx.x & 31;
  const int D4 = D >> 2;
  float4 fi[SB_MAX_F4], acc[SB_MAX_F4];
#pragma unroll
  for (int q = 0; q < SB_MAX_F4; ++q) {
    const int c4 = lane + 32 * q;
    fi[q] = c4 < D4 ? reinterpret_cast<const float4*>(F + (int64_t)i * D)[c4] : make_float4(0, 0, 0, 0);
    acc[q] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int s = 0; s < k; ++s) {
    const int j = nbr[(int64_t)i * k + s];
    const float g = gE[(int64_t)i * k + s];
    if (g == 0.f) continue;
#pragma unroll
    for (int q = 0; q < SB_MAX_F4; ++q) {
      const int c4 = lane + 32 * q;
      if (c4 >= D4) continue;
      const float4 fj = reinterpret_cast<const float4*>(F + (int64_t)j * D)[c4];
      const float dx = g * (fi[q].x - fj.x + 1e-6f), dy = g * (fi[q].y - fj.y + 1e-6f);
      const float dz = g * (fi[q].z - fj.z + 1e-6f), dw = g * (fi[q].w - fj.w + 1e-6f);
      acc[q].x += dx; acc[q].y += dy; acc[q].z += dz; acc[q].w += dw;
      float* o = dF + (int64_t)j * D + 4 * c4;
      atomicAdd(o, -dx);
      atomicAdd(o + 1, -dy);
      atomicAdd(o + 2, -dz);
      atomicAdd(o + 3, -dw);
    }
  }
#pragma unroll
  for (int q = 0; q < SB_MAX_F4; ++q) {
    const int c4 = lane + 32 * q;
    if (c4 >= D4) continue;
    float* o = dF + (int64_t)i * D + 4 * c4;
    atomicAdd(o, acc[q].x);
    atomicAdd(o + 1, acc[q].y);
    atomicAdd(o + 2, acc[q].z);
    atomicAdd(o + 3, acc[q].w);
  }
}

int launch_sim_bwd(const float* F, int D, const uint8_t* valid, const int32_t* nbr, const float* gE,
                   int nn, int k, float* dF, cudaStream_t st) {
  if (D % 4 != 0 || D > 128 * SB_MAX_F4) return R3DFS_E_UNSUPPORTED;
  sim_bwd_kernel<<<(nn + 7) / 8, 256, 0, st>>>(F, D, valid, nbr, gE, nn, k, dF);
  R3DFS_CHECK_LAUNCH();
  return 0;
}

// ---------------------------------------------------------------------------------------------
#define PM_CHUNK_T 2048  // must match PM_CHUNK in proto.cu

__global__ void proto_counts_kernel(const int32_t* __restrict__ pcount,
                                    const int32_t* __restrict__ set_n,
                                    const int32_t* __restrict__ proto_cnt, int m_max, int n_chunks,
                                    int32_t* __restrict__ count_out) {
  const int set = blockIdx.x, p = threadIdx.x;
  if (p >= m_max) return;
  int c = 0;
  if (p < proto_cnt[set]) {
    const int nch = (set_n[set] + PM_CHUNK_T - 1) / PM_CHUNK_T;
    for (int ch = 0; ch < nch; ++ch) c += pcount[((int64_t)set * n_chunks + ch) * m_max + p];
  }
  count_out[set * m_max + p] = c;
}

int launch_proto_counts(const int32_t* pcount, const int32_t* set_n, const int32_t* proto_cnt,
                        int n_sets, int m_max, int n_chunks, int32_t* count_out, cudaStream_t st) {
  if (m_max > 128) return R3DFS_E_UNSUPPORTED;
  proto_counts_kernel<<<n_sets, 128, 0, st>>>(pcount, set_n, proto_cnt, m_max, n_chunks, count_out);
  R3DFS_CHECK_LAUNCH();
  return 0;
}

// CTA per support cloud; same ordered compaction as set_gather_kernel (proto.cu), run backwards:
// support point -> its row in the set buffer -> its prototype -> gradient share.
__global__ __launch_bounds__(256) void support_grad_kernel(
    const float* __restrict__ dFnode, int slot, const int32_t* __restrict__ assign,
    const int32_t* __restrict__ members, const float* __restrict__ dcproto, int cslot,
    const int32_t* __restrict__ cassign, const int32_t* __restrict__ cmembers, int k_shot, int N,
    int D, const int32_t* __restrict__ sy, const int32_t* __restrict__ cloud_bg_off,
    const int32_t* __restrict__ cloud_fg_off, float* __restrict__ dFsup) {
  __shared__ int s_dst[256];
  __shared__ int s_fg[256];
  __shared__ int s_w[2][8];
  const int cloud = blockIdx.x;
  const int way = cloud / k_shot;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  int bg_run = cloud_bg_off[cloud];
  const int fg0 = cloud_fg_off[cloud];
  int fg_run = fg0;
  const int D4 = D >> 2;
  for (int c0 = 0; c0 < N; c0 += 256) {
    const int i = c0 + tid;
    const bool in = i < N;
    const bool fg = in && sy[(int64_t)cloud * N + i] != 0;
    const bool bg = in && !fg;
    const unsigned bf = __ballot_sync(0xffffffffu, fg), bb = __ballot_sync(0xffffffffu, bg);
    if (lane == 0) {
      s_w[0][w] = __popc(bf);
      s_w[1][w] = __popc(bb);
    }
    __syncthreads();
    int offf = 0, offb = 0, totf = 0, totb = 0;
    for (int q = 0; q < 8; ++q) {
      if (q < w) {
        offf += s_w[0][q];
        offb += s_w[1][q];
      }
      totf += s_w[0][q];
      totb += s_w[1][q];
    }
    const unsigned lt = (1u << lane) - 1;
    int dst = -1;
    if (fg && fg0 >= 0) dst = fg_run + offf + __popc(bf & lt);
    if (bg) dst = bg_run + offb + __popc(bb & lt);
    s_dst[tid] = dst;
    s_fg[tid] = fg ? 1 : 0;
    __syncthreads();
    const int cnt = min(256, N - c0);
    for (int r = w; r < cnt; r += 8) {
      const int d = s_dst[r];
      float4* out = reinterpret_cast<float4*>(dFsup + ((int64_t)cloud * N + c0 + r) * D);
      if (d < 0) {
        for (int q = lane; q < D4; q += 32) out[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        continue;
      }
      const int set = s_fg[r] ? 1 + way : 0;
      const int p = assign[d];
      const float inv = 1.f / (float)members[set * slot + p];
      const float4* src = reinterpret_cast<const float4*>(dFnode + ((int64_t)set * slot + p) * D);
      const float4* csrc = nullptr;
      float cinv = 0.f;
      if (dcproto && s_fg[r]) {
        const int cp = cassign[d];
        cinv = 1.f / (float)cmembers[cloud * cslot + cp];
        csrc = reinterpret_cast<const float4*>(dcproto + ((int64_t)cloud * cslot + cp) * D);
      }
      for (int q = lane; q < D4; q += 32) {
        float4 v = src[q];
        v.x *= inv; v.y *= inv; v.z *= inv; v.w *= inv;
        if (csrc) {
          const float4 u = csrc[q];
          v.x += cinv * u.x; v.y += cinv * u.y; v.z += cinv * u.z; v.w += cinv * u.w;
        }
        out[q] = v;
      }
    }
    fg_run += totf;
    bg_run += totb;
    __syncthreads();
  }
}

int launch_support_grad(const float* dFnode, int slot, const int32_t* assign,
                        const int32_t* pcnt_members, const float* dcproto, int cslot,
                        const int32_t* cassign, const int32_t* ccnt_members, int n_way, int k_shot,
                        int N, int D, const int32_t* sy, const int32_t* cloud_bg_off,
                        const int32_t* cloud_fg_off, float* dFsup, cudaStream_t st) {
  support_grad_kernel<<<n_way * k_shot, 256, 0, st>>>(dFnode, slot, assign, pcnt_members, dcproto,
                                                     cslot, cassign, ccnt_members, k_shot, N, D, sy,
                                                     cloud_bg_off, cloud_fg_off, dFsup);
  R3DFS_CHECK_LAUNCH();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Way-contrast loss of ONE way (models/mpti.py:247-309).  Members: every fps_k-prototype of the
// way's shots (label = support_flag[way][shot]) and, when way 0's shots all carry one class
// (:238-244), the prototypes of the first two shots of the next way as negatives (label -1).
// z = normalize(proj(p)); logits = z z^T / temp; SupCon with self-pairs masked.  One CTA.
// ---------------------------------------------------------------------------------------------
#define CT_MAXK 64
#define CT_PD 128  // projection width

__global__ __launch_bounds__(1024) void contrast_kernel(
    const float* __restrict__ cproto, const int32_t* __restrict__ cproto_cnt, int cslot, int D,
    const int32_t* __restrict__ support_flag, int n_way, int k_shot, int way,
    const float* __restrict__ proj_w, const float* __restrict__ proj_b, float temp,
    float* __restrict__ loss_way, int backward, float wscale, float* __restrict__ dproj_w,
    float* __restrict__ dproj_b, float* __restrict__ dcproto) {
  extern __shared__ __align__(16) float smem[];
  float* s_z = smem;                           // [K][128] normalised projections
  float* s_du = s_z + CT_MAXK * CT_PD;         // [K][128] gradient wrt the un-normalised projection
  float* s_l = s_du + CT_MAXK * CT_PD;         // [K][K] logits, then dlogits
  float* s_norm = s_l + CT_MAXK * CT_MAXK;     // [K]
  float* s_lab = s_norm + CT_MAXK;             // [K]
  float* s_row = s_lab + CT_MAXK;              // [K] log-denominator
  float* s_npos = s_row + CT_MAXK;             // [K]
  __shared__ int s_src[CT_MAXK];               // global prototype row (cloud * cslot + p)
  __shared__ int s_K;
  const int tid = threadIdx.x;
  if (tid == 0) {
    int total = 0;
    for (int s = 0; s < k_shot; ++s) total += support_flag[s];
    const bool clean = support_flag[0] * k_shot == total;
    int K = 0;
    for (int s = 0; s < k_shot; ++s) {
      const int cloud = way * k_shot + s;
      for (int p = 0; p < cproto_cnt[cloud] && K < CT_MAXK; ++p) {
        s_src[K] = cloud * cslot + p;
        s_lab[K] = (float)support_flag[cloud];
        ++K;
      }
    }
    if (clean) {
      const int other = way < n_way - 1 ? way + 1 : 0;
      for (int s = 0; s < 2 && s < k_shot; ++s) {
        const int cloud = other * k_shot + s;
        for (int p = 0; p < cproto_cnt[cloud] && K < CT_MAXK; ++p) {
          s_src[K] = cloud * cslot + p;
          s_lab[K] = -1.f;
          ++K;
        }
      }
    }
    s_K = K;
  }
  __syncthreads();
  const int K = s_K;
  // u = W p + b: one warp per output, lanes along the input dimension (coalesced weight rows)
  for (int e = tid >> 5; e < K * CT_PD; e += blockDim.x >> 5) {
    const int a = e / CT_PD, o = e % CT_PD;
    const float* p = cproto + (int64_t)s_src[a] * D;
    const float* wr = proj_w + (int64_t)o * D;
    float s = 0.f;
    for (int d = tid & 31; d < D; d += 32) s = fmaf(wr[d], p[d], s);
    s = warp_sum(s);
    if ((tid & 31) == 0) s_z[e] = s + proj_b[o];
  }
  __syncthreads();
  if (tid < K) {
    float s = 0.f;
    for (int o = 0; o < CT_PD; ++o) s += s_z[tid * CT_PD + o] * s_z[tid * CT_PD + o];
    s_norm[tid] = fmaxf(sqrtf(s), 1e-12f);
  }
  __syncthreads();
  for (int e = tid; e < K * CT_PD; e += blockDim.x) s_z[e] /= s_norm[e / CT_PD];
  __syncthreads();
  for (int e = tid; e < K * K; e += blockDim.x) {
    const int a = e / K, b = e % K;
    float s = 0.f;
    for (int o = 0; o < CT_PD; ++o) s += s_z[a * CT_PD + o] * s_z[b * CT_PD + o];
    s_l[a * CT_MAXK + b] = s / temp;
  }
  __syncthreads();
  if (tid < K) {
    float den = 0.f, npos = 0.f, pos = 0.f;
    for (int b = 0; b < K; ++b) {
      if (b == tid) continue;
      den += expf(s_l[tid * CT_MAXK + b]);
      if (s_lab[b] == s_lab[tid]) {
        npos += 1.f;
        pos += s_l[tid * CT_MAXK + b];
      }
    }
    s_row[tid] = logf(den);
    s_npos[tid] = npos;
    // -(sum_pos (logit - log den)) / npos
    s_du[tid] = -(pos - npos * logf(den)) / npos;  // scratch: per-anchor loss
  }
  __syncthreads();
  if (tid == 0) {
    float s = 0.f;
    for (int a = 0; a < K; ++a) s += s_du[a];
    loss_way[way] = s / (float)K;
  }
  if (!backward) return;
  __syncthreads();
  // dlogits_ab = (w / K) (softmax_ab - mask_ab / npos_a), zero diagonal
  for (int e = tid; e < K * K; e += blockDim.x) {
    const int a = e / K, b = e % K;
    float g = 0.f;
    if (a != b) {
      g = expf(s_l[a * CT_MAXK + b] - s_row[a]);
      if (s_lab[a] == s_lab[b]) g -= 1.f / s_npos[a];
      g *= wscale / (float)K;
    }
    s_l[a * CT_MAXK + b] = g;
  }
  __syncthreads();
  // dz_a = sum_b (dl_ab + dl_ba) z_b / temp ; du_a = (dz_a - z_a (z_a . dz_a)) / norm_a
  for (int e = tid; e < K * CT_PD; e += blockDim.x) {
    const int a = e / CT_PD, o = e % CT_PD;
    float s = 0.f;
    for (int b = 0; b < K; ++b) s += (s_l[a * CT_MAXK + b] + s_l[b * CT_MAXK + a]) * s_z[b * CT_PD + o];
    s_du[e] = s / temp;
  }
  __syncthreads();
  if (tid < K) {
    float dot = 0.f;
    for (int o = 0; o < CT_PD; ++o) dot += s_z[tid * CT_PD + o] * s_du[tid * CT_PD + o];
    s_row[tid] = dot;
  }
  __syncthreads();
  for (int e = tid; e < K * CT_PD; e += blockDim.x) {
    const int a = e / CT_PD;
    s_du[e] = (s_du[e] - s_z[e] * s_row[a]) / s_norm[a];
  }
  __syncthreads();
  // dW[o][d] += sum_a du[a][o] p_a[d] ;  db[o] += sum_a du[a][o] ;  dp_a[d] += sum_o W[o][d] du[a][o]
  for (int e = tid; e < CT_PD * D; e += blockDim.x) {
    const int o = e / D, d = e % D;
    float s = 0.f;
    for (int a = 0; a < K; ++a) s += s_du[a * CT_PD + o] * cproto[(int64_t)s_src[a] * D + d];
    dproj_w[e] += s;
  }
  for (int o = tid; o < CT_PD; o += blockDim.x) {
    float s = 0.f;
    for (int a = 0; a < K; ++a) s += s_du[a * CT_PD + o];
    dproj_b[o] += s;
  }
  for (int e = tid; e < K * D; e += blockDim.x) {
    const int a = e / D, d = e % D;
    float s = 0.f;
    for (int o = 0; o < CT_PD; ++o) s += proj_w[(int64_t)o * D + d] * s_du[a * CT_PD + o];
    dcproto[(int64_t)s_src[a] * D + d] += s;
  }
}

int launch_contrast(const float* cproto, const int32_t* cproto_cnt, int cslot, int D,
                    const int32_t* support_flag, int n_way, int k_shot, int way,
                    const float* proj_w, const float* proj_b, float temp, float* loss_way,
                    int backward, float w, float* dproj_w, float* dproj_b, float* dcproto,
                    cudaStream_t st) {
  if ((k_shot + 2) * cslot > CT_MAXK) return R3DFS_E_UNSUPPORTED;
  const size_t smem = sizeof(float) * (2 * CT_MAXK * CT_PD + CT_MAXK * CT_MAXK + 4 * CT_MAXK);
  cudaError_t e = cudaFuncSetAttribute(contrast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem);
  if (e != cudaSuccess) return (int)e;
  contrast_kernel<<<1, 1024, smem, st>>>(cproto, cproto_cnt, cslot, D, support_flag, n_way, k_shot,
                                        way, proj_w, proj_b, temp, loss_way, backward, w, dproj_w,
                                        dproj_b, dcproto);
  R3DFS_CHECK_LAUNCH();
  return 0;
}
