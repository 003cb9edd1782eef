// ProtoNet + MDNS episode head (reference models/protonet.py:780-858, ProtoNet_Contrast, eval):
// masked average pooling of the support features (:878-890), one prototype per way from the shots
// the multi-scale degree-based noise suppression kept plus one background prototype (:892-915),
// cosine similarity (x 10) of every query point to every prototype (:917-940).  The encoder, the noise suppression and the loss / prediction kernel are the MPTI
// ones; only the three small kernels below are specific to this model.
#include "common.cuh"

// fg / bg masked means of one support cloud: thread d walks the cloud's points in order
__global__ __launch_bounds__(256) void pn_pool_kernel(const float* __restrict__ F, int64_t ep_rows,
                                                      int64_t sup_row_off, int C, int N, int D,
                                                      const int32_t* __restrict__ sy,
                                                      float* __restrict__ fg_out,
                                                      float* __restrict__ bg_out) {
  const int cloud = blockIdx.x, e = cloud / C, c = cloud % C;
  const int d = threadIdx.x;
  if (d >= D) return;
  const float* f = F + ((int64_t)e * ep_rows + sup_row_off + (int64_t)c * N) * D + d;
  const int32_t* m = sy + (int64_t)cloud * N;
  float fs[4] = {0.f, 0.f, 0.f, 0.f}, bs[4] = {0.f, 0.f, 0.f, 0.f};
  int nf = 0;
  int i = 0;
  for (; i + 4 <= N; i += 4) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float v = f[(int64_t)(i + u) * D];
      const int mk = m[i + u];
      fs[u] += v * (float)mk;            // feat * mask, as the reference multiplies
      bs[u] += mk == 0 ? v : 0.f;        // logical_not(mask)
      nf += mk;
    }
  }
  for (; i < N; ++i) {
    const float v = f[(int64_t)i * D];
    const int mk = m[i];
    fs[0] += v * (float)mk;
    bs[0] += mk == 0 ? v : 0.f;
    nf += mk;
  }
  int nb = 0;  // number of zero entries of the mask
  for (int j = 0; j < N; ++j) nb += m[j] == 0;
  fg_out[(int64_t)cloud * D + d] = ((fs[0] + fs[1]) + (fs[2] + fs[3])) / ((float)nf + 1e-5f);
  bg_out[(int64_t)cloud * D + d] = ((bs[0] + bs[1]) + (bs[2] + bs[3])) / ((float)nb + 1e-5f);
}

// prototypes of one episode: class 0 = mean of the bg vectors of all shots, class 1 + w = mean of
// the fg vectors of way w's kept shots (keep == NULL: all shots)
__global__ __launch_bounds__(256) void pn_proto_kernel(const float* __restrict__ fg,
                                                       const float* __restrict__ bg,
                                                       const int32_t* __restrict__ keep, int n_way,
                                                       int k_shot, int D, float* __restrict__ proto) {
  const int e = blockIdx.x, d = threadIdx.x;
  if (d >= D) return;
  const int C = n_way * k_shot;
  float s = 0.f;
  for (int c = 0; c < C; ++c) s += bg[((int64_t)e * C + c) * D + d];
  proto[((int64_t)e * (n_way + 1)) * D + d] = s / (float)C;
  for (int w = 0; w < n_way; ++w) {
    float a = 0.f, n = 0.f;
    for (int k = 0; k < k_shot; ++k) {
      const int c = w * k_shot + k;
      const float kp = keep ? (keep[e * C + c] ? 1.f : 0.f) : 1.f;
      a += fg[((int64_t)e * C + c) * D + d] * kp;
      n += kp;
    }
    proto[((int64_t)e * (n_way + 1) + 1 + w) * D + d] = a / n;
  }
}

// similarity of every query point to every prototype -> rows of Z; one warp per point
__global__ __launch_bounds__(256) void pn_sim_kernel(const float* __restrict__ F, int64_t ep_rows,
                                                     int64_t q_row_off, int nq, int D,
                                                     const float* __restrict__ proto, int nc,
                                                     float* __restrict__ Z, int nn) {
  const int e = blockIdx.y;
  const int q = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (q >= nq) return;
  const int lane = threadIdx.x & 31;
  const float* x = F + ((int64_t)e * ep_rows + q_row_off + q) * D;
  float xx = 0.f;
  for (int d = lane; d < D; d += 32) xx = fmaf(x[d], x[d], xx);
  xx = warp_sum(xx);
  for (int c = 0; c < nc; ++c) {
    const float* p = proto + ((int64_t)e * nc + c) * D;
    float a = 0.f, b = 0.f;
    for (int d = lane; d < D; d += 32) {
      a = fmaf(x[d], p[d], a);
      b = fmaf(p[d], p[d], b);
    }
    a = warp_sum(a);
    b = warp_sum(b);
    if (lane == 0)  // cosine * scaler (F.cosine_similarity clamps the norm product at 1e-8)
      Z[((int64_t)e * nn + q_row_off + q) * nc + c] = a / fmaxf(sqrtf(xx) * sqrtf(b), 1e-8f) * 10.f;
  }
}

int launch_protonet_head(const float* F, int64_t ep_rows, int64_t sup_row_off, int64_t q_row_off,
                         int E, int n_way, int k_shot, int N, int nq, int D, const int32_t* sy,
                         const int32_t* keep, int method, float* fg, float* bg, float* proto,
                         float* Z, int nn, cudaStream_t st) {
  // only 'cosine' runs in the reference: its 'euclidean' branch reduces over the point axis
  // (F.pairwise_distance on (n_queries, feat_dim, n_points)) and fails in the loss, and the
  // scripts' default 'gaussian' raises NotImplementedError (models/protonet.py:933-939)
  if (D > 256 || method != 0) return R3DFS_E_UNSUPPORTED;
  const int C = n_way * k_shot;
  pn_pool_kernel<<<E * C, 256, 0, st>>>(F, ep_rows, sup_row_off, C, N, D, sy, fg, bg);
  R3DFS_CHECK_LAUNCH();
  pn_proto_kernel<<<E, 256, 0, st>>>(fg, bg, keep, n_way, k_shot, D, proto);
  R3DFS_CHECK_LAUNCH();
  pn_sim_kernel<<<dim3((nq + 7) / 8, E), 256, 0, st>>>(F, ep_rows, q_row_off, nq, D, proto,
                                                       n_way + 1, Z, nn);
  R3DFS_CHECK_LAUNCH();
  return 0;
}
