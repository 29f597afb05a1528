// Host-side pieces of the "virtual concat" decoder convolution
//
//   y = act( conv3( GN( cat(enc, nearest_up2x(b)) ) ) )          (Decoder: buildingblocks.py:466-497, DoubleConv conv1)
//
// without materialising the upsampled / concatenated tensor.  The convolution is linear in its input channels:
//   conv3(cat(enc, up(b))) = conv3_enc(enc) + conv3_up(up(b)),
// and a 3x3x3 convolution of a nearest-2x-upsampled tensor is, for every output parity phase, a 2x2x2 convolution of the
// low-res tensor with summed weights (8/27 of the MACs).  The tensor-core launches live in conv_igemm_sm100.cu /
// wgrad_igemm_sm100.cu (b200_conv3_up_*); this file folds the GroupNorm affine into the two weight sets and the phase-aware
// border-class bias table, prepares the dgrad weights, and re-assembles the 27-tap weight gradient.
//
// Per axis, t = tap offset in {-1,0,+1}, p = output parity, j = low-res tap index in {0,1}:
//   p = 0: j=0 <- {t=-1} (low-res offset -1),  j=1 <- {t=0,+1} (offset 0)
//   p = 1: j=0 <- {t=-1,0} (offset 0),         j=1 <- {t=+1}   (offset +1)
// and for the transpose (gradient w.r.t. b / weight gradient), e = r - t in {-1,0,1,2} (r = parity of the upsampled index):
//   e=-1 <- {t=+1},  e=0 <- {t=0,+1},  e=1 <- {t=-1,0},  e=2 <- {t=-1}
#include "common.cuh"

namespace b200 {

// does (parity p, low-res tap j) contain tap index t (0,1,2 <-> -1,0,+1) along one axis?
__host__ __device__ __forceinline__ bool phase_has_tap(int p, int j, int t) {
  return p == 0 ? (j == 0 ? t == 0 : t >= 1) : (j == 0 ? t <= 1 : t == 2);
}
// does transpose offset index ei (0..3 <-> e = -1..2) contain tap index t?
__host__ __device__ __forceinline__ bool offset_has_tap(int ei, int t) {
  return ei == 0 ? t == 2 : (ei == 1 ? t >= 1 : (ei == 2 ? t <= 1 : t == 0));
}

// wf_enc[n][t][co][c0] = bf16(W[co][c0][t] * a[n][c0])
__global__ void upcat_fold_enc_kernel(const float* __restrict__ W, const float* __restrict__ ab, int n_w, int C0, int C, int Cout,
                                      bf16* __restrict__ wf) {
  size_t total = (size_t)n_w * 27 * Cout * C0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    int ci = (int)(i % C0);
    size_t r = i / C0;
    int co = (int)(r % Cout);
    r /= Cout;
    int tap = (int)(r % 27);
    int n = (int)(r / 27);
    float a = ab ? ab[((size_t)n * C + ci) * 2] : 1.f;
    wf[i] = to_act(W[((size_t)co * C + ci) * 27 + tap] * a);
  }
}
// wp[n][phase*8 + j][co][c1] = bf16( sum_{t in S(phase,j)} W[co][C0+c1][t] * a[n][C0+c1] )
__global__ void upcat_fold_phase_kernel(const float* __restrict__ W, const float* __restrict__ ab, int n_w, int C0, int C1, int Cout,
                                        bf16* __restrict__ wp) {
  const int C = C0 + C1;
  size_t total = (size_t)n_w * 64 * Cout * C1;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    int c1 = (int)(i % C1);
    size_t r = i / C1;
    int co = (int)(r % Cout);
    r /= Cout;
    int pj = (int)(r % 64);
    int n = (int)(r / 64);
    const int phase = pj >> 3, j = pj & 7;
    const float* w = W + ((size_t)co * C + C0 + c1) * 27;
    float acc = 0.f;
    for (int td = 0; td < 3; ++td)
      if (phase_has_tap((phase >> 2) & 1, (j >> 2) & 1, td))
        for (int th = 0; th < 3; ++th)
          if (phase_has_tap((phase >> 1) & 1, (j >> 1) & 1, th))
            for (int tw = 0; tw < 3; ++tw)
              if (phase_has_tap(phase & 1, j & 1, tw)) acc += w[(td * 3 + th) * 3 + tw];
    float a = ab ? ab[((size_t)n * C + C0 + c1) * 2] : 1.f;
    wp[i] = to_act(acc * a);
  }
}

// bias table [n][64][Cout] with PHASE-AWARE classes (per axis: 0 low face, 1 interior even, 2 high face, 3 interior odd):
//   conv_bias + sum_{valid taps t} ( sum_c W[co][c][t] * shift[n][c]  +  sum_{c<C0} resid_enc * mean[c] )
//             + sum_{valid low-res taps j of the voxel's phase} sum_{c1} resid_phase[phase][j][c1] * mean[C0+c1]
// resid = (exact folded weight) - (the bf16 the tensor cores multiply with): first-order correction of the weight rounding,
// which the (large) channel means would otherwise amplify.  grid (Cout, n), block 128.
__global__ void upcat_fold_bias_kernel(const float* __restrict__ W, const float* __restrict__ ab, const float* __restrict__ conv_bias,
                                       const double* __restrict__ sums, double count, int C0, int C1, int Cout,
                                       float* __restrict__ biascls) {
  __shared__ float bt[27];
  __shared__ float rk[64];
  const int C = C0 + C1;
  const int co = blockIdx.x, n = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int q = warp; q < 27 + 64; q += nwarps) {  // one warp per reduction
    float acc = 0.f;
    if (ab) {
      if (q < 27) {
        for (int ci = lane; ci < C; ci += 32) {
          const float w = W[((size_t)co * C + ci) * 27 + q];
          const float a = ab[((size_t)n * C + ci) * 2], sh = ab[((size_t)n * C + ci) * 2 + 1];
          acc += w * sh;
          if (ci < C0 && sums) {
            const float wa = w * a;
            acc += (wa - from_act(to_act(wa))) * (float)(sums[((size_t)n * C + ci) * 2] / count);
          }
        }
      } else if (sums) {
        const int pj = q - 27, phase = pj >> 3, j = pj & 7;
        for (int c1 = lane; c1 < C1; c1 += 32) {
          const float* w = W + ((size_t)co * C + C0 + c1) * 27;
          float ws = 0.f;
          for (int td = 0; td < 3; ++td)
            if (phase_has_tap((phase >> 2) & 1, (j >> 2) & 1, td))
              for (int th = 0; th < 3; ++th)
                if (phase_has_tap((phase >> 1) & 1, (j >> 1) & 1, th))
                  for (int tw = 0; tw < 3; ++tw)
                    if (phase_has_tap(phase & 1, j & 1, tw)) ws += w[(td * 3 + th) * 3 + tw];
          const float wa = ws * ab[((size_t)n * C + C0 + c1) * 2];
          acc += (wa - from_act(to_act(wa))) * (float)(sums[((size_t)n * C + C0 + c1) * 2] / count);
        }
      }
    }
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) {
      if (q < 27) bt[q] = acc;
      else rk[q - 27] = acc;
    }
  }
  __syncthreads();
  const float cb = conv_bias ? conv_bias[co] : 0.f;
  for (int cls = threadIdx.x; cls < 64; cls += blockDim.x) {
    const int c3[3] = {cls >> 4, (cls >> 2) & 3, cls & 3};
    int bc[3], ph[3];
    for (int a = 0; a < 3; ++a) {
      bc[a] = c3[a] == 3 ? 1 : c3[a];    // border class for tap validity
      ph[a] = c3[a] >= 2 ? 1 : 0;        // parity: high face and "interior odd" are odd coordinates
    }
    float acc = cb;
    for (int td = 0; td < 3; ++td)
      if (tap_valid(bc[0], td))
        for (int th = 0; th < 3; ++th)
          if (tap_valid(bc[1], th))
            for (int tw = 0; tw < 3; ++tw)
              if (tap_valid(bc[2], tw)) acc += bt[(td * 3 + th) * 3 + tw];
    const int phase = (ph[0] << 2) | (ph[1] << 1) | ph[2];
    for (int j = 0; j < 8; ++j) {
      const int jj[3] = {(j >> 2) & 1, (j >> 1) & 1, j & 1};
      bool ok = true;
      for (int a = 0; a < 3; ++a) {
        // the low-res tap exists iff the full-res taps it carries are in bounds: (p=0,j=0) carries t=-1, (p=1,j=1) carries t=+1
        if (ph[a] == 0 && jj[a] == 0 && !tap_valid(bc[a], 0)) ok = false;
        if (ph[a] == 1 && jj[a] == 1 && !tap_valid(bc[a], 2)) ok = false;
      }
      if (ok) acc += rk[phase * 8 + j];
    }
    biascls[((size_t)n * 64 + cls) * Cout + co] = acc;
  }
}

// dgrad operands: wd_enc[tap'][c0][co] = W[co][c0][26 - tap'];  wd_up[e][c1][co] = sum_{t in E(e)} W[co][C0+c1][t]
__global__ void upcat_prep_dgrad_kernel(const float* __restrict__ W, int C0, int C1, int Cout, bf16* __restrict__ wd_enc,
                                        bf16* __restrict__ wd_up) {
  const int C = C0 + C1;
  const size_t n_enc = (size_t)27 * C0 * Cout, n_up = (size_t)64 * C1 * Cout;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_enc + n_up; i += (size_t)gridDim.x * blockDim.x) {
    if (i < n_enc) {
      int co = (int)(i % Cout);
      size_t r = i / Cout;
      int ci = (int)(r % C0);
      int tap = (int)(r / C0);
      wd_enc[i] = to_act(W[((size_t)co * C + ci) * 27 + (26 - tap)]);
    } else {
      size_t k = i - n_enc;
      int co = (int)(k % Cout);
      size_t r = k / Cout;
      int c1 = (int)(r % C1);
      int e = (int)(r / C1);
      const float* w = W + ((size_t)co * C + C0 + c1) * 27;
      float acc = 0.f;
      for (int td = 0; td < 3; ++td)
        if (offset_has_tap(e >> 4, td))
          for (int th = 0; th < 3; ++th)
            if (offset_has_tap((e >> 2) & 3, th))
              for (int tw = 0; tw < 3; ++tw)
                if (offset_has_tap(e & 3, tw)) acc += w[(td * 3 + th) * 3 + tw];
      wd_up[k] = to_act(acc);
    }
  }
}

// G[n][t][c][co] (one split) from the encoder-channel wgrad G_enc[n][S1][27][C0][co] and Q[n][S2][64][co][c1]:
//   c >= C0: sum over the 8 parities r of Q[e = r - t]  (per axis index e+1 = r - t_off + 1)
__global__ void upcat_assemble_wgrad_kernel(const float* __restrict__ Genc, int S1, const float* __restrict__ Q, int S2, int C0, int C1,
                                            int Cout, float* __restrict__ G) {
  const int C = C0 + C1;
  const int n = blockIdx.y;
  const size_t per_n = (size_t)27 * C * Cout;
  const size_t n_enc = (size_t)27 * C0 * Cout, n_up = (size_t)27 * Cout * C1;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_enc + n_up; i += (size_t)gridDim.x * blockDim.x) {
    if (i < n_enc) {  // (t, c, co), co fastest: coalesced reads of G_enc and writes of G
      int co = (int)(i % Cout);
      size_t r = i / Cout;
      int c = (int)(r % C0);
      int t = (int)(r / C0);
      float acc = 0.f;
      for (int s = 0; s < S1; ++s) acc += Genc[((((size_t)n * S1 + s) * 27 + t) * C0 + c) * Cout + co];
      G[(size_t)n * per_n + ((size_t)t * C + c) * Cout + co] = acc;
    } else {  // (t, co, c1), c1 fastest: coalesced reads of the 8 * S2 Q blocks (the single write per element is strided)
      size_t k = i - n_enc;
      int c1 = (int)(k % C1);
      size_t r = k / C1;
      int co = (int)(r % Cout);
      int t = (int)(r / Cout);
      const int td = t / 9 - 1, th = (t / 3) % 3 - 1, tw = t % 3 - 1;
      float acc = 0.f;
      for (int rr = 0; rr < 8; ++rr) {
        const int e = ((((rr >> 2) & 1) - td + 1) << 4) | ((((rr >> 1) & 1) - th + 1) << 2) | ((rr & 1) - tw + 1);
        for (int s = 0; s < S2; ++s) acc += Q[((((size_t)n * S2 + s) * 64 + e) * Cout + co) * C1 + c1];
      }
      G[(size_t)n * per_n + ((size_t)t * C + C0 + c1) * Cout + co] = acc;
    }
  }
}

}  // namespace b200

using namespace b200;
#define ST(s) ((cudaStream_t)(s))

extern "C" {

int b200_gn_fold_upcat(const double* sums, const float* gamma, const float* beta, int G, double count, const float* W,
                       const float* conv_bias, int N, int C0, int C1, int Cout, void* wf_enc, void* wp, float* biascls, float* mean_rstd,
                       float* ab, b200_stream_t s) {
  const int C = C0 + C1;
  int n_w = 1;
  const float* abp = nullptr;
  if (sums && gamma) {
    int rc = b200_gn_coeffs(sums, gamma, beta, G, count, N, C, mean_rstd, ab, s);
    if (rc) return rc;
    n_w = N;
    abp = ab;
  }
  {
    size_t total = (size_t)n_w * 27 * Cout * C0;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 4096) blocks = 4096;
    upcat_fold_enc_kernel<<<blocks, 256, 0, ST(s)>>>(W, abp, n_w, C0, C, Cout, (bf16*)wf_enc);
    B200_CHECK_LAUNCH("upcat_fold_enc");
  }
  {
    size_t total = (size_t)n_w * 64 * Cout * C1;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 4096) blocks = 4096;
    upcat_fold_phase_kernel<<<blocks, 256, 0, ST(s)>>>(W, abp, n_w, C0, C1, Cout, (bf16*)wp);
    B200_CHECK_LAUNCH("upcat_fold_phase");
  }
  if (biascls && (abp || conv_bias)) {
    dim3 grid(Cout, n_w);
    upcat_fold_bias_kernel<<<grid, 1024, 0, ST(s)>>>(W, abp, conv_bias, abp ? sums : nullptr, count, C0, C1, Cout, biascls);
    B200_CHECK_LAUNCH("upcat_fold_bias");
  }
  return 0;
}

int b200_upcat_prep_dgrad_weights(const float* W, int C0, int C1, int Cout, void* wd_enc, void* wd_up, b200_stream_t s) {
  size_t total = (size_t)27 * C0 * Cout + (size_t)64 * C1 * Cout;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 4096) blocks = 4096;
  upcat_prep_dgrad_kernel<<<blocks, 256, 0, ST(s)>>>(W, C0, C1, Cout, (bf16*)wd_enc, (bf16*)wd_up);
  B200_CHECK_LAUNCH("upcat_prep_dgrad");
  return 0;
}

int b200_upcat_assemble_wgrad(const float* G_enc, int S1, const float* Q, int S2, int N, int C0, int C1, int Cout, float* G,
                              b200_stream_t s) {
  size_t total = (size_t)27 * (C0 + C1) * Cout;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 1024) blocks = 1024;
  dim3 grid(blocks, N);
  upcat_assemble_wgrad_kernel<<<grid, 256, 0, ST(s)>>>(G_enc, S1, Q, S2, C0, C1, Cout, G);
  B200_CHECK_LAUNCH("upcat_assemble_wgrad");
  return 0;
}

}  // extern "C"
