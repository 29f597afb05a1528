// Fused BCEDiceLoss on the network's logits (reference losses.py:187-201, :130-145, :11-37):
//
//   loss = mean(bce_with_logits(x, t)) + alpha * (1 - mean_c 2 * I_c / max(D_c, eps)),   p = sigmoid(x),
//   I_c = sum_{n,v} p*t,  D_c = sum_{n,v} p*p + t*t                                          (x, t: fp32 [N][C][V])
//
// Two passes over the logits instead of the ~20 ATen kernels of the eager formulation: pass 1 reduces the four sums (fixed-order
// partials, fp64 finalise -> deterministic) and a tiny kernel turns them into the loss and the per-channel gradient coefficients;
// pass 2 writes d loss / d x = A (p - t) + (k1_c t + k2_c p) p (1 - p)  with  A = 1/(N C V), k1_c = -2 alpha / (C D_c),
// k2_c = 4 alpha I_c / (C D_c^2)  (k1 = -2 alpha / (C eps), k2 = 0 where D_c is clamped).
#include "common.cuh"

namespace b200 {

constexpr int LOSS_THREADS = 256;

__device__ __forceinline__ void loss_terms(float x, float t, float acc[4]) {
  const float e = __expf(-fabsf(x));
  const float p = x >= 0.f ? 1.f / (1.f + e) : e / (1.f + e);
  acc[0] += fmaxf(x, 0.f) - x * t + log1pf(e);  // binary_cross_entropy_with_logits, the numerically stable form
  acc[1] += p * t;
  acc[2] += p * p;
  acc[3] += t * t;
}

// grid (P, N*C), block 256: block p of row r = (n,c) reduces its slice of the V voxels -> partials[r][p][4]
__global__ void bce_dice_partials_kernel(const float* __restrict__ x, const float* __restrict__ t, long long V, int P,
                                         float* __restrict__ partials) {
  __shared__ float red[LOSS_THREADS][4];
  const int p = blockIdx.x, r = blockIdx.y;
  const long long per = (V + P - 1) / P;
  long long v0 = (long long)p * per, v1 = v0 + per;
  if (v1 > V) v1 = V;
  const float* xr = x + (size_t)r * V;
  const float* tr = t + (size_t)r * V;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (long long v = v0 + threadIdx.x; v < v1; v += LOSS_THREADS) loss_terms(xr[v], tr[v], acc);
#pragma unroll
  for (int i = 0; i < 4; ++i) red[threadIdx.x][i] = acc[i];
  __syncthreads();
  for (int o = LOSS_THREADS / 2; o; o >>= 1) {
    if ((int)threadIdx.x < o) {
#pragma unroll
      for (int i = 0; i < 4; ++i) red[threadIdx.x][i] += red[threadIdx.x + o][i];
    }
    __syncthreads();
  }
  if (threadIdx.x < 4) partials[((size_t)r * P + p) * 4 + threadIdx.x] = red[0][threadIdx.x];
}

// one block: loss[0] and coef = [A, (k1_c, k2_c) x C]
__global__ void bce_dice_finalize_kernel(const float* __restrict__ partials, int N, int C, long long V, int P, float alpha, float eps,
                                         float* __restrict__ loss, float* __restrict__ coef) {
  __shared__ double sh_bce[LOSS_THREADS];
  __shared__ double sh_dice;
  double bce = 0.0;
  for (int i = threadIdx.x; i < N * C * P; i += LOSS_THREADS) bce += (double)partials[(size_t)i * 4];
  sh_bce[threadIdx.x] = bce;
  if (threadIdx.x == 0) sh_dice = 0.0;
  __syncthreads();
  for (int o = LOSS_THREADS / 2; o; o >>= 1) {
    if ((int)threadIdx.x < o) sh_bce[threadIdx.x] += sh_bce[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {  // C is the number of output channels (a handful): serial, fixed order
    double dice_sum = 0.0;
    for (int c = 0; c < C; ++c) {
      double I = 0.0, D = 0.0;
      for (int n = 0; n < N; ++n)
        for (int p = 0; p < P; ++p) {
          const float* q = partials + (((size_t)n * C + c) * P + p) * 4;
          I += (double)q[1];
          D += (double)q[2] + (double)q[3];
        }
      const bool clamped = D < (double)eps;
      const double Dc = clamped ? (double)eps : D;
      dice_sum += 2.0 * I / Dc;
      coef[1 + 2 * c] = (float)(-2.0 * alpha / (C * Dc));
      coef[2 + 2 * c] = clamped ? 0.f : (float)(4.0 * alpha * I / (C * Dc * Dc));
    }
    const double count = (double)N * C * (double)V;
    coef[0] = (float)(1.0 / count);
    loss[0] = (float)(sh_bce[0] / count + alpha * (1.0 - dice_sum / C));
  }
}

__global__ void bce_dice_grad_kernel(const float* __restrict__ x, const float* __restrict__ t, const float* __restrict__ coef, int C,
                                     long long V, float* __restrict__ dx) {
  const int r = blockIdx.y;
  const int c = r % C;
  const float A = coef[0], k1 = coef[1 + 2 * c], k2 = coef[2 + 2 * c];
  const float* xr = x + (size_t)r * V;
  const float* tr = t + (size_t)r * V;
  float* dr = dx + (size_t)r * V;
  for (long long v = (long long)blockIdx.x * LOSS_THREADS + threadIdx.x; v < V; v += (long long)gridDim.x * LOSS_THREADS) {
    const float xv = xr[v], tv = tr[v];
    const float e = __expf(-fabsf(xv));
    const float p = xv >= 0.f ? 1.f / (1.f + e) : e / (1.f + e);
    dr[v] = A * (p - tv) + (k1 * tv + k2 * p) * p * (1.f - p);
  }
}

}  // namespace b200

using namespace b200;

extern "C" {

int b200_bce_dice_partials_count(int N, int C, long long V) {
  (void)N;
  (void)C;
  long long p = (V + (long long)LOSS_THREADS * 16 - 1) / ((long long)LOSS_THREADS * 16);
  return (int)(p > 256 ? 256 : (p < 1 ? 1 : p));
}

int b200_bce_dice_fwd(const float* logits, const float* target, int N, int C, long long V, float alpha, float eps, float* partials,
                      float* loss, float* coef, b200_stream_t s) {
  B200_CHECK_ARG(N >= 1 && C >= 1 && V >= 1 && (long long)N * C <= 65535, "bce_dice_fwd: bad shape N=%d C=%d V=%lld", N, C, V);
  const int P = b200_bce_dice_partials_count(N, C, V);
  dim3 grid(P, N * C);
  bce_dice_partials_kernel<<<grid, LOSS_THREADS, 0, (cudaStream_t)s>>>(logits, target, V, P, partials);
  B200_CHECK_LAUNCH("bce_dice_partials");
  bce_dice_finalize_kernel<<<1, LOSS_THREADS, 0, (cudaStream_t)s>>>(partials, N, C, V, P, alpha, eps, loss, coef);
  B200_CHECK_LAUNCH("bce_dice_finalize");
  return 0;
}

int b200_bce_dice_bwd(const float* logits, const float* target, const float* coef, int N, int C, long long V, float* dlogits,
                      b200_stream_t s) {
  B200_CHECK_ARG(N >= 1 && C >= 1 && V >= 1 && (long long)N * C <= 65535, "bce_dice_bwd: bad shape N=%d C=%d V=%lld", N, C, V);
  long long blocks = (V + (long long)LOSS_THREADS * 8 - 1) / ((long long)LOSS_THREADS * 8);
  if (blocks > 2048) blocks = 2048;
  dim3 grid((unsigned)blocks, N * C);
  bce_dice_grad_kernel<<<grid, LOSS_THREADS, 0, (cudaStream_t)s>>>(logits, target, coef, C, V, dlogits);
  B200_CHECK_LAUNCH("bce_dice_grad");
  return 0;
}

}  // extern "C"
