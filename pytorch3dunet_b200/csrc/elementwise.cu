// HBM-bound kernels of the 3D U-Net path: layout, GroupNorm statistics / folding / backward, MaxPool3d(2),
// nearest-upsample + concat, final 1x1x1 conv + sigmoid/softmax, and the small deterministic reductions.
// Every tensor pass here is a straight 16-byte-vector stream over NDHWC bf16; the roofline for all of them
// is HBM bandwidth (bytes listed per kernel in DESIGN.md).
#include <stdarg.h>

#include <string.h>

#include <stdlib.h>
#include "common.cuh"
#include "ew.cuh"

namespace b200 {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int sm_count() {
  static int cached[64] = {0};  // per device ordinal; benign race (every thread writes the same value)
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (!cached[dev]) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

// ------------------------------------------------------------------------------------------------
// layout
// ------------------------------------------------------------------------------------------------
template <typename OutT>
__global__ void ncdhw_to_ndhwc_kernel(const float* __restrict__ src, OutT* __restrict__ dst, int C, long long voxels) {
  // grid: (blocks over voxels, N); each thread handles one voxel, loops channels (C is small for network inputs;
  // for block-level entry points C is a feature-map count and this is test plumbing, not the hot path)
  int n = blockIdx.y;
  long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= voxels) return;
  const float* s = src + (size_t)n * C * voxels + v;
  OutT* d = dst + ((size_t)n * voxels + v) * C;
  for (int c = 0; c < C; ++c) {
    float f = s[(size_t)c * voxels];
    if constexpr (sizeof(OutT) == 2) d[c] = to_act(f);
    else d[c] = f;
  }
}
__global__ void ndhwc_to_ncdhw_kernel(const bf16* __restrict__ src, float* __restrict__ dst, int C, long long voxels) {
  int n = blockIdx.y;
  long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= voxels) return;
  const bf16* s = src + ((size_t)n * voxels + v) * C;
  float* d = dst + (size_t)n * C * voxels + v;
  for (int c = 0; c < C; ++c) d[(size_t)c * voxels] = from_act(s[c]);
}

// ------------------------------------------------------------------------------------------------
// statistics
// ------------------------------------------------------------------------------------------------
// NCDHW fp32 input planes: grid (P, N*C); partials [N][P][C][2]
__global__ void stats_ncdhw_f32_kernel(const float* __restrict__ x, int C, long long voxels, int P, float* __restrict__ partials) {
  int p = blockIdx.x, nc = blockIdx.y;
  int n = nc / C, c = nc % C;
  long long v0, v1;
  ew_range(voxels, p, P, v0, v1);
  const float* s = x + (size_t)nc * voxels;
  float a = 0.f, q = 0.f;
  for (long long v = v0 + threadIdx.x; v < v1; v += blockDim.x) {
    float f = s[v];
    a += f;
    q += f * f;
  }
  __shared__ float ra[32], rq[32];
  for (int o = 16; o; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    q += __shfl_xor_sync(0xffffffffu, q, o);
  }
  int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) {
    ra[w] = a;
    rq[w] = q;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float ta = 0.f, tq = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) {
      ta += ra[i];
      tq += rq[i];
    }
    float* o = partials + (((size_t)n * P + p) * C + c) * 2;
    o[0] = ta;
    o[1] = tq;
  }
}

__global__ void stats_ndhwc_bf16_kernel(const bf16* __restrict__ x, int C, long long voxels, int P, float* __restrict__ partials) {
  extern __shared__ float red[];
  int p = blockIdx.x, n = blockIdx.y;
  EwMap m = ew_map(C);
  long long v0, v1;
  ew_range(voxels, p, P, v0, v1);
  float s[8] = {0}, q[8] = {0};
  if (m.active) {
    const bf16x8* xp = reinterpret_cast<const bf16x8*>(x + (size_t)n * voxels * C);
    for (long long v = v0 + m.vl; v < v1; v += m.VL) {
      float f[8];
      unpack8(xp[v * m.CG + m.cg], f);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        s[i] += f[i];
        q[i] += f[i] * f[i];
      }
    }
  }
  ew_write_partials(s, q, m, partials + ((size_t)n * P + p) * C * 2, red);
}


// partials of (sum v, sum v*w) over two bf16 NDHWC tensors (GroupNorm-after-conv backward); grid (P, N)
__global__ void stats2_ndhwc_bf16_kernel(const bf16* __restrict__ a, const bf16* __restrict__ b, int C, long long voxels, int P,
                                         float* __restrict__ partials) {
  extern __shared__ float red[];
  int p = blockIdx.x, n = blockIdx.y;
  EwMap m = ew_map(C);
  long long v0, v1;
  ew_range(voxels, p, P, v0, v1);
  float s[8] = {0}, q[8] = {0};
  if (m.active) {
    const bf16x8* ap = reinterpret_cast<const bf16x8*>(a + (size_t)n * voxels * C);
    const bf16x8* bp = reinterpret_cast<const bf16x8*>(b + (size_t)n * voxels * C);
    for (long long v = v0 + m.vl; v < v1; v += m.VL) {
      float f[8], g[8];
      unpack8(ap[v * m.CG + m.cg], f);
      unpack8(bp[v * m.CG + m.cg], g);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        s[i] += f[i];
        q[i] += f[i] * g[i];
      }
    }
  }
  ew_write_partials(s, q, m, partials + ((size_t)n * P + p) * C * 2, red);
}

// GroupNorm-backward sums of the conv INPUT from the wgrad by-products (no pass over dxhat needed):
//   sum_v dxhat[n,v,ci]          = sum_{tap,co} W[co][ci][tap] * T[n][tap][co]
//   sum_v dxhat[n,v,ci]*x[n,v,ci] = sum_{tap,co} W[co][ci][tap] * sum_s G[n][s][tap][ci][co]
// grid (Cin, N), block 128
__global__ void gn_bwd_sums_from_wgrad_kernel(const float* __restrict__ G, int S, const float* __restrict__ T, const float* __restrict__ W,
                                              int Cin, int Cout, double* __restrict__ sums2) {
  __shared__ double r1[128], r2[128];
  int ci = blockIdx.x, n = blockIdx.y;
  double a1 = 0.0, a2 = 0.0;
  for (int idx = threadIdx.x; idx < 27 * Cout; idx += blockDim.x) {
    int tap = idx / Cout, co = idx % Cout;
    double w = (double)W[((size_t)co * Cin + ci) * 27 + tap];
    a1 += w * (double)T[((size_t)n * 27 + tap) * Cout + co];
    double g = 0.0;
    for (int s = 0; s < S; ++s) g += (double)G[((((size_t)n * S + s) * 27 + tap) * Cin + ci) * Cout + co];
    a2 += w * g;
  }
  r1[threadIdx.x] = a1;
  r2[threadIdx.x] = a2;
  __syncthreads();
  for (int o = 64; o; o >>= 1) {
    if ((int)threadIdx.x < o) {
      r1[threadIdx.x] += r1[threadIdx.x + o];
      r2[threadIdx.x] += r2[threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    sums2[((size_t)n * Cin + ci) * 2] = r1[0];
    sums2[((size_t)n * Cin + ci) * 2 + 1] = r2[0];
  }
}

// sums[N][C][2] (double) = sum_p partials[n][p][c][k]; grid (ceil(C*2/32), N), block (32, 32)
__global__ void partials_finalize_kernel(const float* __restrict__ partials, int P, int C, double* __restrict__ sums) {
  __shared__ double red[32][33];
  int n = blockIdx.y;
  int col = blockIdx.x * 32 + threadIdx.x;  // index into C*2
  double acc = 0.0;
  if (col < C * 2) {
    const float* base = partials + (size_t)n * P * C * 2 + col;
    for (int p = threadIdx.y; p < P; p += 32) acc += (double)base[(size_t)p * C * 2];
  }
  red[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && col < C * 2) {
    double t = 0.0;
    for (int i = 0; i < 32; ++i) t += red[i][threadIdx.x];
    sums[(size_t)n * C * 2 + col] = t;
  }
}

// out[K] = sum_p partials[p][K]; grid ceil(K/32), block (32,32)
__global__ void reduce_rows_kernel(const float* __restrict__ partials, int P, int K, float* __restrict__ out) {
  __shared__ double red[32][33];
  int col = blockIdx.x * 32 + threadIdx.x;
  double acc = 0.0;
  if (col < K)
    for (int p = threadIdx.y; p < P; p += 32) acc += (double)partials[(size_t)p * K + col];
  red[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && col < K) {
    double t = 0.0;
    for (int i = 0; i < 32; ++i) t += red[i][threadIdx.x];
    out[col] = (float)t;
  }
}

// ------------------------------------------------------------------------------------------------
// GroupNorm coefficients and folding
// ------------------------------------------------------------------------------------------------
// grid N, block 128.  ab[N][C][2] = (a,b), mean_rstd[N][G][2]
__global__ void gn_coeffs_kernel(const double* __restrict__ sums, const float* __restrict__ gamma, const float* __restrict__ beta,
                                 int G, double count, int C, float* __restrict__ mean_rstd, float* __restrict__ ab) {
  extern __shared__ float sm[];  // [G][2]
  int n = blockIdx.x;
  int cpg = C / G;
  for (int g = threadIdx.x; g < G; g += blockDim.x) {
    double s = 0.0, q = 0.0;
    for (int c = g * cpg; c < (g + 1) * cpg; ++c) {
      s += sums[((size_t)n * C + c) * 2];
      q += sums[((size_t)n * C + c) * 2 + 1];
    }
    double m = count * cpg;
    double mean = s / m;
    double var = q / m - mean * mean;  // biased variance, as native_group_norm
    if (var < 0.0) var = 0.0;
    float rstd = (float)(1.0 / sqrt(var + 1e-5));
    sm[g * 2] = (float)mean;
    sm[g * 2 + 1] = rstd;
    if (mean_rstd) {
      mean_rstd[((size_t)n * G + g) * 2] = (float)mean;
      mean_rstd[((size_t)n * G + g) * 2 + 1] = rstd;
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    int g = c / cpg;
    float a = gamma[c] * sm[g * 2 + 1];
    float b = beta[c] - sm[g * 2] * a;
    ab[((size_t)n * C + c) * 2] = a;
    ab[((size_t)n * C + c) * 2 + 1] = b;
  }
}

// partials_finalize + gn_coeffs in one launch: grid (N, G) -- one block per (sample, GroupNorm group) -- block (32,32).  The column sums
// are formed exactly as partials_finalize_kernel forms them (same order, fp64), so `sums` is bit-identical to the two-kernel route.
__global__ void gn_stats_coeffs_kernel(const float* __restrict__ partials, int P, int C, const float* __restrict__ gamma,
                                       const float* __restrict__ beta, int G, double count, double* __restrict__ sums,
                                       float* __restrict__ mean_rstd, float* __restrict__ ab) {
  extern __shared__ double dsm[];      // [cpg*2] column sums of this group
  __shared__ double red[32][33];
  __shared__ float gst[2];
  const int n = blockIdx.x, g = blockIdx.y;
  const int tid = threadIdx.y * 32 + threadIdx.x;
  const int cpg = C / G;
  const int col0 = g * cpg * 2, ncol = cpg * 2;
  for (int c0 = 0; c0 < ncol; c0 += 32) {
    const int lc = c0 + threadIdx.x;
    double acc = 0.0;
    if (lc < ncol) {
      const float* base = partials + (size_t)n * P * C * 2 + col0 + lc;
      for (int p = threadIdx.y; p < P; p += 32) acc += (double)base[(size_t)p * C * 2];
    }
    red[threadIdx.y][threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.y == 0 && lc < ncol) {
      double t = 0.0;
      for (int i = 0; i < 32; ++i) t += red[i][threadIdx.x];
      dsm[lc] = t;
      sums[(size_t)n * C * 2 + col0 + lc] = t;
    }
    __syncthreads();
  }
  if (tid == 0) {
    double s = 0.0, q = 0.0;
    for (int c = 0; c < cpg; ++c) {
      s += dsm[c * 2];
      q += dsm[c * 2 + 1];
    }
    const double m = count * cpg;
    const double mean = s / m;
    double var = q / m - mean * mean;  // biased variance, as native_group_norm
    if (var < 0.0) var = 0.0;
    const float rstd = (float)(1.0 / sqrt(var + 1e-5));
    gst[0] = (float)mean;
    gst[1] = rstd;
    if (mean_rstd) {
      mean_rstd[((size_t)n * G + g) * 2] = (float)mean;
      mean_rstd[((size_t)n * G + g) * 2 + 1] = rstd;
    }
  }
  __syncthreads();
  for (int lc = tid; lc < cpg; lc += 1024) {
    const int c = g * cpg + lc;
    const float a = gamma[c] * gst[1];
    const float b = beta[c] - gst[0] * a;
    ab[((size_t)n * C + c) * 2] = a;
    ab[((size_t)n * C + c) * 2 + 1] = b;
  }
}

// wf[n][tap][co][ci] = bf16(W[co][ci][tap] * a[n][ci]); one thread per output element
__global__ void fold_weights_kernel(const float* __restrict__ W, const float* __restrict__ ab, int n_w, int Cin, int Cout,
                                    bf16* __restrict__ wf) {
  size_t total = (size_t)n_w * 27 * Cout * Cin;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    int ci = (int)(i % Cin);
    size_t r = i / Cin;
    int co = (int)(r % Cout);
    r /= Cout;
    int tap = (int)(r % 27);
    int n = (int)(r / 27);
    float a = ab ? ab[((size_t)n * Cin + ci) * 2] : 1.f;
    wf[i] = to_act(W[((size_t)co * Cin + ci) * 27 + tap] * a);
  }
}

// biascls[n][cls][co] : grid (Cout, n_b), block 128.
// Besides the GroupNorm shift (W*b) the per-tap term carries the bf16 rounding residual of the folded weight times the
// per-channel mean of x: sum_k (W*a - bf16(W*a)) * mean_x.  Without it the rounding error of W*a multiplies the MEAN of
// the (un-normalised) activation and is amplified by mean/std relative to normalising first; with it only the centred
// part of x sees the rounding error.
__global__ void fold_bias_kernel(const float* __restrict__ W, const float* __restrict__ ab, const float* __restrict__ conv_bias,
                                 const double* __restrict__ sums, double count, int Cin, int Cout, float* __restrict__ biascls) {
  __shared__ float bt[27];
  int co = blockIdx.x, n = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int tap = warp; tap < 27; tap += nwarps) {  // one warp per tap, shuffle reduction over the input channels
    float acc = 0.f;
    if (ab)
      for (int ci = lane; ci < Cin; ci += 32) {
        float w = W[((size_t)co * Cin + ci) * 27 + tap];
        float wa = w * ab[((size_t)n * Cin + ci) * 2];
        float resid = wa - from_act(to_act(wa));
        float mean = sums ? (float)(sums[((size_t)n * Cin + ci) * 2] / count) : 0.f;
        acc += w * ab[((size_t)n * Cin + ci) * 2 + 1] + resid * mean;
      }
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) bt[tap] = acc;
  }
  __syncthreads();
  float cb = conv_bias ? conv_bias[co] : 0.f;
  for (int cls = threadIdx.x; cls < 64; cls += blockDim.x) {
    int cd = cls >> 4, ch = (cls >> 2) & 3, cw = cls & 3;
    float acc = cb;
    for (int td = 0; td < 3; ++td)
      if (tap_valid(cd, td))
        for (int th = 0; th < 3; ++th)
          if (tap_valid(ch, th))
            for (int tw = 0; tw < 3; ++tw)
              if (tap_valid(cw, tw)) acc += bt[(td * 3 + th) * 3 + tw];
    biascls[((size_t)n * 64 + cls) * Cout + co] = acc;
  }
}

// fold_weights + fold_bias in one launch: blocks [0, nbw) fold the weights, block nbw + (n * Cout + co) builds bias row (n, co)
__device__ __forceinline__ void fold_bias_block(const float* __restrict__ W, const float* __restrict__ ab, const float* __restrict__ conv_bias,
                                                const double* __restrict__ sums, double count, int Cin, int Cout, float* __restrict__ biascls,
                                                int co, int n, float* bt) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int tap = warp; tap < 27; tap += nwarps) {
    float acc = 0.f;
    if (ab)
      for (int ci = lane; ci < Cin; ci += 32) {
        float w = W[((size_t)co * Cin + ci) * 27 + tap];
        float wa = w * ab[((size_t)n * Cin + ci) * 2];
        float resid = wa - from_act(to_act(wa));
        float mean = sums ? (float)(sums[((size_t)n * Cin + ci) * 2] / count) : 0.f;
        acc += w * ab[((size_t)n * Cin + ci) * 2 + 1] + resid * mean;
      }
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) bt[tap] = acc;
  }
  __syncthreads();
  float cb = conv_bias ? conv_bias[co] : 0.f;
  for (int cls = threadIdx.x; cls < 64; cls += blockDim.x) {
    int cd = cls >> 4, ch = (cls >> 2) & 3, cw = cls & 3;
    float acc = cb;
    for (int td = 0; td < 3; ++td)
      if (tap_valid(cd, td))
        for (int th = 0; th < 3; ++th)
          if (tap_valid(ch, th))
            for (int tw = 0; tw < 3; ++tw)
              if (tap_valid(cw, tw)) acc += bt[(td * 3 + th) * 3 + tw];
    biascls[((size_t)n * 64 + cls) * Cout + co] = acc;
  }
}
__global__ void fold_all_kernel(const float* __restrict__ W, const float* __restrict__ ab, const float* __restrict__ conv_bias,
                                const double* __restrict__ sums, double count, int n_w, int Cin, int Cout, int nbw, bf16* __restrict__ wf,
                                float* __restrict__ biascls) {
  __shared__ float bt[27];
  if ((int)blockIdx.x >= nbw) {
    const int r = (int)blockIdx.x - nbw;
    fold_bias_block(W, ab, conv_bias, sums, count, Cin, Cout, biascls, r % Cout, r / Cout, bt);
    return;
  }
  const size_t total = (size_t)n_w * 27 * Cout * Cin;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)nbw * blockDim.x) {
    int ci = (int)(i % Cin);
    size_t r = i / Cin;
    int co = (int)(r % Cout);
    r /= Cout;
    int tap = (int)(r % 27);
    int n = (int)(r / 27);
    float a = ab ? ab[((size_t)n * Cin + ci) * 2] : 1.f;
    wf[i] = to_act(W[((size_t)co * Cin + ci) * 27 + tap] * a);
  }
}

// wd[tap'][ci][co] = bf16(W[co][ci][26 - tap'])
__global__ void prep_dgrad_weights_kernel(const float* __restrict__ W, int Cin, int Cout, bf16* __restrict__ wd) {
  size_t total = (size_t)27 * Cin * Cout;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    int co = (int)(i % Cout);
    size_t r = i / Cout;
    int ci = (int)(r % Cin);
    int tap = (int)(r / Cin);
    wd[i] = to_act(W[((size_t)co * Cin + ci) * 27 + (26 - tap)]);
  }
}

// y = act(a*x + b) with partials of y; grid (P, N)
__global__ void gn_apply_act_kernel(const bf16* __restrict__ x, const float* __restrict__ ab, const bf16* __restrict__ residual, int C,
                                    long long voxels, int P, int act, float slope, bf16* __restrict__ y,
                                    float* __restrict__ partials) {
  extern __shared__ float red[];
  int p = blockIdx.x, n = blockIdx.y;
  EwMap m = ew_map(C);
  long long v0, v1;
  ew_range(voxels, p, P, v0, v1);
  float s[8] = {0}, q[8] = {0};
  if (m.active) {
    float a[8], b[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      a[i] = ab[((size_t)n * C + m.cg * 8 + i) * 2];
      b[i] = ab[((size_t)n * C + m.cg * 8 + i) * 2 + 1];
    }
    const bf16x8* xp = reinterpret_cast<const bf16x8*>(x + (size_t)n * voxels * C);
    bf16x8* yp = reinterpret_cast<bf16x8*>(y + (size_t)n * voxels * C);
    const bf16x8* rp = residual ? reinterpret_cast<const bf16x8*>(residual + (size_t)n * voxels * C) : nullptr;
    for (long long v = v0 + m.vl; v < v1; v += m.VL) {
      float f[8], r[8] = {0};
      unpack8(xp[v * m.CG + m.cg], f);
      if (rp) unpack8(rp[v * m.CG + m.cg], r);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        f[i] = bf16_round(act_fwd(a[i] * f[i] + b[i] + r[i], act, slope));
        s[i] += f[i];
        q[i] += f[i] * f[i];
      }
      yp[v * m.CG + m.cg] = pack8(f);
    }
  }
  if (partials) ew_write_partials(s, q, m, partials + ((size_t)n * P + p) * C * 2, red);
}

// ------------------------------------------------------------------------------------------------
// GroupNorm backward
// ------------------------------------------------------------------------------------------------
// grid 1, block 256.  coef[N][C][3], dgamma[C], dbeta[C]
__global__ void gn_bwd_coeffs_kernel(const double* __restrict__ sums2, const float* __restrict__ gamma,
                                     const float* __restrict__ mean_rstd, int G, double count, int N, int C,
                                     float* __restrict__ coef, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  int cpg = C / G;
  // per-channel parameter grads (deterministic serial sum over n)
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    int g = c / cpg;
    double dg = 0.0, db = 0.0;
    for (int n = 0; n < N; ++n) {
      double mean = mean_rstd[((size_t)n * G + g) * 2], rstd = mean_rstd[((size_t)n * G + g) * 2 + 1];
      double s1 = sums2[((size_t)n * C + c) * 2], s2 = sums2[((size_t)n * C + c) * 2 + 1];
      dg += (s2 - mean * s1) * rstd;
      db += s1;
    }
    dgamma[c] = (float)dg;
    dbeta[c] = (float)db;
  }
  // per-(n,group) coefficients expanded per channel
  for (int idx = threadIdx.x; idx < N * G; idx += blockDim.x) {
    int n = idx / G, g = idx % G;
    double mean = mean_rstd[((size_t)n * G + g) * 2], rstd = mean_rstd[((size_t)n * G + g) * 2 + 1];
    double S1 = 0.0, S2x = 0.0;  // sum gamma*dxhat, sum gamma*dxhat*x over the group
    for (int c = g * cpg; c < (g + 1) * cpg; ++c) {
      S1 += (double)gamma[c] * sums2[((size_t)n * C + c) * 2];
      S2x += (double)gamma[c] * sums2[((size_t)n * C + c) * 2 + 1];
    }
    double m = count * cpg;
    double S2 = rstd * (S2x - mean * S1);  // sum gamma*dxhat*xtilde
    double B = -rstd * rstd * S2 / m;
    double Cc = -rstd * S1 / m + rstd * rstd * mean * S2 / m;
    for (int c = g * cpg; c < (g + 1) * cpg; ++c) {
      float* o = coef + ((size_t)n * C + c) * 3;
      o[0] = (float)(rstd * gamma[c]);
      o[1] = (float)B;
      o[2] = (float)Cc;
    }
  }
}

// out = (A*dxhat + B*x + Cc) * act'(x) [+ gadd]; grid (P, N).  gadd: gradient already in dz form (same shape), may alias out
template <bool STATS>
__global__ void gn_bwd_apply_kernel(const bf16* __restrict__ dxhat, const bf16* __restrict__ x, const float* __restrict__ coef,
                                    int C, long long voxels, int P, int act, float slope, const bf16* gadd, bf16* out,
                                    float* __restrict__ partials) {
  extern __shared__ float red[];
  int p = blockIdx.x, n = blockIdx.y;
  EwMap m = ew_map(C);
  long long v0, v1;
  ew_range(voxels, p, P, v0, v1);
  float s[8] = {0}, q[8] = {0};
  if (m.active) {
  float A[8], B[8], Cc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float* cf = coef + ((size_t)n * C + m.cg * 8 + i) * 3;
    A[i] = cf[0];
    B[i] = cf[1];
    Cc[i] = cf[2];
  }
  const bf16x8* dp = reinterpret_cast<const bf16x8*>(dxhat + (size_t)n * voxels * C);
  const bf16x8* xp = reinterpret_cast<const bf16x8*>(x + (size_t)n * voxels * C);
  bf16x8* op = reinterpret_cast<bf16x8*>(out + (size_t)n * voxels * C);
  const bf16x8* gq = gadd ? reinterpret_cast<const bf16x8*>(gadd + (size_t)n * voxels * C) : nullptr;
  // two voxels per iteration: 4..6 independent 16-byte loads in flight per thread
  for (long long v = v0 + m.vl; v < v1; v += 2 * m.VL) {
    const long long vb = v + m.VL;
    const bool two = vb < v1;
    bf16x8 rd[2], rx[2], rg[2];
    rd[0] = dp[v * m.CG + m.cg];
    rx[0] = xp[v * m.CG + m.cg];
    if (gq) rg[0] = gq[v * m.CG + m.cg];
    if (two) {
      rd[1] = dp[vb * m.CG + m.cg];
      rx[1] = xp[vb * m.CG + m.cg];
      if (gq) rg[1] = gq[vb * m.CG + m.cg];
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (u == 1 && !two) break;
      float d[8], f[8], ga[8];
      unpack8(rd[u], d);
      unpack8(rx[u], f);
      if (gq) unpack8(rg[u], ga);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float g = (A[i] * d[i] + B[i] * f[i] + Cc[i]) * act_grad_from_out(f[i], act, slope);
        if (gq) g += ga[i];
        d[i] = STATS ? bf16_round(g) : g;
        if (STATS) {
          s[i] += d[i];
          q[i] += d[i] * d[i];
        }
      }
      op[(u ? vb : v) * m.CG + m.cg] = pack8(d);
    }
  }
  }
  if (STATS) ew_write_partials(s, q, m, partials + ((size_t)n * P + p) * C * 2, red);
}

// out = g[..., g_co : g_co + C] * act'(y) [+ gadd]; g has channel stride g_cs; gadd (dz form, contiguous) may alias out
// STATS: also the per-channel (sum, sum of squares) of the 16-bit result -> partials [N][P][C][2] (as gn_bwd_apply_kernel<true>)
template <bool STATS>
__global__ void act_bwd_kernel(const bf16* g, int g_cs, int g_co, const bf16* __restrict__ y, int C, long long voxels, int P, int act,
                               float slope, const bf16* gadd, bf16* out, float* __restrict__ partials) {
  extern __shared__ float red[];
  int p = blockIdx.x, n = blockIdx.y;
  EwMap m = ew_map(C);
  long long v0, v1;
  ew_range(voxels, p, P, v0, v1);
  if (!STATS && !m.active) return;
  float s[8] = {0}, q[8] = {0};
  if (m.active) {
    const bf16x8* yp = reinterpret_cast<const bf16x8*>(y + (size_t)n * voxels * C);
    bf16x8* op = reinterpret_cast<bf16x8*>(out + (size_t)n * voxels * C);
    for (long long v = v0 + m.vl; v < v1; v += m.VL) {
      float d[8], f[8], ga[8];
      unpack8(*reinterpret_cast<const bf16x8*>(g + ((size_t)n * voxels + v) * g_cs + g_co + m.cg * 8), d);
      unpack8(yp[v * m.CG + m.cg], f);
      if (gadd) unpack8(*reinterpret_cast<const bf16x8*>(gadd + ((size_t)n * voxels + v) * C + m.cg * 8), ga);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float t = d[i] * act_grad_from_out(f[i], act, slope);
        if (gadd) t += ga[i];
        d[i] = STATS ? bf16_round(t) : t;
        if (STATS) {
          s[i] += d[i];
          q[i] += d[i] * d[i];
        }
      }
      op[v * m.CG + m.cg] = pack8(d);
    }
  }
  if (STATS) ew_write_partials(s, q, m, partials + ((size_t)n * P + p) * C * 2, red);
}

// ------------------------------------------------------------------------------------------------
// MaxPool3d(2), floor mode
// ------------------------------------------------------------------------------------------------
__global__ void maxpool_fwd_kernel(const bf16* __restrict__ x, int D, int H, int W, int C, int P, bf16* __restrict__ y,
                                   float* __restrict__ partials) {
  extern __shared__ float red[];
  int p = blockIdx.x, n = blockIdx.y;
  int oD = D / 2, oH = H / 2, oW = W / 2;
  long long ovox = (long long)oD * oH * oW;
  EwMap m = ew_map(C);
  long long v0, v1;
  ew_range(ovox, p, P, v0, v1);
  float s[8] = {0}, q[8] = {0};
  if (m.active) {
    const bf16x8* xp = reinterpret_cast<const bf16x8*>(x + (size_t)n * D * H * W * C);
    bf16x8* yp = reinterpret_cast<bf16x8*>(y + (size_t)n * ovox * C);
    for (long long v = v0 + m.vl; v < v1; v += m.VL) {
      int ow = (int)(v % oW);
      long long r = v / oW;
      int oh = (int)(r % oH), od = (int)(r / oH);
      float mx[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) mx[i] = -INFINITY;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        int dz = k >> 2, dy = (k >> 1) & 1, dx = k & 1;
        size_t iv = ((size_t)(2 * od + dz) * H + (2 * oh + dy)) * W + (2 * ow + dx);
        float f[8];
        unpack8(xp[iv * m.CG + m.cg], f);
#pragma unroll
        for (int i = 0; i < 8; ++i) mx[i] = fmaxf(mx[i], f[i]);
      }
      yp[v * m.CG + m.cg] = pack8(mx);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        s[i] += mx[i];
        q[i] += mx[i] * mx[i];
      }
    }
  }
  if (partials) ew_write_partials(s, q, m, partials + ((size_t)n * P + p) * C * 2, red);
}

// iterates over 2x2x2 cells of the FULL-resolution grid (ceil), so ragged borders still get gadd*mask (or 0)
// GN = true: `gadd` holds the RAW data gradient dxhat of another consumer of x (a GroupNorm -> conv whose backward was deferred) and the
// term added is (A*dxhat + B*x + Cc) * act'(x) (gn_bwd_apply_kernel's formula); the per-channel totals of the 16-bit result go to
// `partials` [N][P][C][2].  One pass over the tensor instead of two.
template <bool GN>
__global__ void maxpool_bwd_kernel(const bf16* __restrict__ dpooled, const bf16* __restrict__ xf, int D, int H, int W, int C, int P,
                                   int act, float slope, const bf16* gadd, bf16* out, const float* __restrict__ coef,
                                   float* __restrict__ partials) {
  extern __shared__ float red[];
  const int p = blockIdx.x, n = blockIdx.y;
  const int oD = D / 2, oH = H / 2, oW = W / 2;
  const int cD = (D + 1) / 2, cH = (H + 1) / 2, cW = (W + 1) / 2;
  const EwMap m = ew_map(C);
  const LineMap lm = line_map(m, cW);
  int l0, l1;
  ew_range_i(cD * cH, p, P, l0, l1);
  if (!GN && !lm.active) return;
  float A[8], B[8], Cc[8], ssum[8] = {0}, ssq[8] = {0};
  if (GN && m.active) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float* cf = coef + ((size_t)n * C + m.cg * 8 + i) * 3;
      A[i] = cf[0];
      B[i] = cf[1];
      Cc[i] = cf[2];
    }
  }
  if (lm.active) {
  const size_t fvox = (size_t)D * H * W;
  const bf16x8* xp = reinterpret_cast<const bf16x8*>(xf + (size_t)n * fvox * C) + m.cg;
  const bf16x8* dp = reinterpret_cast<const bf16x8*>(dpooled + (size_t)n * oD * oH * oW * C) + m.cg;
  const bf16x8* gp = gadd ? reinterpret_cast<const bf16x8*>(gadd + (size_t)n * fvox * C) + m.cg : nullptr;
  bf16x8* op = reinterpret_cast<bf16x8*>(out + (size_t)n * fvox * C) + m.cg;
  for (int l = l0 + lm.ls; l < l1; l += lm.LPB) {
    const int ch = l % cH, cd = l / cH;
    const bool pooled_dh = cd < oD && ch < oH;
    for (int cw = lm.lw; cw < cW; cw += lm.lpl) {
      const bool pooled = pooled_dh && cw < oW;
      bf16x8 xr[8];
      bool inb[8];
      int arg[8];
      float mx[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        mx[i] = -INFINITY;
        arg[i] = 0;
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int z = 2 * cd + (k >> 2), yy = 2 * ch + ((k >> 1) & 1), xx = 2 * cw + (k & 1);
        inb[k] = z < D && yy < H && xx < W;
        if (inb[k]) xr[k] = xp[(((size_t)z * H + yy) * W + xx) * m.CG];
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (!inb[k]) continue;
        float f[8];
        unpack8(xr[k], f);
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (f[i] > mx[i]) {  // strict '>' : first maximum in (d,h,w) scan order wins, as max_pool3d_with_indices
            mx[i] = f[i];
            arg[i] = k;
          }
      }
      float g[8] = {0};
      if (pooled) unpack8(dp[(((size_t)cd * oH + ch) * oW + cw) * m.CG], g);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (!inb[k]) continue;
        const int z = 2 * cd + (k >> 2), yy = 2 * ch + ((k >> 1) & 1), xx = 2 * cw + (k & 1);
        const size_t iv = (((size_t)z * H + yy) * W + xx) * m.CG;
        float ga[8], o[8], f[8];
        unpack8(xr[k], f);
        if (gp) unpack8(gp[iv], ga);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float da = act_grad_from_out(f[i], act, slope);
          float t = (pooled && arg[i] == k) ? g[i] * da : 0.f;
          if (GN) {
            t += (A[i] * ga[i] + B[i] * f[i] + Cc[i]) * da;
            t = bf16_round(t);
            ssum[i] += t;
            ssq[i] += t * t;
          } else if (gp) {
            t += ga[i];
          }
          o[i] = t;
        }
        op[iv] = pack8(o);
      }
    }
  }
  }
  if (GN) ew_write_partials(ssum, ssq, m, partials + ((size_t)n * P + p) * C * 2, red);
}

// Lane-pair forward (C/8 a power of two <= 16), the same mapping as maxpool_bwd_pair_kernel below: a thread reduces the 2x2 (d,h)
// column of one fine w position (4 coalesced 16-byte loads), one shuffle per channel joins the two w positions, the even lane writes.
__global__ void __launch_bounds__(EW_THREADS) maxpool_fwd_pair_kernel(const bf16* __restrict__ x, int D, int H, int W, int C, int P,
                                                                      bf16* __restrict__ y, float* __restrict__ partials) {
  extern __shared__ float red[];
  const int p = blockIdx.x, n = blockIdx.y;
  const int oD = D / 2, oH = H / 2, oW = W / 2;
  const int Wf = 2 * oW;
  const EwMap m = ew_map(C);
  const LineMap lm = line_map(m, Wf);
  int l0, l1;
  ew_range_i(oD * oH, p, P, l0, l1);
  float s[8] = {0}, q[8] = {0};
  const bf16x8* xp = reinterpret_cast<const bf16x8*>(x + (size_t)n * D * H * W * C) + m.cg;
  bf16x8* yp = reinterpret_cast<bf16x8*>(y + (size_t)n * oD * oH * oW * C) + m.cg;
  const bool lane_ok = m.active && lm.active;
  const size_t dplane = (size_t)H * W * m.CG, dline = (size_t)W * m.CG;
  for (int lb = l0; lb < l1; lb += lm.LPB) {
    const int l = lb + lm.ls;
    const int oh = l % oH, od = l / oH;
    for (int fb = 0; fb < Wf; fb += lm.lpl) {
      const int xx = fb + lm.lw;
      const bool mine = lane_ok && l < l1 && xx < Wf;
      float mx[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) mx[i] = -INFINITY;
      if (mine) {
        const size_t iv0 = (((size_t)(2 * od) * H + 2 * oh) * W + xx) * m.CG;
        bf16x8 xr[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) xr[k] = xp[iv0 + (k >> 1) * dplane + (k & 1) * dline];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float f[8];
          unpack8(xr[k], f);
#pragma unroll
          for (int i = 0; i < 8; ++i) mx[i] = fmaxf(mx[i], f[i]);
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) mx[i] = fmaxf(mx[i], __shfl_xor_sync(0xffffffffu, mx[i], m.CG));
      if (mine && !(xx & 1)) {
        yp[(((size_t)od * oH + oh) * oW + (xx >> 1)) * m.CG] = pack8(mx);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          s[i] += mx[i];
          q[i] += mx[i] * mx[i];
        }
      }
    }
  }
  if (partials) ew_write_partials(s, q, m, partials + ((size_t)n * P + p) * C * 2, red);
}

// Lane-pair variant (C/8 a power of two <= 16): a thread owns the 2x2 (d,h) column of ONE fine w position of a cell, its neighbour
// lane (xor C/8) the other w position; consecutive lanes read consecutive 16-byte chunks (fully coalesced, half the registers of
// the one-thread-per-cell kernel above, twice the loads in flight) and the two halves of a cell settle the argmax with one shuffle
// per channel: larger value wins, on a tie the smaller scan index k = (z, y, x) -- exactly max_pool3d_with_indices' first maximum.
template <bool GN>
__global__ void __launch_bounds__(EW_THREADS, 2) maxpool_bwd_pair_kernel(const bf16* __restrict__ dpooled, const bf16* __restrict__ xf, int D, int H, int W,
                                                                      int C, int P, int act, float slope, const bf16* gadd, bf16* out,
                                                                      const float* __restrict__ coef, float* __restrict__ partials) {
  extern __shared__ float red[];
  const int p = blockIdx.x, n = blockIdx.y;
  const int oD = D / 2, oH = H / 2, oW = W / 2;
  const int cD = (D + 1) / 2, cH = (H + 1) / 2, cW = (W + 1) / 2;
  const int Wf = 2 * cW;  // fine w positions of a cell row (the last one may lie outside the tensor)
  const EwMap m = ew_map(C);
  const LineMap lm = line_map(m, Wf);
  int l0, l1;
  ew_range_i(cD * cH, p, P, l0, l1);
  float A[8], B[8], Cc[8], ssum[8] = {0}, ssq[8] = {0};
  if (GN) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float* cf = coef + ((size_t)n * C + (m.active ? m.cg : 0) * 8 + i) * 3;
      A[i] = cf[0];
      B[i] = cf[1];
      Cc[i] = cf[2];
    }
  }
  const size_t fvox = (size_t)D * H * W;
  const bf16x8* xp = reinterpret_cast<const bf16x8*>(xf + (size_t)n * fvox * C) + m.cg;
  const bf16x8* dp = reinterpret_cast<const bf16x8*>(dpooled + (size_t)n * oD * oH * oW * C) + m.cg;
  const bf16x8* gp = gadd ? reinterpret_cast<const bf16x8*>(gadd + (size_t)n * fvox * C) + m.cg : nullptr;
  bf16x8* op = reinterpret_cast<bf16x8*>(out + (size_t)n * fvox * C) + m.cg;
  const bool lane_ok = m.active && lm.active;
  // uniform trip counts: every lane of a warp reaches the shuffles
  for (int lb = l0; lb < l1; lb += lm.LPB) {
    const int l = lb + lm.ls;
    const int ch = l % cH, cd = l / cH;
    for (int fb = 0; fb < Wf; fb += lm.lpl) {
      const int xx = fb + lm.lw, cw = xx >> 1;
      const bool mine = lane_ok && l < l1 && xx < Wf;
      const bool pooled = mine && cd < oD && ch < oH && cw < oW;
      bf16x8 xr[4], gr[4];
      bool inb[4];
      const size_t iv0 = (((size_t)(2 * cd) * H + 2 * ch) * W + xx) * m.CG;  // voxel k: + (k>>1) planes + (k&1) lines
      const size_t dplane = (size_t)H * W * m.CG, dline = (size_t)W * m.CG;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        inb[k] = mine && 2 * cd + (k >> 1) < D && 2 * ch + (k & 1) < H && xx < W;
        const size_t iv = iv0 + (k >> 1) * dplane + (k & 1) * dline;
        if (inb[k]) {
          xr[k] = xp[iv];
          if (gp) gr[k] = gp[iv];
        }
      }
      float g[8] = {0};
      if (pooled) unpack8(dp[(((size_t)cd * oH + ch) * oW + cw) * m.CG], g);
      float mx[8];
      int arg[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        mx[i] = -INFINITY;
        arg[i] = 8;
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (!inb[k]) continue;
        float f[8];
        unpack8(xr[k], f);
        const int kk = (k << 1) | (xx & 1);  // scan index (z, y, x) of this voxel inside the cell
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (f[i] > mx[i]) {
            mx[i] = f[i];
            arg[i] = kk;
          }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {  // the other w position of the cell
        const float om = __shfl_xor_sync(0xffffffffu, mx[i], m.CG);
        const int oa = __shfl_xor_sync(0xffffffffu, arg[i], m.CG);
        if (om > mx[i] || (om == mx[i] && oa < arg[i])) arg[i] = oa;  // (value itself no longer needed)
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (!inb[k]) continue;
        const int kk = (k << 1) | (xx & 1);
        float ga[8], o[8], f[8];
        unpack8(xr[k], f);
        if (gp) unpack8(gr[k], ga);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float da = act_grad_from_out(f[i], act, slope);
          float t = (pooled && arg[i] == kk) ? g[i] * da : 0.f;
          if (GN) {
            t += (A[i] * ga[i] + B[i] * f[i] + Cc[i]) * da;
            t = bf16_round(t);
            ssum[i] += t;
            ssq[i] += t * t;
          } else if (gp) {
            t += ga[i];
          }
          o[i] = t;
        }
        op[iv0 + (k >> 1) * dplane + (k & 1) * dline] = pack8(o);
      }
    }
  }
  if (GN) ew_write_partials(ssum, ssq, m, partials + ((size_t)n * P + p) * C * 2, red);
}

// ------------------------------------------------------------------------------------------------
// nearest upsample (to the encoder feature size) + concat (encoder channels first)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int nearest_src(int dst, int in, int out) {
  // upsample_nearest3d: src = min(floor(dst * (float)in / out), in - 1)
  float scale = (float)in / (float)out;
  int s = (int)floorf((float)dst * scale);
  return s < in - 1 ? s : in - 1;
}

__global__ void upcat_fwd_kernel(const bf16* __restrict__ enc, int C0, const bf16* __restrict__ x, int C1, int D, int H, int W, int d,
                                 int h, int w, int P, bf16* __restrict__ cat, float* __restrict__ partials) {
  extern __shared__ float red[];
  const int p = blockIdx.x, n = blockIdx.y;
  const int C = C0 + C1;
  const size_t vox = (size_t)D * H * W;
  const EwMap m = ew_map(C);
  const LineMap lm = line_map(m, W);
  int l0, l1;
  ew_range_i(D * H, p, P, l0, l1);
  float s[8] = {0}, q[8] = {0};
  if (lm.active) {
    const int c = m.cg * 8;
    const bool from_enc = c < C0;
    const bf16* encp = enc + (size_t)n * vox * C0 + c;
    const bf16* xp = x + (size_t)n * d * h * w * C1 + (c - C0);
    bf16x8* op = reinterpret_cast<bf16x8*>(cat + (size_t)n * vox * C) + m.cg;
    for (int l = l0 + lm.ls; l < l1; l += lm.LPB) {
      const int xh = l % H, xd = l / H;
      const size_t srow = ((size_t)nearest_src(xd, d, D) * h + nearest_src(xh, h, H)) * w;
      for (int xw = lm.lw; xw < W; xw += lm.lpl) {
        const size_t v = (size_t)l * W + xw;
        bf16x8 val;
        if (from_enc)
          val = *reinterpret_cast<const bf16x8*>(encp + v * C0);
        else
          val = *reinterpret_cast<const bf16x8*>(xp + (srow + nearest_src(xw, w, W)) * C1);
        op[v * m.CG] = val;
        float f[8];
        unpack8(val, f);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          s[i] += f[i];
          q[i] += f[i] * f[i];
        }
      }
    }
  }
  if (partials) ew_write_partials(s, q, m, partials + ((size_t)n * P + p) * C * 2, red);
}

// destination index range [lo,hi] along one axis that maps to source index s (empty if lo>hi)
__device__ __forceinline__ void nearest_dst_range(int s, int in, int out, int& lo, int& hi) {
  if (out == 2 * in) {  // the usual case: exact 2x
    lo = 2 * s;
    hi = 2 * s + 1;
    return;
  }
  float inv = (float)out / (float)in;
  int a = (int)floorf((float)s * inv) - 2, b = (int)ceilf((float)(s + 1) * inv) + 2;
  if (a < 0) a = 0;
  if (b > out - 1) b = out - 1;
  lo = out;
  hi = -1;
  for (int t = a; t <= b; ++t)
    if (nearest_src(t, in, out) == s) {
      if (t < lo) lo = t;
      if (t > hi) hi = t;
    }
}

__global__ void upcat_bwd_kernel(const bf16* __restrict__ dcat, int C0, int C1, const bf16* __restrict__ xs, int D, int H, int W, int d,
                                 int h, int w, int P, int act, float slope, bf16* __restrict__ out) {
  const int p = blockIdx.x, n = blockIdx.y;
  const int C = C0 + C1;
  const size_t svox = (size_t)d * h * w, vox = (size_t)D * H * W;
  const EwMap m = ew_map(C1);
  const LineMap lm = line_map(m, w);
  int l0, l1;
  ew_range_i(d * h, p, P, l0, l1);
  if (!lm.active) return;
  const bf16* gp = dcat + (size_t)n * vox * C + C0 + m.cg * 8;
  for (int l = l0 + lm.ls; l < l1; l += lm.LPB) {
    const int sh = l % h, sd = l / h;
    int d0, d1, h0, h1;
    nearest_dst_range(sd, d, D, d0, d1);
    nearest_dst_range(sh, h, H, h0, h1);
    for (int sw = lm.lw; sw < w; sw += lm.lpl) {
      int w0, w1;
      nearest_dst_range(sw, w, W, w0, w1);
      float acc[8] = {0};
      for (int z = d0; z <= d1; ++z)
        for (int y = h0; y <= h1; ++y)
          for (int x = w0; x <= w1; ++x) {
            const size_t dv = ((size_t)z * H + y) * W + x;
            float f[8];
            unpack8(*reinterpret_cast<const bf16x8*>(gp + dv * C), f);
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] += f[i];
          }
      const size_t v = (size_t)l * w + sw;
      float xv[8];
      unpack8(*reinterpret_cast<const bf16x8*>(xs + ((size_t)n * svox + v) * C1 + m.cg * 8), xv);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] *= act_grad_from_out(xv[i], act, slope);
      *reinterpret_cast<bf16x8*>(out + ((size_t)n * svox + v) * C1 + m.cg * 8) = pack8(acc);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// border tap sums: T[n][tap][c] = sum_{v : v+tap-1 in bounds} dz[n,v,c]
// ------------------------------------------------------------------------------------------------
// stage 1a: per-channel TOTAL sums of dz: stats_ndhwc_bf16_kernel (a branch-free streaming pass at HBM speed).
// stage 1b: class sums of the BORDER voxels only (4-5 % of a 128^3 volume); the interior class is total - sum(border).
// grid (P, N, C/BT_CC).  Two evenly distributed phases per block:
//   A  the middles (w = 1..W-2) of the lines that lie on a d/h face, dealt round-robin over the blocks (one class per line:
//      registers -> one shuffle reduction -> shared-memory bins);
//   B  the two end voxels (w = 0, W-1) of a contiguous range of ALL lines, one (line, end, channel group) item per thread per
//      pass; lines off the d/h faces (almost all) accumulate in registers, face lines go straight to the bins.
// Partials Rp [N][P][64][C] (interior slot left 0).
constexpr int BT_CC = 64;
__global__ void border_class_sums_kernel(const bf16* __restrict__ dz, int D, int H, int W, int C, int P, float* __restrict__ Rp) {
  __shared__ float bins[64][BT_CC];
  const int p = blockIdx.x, n = blockIdx.y, c0 = blockIdx.z * BT_CC;
  const int CC = min(BT_CC, C - c0);
  const int CG = CC >> 3;
  for (int i = threadIdx.x; i < 64 * BT_CC; i += EW_THREADS) (&bins[0][0])[i] = 0.f;
  __syncthreads();
  const size_t vox = (size_t)D * H * W;
  const bf16* base = dz + (size_t)n * vox * C + c0;
  const int lane = threadIdx.x & 31;
  const int interior_dh = (1 << 4) | (1 << 2);
  // ---- phase A
  if (W > 2) {
    const int VL = EW_THREADS / CG;
    const int cg = threadIdx.x % CG, vl = threadIdx.x / CG;
    const int nd_face = D < 2 ? D : 2, nh_face = H < 2 ? H : 2;
    const int linesA = nd_face * H, nface = linesA + (D - nd_face) * nh_face;
    for (int f = p; f < nface; f += P) {
      int xd, xh;
      if (f < linesA) {
        xd = (f < H) ? 0 : D - 1;
        xh = f - (f < H ? 0 : H);
      } else {
        const int g = f - linesA;
        xd = 1 + g / nh_face;
        xh = (g % nh_face == 0) ? 0 : H - 1;
      }
      const int cls = (axis_cls(xd, D) << 4) | (axis_cls(xh, H) << 2) | 1;
      const bf16* line = base + ((size_t)xd * H + xh) * W * C + cg * 8;
      float acc[8] = {0};
      if (vl < VL)
        for (int xw = 1 + vl; xw < W - 1; xw += VL) {
          float v[8];
          unpack8(*reinterpret_cast<const bf16x8*>(line + (size_t)xw * C), v);
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[i] += v[i];
        }
      if ((32 % CG) == 0) {  // lanes of a warp that share cg differ by multiples of CG
#pragma unroll
        for (int i = 0; i < 8; ++i)
          for (int o = CG; o < 32; o <<= 1) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], o);
        if (lane < CG) {
#pragma unroll
          for (int i = 0; i < 8; ++i) atomicAdd(&bins[cls][cg * 8 + i], acc[i]);
        }
      } else if (vl < VL) {
#pragma unroll
        for (int i = 0; i < 8; ++i) atomicAdd(&bins[cls][cg * 8 + i], acc[i]);
      }
    }
  }
  // ---- phase B
  {
    const int ends = W >= 2 ? 2 : 1;
    const int per_line = ends * CG;
    const int LPP = EW_THREADS / per_line;  // lines per pass
    const int cg = threadIdx.x % CG, e = (threadIdx.x / CG) % ends, ls = threadIdx.x / per_line;
    const int xw = e == 0 ? 0 : W - 1;
    const int wcls = axis_cls(xw, W);
    int l0, l1;
    ew_range_i(D * H, p, P, l0, l1);
    float acc[8] = {0};
    if (ls < LPP) {
      for (int l = l0 + ls; l < l1; l += LPP) {
        const int xh = l % H, xd = l / H;
        const int cdh = (axis_cls(xd, D) << 4) | (axis_cls(xh, H) << 2);
        float v[8];
        unpack8(*reinterpret_cast<const bf16x8*>(base + ((size_t)l * W + xw) * C + cg * 8), v);
        if (cdh == interior_dh) {
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[i] += v[i];
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) atomicAdd(&bins[cdh | wcls][cg * 8 + i], v[i]);
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) atomicAdd(&bins[interior_dh | wcls][cg * 8 + i], acc[i]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 64 * CC; i += EW_THREADS) {
    int cls = i / CC, c = i % CC;
    Rp[(((size_t)n * P + p) * 64 + cls) * C + c0 + c] = bins[cls][c];
  }
}
// stage 2 (after the border class sums were reduced over P into R[n][cls][c] and the totals into tot[n][c][2], double):
// T[n][tap][c] = sum_{cls: tap valid} R, with R[interior] = total - sum(border classes); grid (ceil(27*C/256), N), block 256
__global__ void border_tap_from_class_kernel(const double* __restrict__ R, const double* __restrict__ tot, int C, float* __restrict__ T) {
  // one thread per (n, c): the 64 class sums are read once (independent, coalesced loads), then contracted axis by axis with the
  // tap-validity table:  T[td][th][tw] = sum_cd V(cd,td) sum_ch V(ch,th) sum_cw V(cw,tw) R[cd][ch][cw]
  const int n = blockIdx.y;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const int interior = (1 << 4) | (1 << 2) | 1;
  double r[64];
  double border_all = 0.0;
#pragma unroll
  for (int cls = 0; cls < 64; ++cls) {
    r[cls] = cls == interior ? 0.0 : R[((size_t)n * 64 + cls) * C + c];
    border_all += r[cls];
  }
  r[interior] = tot[((size_t)n * C + c) * 2] - border_all;
  double s1[16][3];  // [cd][ch][tw]
#pragma unroll
  for (int i = 0; i < 16; ++i)
#pragma unroll
    for (int tw = 0; tw < 3; ++tw) {
      double acc = 0.0;
#pragma unroll
      for (int cw = 0; cw < 4; ++cw)
        if (tap_valid(cw, tw)) acc += r[i * 4 + cw];
      s1[i][tw] = acc;
    }
  double s2[4][9];  // [cd][th][tw]
#pragma unroll
  for (int cd = 0; cd < 4; ++cd)
#pragma unroll
    for (int th = 0; th < 3; ++th)
#pragma unroll
      for (int tw = 0; tw < 3; ++tw) {
        double acc = 0.0;
#pragma unroll
        for (int ch = 0; ch < 4; ++ch)
          if (tap_valid(ch, th)) acc += s1[cd * 4 + ch][tw];
        s2[cd][th * 3 + tw] = acc;
      }
#pragma unroll
  for (int td = 0; td < 3; ++td)
#pragma unroll
    for (int t9 = 0; t9 < 9; ++t9) {
      double acc = 0.0;
#pragma unroll
      for (int cd = 0; cd < 4; ++cd)
        if (tap_valid(cd, td)) acc += s2[cd][t9];
      T[((size_t)n * 27 + td * 9 + t9) * C + c] = (float)acc;
    }
}

// dW[co][ci][tap] = sum_n ( a[n][ci] * sum_s G[n][s][tap][ci][co] + b[n][ci] * T[n][tap][co] )
__global__ void wgrad_finalize_kernel(const float* __restrict__ G, int N, int S, int Cin, int Cout, const float* __restrict__ ab,
                                      const float* __restrict__ T, float* __restrict__ dW, float* __restrict__ Gsum) {
  size_t total = (size_t)27 * Cin * Cout;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    int co = (int)(i % Cout);
    size_t r = i / Cout;
    int ci = (int)(r % Cin);
    int tap = (int)(r / Cin);
    double acc = 0.0;
    for (int n = 0; n < N; ++n) {
      double g = 0.0;
      for (int s = 0; s < S; ++s) g += (double)G[((((size_t)n * S + s) * 27 + tap) * Cin + ci) * Cout + co];
      if (Gsum) Gsum[(((size_t)n * 27 + tap) * Cin + ci) * Cout + co] = (float)g;  // split-reduced raw wgrad, reused by GN backward
      if (ab) {
        acc += (double)ab[((size_t)n * Cin + ci) * 2] * g;
        if (T) acc += (double)ab[((size_t)n * Cin + ci) * 2 + 1] * (double)T[((size_t)n * 27 + tap) * Cout + co];
      } else {
        acc += g;
      }
    }
    dW[((size_t)co * Cin + ci) * 27 + tap] = (float)acc;
  }
}

__global__ void bias_grad_from_T_kernel(const float* __restrict__ T, int N, int C, float* __restrict__ db) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double acc = 0.0;
  for (int n = 0; n < N; ++n) acc += (double)T[((size_t)n * 27 + 13) * C + c];  // centre tap is valid everywhere
  db[c] = (float)acc;
}

// ------------------------------------------------------------------------------------------------
// final 1x1x1 conv (+bias) + sigmoid / softmax; logits & probs are NCDHW fp32 (predictor.py:169 needs fp32)
// ------------------------------------------------------------------------------------------------
constexpr int FC_MAXO = 16;
__global__ void final_conv_fwd_kernel(const bf16* __restrict__ x, long long voxels, int C, const float* __restrict__ Wt,
                                      const float* __restrict__ bias, int Cout, int final_act, float* __restrict__ logits,
                                      float* __restrict__ probs) {
  extern __shared__ float wsm[];  // [Cout][C] + [Cout]
  int n = blockIdx.y;
  for (int i = threadIdx.x; i < Cout * C; i += blockDim.x) wsm[i] = Wt[i];
  for (int i = threadIdx.x; i < Cout; i += blockDim.x) wsm[Cout * C + i] = bias ? bias[i] : 0.f;
  __syncthreads();
  long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= voxels) return;
  float acc[FC_MAXO];
#pragma unroll
  for (int o = 0; o < FC_MAXO; ++o) acc[o] = o < Cout ? wsm[Cout * C + o] : 0.f;
  const bf16x8* xp = reinterpret_cast<const bf16x8*>(x + ((size_t)n * voxels + v) * C);
  for (int cg = 0; cg < C / 8; ++cg) {
    float f[8];
    unpack8(xp[cg], f);
#pragma unroll
    for (int o = 0; o < FC_MAXO; ++o)
      if (o < Cout) {
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[o] += f[i] * wsm[o * C + cg * 8 + i];
      }
  }
  float mx = -INFINITY, den = 0.f;
  if (final_act == B200_FINAL_SOFTMAX) {
#pragma unroll
    for (int o = 0; o < FC_MAXO; ++o)
      if (o < Cout) mx = fmaxf(mx, acc[o]);
#pragma unroll
    for (int o = 0; o < FC_MAXO; ++o)
      if (o < Cout) den += expf(acc[o] - mx);
  }
#pragma unroll
  for (int o = 0; o < FC_MAXO; ++o)
    if (o < Cout) {
      size_t idx = ((size_t)n * Cout + o) * voxels + v;
      logits[idx] = acc[o];
      if (probs) {
        float pr = acc[o];
        if (final_act == B200_FINAL_SIGMOID) pr = 1.f / (1.f + expf(-acc[o]));
        else if (final_act == B200_FINAL_SOFTMAX) pr = expf(acc[o] - mx) / den;
        probs[idx] = pr;
      }
    }
}

// dz = (sum_o dl[o] W[o][c]) * act'(x); partial sums of dW[o][c] and db[o]; grid (P, N)
// partials row layout: [Cout*C] dW then [Cout] db
__global__ void final_conv_bwd_kernel(const float* __restrict__ dl, const bf16* __restrict__ x, long long voxels, int C,
                                      const float* __restrict__ Wt, int Cout, int P, int act, float slope, bf16* __restrict__ dz,
                                      float* __restrict__ partials) {
  extern __shared__ float sm[];  // red[EW_THREADS*16] then W[Cout*C]
  float* red = sm;
  float* wsm = sm + EW_THREADS * 16;
  int p = blockIdx.x, n = blockIdx.y;
  for (int i = threadIdx.x; i < Cout * C; i += blockDim.x) wsm[i] = Wt[i];
  __syncthreads();
  EwMap m = ew_map(C);
  long long v0, v1;
  ew_range(voxels, p, P, v0, v1);
  float* prow = partials + ((size_t)n * P + p) * ((size_t)Cout * C + Cout);
  const bf16x8* xp = reinterpret_cast<const bf16x8*>(x + (size_t)n * voxels * C);
  bf16x8* zp = reinterpret_cast<bf16x8*>(dz + (size_t)n * voxels * C);
  // pass over output channels in pairs; dz is produced in the first pass (needs all o, cheap recompute)
  for (int o0 = 0; o0 < Cout; o0 += 2) {
    float s[8] = {0}, q[8] = {0};  // dW[o0][8ch], dW[o0+1][8ch]
    float db0 = 0.f, db1 = 0.f;
    if (m.active) {
      // two voxels per iteration (independent loads in flight); accumulation order per thread unchanged (v, then v + VL)
      for (long long va = v0 + m.vl; va < v1; va += 2 * m.VL) {
        const long long vb = va + m.VL;
        const bool two = vb < v1;
        bf16x8 xr[2];
        float d0r[2], d1r[2];
        xr[0] = xp[va * m.CG + m.cg];
        d0r[0] = dl[((size_t)n * Cout + o0) * voxels + va];
        d1r[0] = (o0 + 1 < Cout) ? dl[((size_t)n * Cout + o0 + 1) * voxels + va] : 0.f;
        if (two) {
          xr[1] = xp[vb * m.CG + m.cg];
          d0r[1] = dl[((size_t)n * Cout + o0) * voxels + vb];
          d1r[1] = (o0 + 1 < Cout) ? dl[((size_t)n * Cout + o0 + 1) * voxels + vb] : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          if (u == 1 && !two) break;
          const long long v = u ? vb : va;
          float f[8];
          unpack8(xr[u], f);
          const float d0 = d0r[u], d1 = d1r[u];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            s[i] += d0 * f[i];
            q[i] += d1 * f[i];
          }
          if (m.cg == 0) {
            db0 += d0;
            db1 += d1;
          }
          if (o0 == 0) {
            float g[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) g[i] = 0.f;
            for (int o = 0; o < Cout; ++o) {
              float d = o == 0 ? d0 : (o == 1 ? d1 : dl[((size_t)n * Cout + o) * voxels + v]);
#pragma unroll
              for (int i = 0; i < 8; ++i) g[i] += d * wsm[o * C + m.cg * 8 + i];
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) g[i] *= act_grad_from_out(f[i], act, slope);
            zp[v * m.CG + m.cg] = pack8(g);
          }
        }
      }
    }
    // reduce over voxel lanes: reuse the (s,q) -> [C][2] machinery into a temp in smem-free fashion
    if (m.active) {
      float* r = red + (size_t)(m.vl * m.CG + m.cg) * 16;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        r[i] = s[i];
        r[8 + i] = q[i];
      }
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < m.CG * 16; idx += EW_THREADS) {
      int cg = idx >> 4, i = idx & 15;
      float acc = 0.f;
      for (int vl = 0; vl < m.VL; ++vl) acc += red[(size_t)(vl * m.CG + cg) * 16 + i];
      int c = cg * 8 + (i & 7), k = i >> 3;
      if (o0 + k < Cout) prow[(size_t)(o0 + k) * C + c] = acc;
    }
    __syncthreads();
    // bias partials: lanes with cg==0 hold db; reduce through smem
    if (m.active && m.cg == 0) {
      red[m.vl * 2] = db0;
      red[m.vl * 2 + 1] = db1;
    }
    __syncthreads();
    if (threadIdx.x < 2 && o0 + (int)threadIdx.x < Cout) {
      float acc = 0.f;
      for (int vl = 0; vl < m.VL; ++vl) acc += red[vl * 2 + threadIdx.x];
      prow[(size_t)Cout * C + o0 + threadIdx.x] = acc;
    }
    __syncthreads();
  }
}

}  // namespace b200

// ================================================================================================
// C-ABI
// ================================================================================================
using namespace b200;
#define ST(s) ((cudaStream_t)(s))

extern "C" {

int b200_version(void) { return 1; }

int b200_last_error(char* buf, size_t len) {
  size_t n = strlen(g_err);
  if (buf && len) {
    size_t k = n < len - 1 ? n : len - 1;
    memcpy(buf, g_err, k);
    buf[k] = 0;
  }
  return (int)n;
}

int b200_device_is_sm100(void) {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
  return major == 10;
}

int b200_ncdhw_f32_to_ndhwc_f32(const float* src, float* dst, int N, int C, int D, int H, int W, b200_stream_t s) {
  long long vox = (long long)D * H * W;
  dim3 grid(ceil_div(vox, 256), N);
  ncdhw_to_ndhwc_kernel<float><<<grid, 256, 0, ST(s)>>>(src, dst, C, vox);
  B200_CHECK_LAUNCH("ncdhw_f32_to_ndhwc_f32");
  return 0;
}
int b200_ncdhw_f32_to_ndhwc_bf16(const float* src, void* dst, int N, int C, int D, int H, int W, b200_stream_t s) {
  long long vox = (long long)D * H * W;
  dim3 grid(ceil_div(vox, 256), N);
  ncdhw_to_ndhwc_kernel<bf16><<<grid, 256, 0, ST(s)>>>(src, (bf16*)dst, C, vox);
  B200_CHECK_LAUNCH("ncdhw_f32_to_ndhwc_bf16");
  return 0;
}
int b200_ndhwc_bf16_to_ncdhw_f32(const void* src, float* dst, int N, int C, int D, int H, int W, b200_stream_t s) {
  long long vox = (long long)D * H * W;
  dim3 grid(ceil_div(vox, 256), N);
  ndhwc_to_ncdhw_kernel<<<grid, 256, 0, ST(s)>>>((const bf16*)src, dst, C, vox);
  B200_CHECK_LAUNCH("ndhwc_bf16_to_ncdhw_f32");
  return 0;
}

int b200_stats_partials_count(int N, int C, long long voxels) {
  (void)N;
  if (C % 8 == 0) return ew_blocks(voxels, C);
  long long p = (voxels + 8191) / 8192;
  return (int)(p > 1024 ? 1024 : (p < 1 ? 1 : p));
}
int b200_stats_ncdhw_f32(const float* x, int N, int C, long long voxels, float* partials, b200_stream_t s) {
  long long p = (voxels + 8191) / 8192;
  int P = (int)(p > 1024 ? 1024 : (p < 1 ? 1 : p));
  dim3 grid(P, N * C);
  stats_ncdhw_f32_kernel<<<grid, 256, 0, ST(s)>>>(x, C, voxels, P, partials);
  B200_CHECK_LAUNCH("stats_ncdhw_f32");
  return 0;
}
int b200_stats_ncdhw_f32_partials_count(long long voxels) {
  long long p = (voxels + 8191) / 8192;
  return (int)(p > 1024 ? 1024 : (p < 1 ? 1 : p));
}
int b200_stats_ndhwc_bf16(const void* x, int N, int C, long long voxels, float* partials, b200_stream_t s) {
  B200_CHECK_ARG(C % 8 == 0 && C <= 2048, "stats_ndhwc_bf16: C=%d must be a multiple of 8 and <= 2048", C);
  int P = ew_blocks(voxels, C);
  dim3 grid(P, N);
  stats_ndhwc_bf16_kernel<<<grid, EW_THREADS, EW_THREADS * 16 * sizeof(float), ST(s)>>>((const bf16*)x, C, voxels, P, partials);
  B200_CHECK_LAUNCH("stats_ndhwc_bf16");
  return 0;
}
int b200_stats2_ndhwc_bf16(const void* a, const void* b, int N, int C, long long voxels, float* partials, b200_stream_t s) {
  B200_CHECK_ARG(C % 8 == 0 && C <= 2048, "stats2_ndhwc_bf16: C=%d must be a multiple of 8 and <= 2048", C);
  int P = ew_blocks(voxels, C);
  dim3 grid(P, N);
  stats2_ndhwc_bf16_kernel<<<grid, EW_THREADS, EW_THREADS * 16 * sizeof(float), ST(s)>>>((const bf16*)a, (const bf16*)b, C, voxels, P,
                                                                                   partials);
  B200_CHECK_LAUNCH("stats2_ndhwc_bf16");
  return 0;
}
int b200_gn_bwd_sums_from_wgrad(const float* G, int S, const float* T, const float* W, int N, int Cin, int Cout, double* sums2,
                                b200_stream_t s) {
  dim3 grid(Cin, N);
  gn_bwd_sums_from_wgrad_kernel<<<grid, 128, 0, ST(s)>>>(G, S, T, W, Cin, Cout, sums2);
  B200_CHECK_LAUNCH("gn_bwd_sums_from_wgrad");
  return 0;
}
int b200_partials_finalize(const float* partials, int N, int P, int C, double* sums, b200_stream_t s) {
  dim3 grid(ceil_div(C * 2, 32), N), block(32, 32);
  partials_finalize_kernel<<<grid, block, 0, ST(s)>>>(partials, P, C, sums);
  B200_CHECK_LAUNCH("partials_finalize");
  return 0;
}
int b200_reduce_rows(const float* partials, int P, int K, float* out, b200_stream_t s) {
  dim3 grid(ceil_div(K, 32)), block(32, 32);
  reduce_rows_kernel<<<grid, block, 0, ST(s)>>>(partials, P, K, out);
  B200_CHECK_LAUNCH("reduce_rows");
  return 0;
}

int b200_gn_coeffs(const double* sums, const float* gamma, const float* beta, int G, double count, int N, int C,
                   float* mean_rstd, float* ab, b200_stream_t s) {
  B200_CHECK_ARG(G > 0 && C % G == 0, "gn_coeffs: C=%d not divisible by G=%d", C, G);
  gn_coeffs_kernel<<<N, 128, G * 2 * sizeof(float), ST(s)>>>(sums, gamma, beta, G, count, C, mean_rstd, ab);
  B200_CHECK_LAUNCH("gn_coeffs");
  return 0;
}

int b200_gn_fold(const double* sums, const float* gamma, const float* beta, int G, double count, const float* W,
                 const float* conv_bias, int N, int Cin, int Cout, void* wf, float* biascls, float* mean_rstd, float* ab,
                 b200_stream_t s) {
  int n_w = 1;
  const float* abp = nullptr;
  if (sums) {
    int rc = b200_gn_coeffs(sums, gamma, beta, G, count, N, Cin, mean_rstd, ab, s);
    if (rc) return rc;
    n_w = N;
    abp = ab;
  }
  size_t total = (size_t)n_w * 27 * Cout * Cin;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 4096) blocks = 4096;
  fold_weights_kernel<<<blocks, 256, 0, ST(s)>>>(W, abp, n_w, Cin, Cout, (bf16*)wf);
  B200_CHECK_LAUNCH("fold_weights");
  if (biascls && (abp || conv_bias)) {
    dim3 grid(Cout, n_w);
    fold_bias_kernel<<<grid, 256, 0, ST(s)>>>(W, abp, conv_bias, sums, count, Cin, Cout, biascls);
    B200_CHECK_LAUNCH("fold_bias");
  }
  return 0;
}

// partials -> sums (fp64, as b200_partials_finalize) -> GroupNorm (mean, rstd) and the per-channel (a, b) in ONE launch
int b200_gn_stats_coeffs(const float* partials, int N, int P, int C, const float* gamma, const float* beta, int G, double count,
                         double* sums, float* mean_rstd, float* ab, b200_stream_t s) {
  B200_CHECK_ARG(G > 0 && C % G == 0 && C <= 2048, "gn_stats_coeffs: C=%d G=%d unsupported", C, G);
  dim3 grid(N, G), block(32, 32);
  size_t smem = (size_t)(C / G) * 2 * sizeof(double);
  gn_stats_coeffs_kernel<<<grid, block, smem, ST(s)>>>(partials, P, C, gamma, beta, G, count, sums, mean_rstd, ab);
  B200_CHECK_LAUNCH("gn_stats_coeffs");
  return 0;
}
// folded weights wf[n_w][27][Cout][Cin] and border-class bias table [n_w][64][Cout] in ONE launch (ab == NULL: n_w = 1, plain cast)
int b200_fold_weights_bias(const float* W, const float* ab, const float* conv_bias, const double* sums, double count, int N, int Cin,
                           int Cout, void* wf, float* biascls, b200_stream_t s) {
  const int n_w = ab ? N : 1;
  size_t total = (size_t)n_w * 27 * Cout * Cin;
  int nbw = (int)((total + 255) / 256);
  if (nbw > 4096) nbw = 4096;
  const int nbb = (biascls && (ab || conv_bias)) ? Cout * n_w : 0;
  fold_all_kernel<<<nbw + nbb, 256, 0, ST(s)>>>(W, ab, conv_bias, sums, count, n_w, Cin, Cout, nbw, (bf16*)wf, biascls);
  B200_CHECK_LAUNCH("fold_weights_bias");
  return 0;
}

int b200_prep_dgrad_weights(const float* W, int Cin, int Cout, void* wd, b200_stream_t s) {
  size_t total = (size_t)27 * Cin * Cout;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 4096) blocks = 4096;
  prep_dgrad_weights_kernel<<<blocks, 256, 0, ST(s)>>>(W, Cin, Cout, (bf16*)wd);
  B200_CHECK_LAUNCH("prep_dgrad_weights");
  return 0;
}

int b200_gn_apply_act_res(const void* x, const float* ab, const void* residual, int N, int C, long long voxels, int act, float slope,
                          void* y, float* partials, b200_stream_t s) {
  B200_CHECK_ARG(C % 8 == 0 && C <= 2048, "gn_apply_act: C=%d must be a multiple of 8", C);
  int P = ew_blocks(voxels, C);
  dim3 grid(P, N);
  gn_apply_act_kernel<<<grid, EW_THREADS, EW_THREADS * 16 * sizeof(float), ST(s)>>>((const bf16*)x, ab, (const bf16*)residual, C, voxels,
                                                                                  P, act, slope, (bf16*)y, partials);
  B200_CHECK_LAUNCH("gn_apply_act");
  return 0;
}
int b200_gn_apply_act(const void* x, const float* ab, int N, int C, long long voxels, int act, float slope, void* y,
                      float* partials, b200_stream_t s) {
  return b200_gn_apply_act_res(x, ab, nullptr, N, C, voxels, act, slope, y, partials, s);
}

int b200_gn_bwd_coeffs(const double* sums2, const float* gamma, const float* mean_rstd, int G, double count, int N, int C,
                       float* coef, float* dgamma, float* dbeta, b200_stream_t s) {
  gn_bwd_coeffs_kernel<<<1, 256, 0, ST(s)>>>(sums2, gamma, mean_rstd, G, count, N, C, coef, dgamma, dbeta);
  B200_CHECK_LAUNCH("gn_bwd_coeffs");
  return 0;
}

int b200_gn_bwd_apply(const void* dxhat, const void* x, const float* coef, int N, int C, long long voxels, int act, float slope,
                      const void* gadd, void* out, b200_stream_t s) {
  B200_CHECK_ARG(C % 8 == 0 && C <= 2048, "gn_bwd_apply: C=%d must be a multiple of 8", C);
  int P = ew_blocks(voxels, C);
  dim3 grid(P, N);
  gn_bwd_apply_kernel<false><<<grid, EW_THREADS, 0, ST(s)>>>((const bf16*)dxhat, (const bf16*)x, coef, C, voxels, P, act, slope,
                                                             (const bf16*)gadd, (bf16*)out, nullptr);
  B200_CHECK_LAUNCH("gn_bwd_apply");
  return 0;
}
int b200_gn_bwd_apply_stats(const void* dxhat, const void* x, const float* coef, int N, int C, long long voxels, int act, float slope,
                            const void* gadd, void* out, float* partials, b200_stream_t s) {
  B200_CHECK_ARG(C % 8 == 0 && C <= 2048 && partials, "gn_bwd_apply_stats: C=%d must be a multiple of 8, partials required", C);
  int P = ew_blocks(voxels, C);  // == b200_stats_partials_count
  dim3 grid(P, N);
  gn_bwd_apply_kernel<true><<<grid, EW_THREADS, EW_THREADS * 16 * sizeof(float), ST(s)>>>((const bf16*)dxhat, (const bf16*)x, coef, C, voxels, P,
                                                                                        act, slope, (const bf16*)gadd, (bf16*)out, partials);
  B200_CHECK_LAUNCH("gn_bwd_apply_stats");
  return 0;
}

int b200_act_bwd(const void* g, int g_cs, int g_co, const void* y, int N, int C, long long voxels, int act, float slope,
                 const void* gadd, void* out, b200_stream_t s) {
  B200_CHECK_ARG(C % 8 == 0 && C <= 2048, "act_bwd: C=%d must be a multiple of 8", C);
  B200_CHECK_ARG(g && g_cs % 8 == 0 && g_co % 8 == 0 && g_co + C <= g_cs, "act_bwd: bad gradient slice (cs=%d co=%d C=%d)", g_cs, g_co, C);
  int P = ew_blocks(voxels, C);
  dim3 grid(P, N);
  act_bwd_kernel<false><<<grid, EW_THREADS, 0, ST(s)>>>((const bf16*)g, g_cs, g_co, (const bf16*)y, C, voxels, P, act, slope,
                                                        (const bf16*)gadd, (bf16*)out, nullptr);
  B200_CHECK_LAUNCH("act_bwd");
  return 0;
}
// the same + per-channel totals of the result: partials [N][b200_stats_partials_count][C][2]
int b200_act_bwd_stats(const void* g, int g_cs, int g_co, const void* y, int N, int C, long long voxels, int act, float slope,
                       const void* gadd, void* out, float* partials, b200_stream_t s) {
  B200_CHECK_ARG(C % 8 == 0 && C <= 2048 && partials, "act_bwd_stats: C=%d must be a multiple of 8, partials required", C);
  B200_CHECK_ARG(g && g_cs % 8 == 0 && g_co % 8 == 0 && g_co + C <= g_cs, "act_bwd_stats: bad gradient slice (cs=%d co=%d C=%d)", g_cs, g_co, C);
  int P = ew_blocks(voxels, C);
  dim3 grid(P, N);
  act_bwd_kernel<true><<<grid, EW_THREADS, EW_THREADS * 16 * sizeof(float), ST(s)>>>((const bf16*)g, g_cs, g_co, (const bf16*)y, C, voxels, P,
                                                                                    act, slope, (const bf16*)gadd, (bf16*)out, partials);
  B200_CHECK_LAUNCH("act_bwd_stats");
  return 0;
}

static bool maxpool_pair_ok(int C);  // lane-pair kernels: the two w positions of a cell sit C/8 lanes apart in one warp
int b200_maxpool_partials_count(int N, int D, int H, int W, int C) {
  (void)N;
  return ew_blocks((long long)(D / 2) * (H / 2) * (W / 2), C);
}
int b200_maxpool_fwd(const void* x, int N, int D, int H, int W, int C, void* y, float* partials, b200_stream_t s) {
  B200_CHECK_ARG(C % 8 == 0 && C <= 2048, "maxpool_fwd: C=%d must be a multiple of 8", C);
  B200_CHECK_ARG(D >= 2 && H >= 2 && W >= 2, "maxpool_fwd: spatial size (%d,%d,%d) too small for MaxPool3d(2)", D, H, W);
  int P = b200_maxpool_partials_count(N, D, H, W, C);
  dim3 grid(P, N);
  if (maxpool_pair_ok(C))
    maxpool_fwd_pair_kernel<<<grid, EW_THREADS, EW_THREADS * 16 * sizeof(float), ST(s)>>>((const bf16*)x, D, H, W, C, P, (bf16*)y, partials);
  else
    maxpool_fwd_kernel<<<grid, EW_THREADS, EW_THREADS * 16 * sizeof(float), ST(s)>>>((const bf16*)x, D, H, W, C, P, (bf16*)y,
                                                                                   partials);
  B200_CHECK_LAUNCH("maxpool_fwd");
  return 0;
}
// lane-pair kernel: the two w positions of a cell sit C/8 lanes apart in one warp
static bool maxpool_pair_ok(int C) {
  const int cg = C / 8;
  const char* e = getenv("B200UNET_MAXPOOL_PAIR");
  return (cg & (cg - 1)) == 0 && cg <= 16 && !(e && e[0] == '0');
}
static int maxpool_bwd_blocks(int D, int H, int W, int C) {
  long long cells = (long long)((D + 1) / 2) * ((H + 1) / 2) * ((W + 1) / 2);
  return maxpool_pair_ok(C) ? ew_blocks_dense(2 * cells, C) : ew_blocks_dense(cells, C);
}
int b200_maxpool_bwd(const void* dpooled, const void* x_full, int N, int D, int H, int W, int C, int act, float slope,
                     const void* gadd, void* dz_full, b200_stream_t s) {
  B200_CHECK_ARG(C % 8 == 0 && C <= 2048, "maxpool_bwd: C=%d must be a multiple of 8", C);
  int P = maxpool_bwd_blocks(D, H, W, C);
  dim3 grid(P, N);
  if (maxpool_pair_ok(C))
    maxpool_bwd_pair_kernel<false><<<grid, EW_THREADS, 0, ST(s)>>>((const bf16*)dpooled, (const bf16*)x_full, D, H, W, C, P, act, slope,
                                                                   (const bf16*)gadd, (bf16*)dz_full, nullptr, nullptr);
  else
    maxpool_bwd_kernel<false><<<grid, EW_THREADS, 0, ST(s)>>>((const bf16*)dpooled, (const bf16*)x_full, D, H, W, C, P, act, slope,
                                                              (const bf16*)gadd, (bf16*)dz_full, nullptr, nullptr);
  B200_CHECK_LAUNCH("maxpool_bwd");
  return 0;
}
int b200_maxpool_bwd_partials_count(int N, int D, int H, int W, int C) {
  (void)N;
  return maxpool_bwd_blocks(D, H, W, C);
}
// dz_full = scatter(dpooled) * act'(x) + (A*dxhat + B*x + Cc) * act'(x)  (dz_full may alias dxhat); partials [N][P][C][2] of the result
int b200_maxpool_bwd_gn(const void* dpooled, const void* x_full, int N, int D, int H, int W, int C, int act, float slope, const void* dxhat,
                        const float* coef, void* dz_full, float* partials, b200_stream_t s) {
  B200_CHECK_ARG(C % 8 == 0 && C <= 2048 && dxhat && coef && partials, "maxpool_bwd_gn: C=%d must be a multiple of 8; dxhat, coef, partials required", C);
  int P = b200_maxpool_bwd_partials_count(N, D, H, W, C);
  dim3 grid(P, N);
  if (maxpool_pair_ok(C))
    maxpool_bwd_pair_kernel<true><<<grid, EW_THREADS, EW_THREADS * 16 * sizeof(float), ST(s)>>>(
        (const bf16*)dpooled, (const bf16*)x_full, D, H, W, C, P, act, slope, (const bf16*)dxhat, (bf16*)dz_full, coef, partials);
  else
    maxpool_bwd_kernel<true><<<grid, EW_THREADS, EW_THREADS * 16 * sizeof(float), ST(s)>>>(
        (const bf16*)dpooled, (const bf16*)x_full, D, H, W, C, P, act, slope, (const bf16*)dxhat, (bf16*)dz_full, coef, partials);
  B200_CHECK_LAUNCH("maxpool_bwd_gn");
  return 0;
}

int b200_upcat_partials_count(int N, int D, int H, int W, int C) {
  (void)N;
  return ew_blocks((long long)D * H * W, C);
}
int b200_upcat_fwd(const void* enc, int C0, const void* x, int C1, int N, int D, int H, int W, int d, int h, int w, void* cat,
                   float* partials, b200_stream_t s) {
  B200_CHECK_ARG(C0 % 8 == 0 && C1 % 8 == 0 && C0 + C1 <= 2048, "upcat_fwd: channel counts %d,%d must be multiples of 8", C0, C1);
  int P = b200_upcat_partials_count(N, D, H, W, C0 + C1);
  dim3 grid(P, N);
  upcat_fwd_kernel<<<grid, EW_THREADS, EW_THREADS * 16 * sizeof(float), ST(s)>>>((const bf16*)enc, C0, (const bf16*)x, C1, D, H, W,
                                                                               d, h, w, P, (bf16*)cat, partials);
  B200_CHECK_LAUNCH("upcat_fwd");
  return 0;
}
int b200_upcat_bwd(const void* dcat, int C0, int C1, const void* x_small, int N, int D, int H, int W, int d, int h, int w, int act,
                   float slope, void* dx_small, b200_stream_t s) {
  B200_CHECK_ARG(C0 % 8 == 0 && C1 % 8 == 0, "upcat_bwd: channel counts %d,%d must be multiples of 8", C0, C1);
  int P = ew_blocks_dense((long long)d * h * w, C1);
  dim3 grid(P, N);
  upcat_bwd_kernel<<<grid, EW_THREADS, 0, ST(s)>>>((const bf16*)dcat, C0, C1, (const bf16*)x_small, D, H, W, d, h, w, P, act, slope,
                                                   (bf16*)dx_small);
  B200_CHECK_LAUNCH("upcat_bwd");
  return 0;
}

static int border_blocks(int D, int H) {
  long long lines = (long long)D * H;
  long long p = (lines + 63) / 64;
  const int sms = sm_count();
  return (int)(p > sms ? sms : (p < 1 ? 1 : p));
}
int b200_border_tap_sums_workspace(int N, int D, int H, int W, int C) {
  // floats: border class partials [N][P][64][C] | totals partials [N][P2][C][2] | (doubles) R [N][64][C] | tot [N][C][2]
  int P = border_blocks(D, H);
  int P2 = ew_blocks((long long)D * H * W, C);
  size_t f = (size_t)N * P * 64 * C + (size_t)N * P2 * C * 2;
  f += f & 1;
  return (int)(f + 2 * ((size_t)N * 64 * C + (size_t)N * C * 2));
}
int b200_border_tap_sums(const void* dz, int N, int D, int H, int W, int C, float* T, float* scratch, b200_stream_t s) {
  return b200_border_tap_sums_pre(dz, N, D, H, W, C, nullptr, 0, T, scratch, s);
}
// same, with the per-channel totals of dz given as partial sums [N][Ptot][C][2] (column 0) by the kernel that produced dz
// (tot_partials == NULL: computed here by one streaming pass over dz)
int b200_border_tap_sums_pre(const void* dz, int N, int D, int H, int W, int C, const float* tot_partials, int Ptot, float* T, float* scratch,
                             b200_stream_t s) {
  B200_CHECK_ARG(C % 8 == 0 && C <= 2048, "border_tap_sums: C=%d must be a multiple of 8", C);
  const long long vox = (long long)D * H * W;
  int P = border_blocks(D, H);
  int P2 = ew_blocks(vox, C);
  float* Rp = scratch;
  float* totp = scratch + (size_t)N * P * 64 * C;
  size_t f = (size_t)N * P * 64 * C + (size_t)N * P2 * C * 2;
  f += f & 1;
  double* R = reinterpret_cast<double*>(scratch + f);
  double* tot = R + (size_t)N * 64 * C;
  {
    dim3 g2(ceil_div(C * 2, 32), N), b2(32, 32);
    if (tot_partials) {
      partials_finalize_kernel<<<g2, b2, 0, ST(s)>>>(tot_partials, Ptot, C, tot);
    } else {
      dim3 grid(P2, N);
      stats_ndhwc_bf16_kernel<<<grid, EW_THREADS, EW_THREADS * 16 * sizeof(float), ST(s)>>>((const bf16*)dz, C, vox, P2, totp);
      B200_CHECK_LAUNCH("border_totals");
      partials_finalize_kernel<<<g2, b2, 0, ST(s)>>>(totp, P2, C, tot);
    }
    B200_CHECK_LAUNCH("border_totals_reduce");
  }
  {
    dim3 grid(P, N, ceil_div(C, BT_CC));
    border_class_sums_kernel<<<grid, EW_THREADS, 0, ST(s)>>>((const bf16*)dz, D, H, W, C, P, Rp);
    B200_CHECK_LAUNCH("border_class_sums");
    dim3 g2(ceil_div(64 * C, 32), N), b2(32, 32);  // [N][P][64*C] viewed as [N][P][C'][2] with C' = 32*C
    partials_finalize_kernel<<<g2, b2, 0, ST(s)>>>(Rp, P, 32 * C, R);
    B200_CHECK_LAUNCH("border_class_reduce");
  }
  dim3 g3(ceil_div(C, 64), N);
  border_tap_from_class_kernel<<<g3, 64, 0, ST(s)>>>(R, tot, C, T);
  B200_CHECK_LAUNCH("border_tap_from_class");
  return 0;
}

int b200_wgrad_finalize(const float* G, int N, int S, int Cin, int Cout, const float* ab, const float* T, float* dW, float* Gsum,
                        b200_stream_t s) {
  size_t total = (size_t)27 * Cin * Cout;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 4096) blocks = 4096;
  wgrad_finalize_kernel<<<blocks, 256, 0, ST(s)>>>(G, N, S, Cin, Cout, ab, T, dW, Gsum);
  B200_CHECK_LAUNCH("wgrad_finalize");
  return 0;
}
int b200_bias_grad_from_T(const float* T, int N, int C, float* db, b200_stream_t s) {
  bias_grad_from_T_kernel<<<ceil_div(C, 128), 128, 0, ST(s)>>>(T, N, C, db);
  B200_CHECK_LAUNCH("bias_grad_from_T");
  return 0;
}

int b200_final_conv_fwd(const void* x, int N, long long voxels, int C, const float* W, const float* bias, int Cout,
                        int final_act, float* logits, float* probs, b200_stream_t s) {
  B200_CHECK_ARG(C % 8 == 0, "final_conv_fwd: C=%d must be a multiple of 8", C);
  B200_CHECK_ARG(Cout >= 1 && Cout <= FC_MAXO, "final_conv_fwd: out_channels=%d unsupported (max %d)", Cout, FC_MAXO);
  size_t smem = ((size_t)Cout * C + Cout) * sizeof(float);
  dim3 grid(ceil_div(voxels, 128), N);
  final_conv_fwd_kernel<<<grid, 128, smem, ST(s)>>>((const bf16*)x, voxels, C, W, bias, Cout, final_act, logits, probs);
  B200_CHECK_LAUNCH("final_conv_fwd");
  return 0;
}
int b200_final_conv_bwd_partials_count(int N, long long voxels, int C, int Cout) {
  (void)N;
  (void)Cout;
  return ew_blocks(voxels, C);
}
int b200_final_conv_bwd(const float* dlogits, const void* x, int N, long long voxels, int C, const float* W, int Cout, int act,
                        float slope, void* dz, float* partials, b200_stream_t s) {
  B200_CHECK_ARG(C % 8 == 0 && C <= 2048, "final_conv_bwd: C=%d must be a multiple of 8", C);
  int P = ew_blocks(voxels, C);
  dim3 grid(P, N);
  size_t smem = ((size_t)EW_THREADS * 16 + (size_t)Cout * C) * sizeof(float);
  B200_CHECK_ARG(smem <= 48 * 1024, "final_conv_bwd: Cout*C=%d too large", Cout * C);
  final_conv_bwd_kernel<<<grid, EW_THREADS, smem, ST(s)>>>(dlogits, (const bf16*)x, voxels, C, W, Cout, P, act, slope, (bf16*)dz,
                                                          partials);
  B200_CHECK_LAUNCH("final_conv_bwd");
  return 0;
}

}  // extern "C"
