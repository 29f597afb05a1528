// Kernels that only the residual model family needs (ResidualUNet3D / ResidualUNetSE3D):
//   * 1x1x1 convolution with bias (ResNetBlock.conv1, buildingblocks.py:251) forward / dgrad / wgrad
//   * ConvTranspose3d(k=3, stride=2, padding=1, bias=False) + nearest resize to the encoder size + sum-join
//     (TransposeConvUpsampling buildingblocks.py:617-664, Decoder._joining :493) forward / backward
// These carry ~1/27 (pointwise) resp. a few % (deconv) of the model's FLOPs; they are CUDA-core kernels in this
// round (the 3x3x3 convolutions of the residual blocks run on the tcgen05 kernels with the residual add + activation
// fused into the epilogue).
#include "common.cuh"
#include "ew.cuh"

namespace b200 {
void deconv_phase_table(signed char* k3, signed char* off, signed char* ntaps);  // conv_igemm_sm100.cu
}

namespace b200 {

template <typename T>
__device__ __forceinline__ float ldv(const T* p);
template <>
__device__ __forceinline__ float ldv<float>(const float* p) {
  return *p;
}
template <>
__device__ __forceinline__ float ldv<bf16>(const bf16* p) {
  return from_act(*p);
}

// ---------------------------------------------------------------------------------------------------------------
// y[v,co] = sum_ci Wq[co][ci] * x[v,ci] + bias[co]      (Wq: fp32 [Cout][Cin] or, transposed=1, [Cin_w][Cout_w] read as W^T)
// grid (P, N, ceil(Cout/64)); a block keeps its 64 x Cin weight slab in shared memory.
// ---------------------------------------------------------------------------------------------------------------
static int pw_co_chunk(int Cin) {  // output channels per block: the fp32 weight slab [chunk][Cin] must fit in shared memory
  int c = 64;
  while (c > 8 && (size_t)c * Cin * sizeof(float) > 160 * 1024) c >>= 1;
  return c;
}
template <typename InT>
__global__ void pointwise_fwd_kernel(const InT* __restrict__ x, const float* __restrict__ Wq, int transposed, const float* __restrict__ bias,
                                     long long vox, int Cin, int Cout, int P, int PW_CO, bf16* __restrict__ y, float* __restrict__ partials) {
  extern __shared__ float sm[];  // w[CC][Cin] | red[EW_THREADS*16]
  const int p = blockIdx.x, n = blockIdx.y, c0 = blockIdx.z * PW_CO;
  const int CC = min(PW_CO, Cout - c0);
  float* wsm = sm;
  float* red = sm + (size_t)PW_CO * Cin;
  for (int i = threadIdx.x; i < CC * Cin; i += EW_THREADS) {
    int co = i / Cin, ci = i % Cin;
    wsm[i] = transposed ? Wq[(size_t)ci * Cout + c0 + co] : Wq[(size_t)(c0 + co) * Cin + ci];
  }
  __syncthreads();
  const int CG = CC >> 3, VL = EW_THREADS / CG;
  const int cg = threadIdx.x % CG, vl = threadIdx.x / CG;
  long long v0, v1;
  ew_range(vox, p, P, v0, v1);
  float s[8] = {0}, q[8] = {0};
  if (vl < VL) {
    float b8[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) b8[i] = bias ? bias[c0 + cg * 8 + i] : 0.f;
    const InT* xn = x + (size_t)n * vox * Cin;
    for (long long v = v0 + vl; v < v1; v += VL) {
      float acc[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = b8[i];
      const InT* xp = xn + (size_t)v * Cin;
      const float* wp = wsm + (size_t)cg * 8 * Cin;
      for (int ci = 0; ci < Cin; ++ci) {
        float xv = ldv<InT>(xp + ci);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = fmaf(xv, wp[i * Cin + ci], acc[i]);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        acc[i] = bf16_round(acc[i]);
        s[i] += acc[i];
        q[i] += acc[i] * acc[i];
      }
      *reinterpret_cast<bf16x8*>(y + ((size_t)n * vox + v) * Cout + c0 + cg * 8) = pack8(acc);
    }
  }
  if (partials) {
    // per-channel partial sums of this block's channel slab: [N][P][Cout][2]
    EwMap m;
    m.CG = CG; m.VL = VL; m.cg = cg; m.vl = vl; m.active = vl < VL;
    ew_write_partials(s, q, m, partials + (((size_t)n * P + p) * Cout + c0) * 2, red);
  }
}

// dW[co][ci] = sum_v dy[v,co] * x[v,ci], db[co] = sum_v dy[v,co]; partial rows [N*P][Cout*Cin + Cout]; grid (P, N, tiles of 16x16 (co,ci))
template <typename InT>
__global__ void pointwise_wgrad_kernel(const InT* __restrict__ x, const bf16* __restrict__ dy, long long vox, int Cin, int Cout, int P,
                                       float* __restrict__ partials) {
  __shared__ float xs[64][17], ds[64][17];
  const int p = blockIdx.x, n = blockIdx.y;
  const int tiles_ci = (Cin + 15) / 16;
  const int co0 = (blockIdx.z / tiles_ci) * 16, ci0 = (blockIdx.z % tiles_ci) * 16;
  const int tco = threadIdx.x / 16, tci = threadIdx.x % 16;  // 256 threads = 16 x 16 outputs
  long long v0, v1;
  ew_range(vox, p, P, v0, v1);
  float acc = 0.f, accb = 0.f;
  for (long long vb = v0; vb < v1; vb += 64) {
    __syncthreads();
    for (int i = threadIdx.x; i < 64 * 16; i += 256) {
      int vv = i / 16, c = i % 16;
      long long v = vb + vv;
      bool ok = v < v1;
      xs[vv][c] = (ok && ci0 + c < Cin) ? ldv<InT>(x + ((size_t)n * vox + v) * Cin + ci0 + c) : 0.f;
      ds[vv][c] = (ok && co0 + c < Cout) ? from_act(dy[((size_t)n * vox + v) * Cout + co0 + c]) : 0.f;
    }
    __syncthreads();
#pragma unroll 16
    for (int vv = 0; vv < 64; ++vv) {
      acc = fmaf(ds[vv][tco], xs[vv][tci], acc);
      accb += ds[vv][tco];
    }
  }
  float* row = partials + ((size_t)n * P + p) * ((size_t)Cout * Cin + Cout);
  if (co0 + tco < Cout && ci0 + tci < Cin) row[(size_t)(co0 + tco) * Cin + ci0 + tci] = acc;
  if (ci0 == 0 && tci == 0 && co0 + tco < Cout) row[(size_t)Cout * Cin + co0 + tco] = accb;
}

// the same for the fp32 network input with C_in <= 4 (ResNetBlock.conv1 of the first encoder, buildingblocks.py:251): one streaming pass,
// a thread keeps 8 output channels x C_in products + the bias sums in registers; grid (P, N); HBM-bound (reads dy once)
template <int CIN>
__global__ void pointwise_wgrad_small_kernel(const float* __restrict__ x, const bf16* __restrict__ dy, long long vox, int Cout, int P,
                                             float* __restrict__ partials) {
  extern __shared__ float red[];  // [EW_THREADS][8]
  const int p = blockIdx.x, n = blockIdx.y;
  const EwMap m = ew_map(Cout);
  long long v0, v1;
  ew_range(vox, p, P, v0, v1);
  float aw[CIN][8], ab[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    ab[i] = 0.f;
#pragma unroll
    for (int c = 0; c < CIN; ++c) aw[c][i] = 0.f;
  }
  if (m.active) {
    const bf16x8* dp = reinterpret_cast<const bf16x8*>(dy + (size_t)n * vox * Cout) + m.cg;
    const float* xp = x + (size_t)n * vox * CIN;
    for (long long v = v0 + m.vl; v < v1; v += m.VL) {
      float f[8], xv[CIN];
      unpack8(dp[v * m.CG], f);
#pragma unroll
      for (int c = 0; c < CIN; ++c) xv[c] = xp[v * CIN + c];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        ab[i] += f[i];
#pragma unroll
        for (int c = 0; c < CIN; ++c) aw[c][i] = fmaf(f[i], xv[c], aw[c][i]);
      }
    }
  }
  float* row = partials + ((size_t)n * P + p) * ((size_t)Cout * CIN + Cout);
#pragma unroll
  for (int pass = 0; pass <= CIN; ++pass) {
    __syncthreads();
    if (m.active) {
      float* r = red + (size_t)(m.vl * m.CG + m.cg) * 8;
#pragma unroll
      for (int i = 0; i < 8; ++i) r[i] = pass < CIN ? aw[pass < CIN ? pass : 0][i] : ab[i];
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < m.CG * 8; idx += EW_THREADS) {  // idx = output channel
      float acc = 0.f;
      for (int vl = 0; vl < m.VL; ++vl) acc += red[(size_t)(vl * m.CG) * 8 + idx];
      if (pass < CIN) row[(size_t)idx * CIN + pass] = acc;
      else row[(size_t)Cout * CIN + idx] = acc;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// transposed conv by output parity phases (conv_igemm_sm100.cu: b200_deconv_phase_*): operand packing, weight-gradient assembly, and
// the join  out[j] = enc[j] + P[max(j, 1)]  (= enc + nearest-resize(T) for an encoder feature of exactly twice the low-res size)
// ---------------------------------------------------------------------------------------------------------------
struct DeconvK {
  signed char k3[27 * 3];
};
// wq[r][co][ci] = bf16(Wt[ci][co][k(r)])  (forward, (phase, tap) order);  wd[e][ci][co] = bf16(Wt[ci][co][e])  (adjoint)
__global__ void deconv_phase_weights_kernel(const float* __restrict__ Wt, int Cin, int Cout, DeconvK tab, bf16* __restrict__ wq,
                                            bf16* __restrict__ wd) {
  const size_t per = (size_t)Cin * Cout, total = 27 * per;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < 2 * total; i += (size_t)gridDim.x * blockDim.x) {
    if (i < total) {
      const int ci = (int)(i % Cin);
      const size_t r2 = i / Cin;
      const int co = (int)(r2 % Cout), r = (int)(r2 / Cout);
      const int k = (tab.k3[3 * r] * 3 + tab.k3[3 * r + 1]) * 3 + tab.k3[3 * r + 2];
      wq[i] = to_act(Wt[((size_t)ci * Cout + co) * 27 + k]);
    } else {
      const size_t t = i - total;
      const int co = (int)(t % Cout);
      const size_t r2 = t / Cout;
      const int ci = (int)(r2 % Cin), e = (int)(r2 / Cin);
      wd[t] = to_act(Wt[((size_t)ci * Cout + co) * 27 + e]);
    }
  }
}
// dWt[ci][co][e] = sum_rows Q[row][e][co][ci]   (rows = N * splits)
__global__ void deconv_phase_wgrad_finalize_kernel(const float* __restrict__ Q, int rows, int Cin, int Cout, float* __restrict__ dWt) {
  const size_t per = (size_t)27 * Cout * Cin;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < per; i += (size_t)gridDim.x * blockDim.x) {
    const int ci = (int)(i % Cin);   // ci fastest: coalesced reads of Q
    const size_t r2 = i / Cin;
    const int co = (int)(r2 % Cout), e = (int)(r2 / Cout);
    float acc = 0.f;
    for (int r = 0; r < rows; ++r) acc += Q[(size_t)r * per + i];
    dWt[((size_t)ci * Cout + co) * 27 + e] = acc;
  }
}
// out[j] = enc[j] + P[max(jd,1), max(jh,1), max(jw,1)]; partials of out.  grid (P, N)
__global__ void shift_add_fwd_kernel(const bf16* __restrict__ Pt, const bf16* __restrict__ enc, int D, int H, int W, int C, int P,
                                     bf16* __restrict__ out, float* __restrict__ partials) {
  extern __shared__ float red[];
  const int p = blockIdx.x, n = blockIdx.y;
  const EwMap m = ew_map(C);
  const LineMap lm = line_map(m, W);
  int l0, l1;
  ew_range_i(D * H, p, P, l0, l1);
  float s[8] = {0}, q[8] = {0};
  if (lm.active) {
    const size_t base = (size_t)n * D * H * W;
    const bf16x8* ep = reinterpret_cast<const bf16x8*>(enc + base * C) + m.cg;
    const bf16x8* pp = reinterpret_cast<const bf16x8*>(Pt + base * C) + m.cg;
    bf16x8* op = reinterpret_cast<bf16x8*>(out + base * C) + m.cg;
    for (int l = l0 + lm.ls; l < l1; l += lm.LPB) {
      const int xh = l % H, xd = l / H;
      const size_t srow = ((size_t)(xd > 1 ? xd : 1) * H + (xh > 1 ? xh : 1)) * W;
      for (int xw = lm.lw; xw < W; xw += lm.lpl) {
        float a[8], b[8];
        unpack8(ep[((size_t)l * W + xw) * m.CG], a);
        unpack8(pp[(srow + (xw > 1 ? xw : 1)) * m.CG], b);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          a[i] = bf16_round(a[i] + b[i]);
          s[i] += a[i];
          q[i] += a[i] * a[i];
        }
        op[((size_t)l * W + xw) * m.CG] = pack8(a);
      }
    }
  }
  if (partials) ew_write_partials(s, q, m, partials + ((size_t)n * P + p) * C * 2, red);
}
// adjoint of j -> max(j, 1) per axis: gp[j] = 0 where any coordinate is 0, else the sum of g over {0,1} along every axis where j == 1
__global__ void shift_fold_bwd_kernel(const bf16* __restrict__ g, int D, int H, int W, int C, int P, bf16* __restrict__ gp) {
  const int p = blockIdx.x, n = blockIdx.y;
  const EwMap m = ew_map(C);
  const LineMap lm = line_map(m, W);
  int l0, l1;
  ew_range_i(D * H, p, P, l0, l1);
  if (!lm.active) return;
  const size_t base = (size_t)n * D * H * W;
  const bf16x8* ip = reinterpret_cast<const bf16x8*>(g + base * C) + m.cg;
  bf16x8* op = reinterpret_cast<bf16x8*>(gp + base * C) + m.cg;
  for (int l = l0 + lm.ls; l < l1; l += lm.LPB) {
    const int xh = l % H, xd = l / H;
    for (int xw = lm.lw; xw < W; xw += lm.lpl) {
      float acc[8] = {0};
      if (xd > 0 && xh > 0 && xw > 0) {
        for (int zd = (xd == 1 ? 0 : xd); zd <= xd; ++zd)
          for (int zh = (xh == 1 ? 0 : xh); zh <= xh; ++zh)
            for (int zw = (xw == 1 ? 0 : xw); zw <= xw; ++zw) {
              float f[8];
              unpack8(ip[(((size_t)zd * H + zh) * W + zw) * m.CG], f);
#pragma unroll
              for (int i = 0; i < 8; ++i) acc[i] += f[i];
            }
      }
      op[((size_t)l * W + xw) * m.CG] = pack8(acc);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// helpers of the transposed-conv join (the conv itself runs on the tcgen05 kernels over the zero-inserted input, see
// Engine.deconv_up_add): nearest resize (2n-1 -> encoder size), its adjoint, zero-insert and its adjoint.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int nearest_src_i(int dst, int in, int out) {
  float scale = (float)in / (float)out;
  int s = (int)floorf((float)dst * scale);
  return s < in - 1 ? s : in - 1;
}
// dT[s,co] = sum over destination voxels o with s(o) == s of dout[o,co]   (adjoint of the nearest resize 2n-1 -> size);
// grid (P, N) over the (2d-1)(2h-1)(2w-1) deconv grid
__device__ __forceinline__ void dst_range(int s, int in, int out, int& lo, int& hi) {
  float inv = (float)out / (float)in;
  int a = (int)floorf((float)s * inv) - 2, b = (int)ceilf((float)(s + 1) * inv) + 2;
  if (a < 0) a = 0;
  if (b > out - 1) b = out - 1;
  lo = out;
  hi = -1;
  for (int t = a; t <= b; ++t)
    if (nearest_src_i(t, in, out) == s) {
      if (t < lo) lo = t;
      if (t > hi) hi = t;
    }
}
__global__ void deconv_gather_kernel(const bf16* __restrict__ dout, int sd, int sh, int sw, int D, int H, int W, int C, int P,
                                     bf16* __restrict__ dT) {
  const int p = blockIdx.x, n = blockIdx.y;
  EwMap m = ew_map(C);
  const long long svox = (long long)sd * sh * sw, vox = (long long)D * H * W;
  long long v0, v1;
  ew_range(svox, p, P, v0, v1);
  if (!m.active) return;
  for (long long v = v0 + m.vl; v < v1; v += m.VL) {
    const int xw = (int)(v % sw);
    const long long r = v / sw;
    const int xh = (int)(r % sh), xd = (int)(r / sh);
    int d0, d1, h0, h1, w0, w1;
    dst_range(xd, sd, D, d0, d1);
    dst_range(xh, sh, H, h0, h1);
    dst_range(xw, sw, W, w0, w1);
    float acc[8] = {0};
    for (int z = d0; z <= d1; ++z)
      for (int y = h0; y <= h1; ++y)
        for (int xx = w0; xx <= w1; ++xx) {
          float f[8];
          unpack8(*reinterpret_cast<const bf16x8*>(dout + ((size_t)n * vox + ((size_t)z * H + y) * W + xx) * C + m.cg * 8), f);
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[i] += f[i];
        }
    *reinterpret_cast<bf16x8*>(dT + ((size_t)n * svox + v) * C + m.cg * 8) = pack8(acc);
  }
}

// ---- transposed conv as a tensor-core 3x3x3 conv over the zero-inserted input -------------------------------------------------
// conv_transpose3d(x, Wt, stride 2, padding 1) == conv3d(zero_insert(x), Wc, padding 1) with Wc[co][ci][k] = Wt[ci][co][26-k];
// zero_insert(x)[2i] = x[i] on the (2d-1, 2h-1, 2w-1) grid.  7/8 of that tensor is zeros (8x redundant FLOPs), but the
// convolution, its dgrad and its wgrad then run on the tcgen05 kernels (two orders of magnitude faster than the CUDA-core
// gather formulation this replaced).

// xz[s] = (s all even) ? x[s/2] : 0 ; grid (P, N) over the (2d-1)(2h-1)(2w-1) grid
__global__ void zero_insert_kernel(const bf16* __restrict__ x, int d, int h, int w, int C, int P, bf16* __restrict__ xz) {
  const int p = blockIdx.x, n = blockIdx.y;
  EwMap m = ew_map(C);
  const int sd = 2 * d - 1, sh = 2 * h - 1, sw = 2 * w - 1;
  const long long svox = (long long)sd * sh * sw, vox = (long long)d * h * w;
  long long v0, v1;
  ew_range(svox, p, P, v0, v1);
  if (!m.active) return;
  bf16x8 zero;
  zero.u = make_uint4(0, 0, 0, 0);
  for (long long v = v0 + m.vl; v < v1; v += m.VL) {
    const int xw = (int)(v % sw);
    const long long r = v / sw;
    const int xh = (int)(r % sh), xd = (int)(r / sh);
    bf16x8 val = zero;
    if (!((xw | xh | xd) & 1))
      val = *reinterpret_cast<const bf16x8*>(x + ((size_t)n * vox + ((size_t)(xd >> 1) * h + (xh >> 1)) * w + (xw >> 1)) * C + m.cg * 8);
    *reinterpret_cast<bf16x8*>(xz + ((size_t)n * svox + v) * C + m.cg * 8) = val;
  }
}
// out[i] = dxz[2i] * act'(x[i]) [+ gadd]
__global__ void subsample2_bwd_kernel(const bf16* __restrict__ dxz, const bf16* __restrict__ x, int d, int h, int w, int C, int P, int act,
                                      float slope, const bf16* gadd, bf16* out) {
  const int p = blockIdx.x, n = blockIdx.y;
  EwMap m = ew_map(C);
  const int sh = 2 * h - 1, sw = 2 * w - 1;
  const long long svox = (long long)(2 * d - 1) * sh * sw, vox = (long long)d * h * w;
  long long v0, v1;
  ew_range(vox, p, P, v0, v1);
  if (!m.active) return;
  for (long long v = v0 + m.vl; v < v1; v += m.VL) {
    const int iw = (int)(v % w);
    const long long r = v / w;
    const int ih = (int)(r % h), id = (int)(r / h);
    float g[8], xv[8], ga[8];
    unpack8(*reinterpret_cast<const bf16x8*>(dxz + ((size_t)n * svox + ((size_t)(2 * id) * sh + 2 * ih) * sw + 2 * iw) * C + m.cg * 8), g);
    const size_t o = ((size_t)n * vox + v) * C + m.cg * 8;
    unpack8(*reinterpret_cast<const bf16x8*>(x + o), xv);
    if (gadd) unpack8(*reinterpret_cast<const bf16x8*>(gadd + o), ga);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      g[i] *= act_grad_from_out(xv[i], act, slope);
      if (gadd) g[i] += ga[i];
    }
    *reinterpret_cast<bf16x8*>(out + o) = pack8(g);
  }
}
// out[o] = enc[o] + T[nearest_src(o)] (+ partial sums of out for the next GroupNorm); T on the (sd,sh,sw) grid
__global__ void resize_add_fwd_kernel(const bf16* __restrict__ T, const bf16* __restrict__ enc, int sd, int sh, int sw, int D, int H, int W, int C,
                                      int P, bf16* __restrict__ out, float* __restrict__ partials) {
  extern __shared__ float red[];
  const int p = blockIdx.x, n = blockIdx.y;
  EwMap m = ew_map(C);
  const long long vox = (long long)D * H * W, svox = (long long)sd * sh * sw;
  long long v0, v1;
  ew_range(vox, p, P, v0, v1);
  float s[8] = {0}, q[8] = {0};
  if (m.active) {
    for (long long v = v0 + m.vl; v < v1; v += m.VL) {
      const int ow = (int)(v % W);
      const long long r = v / W;
      const int oh = (int)(r % H), od = (int)(r / H);
      const size_t sv = ((size_t)nearest_src_i(od, sd, D) * sh + nearest_src_i(oh, sh, H)) * sw + nearest_src_i(ow, sw, W);
      float a[8], b[8];
      unpack8(*reinterpret_cast<const bf16x8*>(enc + ((size_t)n * vox + v) * C + m.cg * 8), a);
      unpack8(*reinterpret_cast<const bf16x8*>(T + ((size_t)n * svox + sv) * C + m.cg * 8), b);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        a[i] = bf16_round(a[i] + b[i]);
        s[i] += a[i];
        q[i] += a[i] * a[i];
      }
      *reinterpret_cast<bf16x8*>(out + ((size_t)n * vox + v) * C + m.cg * 8) = pack8(a);
    }
  }
  if (partials) ew_write_partials(s, q, m, partials + ((size_t)n * P + p) * C * 2, red);
}
// Wc[co][ci][k] = Wt[ci][co][26-k]  (to_conv = 1)   or   dWt[ci][co][k] = dWc[co][ci][26-k]  (to_conv = 0)
__global__ void deconv_weight_permute_kernel(const float* __restrict__ src, int Cin, int Cout, int to_conv, float* __restrict__ dst) {
  size_t total = (size_t)27 * Cin * Cout;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    int k = (int)(i % 27);
    size_t r = i / 27;
    if (to_conv) {  // i indexes Wc[co][ci][k]
      int ci = (int)(r % Cin), co = (int)(r / Cin);
      dst[i] = src[((size_t)ci * Cout + co) * 27 + (26 - k)];
    } else {  // i indexes dWt[ci][co][k]
      int co = (int)(r % Cout), ci = (int)(r / Cout);
      dst[i] = src[((size_t)co * Cin + ci) * 27 + (26 - k)];
    }
  }
}

}  // namespace b200

using namespace b200;
#define ST(s) ((cudaStream_t)(s))

namespace b200 {
// bf16 weights of the 1x1x1 conv for the tensor-core path: wq[co][ci] (fprop) or wq[ci][co] (dgrad: roles of the channels swap)
__global__ void pointwise_prep_weights_kernel(const float* __restrict__ W, int Cin, int Cout, int transposed, bf16* __restrict__ wq) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Cin * Cout) return;
  if (!transposed) {
    wq[i] = to_act(W[i]);
  } else {
    int ci = i / Cout, co = i % Cout;
    wq[i] = to_act(W[(size_t)co * Cin + ci]);
  }
}
}  // namespace b200

extern "C" {

int b200_pointwise_prep_weights(const float* W, int Cin, int Cout, int transposed, void* wq, b200_stream_t s) {
  int total = Cin * Cout;
  b200::pointwise_prep_weights_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)s>>>(W, Cin, Cout, transposed, (bf16*)wq);
  B200_CHECK_LAUNCH("pointwise_prep_weights");
  return 0;
}


int b200_pointwise_partials_count(int N, long long voxels, int Cout) {
  (void)N;
  return ew_blocks(voxels, Cout < 8 ? 8 : (Cout < 64 ? Cout : 64));
}

int b200_pointwise_fwd(const void* x, int x_is_f32, const float* W, int transposed, const float* bias, int N, long long voxels, int Cin,
                       int Cout, void* y, float* partials, b200_stream_t s) {
  B200_CHECK_ARG(Cout % 8 == 0 && Cout <= 4096, "pointwise_fwd: Cout=%d must be a multiple of 8", Cout);
  int P = b200_pointwise_partials_count(N, voxels, Cout);
  const int PW_CO = pw_co_chunk(Cin);
  dim3 grid(P, N, ceil_div(Cout, PW_CO));
  size_t smem = ((size_t)PW_CO * Cin + EW_THREADS * 16) * sizeof(float);
  B200_CHECK_ARG(smem <= 200 * 1024, "pointwise_fwd: Cin=%d too large", Cin);
  if (x_is_f32) {
    cudaFuncSetAttribute(pointwise_fwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    pointwise_fwd_kernel<float><<<grid, EW_THREADS, smem, ST(s)>>>((const float*)x, W, transposed, bias, voxels, Cin, Cout, P, PW_CO, (bf16*)y,
                                                                  partials);
  } else {
    cudaFuncSetAttribute(pointwise_fwd_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    pointwise_fwd_kernel<bf16><<<grid, EW_THREADS, smem, ST(s)>>>((const bf16*)x, W, transposed, bias, voxels, Cin, Cout, P, PW_CO, (bf16*)y,
                                                                 partials);
  }
  B200_CHECK_LAUNCH("pointwise_fwd");
  return 0;
}

int b200_pointwise_wgrad_partials_count(int N, long long voxels) {
  (void)N;
  long long p = (voxels + 2047) / 2048;
  return (int)(p > 256 ? 256 : (p < 1 ? 1 : p));
}
int b200_pointwise_wgrad(const void* x, int x_is_f32, const void* dy, int N, long long voxels, int Cin, int Cout, float* partials,
                         b200_stream_t s) {
  int P = b200_pointwise_wgrad_partials_count(N, voxels);
  if (x_is_f32 && Cin <= 4 && Cout % 8 == 0 && Cout <= 2048) {
    dim3 g2(P, N);
    const size_t sm = EW_THREADS * 8 * sizeof(float);
    switch (Cin) {
      case 1: pointwise_wgrad_small_kernel<1><<<g2, EW_THREADS, sm, ST(s)>>>((const float*)x, (const bf16*)dy, voxels, Cout, P, partials); break;
      case 2: pointwise_wgrad_small_kernel<2><<<g2, EW_THREADS, sm, ST(s)>>>((const float*)x, (const bf16*)dy, voxels, Cout, P, partials); break;
      case 3: pointwise_wgrad_small_kernel<3><<<g2, EW_THREADS, sm, ST(s)>>>((const float*)x, (const bf16*)dy, voxels, Cout, P, partials); break;
      default: pointwise_wgrad_small_kernel<4><<<g2, EW_THREADS, sm, ST(s)>>>((const float*)x, (const bf16*)dy, voxels, Cout, P, partials); break;
    }
    B200_CHECK_LAUNCH("pointwise_wgrad_small");
    return 0;
  }
  dim3 grid(P, N, ceil_div(Cout, 16) * ceil_div(Cin, 16));
  if (x_is_f32) pointwise_wgrad_kernel<float><<<grid, 256, 0, ST(s)>>>((const float*)x, (const bf16*)dy, voxels, Cin, Cout, P, partials);
  else pointwise_wgrad_kernel<bf16><<<grid, 256, 0, ST(s)>>>((const bf16*)x, (const bf16*)dy, voxels, Cin, Cout, P, partials);
  B200_CHECK_LAUNCH("pointwise_wgrad");
  return 0;
}

int b200_deconv_phase_weights(const float* Wt, int Cin, int Cout, void* wq, void* wd, b200_stream_t s) {
  DeconvK tab;
  signed char off[27 * 3], nt[8];
  deconv_phase_table(tab.k3, off, nt);
  const size_t total = (size_t)2 * 27 * Cin * Cout;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 4096) blocks = 4096;
  deconv_phase_weights_kernel<<<blocks, 256, 0, ST(s)>>>(Wt, Cin, Cout, tab, (bf16*)wq, (bf16*)wd);
  B200_CHECK_LAUNCH("deconv_phase_weights");
  return 0;
}
int b200_deconv_phase_wgrad_finalize(const float* Q, int rows, int Cin, int Cout, float* dWt, b200_stream_t s) {
  const size_t per = (size_t)27 * Cin * Cout;
  int blocks = (int)((per + 255) / 256);
  if (blocks > 4096) blocks = 4096;
  deconv_phase_wgrad_finalize_kernel<<<blocks, 256, 0, ST(s)>>>(Q, rows, Cin, Cout, dWt);
  B200_CHECK_LAUNCH("deconv_phase_wgrad_finalize");
  return 0;
}
int b200_shift_add_fwd(const void* P, const void* enc, int N, int D, int H, int W, int C, void* out, float* partials, b200_stream_t s) {
  B200_CHECK_ARG(C % 8 == 0 && C <= 2048 && D >= 2 && H >= 2 && W >= 2, "shift_add_fwd: bad shape C=%d %dx%dx%d", C, D, H, W);
  int Pn = b200_upcat_partials_count(N, D, H, W, C);
  dim3 grid(Pn, N);
  shift_add_fwd_kernel<<<grid, EW_THREADS, EW_THREADS * 16 * sizeof(float), ST(s)>>>((const bf16*)P, (const bf16*)enc, D, H, W, C, Pn, (bf16*)out,
                                                                                 partials);
  B200_CHECK_LAUNCH("shift_add_fwd");
  return 0;
}
int b200_shift_fold_bwd(const void* g, int N, int D, int H, int W, int C, void* gp, b200_stream_t s) {
  B200_CHECK_ARG(C % 8 == 0 && C <= 2048, "shift_fold_bwd: C=%d must be a multiple of 8", C);
  int Pn = ew_blocks_dense((long long)D * H * W, C);
  dim3 grid(Pn, N);
  shift_fold_bwd_kernel<<<grid, EW_THREADS, 0, ST(s)>>>((const bf16*)g, D, H, W, C, Pn, (bf16*)gp);
  B200_CHECK_LAUNCH("shift_fold_bwd");
  return 0;
}

int b200_deconv_gather(const void* dout, int N, int d, int h, int w, int D, int H, int W, int C, void* dT, b200_stream_t s) {
  B200_CHECK_ARG(C % 8 == 0 && C <= 2048, "deconv_gather: C=%d must be a multiple of 8", C);
  int sd = 2 * d - 1, sh = 2 * h - 1, sw = 2 * w - 1;
  int P = ew_blocks((long long)sd * sh * sw, C);
  dim3 grid(P, N);
  deconv_gather_kernel<<<grid, EW_THREADS, 0, ST(s)>>>((const bf16*)dout, sd, sh, sw, D, H, W, C, P, (bf16*)dT);
  B200_CHECK_LAUNCH("deconv_gather");
  return 0;
}

int b200_zero_insert(const void* x, int N, int d, int h, int w, int C, void* xz, b200_stream_t s) {
  B200_CHECK_ARG(C % 8 == 0 && C <= 2048, "zero_insert: C=%d must be a multiple of 8", C);
  int P = ew_blocks((long long)(2 * d - 1) * (2 * h - 1) * (2 * w - 1), C);
  dim3 grid(P, N);
  zero_insert_kernel<<<grid, EW_THREADS, 0, ST(s)>>>((const bf16*)x, d, h, w, C, P, (bf16*)xz);
  B200_CHECK_LAUNCH("zero_insert");
  return 0;
}
int b200_subsample2_bwd(const void* dxz, const void* x, int N, int d, int h, int w, int C, int act, float slope, const void* gadd, void* out,
                        b200_stream_t s) {
  B200_CHECK_ARG(C % 8 == 0 && C <= 2048, "subsample2_bwd: C=%d must be a multiple of 8", C);
  int P = ew_blocks((long long)d * h * w, C);
  dim3 grid(P, N);
  subsample2_bwd_kernel<<<grid, EW_THREADS, 0, ST(s)>>>((const bf16*)dxz, (const bf16*)x, d, h, w, C, P, act, slope, (const bf16*)gadd, (bf16*)out);
  B200_CHECK_LAUNCH("subsample2_bwd");
  return 0;
}
int b200_resize_add_partials_count(int N, int D, int H, int W, int C) {
  (void)N;
  return ew_blocks((long long)D * H * W, C);
}
int b200_resize_add_fwd(const void* T, const void* enc, int N, int sd, int sh, int sw, int D, int H, int W, int C, void* out, float* partials,
                        b200_stream_t s) {
  B200_CHECK_ARG(C % 8 == 0 && C <= 2048, "resize_add_fwd: C=%d must be a multiple of 8", C);
  int P = b200_resize_add_partials_count(N, D, H, W, C);
  dim3 grid(P, N);
  resize_add_fwd_kernel<<<grid, EW_THREADS, EW_THREADS * 16 * sizeof(float), ST(s)>>>((const bf16*)T, (const bf16*)enc, sd, sh, sw, D, H, W, C, P,
                                                                                  (bf16*)out, partials);
  B200_CHECK_LAUNCH("resize_add_fwd");
  return 0;
}
int b200_deconv_weight_permute(const float* src, int Cin, int Cout, int to_conv, float* dst, b200_stream_t s) {
  size_t total = (size_t)27 * Cin * Cout;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 4096) blocks = 4096;
  deconv_weight_permute_kernel<<<blocks, 256, 0, ST(s)>>>(src, Cin, Cout, to_conv, dst);
  B200_CHECK_LAUNCH("deconv_weight_permute");
  return 0;
}

}  // extern "C"

// =================================================================================================================
// Concurrent spatial + channel squeeze-and-excitation (ChannelSpatialSELayer3D, reduction_ratio = 1; se.py:18-114):
//   s = mean_v y;  h = relu(W1 s + b1);  g = sigmoid(W2 h + b2)                  (cSE gates, per sample & channel)
//   q[v] = sigmoid(sum_c ws[c] y[v,c] + bs)                                      (sSE gate, per voxel)
//   out[v,c] = max(y[v,c]*g[c], y[v,c]*q[v])
// The channel means come for free from the (sum y) partials the producing conv's epilogue already emits.  The apply
// pass is one read of y and one write of out (the reference makes 9 full-tensor passes).
// =================================================================================================================
namespace b200 {

// cSE gates (ChannelSELayer3D, se.py:12-52): s = mean_v y ; h = relu(W1 s + b1) ; g = sigmoid(W2 h + b2).
// One fully-connected layer per launch, grid (ceil(C/8), N), block 256: warp w of block b owns output row 8b+w, its lanes
// stride over the input vector (coalesced reads of the weight row) and reduce with shuffles.
// mode 0: in = sums (double [N][C][2]) / count, writes the mean to `aux` and relu(.) to out; mode 1: in = vec, sigmoid(.) to out
__global__ void se_fc_kernel(const double* __restrict__ sums, double count, const float* __restrict__ vec, const float* __restrict__ W,
                             const float* __restrict__ b, int C, int mode, float* __restrict__ aux, float* __restrict__ out) {
  extern __shared__ float sh[];  // in[C]
  const int n = blockIdx.y;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float m = mode == 0 ? (float)(sums[((size_t)n * C + c) * 2] / count) : vec[(size_t)n * C + c];
    sh[c] = m;
    if (mode == 0 && blockIdx.x == 0) aux[(size_t)n * C + c] = m;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = blockIdx.x * 8 + warp;
  if (j >= C) return;
  const float* wr = W + (size_t)j * C;
  float acc = 0.f;
  for (int c = lane; c < C; c += 32) acc = fmaf(wr[c], sh[c], acc);
  for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) {
    acc += b[j];
    out[(size_t)n * C + j] = mode == 0 ? fmaxf(acc, 0.f) : 1.f / (1.f + expf(-acc));
  }
}

// lane mapping shared by the apply / backward kernels: a warp handles VPW voxels, each by GL = min(CG,32) lanes that own
// CPL = ceil(CG/32) 16-byte channel chunks each; dot products over channels reduce with xor shuffles inside the GL lanes.
struct SeMap {
  int CG, GL, VPW, CPL, sub, gl;
};
__device__ __forceinline__ SeMap se_map(int C, int lane) {
  SeMap m;
  m.CG = C >> 3;
  m.GL = m.CG < 32 ? m.CG : 32;
  m.VPW = 32 / m.GL;
  m.CPL = (m.CG + 31) / 32;
  m.sub = lane / m.GL;
  m.gl = lane % m.GL;
  return m;
}
__device__ __forceinline__ float se_group_sum(float v, int GL) {
  for (int o = GL >> 1; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
constexpr int SE_MAX_CPL = 4;  // C <= 1024

// grid (P, N), block 256 (8 warps).  q: float [N][V]
__global__ void scse_apply_fwd_kernel(const bf16* __restrict__ y, const float* __restrict__ g, const float* __restrict__ ws,
                                      const float* __restrict__ bs_ptr, int C,
                                      long long vox, int P, bf16* __restrict__ out, float* __restrict__ q) {
  const int p = blockIdx.x, n = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const SeMap m = se_map(C, lane);
  const float bs = bs_ptr[0];
  long long v0, v1;
  ew_range(vox, p, P, v0, v1);
  float gg[SE_MAX_CPL][8], ww[SE_MAX_CPL][8];
#pragma unroll
  for (int k = 0; k < SE_MAX_CPL; ++k) {
    const int cg = m.gl + 32 * k;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const bool ok = k < m.CPL && cg < m.CG;
      gg[k][i] = ok ? g[(size_t)n * C + cg * 8 + i] : 0.f;
      ww[k][i] = ok ? ws[cg * 8 + i] : 0.f;
    }
  }
  const int vstep = 8 * m.VPW;
  for (long long vb = v0 + warp * m.VPW; vb < v1; vb += vstep) {
    const long long v = vb + m.sub;
    const bool okv = v < v1;
    float f[SE_MAX_CPL][8];
    float dot = 0.f;
#pragma unroll
    for (int k = 0; k < SE_MAX_CPL; ++k) {
      const int cg = m.gl + 32 * k;
      if (k < m.CPL && cg < m.CG && okv) {
        unpack8(*reinterpret_cast<const bf16x8*>(y + ((size_t)n * vox + v) * C + cg * 8), f[k]);
#pragma unroll
        for (int i = 0; i < 8; ++i) dot = fmaf(f[k][i], ww[k][i], dot);
      }
    }
    dot = se_group_sum(dot, m.GL);
    const float qv = 1.f / (1.f + expf(-(dot + bs)));
    if (okv && m.gl == 0) q[(size_t)n * vox + v] = qv;
#pragma unroll
    for (int k = 0; k < SE_MAX_CPL; ++k) {
      const int cg = m.gl + 32 * k;
      if (k < m.CPL && cg < m.CG && okv) {
        float o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = fmaxf(f[k][i] * gg[k][i], f[k][i] * qv);
        *reinterpret_cast<bf16x8*>(out + ((size_t)n * vox + v) * C + cg * 8) = pack8(o);
      }
    }
  }
}

// backward pass 1.  tmp[v,c] = dout*(sel_c*g + sel_s*q) + dlogit*ws ;  partials [N][P][C][2] = (sum dout*y*sel_c, sum dlogit*y);
// dbs partial [N][P] = sum dlogit.   sel_c = (y*g > y*q) + 0.5*(y*g == y*q)   (torch.max splits the gradient at ties)
__global__ void scse_bwd1_kernel(const bf16* __restrict__ dout, const bf16* __restrict__ y, const float* __restrict__ g, const float* __restrict__ q,
                                 const float* __restrict__ ws, int C, long long vox, int P, bf16* __restrict__ tmp, float* __restrict__ partials,
                                 float* __restrict__ dbs_part) {
  extern __shared__ float red[];  // [8 warps][C][2] + [8]
  const int p = blockIdx.x, n = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const SeMap m = se_map(C, lane);
  long long v0, v1;
  ew_range(vox, p, P, v0, v1);
  float gg[SE_MAX_CPL][8], ww[SE_MAX_CPL][8], adg[SE_MAX_CPL][8], adw[SE_MAX_CPL][8];
  float adb = 0.f;
#pragma unroll
  for (int k = 0; k < SE_MAX_CPL; ++k) {
    const int cg = m.gl + 32 * k;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const bool ok = k < m.CPL && cg < m.CG;
      gg[k][i] = ok ? g[(size_t)n * C + cg * 8 + i] : 0.f;
      ww[k][i] = ok ? ws[cg * 8 + i] : 0.f;
      adg[k][i] = 0.f;
      adw[k][i] = 0.f;
    }
  }
  const int vstep = 8 * m.VPW;
  for (long long vb = v0 + warp * m.VPW; vb < v1; vb += vstep) {
    const long long v = vb + m.sub;
    const bool okv = v < v1;
    const float qv = okv ? q[(size_t)n * vox + v] : 0.f;
    float fy[SE_MAX_CPL][8], fd[SE_MAX_CPL][8], sc[SE_MAX_CPL][8];
    float dq = 0.f;
#pragma unroll
    for (int k = 0; k < SE_MAX_CPL; ++k) {
      const int cg = m.gl + 32 * k;
      if (k < m.CPL && cg < m.CG && okv) {
        const size_t o = ((size_t)n * vox + v) * C + cg * 8;
        unpack8(*reinterpret_cast<const bf16x8*>(y + o), fy[k]);
        unpack8(*reinterpret_cast<const bf16x8*>(dout + o), fd[k]);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float a = fy[k][i] * gg[k][i], b = fy[k][i] * qv;
          sc[k][i] = a > b ? 1.f : (a == b ? 0.5f : 0.f);
          dq = fmaf(fd[k][i] * fy[k][i], 1.f - sc[k][i], dq);
        }
      }
    }
    dq = se_group_sum(dq, m.GL);
    const float dl = dq * qv * (1.f - qv);
    if (okv && m.gl == 0) adb += dl;
#pragma unroll
    for (int k = 0; k < SE_MAX_CPL; ++k) {
      const int cg = m.gl + 32 * k;
      if (k < m.CPL && cg < m.CG && okv) {
        float o8[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          o8[i] = fd[k][i] * (sc[k][i] * gg[k][i] + (1.f - sc[k][i]) * qv) + dl * ww[k][i];
          adg[k][i] = fmaf(fd[k][i] * fy[k][i], sc[k][i], adg[k][i]);
          adw[k][i] = fmaf(dl, fy[k][i], adw[k][i]);
        }
        *reinterpret_cast<bf16x8*>(tmp + ((size_t)n * vox + v) * C + cg * 8) = pack8(o8);
      }
    }
  }
  // reduce over the voxel sub-groups of the warp (lanes with equal gl), then over the 8 warps through shared memory
#pragma unroll
  for (int k = 0; k < SE_MAX_CPL; ++k)
#pragma unroll
    for (int i = 0; i < 8; ++i)
      for (int o = m.GL; o < 32; o <<= 1) {
        adg[k][i] += __shfl_xor_sync(0xffffffffu, adg[k][i], o);
        adw[k][i] += __shfl_xor_sync(0xffffffffu, adw[k][i], o);
      }
  for (int o = 16; o; o >>= 1) adb += __shfl_xor_sync(0xffffffffu, adb, o);
  float* rw = red + (size_t)warp * C * 2;
  if (m.sub == 0) {
#pragma unroll
    for (int k = 0; k < SE_MAX_CPL; ++k) {
      const int cg = m.gl + 32 * k;
      if (k < m.CPL && cg < m.CG)
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          rw[(cg * 8 + i) * 2] = adg[k][i];
          rw[(cg * 8 + i) * 2 + 1] = adw[k][i];
        }
    }
  }
  float* rb = red + (size_t)8 * C * 2;
  if (lane == 0) rb[warp] = adb;
  __syncthreads();
  for (int i = threadIdx.x; i < C * 2; i += blockDim.x) {
    float a = 0.f;
    for (int wv = 0; wv < 8; ++wv) a += red[(size_t)wv * C * 2 + i];
    partials[((size_t)n * P + p) * C * 2 + i] = a;
  }
  if (threadIdx.x == 0) {
    float a = 0.f;
    for (int wv = 0; wv < 8; ++wv) a += rb[wv];
    dbs_part[(size_t)n * P + p] = a;
  }
}

// ---- gates backward (several small launches).  sums2 double [N][C][2] = (dg, dws_n).
// dl2[n][j] = dg*g*(1-g)
__global__ void se_dl2_kernel(const double* __restrict__ sums2, const float* __restrict__ g, int NC, float* __restrict__ dl2) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= NC) return;
  const float gv = g[i];
  dl2[i] = (float)sums2[(size_t)i * 2] * gv * (1.f - gv);
}
// transposed mat-vec out[n][c] = sum_j W[j][c] * vec[n][j]; grid (ceil(C/64), N), block 256 = 64 columns x 4 slices of j.
// mode 0: out = (h[n][c] > 0) ? acc : 0  (dl1);  mode 1: coef[n][c] = (1, 0, acc / count)
__global__ void se_matvec_t_kernel(const float* __restrict__ W, const float* __restrict__ vec, const float* __restrict__ h, int C, int mode,
                                   double count, float* __restrict__ out) {
  extern __shared__ float sh[];  // vec[C] | red[4][64]
  float* red = sh + C;
  const int n = blockIdx.y;
  for (int j = threadIdx.x; j < C; j += blockDim.x) sh[j] = vec[(size_t)n * C + j];
  __syncthreads();
  const int cl = threadIdx.x & 63, sl = threadIdx.x >> 6;
  const int c = blockIdx.x * 64 + cl;
  float acc = 0.f;
  if (c < C) {
    const int per = (C + 3) / 4;
    const int j0 = sl * per, j1 = min(C, j0 + per);
#pragma unroll 4
    for (int j = j0; j < j1; ++j) acc = fmaf(W[(size_t)j * C + c], sh[j], acc);
  }
  red[sl * 64 + cl] = acc;
  __syncthreads();
  if (sl == 0 && c < C) {
    acc = red[cl] + red[64 + cl] + red[128 + cl] + red[192 + cl];
    const size_t i = (size_t)n * C + c;
    if (mode == 0) {
      out[i] = h[i] > 0.f ? acc : 0.f;
    } else {
      out[i * 3] = 1.f;
      out[i * 3 + 1] = 0.f;
      out[i * 3 + 2] = (float)(acc / count);
    }
  }
}
// dW2[j][c] = sum_n dl2[n][j] h[n][c];  dW1[j][c] = sum_n dl1[n][j] s[n][c];  block (j = 0, first C threads) also db2, db1, dws
__global__ void se_outer_kernel(const float* __restrict__ dl2, const float* __restrict__ dl1, const float* __restrict__ h,
                                const float* __restrict__ sm, const double* __restrict__ sums2, int N, int C, float* __restrict__ dW1,
                                float* __restrict__ db1, float* __restrict__ dW2, float* __restrict__ db2, float* __restrict__ dws) {
  const int j = blockIdx.y;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float a2 = 0.f, a1 = 0.f;
  for (int n = 0; n < N; ++n) {
    a2 = fmaf(dl2[(size_t)n * C + j], h[(size_t)n * C + c], a2);
    a1 = fmaf(dl1[(size_t)n * C + j], sm[(size_t)n * C + c], a1);
  }
  dW2[(size_t)j * C + c] = a2;
  dW1[(size_t)j * C + c] = a1;
  if (j == 0) {
    float b2 = 0.f, b1 = 0.f, aw = 0.f;
    for (int n = 0; n < N; ++n) {
      b2 += dl2[(size_t)n * C + c];
      b1 += dl1[(size_t)n * C + c];
      aw += (float)sums2[((size_t)n * C + c) * 2 + 1];
    }
    db2[c] = b2;
    db1[c] = b1;
    dws[c] = aw;
  }
}

}  // namespace b200

extern "C" {

int b200_se_gates_fwd(const double* sums, double count, const float* W1, const float* b1, const float* W2, const float* b2, int N, int C,
                      float* smean, float* h, float* g, b200_stream_t s) {
  dim3 grid((C + 7) / 8, N);
  b200::se_fc_kernel<<<grid, 256, C * sizeof(float), ST(s)>>>(sums, count, nullptr, W1, b1, C, 0, smean, h);
  B200_CHECK_LAUNCH("se_fc1");
  b200::se_fc_kernel<<<grid, 256, C * sizeof(float), ST(s)>>>(nullptr, 1.0, h, W2, b2, C, 1, nullptr, g);
  B200_CHECK_LAUNCH("se_fc2");
  return 0;
}
int b200_scse_partials_count(int N, long long voxels, int C) {
  (void)N;
  // a block of 8 warps handles 8 * VPW voxels per pass (VPW = voxels per warp, see se_map); >= 4 passes per block
  int CG = C / 8;
  int GL = CG < 32 ? CG : 32;
  long long per_block = (long long)8 * (32 / (GL < 1 ? 1 : GL)) * 4;
  long long p = (voxels + per_block - 1) / per_block;
  return (int)(p > 1024 ? 1024 : (p < 1 ? 1 : p));
}
int b200_scse_apply_fwd(const void* y, const float* g, const float* ws, const float* bs, int N, long long voxels, int C, void* out, float* q,
                        b200_stream_t s) {
  int CG = C / 8;
  B200_CHECK_ARG(C % 8 == 0 && C <= 1024 && (CG & (CG - 1)) == 0, "scse_apply_fwd: C=%d must be 8 * a power of two, <= 1024", C);
  int P = b200_scse_partials_count(N, voxels, C);
  dim3 grid(P, N);
  b200::scse_apply_fwd_kernel<<<grid, 256, 0, ST(s)>>>((const bf16*)y, g, ws, bs, C, voxels, P, (bf16*)out, q);
  B200_CHECK_LAUNCH("scse_apply_fwd");
  return 0;
}
int b200_scse_bwd1(const void* dout, const void* y, const float* g, const float* q, const float* ws, int N, long long voxels, int C, void* tmp,
                   float* partials, float* dbs_part, b200_stream_t s) {
  int CG = C / 8;
  B200_CHECK_ARG(C % 8 == 0 && C <= 1024 && (CG & (CG - 1)) == 0, "scse_bwd1: C=%d must be 8 * a power of two, <= 1024", C);
  int P = b200_scse_partials_count(N, voxels, C);
  dim3 grid(P, N);
  size_t smem = ((size_t)8 * C * 2 + 8) * sizeof(float);
  cudaFuncSetAttribute(b200::scse_bwd1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  b200::scse_bwd1_kernel<<<grid, 256, smem, ST(s)>>>((const bf16*)dout, (const bf16*)y, g, q, ws, C, voxels, P, (bf16*)tmp, partials, dbs_part);
  B200_CHECK_LAUNCH("scse_bwd1");
  return 0;
}
int b200_se_gates_bwd(const double* sums2, const float* smean, const float* h, const float* g, const float* W1, const float* W2, int N, int C,
                      double count, float* coef, float* dW1, float* db1, float* dW2, float* db2, float* dws, float* scratch, b200_stream_t s) {
  float* dl2 = scratch;                  // [N][C]
  float* dl1 = scratch + (size_t)N * C;  // [N][C]
  b200::se_dl2_kernel<<<(N * C + 255) / 256, 256, 0, ST(s)>>>(sums2, g, N * C, dl2);
  B200_CHECK_LAUNCH("se_dl2");
  dim3 gt((C + 63) / 64, N);
  size_t smem = ((size_t)C + 256) * sizeof(float);
  b200::se_matvec_t_kernel<<<gt, 256, smem, ST(s)>>>(W2, dl2, h, C, 0, count, dl1);
  B200_CHECK_LAUNCH("se_dh");
  dim3 go((C + 255) / 256, C);
  b200::se_outer_kernel<<<go, 256, 0, ST(s)>>>(dl2, dl1, h, smean, sums2, N, C, dW1, db1, dW2, db2, dws);
  B200_CHECK_LAUNCH("se_outer");
  b200::se_matvec_t_kernel<<<gt, 256, smem, ST(s)>>>(W1, dl1, h, C, 1, count, coef);
  B200_CHECK_LAUNCH("se_ds");
  return 0;
}

}  // extern "C"
