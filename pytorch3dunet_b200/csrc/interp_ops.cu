// Decoder joins and pooling variants beyond the default nearest/max pair (HBM-bound vector kernels):
//   * trilinear upsampling to the encoder feature size + channel concat  (InterpolateUpsampling(mode='trilinear'),
//     reference buildingblocks.py:598-614 -> F.interpolate(x, size, mode) = upsample_trilinear3d, align_corners=False;
//     Decoder._joining concat :488-491) and its adjoint;
//   * AvgPool3d(2) (Encoder pool_type='avg', buildingblocks.py:358-363) and its adjoint.
// Same thread mapping and partial-sum conventions as elementwise.cu (ew.cuh).
#include "ew.cuh"

namespace b200 {

// upsample_trilinear3d source coordinate (ATen area_pixel_compute_source_index, align_corners=False, no scale_factor):
// src = max(scale*(dst+0.5)-0.5, 0), i0 = floor(src), i1 = i0 + (i0 < in-1), l1 = src - i0, l0 = 1 - l1; all in fp32
struct Lerp {
  int i0, i1;
  float l0, l1;
};
__device__ __forceinline__ Lerp lerp_src(int dst, int in, int out) {
  float scale = (float)in / (float)out;
  float src = scale * ((float)dst + 0.5f) - 0.5f;
  if (src < 0.f) src = 0.f;
  Lerp r;
  r.i0 = (int)src;
  if (r.i0 > in - 1) r.i0 = in - 1;
  r.i1 = r.i0 + (r.i0 < in - 1 ? 1 : 0);
  r.l1 = src - (float)r.i0;
  r.l0 = 1.f - r.l1;
  return r;
}
// weight with which source index s contributes to destination dst along one axis (0 when it does not)
__device__ __forceinline__ float lerp_weight(int dst, int s, int in, int out) {
  Lerp r = lerp_src(dst, in, out);
  float wgt = 0.f;
  if (r.i0 == s) wgt += r.l0;
  if (r.i1 == s) wgt += r.l1;
  return wgt;
}
// conservative destination range [lo,hi] that can reference source index s
__device__ __forceinline__ void lerp_dst_range(int s, int in, int out, int& lo, int& hi) {
  float inv = (float)out / (float)in;
  lo = (int)floorf(((float)s - 0.5f) * inv - 0.5f) - 1;
  hi = (int)ceilf(((float)s + 1.5f) * inv - 0.5f) + 1;
  if (lo < 0) lo = 0;
  if (hi > out - 1) hi = out - 1;
}

// cat[..., :C0] = enc, cat[..., C0:] = trilinear(x -> (D,H,W)); partials of cat.  grid (P, N)
__global__ void upcat_trilinear_fwd_kernel(const bf16* __restrict__ enc, int C0, const bf16* __restrict__ x, int C1, int D, int H, int W,
                                           int d, int h, int w, int P, bf16* __restrict__ cat, float* __restrict__ partials) {
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
      const Lerp ld = lerp_src(xd, d, D), lh = lerp_src(xh, h, H);
      for (int xw = lm.lw; xw < W; xw += lm.lpl) {
        const size_t v = (size_t)l * W + xw;
        float f[8];
        if (from_enc) {
          unpack8(*reinterpret_cast<const bf16x8*>(encp + v * C0), f);
        } else {
          const Lerp lw_ = lerp_src(xw, w, W);
#pragma unroll
          for (int i = 0; i < 8; ++i) f[i] = 0.f;
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int zi = (k & 4) ? ld.i1 : ld.i0, yi = (k & 2) ? lh.i1 : lh.i0, xi = (k & 1) ? lw_.i1 : lw_.i0;
            const float wgt = ((k & 4) ? ld.l1 : ld.l0) * ((k & 2) ? lh.l1 : lh.l0) * ((k & 1) ? lw_.l1 : lw_.l0);
            float t[8];
            unpack8(*reinterpret_cast<const bf16x8*>(xp + (((size_t)zi * h + yi) * w + xi) * C1), t);
#pragma unroll
            for (int i = 0; i < 8; ++i) f[i] += wgt * t[i];
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) f[i] = bf16_round(f[i]);
        }
        op[v * m.CG] = pack8(f);
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

// dx_small[s] = ( sum over destinations of weight(dst, s) * dcat[dst, C0:] ) * act'(x_small): gather form of the adjoint. grid (P, N)
__global__ void upcat_trilinear_bwd_kernel(const bf16* __restrict__ dcat, int C0, int C1, const bf16* __restrict__ xs, int D, int H, int W,
                                           int d, int h, int w, int P, int act, float slope, bf16* __restrict__ out) {
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
    lerp_dst_range(sd, d, D, d0, d1);
    lerp_dst_range(sh, h, H, h0, h1);
    for (int sw = lm.lw; sw < w; sw += lm.lpl) {
      int w0, w1;
      lerp_dst_range(sw, w, W, w0, w1);
      float acc[8] = {0};
      for (int z = d0; z <= d1; ++z) {
        const float wz = lerp_weight(z, sd, d, D);
        if (wz == 0.f) continue;
        for (int y = h0; y <= h1; ++y) {
          const float wy = wz * lerp_weight(y, sh, h, H);
          if (wy == 0.f) continue;
          for (int xx = w0; xx <= w1; ++xx) {
            const float wx = wy * lerp_weight(xx, sw, w, W);
            if (wx == 0.f) continue;
            float f[8];
            unpack8(*reinterpret_cast<const bf16x8*>(gp + (((size_t)z * H + y) * W + xx) * C), f);
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] += wx * f[i];
          }
        }
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

// AvgPool3d(2), floor mode; partials of y. grid (P, N)
__global__ void avgpool_fwd_kernel(const bf16* __restrict__ x, int D, int H, int W, int C, int P, bf16* __restrict__ y,
                                   float* __restrict__ partials) {
  extern __shared__ float red[];
  const int p = blockIdx.x, n = blockIdx.y;
  const int oD = D / 2, oH = H / 2, oW = W / 2;
  const long long ovox = (long long)oD * oH * oW;
  const EwMap m = ew_map(C);
  long long v0, v1;
  ew_range(ovox, p, P, v0, v1);
  float s[8] = {0}, q[8] = {0};
  if (m.active) {
    const bf16x8* xp = reinterpret_cast<const bf16x8*>(x + (size_t)n * D * H * W * C);
    bf16x8* yp = reinterpret_cast<bf16x8*>(y + (size_t)n * ovox * C);
    for (long long v = v0 + m.vl; v < v1; v += m.VL) {
      const int ow = (int)(v % oW);
      const long long r = v / oW;
      const int oh = (int)(r % oH), od = (int)(r / oH);
      float a[8] = {0};
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const size_t iv = ((size_t)(2 * od + (k >> 2)) * H + (2 * oh + ((k >> 1) & 1))) * W + (2 * ow + (k & 1));
        float f[8];
        unpack8(xp[iv * m.CG + m.cg], f);
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] += f[i];
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        a[i] = bf16_round(a[i] * 0.125f);
        s[i] += a[i];
        q[i] += a[i] * a[i];
      }
      yp[v * m.CG + m.cg] = pack8(a);
    }
  }
  if (partials) ew_write_partials(s, q, m, partials + ((size_t)n * P + p) * C * 2, red);
}

// dz_full[v] = dpooled[v/2]/8 * act'(x_full[v]) (0 on the ragged border that the floor-mode pool dropped) [+ gadd]. grid (P, N)
__global__ void avgpool_bwd_kernel(const bf16* __restrict__ dpooled, const bf16* __restrict__ xf, int D, int H, int W, int C, int P, int act,
                                   float slope, const bf16* gadd, bf16* out) {
  const int p = blockIdx.x, n = blockIdx.y;
  const int oD = D / 2, oH = H / 2, oW = W / 2;
  const EwMap m = ew_map(C);
  const LineMap lm = line_map(m, W);
  int l0, l1;
  ew_range_i(D * H, p, P, l0, l1);
  if (!lm.active) return;
  const size_t fvox = (size_t)D * H * W;
  const bf16x8* xp = reinterpret_cast<const bf16x8*>(xf + (size_t)n * fvox * C) + m.cg;
  const bf16x8* dp = reinterpret_cast<const bf16x8*>(dpooled + (size_t)n * oD * oH * oW * C) + m.cg;
  const bf16x8* gp = gadd ? reinterpret_cast<const bf16x8*>(gadd + (size_t)n * fvox * C) + m.cg : nullptr;
  bf16x8* op = reinterpret_cast<bf16x8*>(out + (size_t)n * fvox * C) + m.cg;
  for (int l = l0 + lm.ls; l < l1; l += lm.LPB) {
    const int yh = l % H, zd = l / H;
    const bool in_dh = (zd >> 1) < oD && (yh >> 1) < oH;
    for (int xw = lm.lw; xw < W; xw += lm.lpl) {
      const size_t iv = ((size_t)l * W + xw) * m.CG;
      float g[8] = {0}, f[8], o[8], ga[8];
      if (in_dh && (xw >> 1) < oW) unpack8(dp[(((size_t)(zd >> 1) * oH + (yh >> 1)) * oW + (xw >> 1)) * m.CG], g);
      unpack8(xp[iv], f);
      if (gp) unpack8(gp[iv], ga);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        o[i] = 0.125f * g[i] * act_grad_from_out(f[i], act, slope);
        if (gp) o[i] += ga[i];
      }
      op[iv] = pack8(o);
    }
  }
}

}  // namespace b200

using namespace b200;
#define ST(s) ((cudaStream_t)(s))

extern "C" {

int b200_upcat_trilinear_fwd(const void* enc, int C0, const void* x, int C1, int N, int D, int H, int W, int d, int h, int w, void* cat,
                             float* partials, b200_stream_t s) {
  B200_CHECK_ARG(C0 % 8 == 0 && C1 % 8 == 0 && C0 + C1 <= 2048, "upcat_trilinear_fwd: channel counts %d,%d must be multiples of 8", C0, C1);
  int P = b200_upcat_partials_count(N, D, H, W, C0 + C1);
  dim3 grid(P, N);
  upcat_trilinear_fwd_kernel<<<grid, EW_THREADS, EW_THREADS * 16 * sizeof(float), ST(s)>>>((const bf16*)enc, C0, (const bf16*)x, C1, D, H, W,
                                                                                         d, h, w, P, (bf16*)cat, partials);
  B200_CHECK_LAUNCH("upcat_trilinear_fwd");
  return 0;
}
int b200_upcat_trilinear_bwd(const void* dcat, int C0, int C1, const void* x_small, int N, int D, int H, int W, int d, int h, int w, int act,
                             float slope, void* dx_small, b200_stream_t s) {
  B200_CHECK_ARG(C0 % 8 == 0 && C1 % 8 == 0, "upcat_trilinear_bwd: channel counts %d,%d must be multiples of 8", C0, C1);
  int P = ew_blocks_dense((long long)d * h * w, C1);
  dim3 grid(P, N);
  upcat_trilinear_bwd_kernel<<<grid, EW_THREADS, 0, ST(s)>>>((const bf16*)dcat, C0, C1, (const bf16*)x_small, D, H, W, d, h, w, P, act,
                                                             slope, (bf16*)dx_small);
  B200_CHECK_LAUNCH("upcat_trilinear_bwd");
  return 0;
}

int b200_avgpool_fwd(const void* x, int N, int D, int H, int W, int C, void* y, float* partials, b200_stream_t s) {
  B200_CHECK_ARG(C % 8 == 0 && C <= 2048, "avgpool_fwd: C=%d must be a multiple of 8", C);
  B200_CHECK_ARG(D >= 2 && H >= 2 && W >= 2, "avgpool_fwd: spatial size (%d,%d,%d) too small for AvgPool3d(2)", D, H, W);
  int P = b200_maxpool_partials_count(N, D, H, W, C);
  dim3 grid(P, N);
  avgpool_fwd_kernel<<<grid, EW_THREADS, EW_THREADS * 16 * sizeof(float), ST(s)>>>((const bf16*)x, D, H, W, C, P, (bf16*)y, partials);
  B200_CHECK_LAUNCH("avgpool_fwd");
  return 0;
}
int b200_avgpool_bwd(const void* dpooled, const void* x_full, int N, int D, int H, int W, int C, int act, float slope, const void* gadd,
                     void* dz_full, b200_stream_t s) {
  B200_CHECK_ARG(C % 8 == 0 && C <= 2048, "avgpool_bwd: C=%d must be a multiple of 8", C);
  int P = ew_blocks_dense((long long)D * H * W, C);
  dim3 grid(P, N);
  avgpool_bwd_kernel<<<grid, EW_THREADS, 0, ST(s)>>>((const bf16*)dpooled, (const bf16*)x_full, D, H, W, C, P, act, slope,
                                                     (const bf16*)gadd, (bf16*)dz_full);
  B200_CHECK_LAUNCH("avgpool_bwd");
  return 0;
}

}  // extern "C"
