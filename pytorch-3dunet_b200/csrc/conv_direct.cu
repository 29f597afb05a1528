// Direct (CUDA-core) 3x3x3 convolution kernels.
//   * the production path for the network stem (C_in < 16: 25 FLOP/B, HBM-bound, no tensor-core shape) and
//   * the general on-GPU fallback / on-device checker for shapes the tcgen05 kernels do not take
//     (odd channel counts, tiny spatial sizes).  Same operands and epilogue contract as the tcgen05 kernel.
#include "common.cuh"
#include "ew.cuh"

namespace b200 {

template <typename T>
__device__ __forceinline__ float ldf(const T* p);
template <>
__device__ __forceinline__ float ldf<float>(const float* p) {
  return *p;
}
template <>
__device__ __forceinline__ float ldf<bf16>(const bf16* p) {
  return __bfloat162float(*p);
}

// one thread = one output voxel x 8 output channels.  grid (P, N)
template <typename InT>
__global__ void conv3_direct_fwd_kernel(const InT* __restrict__ x, const bf16* __restrict__ wf, int n_w, const float* __restrict__ biascls,
                                        int n_b, const bf16* __restrict__ residual, int act, float slope, int D, int H, int W, int Cin,
                                        int Cout, int P, bf16* __restrict__ y, int pmode, const bf16* __restrict__ aux,
                                        float* __restrict__ partials) {
  extern __shared__ float red[];
  int p = blockIdx.x, n = blockIdx.y;
  EwMap m = ew_map(Cout);
  long long vox = (long long)D * H * W, v0, v1;
  ew_range(vox, p, P, v0, v1);
  float s[8] = {0}, q[8] = {0};
  if (m.active) {
    const InT* xn = x + (size_t)n * vox * Cin;
    const bf16* wn = wf + (size_t)(n_w > 1 ? n : 0) * 27 * Cout * Cin;
    for (long long v = v0 + m.vl; v < v1; v += m.VL) {
      int xw = (int)(v % W);
      long long r = v / W;
      int xh = (int)(r % H), xd = (int)(r / H);
      float acc[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = 0.f;
      for (int td = 0; td < 3; ++td) {
        int zd = xd + td - 1;
        if (zd < 0 || zd >= D) continue;
        for (int th = 0; th < 3; ++th) {
          int zh = xh + th - 1;
          if (zh < 0 || zh >= H) continue;
          for (int tw = 0; tw < 3; ++tw) {
            int zw = xw + tw - 1;
            if (zw < 0 || zw >= W) continue;
            int tap = (td * 3 + th) * 3 + tw;
            const InT* xp = xn + (((size_t)zd * H + zh) * W + zw) * Cin;
            const bf16* wp = wn + ((size_t)tap * Cout + m.cg * 8) * Cin;
            for (int ci = 0; ci < Cin; ++ci) {
              float xv = ldf<InT>(xp + ci);
#pragma unroll
              for (int i = 0; i < 8; ++i) acc[i] += xv * __bfloat162float(wp[(size_t)i * Cin + ci]);
            }
          }
        }
      }
      if (n_b) {
        int cls = (axis_cls(xd, D) << 4) | (axis_cls(xh, H) << 2) | axis_cls(xw, W);
        const float* bp = biascls + ((size_t)(n_b > 1 ? n : 0) * 64 + cls) * Cout + m.cg * 8;
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] += bp[i];
      }
      size_t oidx = ((size_t)n * vox + v) * Cout + m.cg * 8;
      if (residual) {
        float rv[8];
        unpack8(*reinterpret_cast<const bf16x8*>(residual + oidx), rv);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] += rv[i];
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = bf16_round(act_fwd(acc[i], act, slope));
      *reinterpret_cast<bf16x8*>(y + oidx) = pack8(acc);
      if (pmode == 1) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          s[i] += acc[i];
          q[i] += acc[i] * acc[i];
        }
      } else if (pmode == 2) {
        float av[8];
        unpack8(*reinterpret_cast<const bf16x8*>(aux + oidx), av);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          s[i] += acc[i];
          q[i] += acc[i] * av[i];
        }
      }
    }
  }
  if (pmode) ew_write_partials(s, q, m, partials + ((size_t)n * P + p) * Cout * 2, red);
}

// G[n][0][tap][ci][co] += sum_{v in chunk} dz[v,co] * x[v+tap-1,ci]   (atomic across chunks; G pre-zeroed)
// grid (chunks, N, ceil(27*Cin*Cout / (256*OPT)))
constexpr int WG_CHUNK = 2048;
template <typename InT>
__global__ void conv3_direct_wgrad_kernel(const InT* __restrict__ x, const bf16* __restrict__ dz, int D, int H, int W, int Cin, int Cout,
                                          float* __restrict__ G) {
  int n = blockIdx.y;
  long long vox = (long long)D * H * W;
  long long v0 = (long long)blockIdx.x * WG_CHUNK, v1 = v0 + WG_CHUNK;
  if (v1 > vox) v1 = vox;
  int total = 27 * Cin * Cout;
  int o = blockIdx.z * blockDim.x + threadIdx.x;
  if (o >= total) return;
  int co = o % Cout;
  int r = o / Cout;
  int ci = r % Cin;
  int tap = r / Cin;
  int td = tap / 9 - 1, th = (tap / 3) % 3 - 1, tw = tap % 3 - 1;
  const InT* xn = x + (size_t)n * vox * Cin;
  const bf16* dn = dz + (size_t)n * vox * Cout;
  float acc = 0.f;
  int xw = (int)(v0 % W);
  long long rr = v0 / W;
  int xh = (int)(rr % H), xd = (int)(rr / H);
  for (long long v = v0; v < v1; ++v) {
    int zd = xd + td, zh = xh + th, zw = xw + tw;
    if (zd >= 0 && zd < D && zh >= 0 && zh < H && zw >= 0 && zw < W)
      acc += __bfloat162float(dn[(size_t)v * Cout + co]) * ldf<InT>(xn + (((size_t)zd * H + zh) * W + zw) * Cin + ci);
    if (++xw == W) {
      xw = 0;
      if (++xh == H) {
        xh = 0;
        ++xd;
      }
    }
  }
  atomicAdd(&G[(((size_t)n * 27 + tap) * Cin + ci) * Cout + co], acc);
}


// ------------------------------------------------------------------------------------------------
// network stem (fp32 input, C_in <= 4): 25 FLOP/B, HBM-bound.  One thread = one voxel x all COUT channels,
// folded weights + border-class bias staged in shared memory as fp32, 27*C_in broadcast-free scalar loads of x
// (neighbouring threads share cache lines), one 2*COUT-byte vector store.  grid (P, N), block 256.
// ------------------------------------------------------------------------------------------------
template <int COUT>
__global__ void __launch_bounds__(256, 2) stem_conv_fwd_kernel(const float* __restrict__ x, const bf16* __restrict__ wf, int n_w,
                                                            const float* __restrict__ biascls, int n_b, int act, float slope, int D, int H,
                                                            int W, int Cin, int P, bf16* __restrict__ y, int pmode,
                                                            float* __restrict__ partials) {
  extern __shared__ float sm[];  // w[27*Cin][COUT] | bias[64][COUT] | red[8][COUT*2]
  float* wsm = sm;
  float* bsm = wsm + 27 * Cin * COUT;
  float* red = bsm + 64 * COUT;
  const int p = blockIdx.x, n = blockIdx.y;
  const bf16* wn = wf + (size_t)(n_w > 1 ? n : 0) * 27 * COUT * Cin;
  for (int i = threadIdx.x; i < 27 * Cin * COUT; i += 256) {
    int co = i % COUT, r = i / COUT, ci = r % Cin, tap = r / Cin;
    wsm[i] = __bfloat162float(wn[((size_t)tap * COUT + co) * Cin + ci]);
  }
  for (int i = threadIdx.x; i < 64 * COUT; i += 256) bsm[i] = n_b ? biascls[(size_t)(n_b > 1 ? n : 0) * 64 * COUT + i] : 0.f;
  __syncthreads();
  const long long vox = (long long)D * H * W;
  long long v0, v1;
  ew_range(vox, p, P, v0, v1);
  const float* xn = x + (size_t)n * vox * Cin;
  float s[COUT], q[COUT];
#pragma unroll
  for (int i = 0; i < COUT; ++i) s[i] = q[i] = 0.f;
  for (long long v = v0 + threadIdx.x; v < v1; v += 256) {
    int xw = (int)(v % W);
    long long r = v / W;
    int xh = (int)(r % H), xd = (int)(r / H);
    float acc[COUT];
    const float* bp = bsm + ((axis_cls(xd, D) << 4) | (axis_cls(xh, H) << 2) | axis_cls(xw, W)) * COUT;
#pragma unroll
    for (int i = 0; i < COUT; ++i) acc[i] = bp[i];
    for (int td = 0; td < 3; ++td) {
      int zd = xd + td - 1;
      if (zd < 0 || zd >= D) continue;
      for (int th = 0; th < 3; ++th) {
        int zh = xh + th - 1;
        if (zh < 0 || zh >= H) continue;
#pragma unroll
        for (int tw = 0; tw < 3; ++tw) {
          int zw = xw + tw - 1;
          if (zw < 0 || zw >= W) continue;
          const float* xp = xn + (((size_t)zd * H + zh) * W + zw) * Cin;
          const float* wp = wsm + (size_t)((td * 3 + th) * 3 + tw) * Cin * COUT;
          for (int ci = 0; ci < Cin; ++ci) {
            float xv = __ldg(xp + ci);
#pragma unroll
            for (int i = 0; i < COUT; ++i) acc[i] = fmaf(xv, wp[ci * COUT + i], acc[i]);
          }
        }
      }
    }
#pragma unroll
    for (int i = 0; i < COUT; ++i) {
      acc[i] = bf16_round(act_fwd(acc[i], act, slope));
      s[i] += acc[i];
      q[i] += acc[i] * acc[i];
    }
    bf16x8* op = reinterpret_cast<bf16x8*>(y + ((size_t)n * vox + v) * COUT);
#pragma unroll
    for (int i = 0; i < COUT / 8; ++i) op[i] = pack8(&acc[8 * i]);
  }
  if (pmode) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int i = 0; i < COUT; ++i) {
      float a = s[i], b = q[i];
      for (int o = 16; o; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
      }
      if (lane == 0) {
        red[(warp * COUT + i) * 2] = a;
        red[(warp * COUT + i) * 2 + 1] = b;
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < COUT * 2; i += 256) {
      float a = 0.f;
      for (int wv = 0; wv < 8; ++wv) a += red[wv * COUT * 2 + i];
      partials[((size_t)n * P + p) * COUT * 2 + i] = a;
    }
  }
}

// stem weight gradient: G[n][0][tap][ci][co] += sum_{v in chunk} dz[v,co] * x[v+tap-1,ci]   (fp32 x, C_in <= 4)
// thread = (co, voxel lane); 27 register accumulators per input channel; block reduce, one atomicAdd per output.
constexpr int STEM_WG_CHUNK = 4096;
template <int COUT>
__global__ void __launch_bounds__(256, 2) stem_wgrad_kernel(const float* __restrict__ x, const bf16* __restrict__ dz, int D, int H, int W, int Cin,
                                                         float* __restrict__ G) {
  __shared__ float red[256 / COUT][27][COUT + 1];
  const int n = blockIdx.y;
  const long long vox = (long long)D * H * W;
  long long v0 = (long long)blockIdx.x * STEM_WG_CHUNK, v1 = v0 + STEM_WG_CHUNK;
  if (v1 > vox) v1 = vox;
  constexpr int VL = 256 / COUT;
  const int co = threadIdx.x % COUT, vl = threadIdx.x / COUT;
  const float* xn = x + (size_t)n * vox * Cin;
  const bf16* dn = dz + (size_t)n * vox * COUT;
  for (int ci = 0; ci < Cin; ++ci) {
    float acc[27];
#pragma unroll
    for (int t = 0; t < 27; ++t) acc[t] = 0.f;
    for (long long v = v0 + vl; v < v1; v += VL) {
      float g = __bfloat162float(dn[(size_t)v * COUT + co]);
      int xw = (int)(v % W);
      long long r = v / W;
      int xh = (int)(r % H), xd = (int)(r / H);
#pragma unroll
      for (int td = 0; td < 3; ++td) {
        int zd = xd + td - 1;
        bool okd = zd >= 0 && zd < D;
#pragma unroll
        for (int th = 0; th < 3; ++th) {
          int zh = xh + th - 1;
          bool okh = okd && zh >= 0 && zh < H;
#pragma unroll
          for (int tw = 0; tw < 3; ++tw) {
            int zw = xw + tw - 1;
            float xv = (okh && zw >= 0 && zw < W) ? __ldg(xn + (((size_t)zd * H + zh) * W + zw) * Cin + ci) : 0.f;
            acc[(td * 3 + th) * 3 + tw] = fmaf(g, xv, acc[(td * 3 + th) * 3 + tw]);
          }
        }
      }
    }
#pragma unroll
    for (int t = 0; t < 27; ++t) red[vl][t][co] = acc[t];
    __syncthreads();
    for (int i = threadIdx.x; i < 27 * COUT; i += 256) {
      int t = i / COUT, c = i % COUT;
      float a = 0.f;
      for (int l = 0; l < VL; ++l) a += red[l][t][c];
      atomicAdd(&G[(((size_t)n * 27 + t) * Cin + ci) * COUT + c], a);
    }
    __syncthreads();
  }
}

}  // namespace b200

using namespace b200;
#define ST(s) ((cudaStream_t)(s))

extern "C" {

int b200_conv3_direct_partials_count(int N, int D, int H, int W, int Cout) {
  (void)N;
  return ew_blocks((long long)D * H * W, Cout);
}

int b200_conv3_direct_fwd(const void* x, int x_is_f32, const void* wf, int n_w, const float* biascls, int n_b, const void* residual,
                          int act, float slope, int N, int D, int H, int W, int Cin, int Cout, void* y, int pmode, const void* aux,
                          float* partials, b200_stream_t s) {
  B200_CHECK_ARG(Cout % 8 == 0 && Cout <= 2048, "conv3_direct_fwd: Cout=%d must be a multiple of 8", Cout);
  int P = ew_blocks((long long)D * H * W, Cout);
  dim3 grid(P, N);
  size_t smem = EW_THREADS * 16 * sizeof(float);
  if (x_is_f32 && Cin <= 4 && (Cout == 8 || Cout == 16 || Cout == 32) && !residual && pmode != 2) {
    size_t sm2 = ((size_t)27 * Cin * Cout + 64 * Cout + 8 * Cout * 2) * sizeof(float);
    const float* xf = (const float*)x;
    if (Cout == 8)
      stem_conv_fwd_kernel<8><<<grid, 256, sm2, ST(s)>>>(xf, (const bf16*)wf, n_w, biascls, n_b, act, slope, D, H, W, Cin, P, (bf16*)y, pmode, partials);
    else if (Cout == 16)
      stem_conv_fwd_kernel<16><<<grid, 256, sm2, ST(s)>>>(xf, (const bf16*)wf, n_w, biascls, n_b, act, slope, D, H, W, Cin, P, (bf16*)y, pmode, partials);
    else
      stem_conv_fwd_kernel<32><<<grid, 256, sm2, ST(s)>>>(xf, (const bf16*)wf, n_w, biascls, n_b, act, slope, D, H, W, Cin, P, (bf16*)y, pmode, partials);
    B200_CHECK_LAUNCH("stem_conv_fwd");
    return 0;
  }
  if (x_is_f32)
    conv3_direct_fwd_kernel<float><<<grid, EW_THREADS, smem, ST(s)>>>((const float*)x, (const bf16*)wf, n_w, biascls, n_b,
                                                                     (const bf16*)residual, act, slope, D, H, W, Cin, Cout, P, (bf16*)y,
                                                                     pmode, (const bf16*)aux, partials);
  else
    conv3_direct_fwd_kernel<bf16><<<grid, EW_THREADS, smem, ST(s)>>>((const bf16*)x, (const bf16*)wf, n_w, biascls, n_b,
                                                                    (const bf16*)residual, act, slope, D, H, W, Cin, Cout, P, (bf16*)y,
                                                                    pmode, (const bf16*)aux, partials);
  B200_CHECK_LAUNCH("conv3_direct_fwd");
  return 0;
}

int b200_conv3_direct_wgrad(const void* x, int x_is_f32, const void* dz, int N, int D, int H, int W, int Cin, int Cout, float* G,
                            b200_stream_t s) {
  size_t bytes = (size_t)N * 27 * Cin * Cout * sizeof(float);
  cudaError_t e = cudaMemsetAsync(G, 0, bytes, ST(s));
  B200_CHECK_ARG(e == cudaSuccess, "conv3_direct_wgrad: memset failed: %s", cudaGetErrorString(e));
  long long vox = (long long)D * H * W;
  int total = 27 * Cin * Cout;
  if (x_is_f32 && Cin <= 4 && (Cout == 8 || Cout == 16 || Cout == 32)) {
    dim3 g2(ceil_div(vox, STEM_WG_CHUNK), N);
    if (Cout == 8) stem_wgrad_kernel<8><<<g2, 256, 0, ST(s)>>>((const float*)x, (const bf16*)dz, D, H, W, Cin, G);
    else if (Cout == 16) stem_wgrad_kernel<16><<<g2, 256, 0, ST(s)>>>((const float*)x, (const bf16*)dz, D, H, W, Cin, G);
    else stem_wgrad_kernel<32><<<g2, 256, 0, ST(s)>>>((const float*)x, (const bf16*)dz, D, H, W, Cin, G);
    B200_CHECK_LAUNCH("stem_wgrad");
    return 0;
  }
  dim3 grid(ceil_div(vox, WG_CHUNK), N, ceil_div(total, 256));
  if (x_is_f32)
    conv3_direct_wgrad_kernel<float><<<grid, 256, 0, ST(s)>>>((const float*)x, (const bf16*)dz, D, H, W, Cin, Cout, G);
  else
    conv3_direct_wgrad_kernel<bf16><<<grid, 256, 0, ST(s)>>>((const bf16*)x, (const bf16*)dz, D, H, W, Cin, Cout, G);
  B200_CHECK_LAUNCH("conv3_direct_wgrad");
  return 0;
}

}  // extern "C"
