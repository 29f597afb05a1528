// Direct (CUDA-core) 3x3x3 convolution kernels.
//   * the production path for the network stem (C_in < 16: 25 FLOP/B, HBM-bound, no tensor-core shape) and
//   * the general on-GPU fallback / on-device checker for shapes the tcgen05 kernels do not take
//     (odd channel counts, tiny spatial sizes).  Same operands and epilogue contract as the tcgen05 kernel.
#include <cstdlib>
#include "common.cuh"
#include "ew.cuh"
#include "conv_common.cuh"

namespace b200 {

template <typename T>
__device__ __forceinline__ float ldf(const T* p);
template <>
__device__ __forceinline__ float ldf<float>(const float* p) {
  return *p;
}
template <>
__device__ __forceinline__ float ldf<bf16>(const bf16* p) {
  return from_act(*p);
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
              for (int i = 0; i < 8; ++i) acc[i] += xv * from_act(wp[(size_t)i * Cin + ci]);
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
  long long v0 = (long long)blockIdx.z * WG_CHUNK, v1 = v0 + WG_CHUNK;
  if (v1 > vox) v1 = vox;
  int total = 27 * Cin * Cout;
  int o = blockIdx.x * blockDim.x + threadIdx.x;
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
      acc += from_act(dn[(size_t)v * Cout + co]) * ldf<InT>(xn + (((size_t)zd * H + zh) * W + zw) * Cin + ci);
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
    wsm[i] = from_act(wn[((size_t)tap * COUT + co) * Cin + ci]);
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

// register-tiled variant for C_in == 1, W % VW == 0: one thread = VW consecutive voxels of a line x all COUT channels.  The 3x3x(VW+2)
// halo of x lives in registers (loaded once: 9*(VW+2) loads for VW voxels instead of 27 per voxel), every weight row read from shared
// memory feeds VW*COUT FMAs: the loop is FMA-issue bound instead of load-latency bound (1 -> 16 @ 2x128^3: 0.38 -> ~0.1 ms).
template <int COUT, int VW>
__global__ void __launch_bounds__(256, 1) stem_conv_fwd_tiled_kernel(const float* __restrict__ x, const bf16* __restrict__ wf, int n_w,
                                                                  const float* __restrict__ biascls, int n_b, int act, float slope, int D,
                                                                  int H, int W, int P, bf16* __restrict__ y, int pmode,
                                                                  float* __restrict__ partials) {
  extern __shared__ float sm[];  // w[27][COUT] | bias[64][COUT] | red[8][COUT*2]
  float* wsm = sm;
  float* bsm = wsm + 27 * COUT;
  float* red = bsm + 64 * COUT;
  const int p = blockIdx.x, n = blockIdx.y;
  const bf16* wn = wf + (size_t)(n_w > 1 ? n : 0) * 27 * COUT;
  for (int i = threadIdx.x; i < 27 * COUT; i += 256) wsm[i] = from_act(wn[i]);  // [tap][co] (C_in == 1)
  for (int i = threadIdx.x; i < 64 * COUT; i += 256) bsm[i] = n_b ? biascls[(size_t)(n_b > 1 ? n : 0) * 64 * COUT + i] : 0.f;
  __syncthreads();
  const long long vox = (long long)D * H * W;
  const long long quads = vox / VW;
  long long q0, q1;
  ew_range(quads, p, P, q0, q1);
  const float* xn = x + (size_t)n * vox;
  const int WQ = W / VW;
  float s[COUT], q[COUT];
#pragma unroll
  for (int i = 0; i < COUT; ++i) s[i] = q[i] = 0.f;
  for (long long qi = q0 + threadIdx.x; qi < q1; qi += 256) {
    const int xw0 = (int)(qi % WQ) * VW;
    const long long r = qi / WQ;
    const int xh = (int)(r % H), xd = (int)(r / H);
    float xr[3][3][VW + 2];
#pragma unroll
    for (int td = 0; td < 3; ++td) {
      const int zd = xd + td - 1;
#pragma unroll
      for (int th = 0; th < 3; ++th) {
        const int zh = xh + th - 1;
        const bool ok = zd >= 0 && zd < D && zh >= 0 && zh < H;
        const float* xp = xn + ((size_t)(ok ? zd : 0) * H + (ok ? zh : 0)) * W + xw0;
#pragma unroll
        for (int j = 0; j < VW + 2; ++j) {
          const int zw = xw0 + j - 1;
          xr[td][th][j] = (ok && zw >= 0 && zw < W) ? __ldg(xp + j - 1) : 0.f;
        }
      }
    }
    float acc[VW][COUT];
    const int cdh = (axis_cls(xd, D) << 4) | (axis_cls(xh, H) << 2);
#pragma unroll
    for (int v = 0; v < VW; ++v) {
      const float* bp = bsm + (cdh | axis_cls(xw0 + v, W)) * COUT;
#pragma unroll
      for (int i = 0; i < COUT; ++i) acc[v][i] = bp[i];
    }
#pragma unroll
    for (int td = 0; td < 3; ++td)
#pragma unroll
      for (int th = 0; th < 3; ++th)
#pragma unroll
        for (int tw = 0; tw < 3; ++tw) {
          const float4* wp = reinterpret_cast<const float4*>(wsm + ((td * 3 + th) * 3 + tw) * COUT);
#pragma unroll
          for (int i4 = 0; i4 < COUT / 4; ++i4) {
            const float4 w4 = wp[i4];
#pragma unroll
            for (int v = 0; v < VW; ++v) {
              const float xv = xr[td][th][tw + v];
              acc[v][4 * i4] = fmaf(xv, w4.x, acc[v][4 * i4]);
              acc[v][4 * i4 + 1] = fmaf(xv, w4.y, acc[v][4 * i4 + 1]);
              acc[v][4 * i4 + 2] = fmaf(xv, w4.z, acc[v][4 * i4 + 2]);
              acc[v][4 * i4 + 3] = fmaf(xv, w4.w, acc[v][4 * i4 + 3]);
            }
          }
        }
    bf16x8* op = reinterpret_cast<bf16x8*>(y + ((size_t)n * vox + qi * VW) * COUT);
#pragma unroll
    for (int v = 0; v < VW; ++v) {
#pragma unroll
      for (int i = 0; i < COUT; ++i) {
        acc[v][i] = bf16_round(act_fwd(acc[v][i], act, slope));
        s[i] += acc[v][i];
        q[i] += acc[v][i] * acc[v][i];
      }
#pragma unroll
      for (int i = 0; i < COUT / 8; ++i) op[v * (COUT / 8) + i] = pack8(&acc[v][8 * i]);
    }
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

// stem weight gradient: G[n][0][tap][ci][co] += sum_v dz[v,co] * x[v+tap-1,ci]   (fp32 x, C_in <= 4).
// Shared-memory tiled: a block walks tiles of 2x8x32 voxels; per tile the fp32 x halo (4x10x34) and the dz tile (as fp32)
// are staged in shared memory; a thread owns one (dd,dh) tap row x 8 output channels = 24 accumulators (3 dw taps) and
// strides over the tile's voxels: 2 LDS.128 + 3 LDS per 24 FMA.  One block-level reduction at the very end into the block's own split slot
// (no atomics: the weight gradient is bit-reproducible).
constexpr int SW_TD = 2, SW_TH = 8, SW_TW = 32, SW_VOX = SW_TD * SW_TH * SW_TW;
constexpr int SW_HD = SW_TD + 2, SW_HH = SW_TH + 2, SW_HW = SW_TW + 2, SW_HALO = SW_HD * SW_HH * SW_HW;
template <int COUT>
__global__ void __launch_bounds__(256, 2) stem_wgrad_kernel(const float* __restrict__ x, const bf16* __restrict__ dz, int D, int H, int W, int Cin,
                                                            int tiles_per_block, float* __restrict__ G) {
  extern __shared__ float sm[];
  float* xs = sm;                 // [SW_HALO]
  float* dzs = sm + SW_HALO;  // [SW_VOX][COUT]
  static_assert(SW_HALO % 4 == 0, "dz tile must stay 16-byte aligned");
  constexpr int GROUPS = 9 * (COUT / 8);
  constexpr int LANES = 256 / GROUPS;
  const int n = blockIdx.y;
  const int tD = (D + SW_TD - 1) / SW_TD, tH = (H + SW_TH - 1) / SW_TH, tW = (W + SW_TW - 1) / SW_TW;
  const int ntiles = tD * tH * tW;
  const int t_begin = blockIdx.x * tiles_per_block;
  const int t_end = min(ntiles, t_begin + tiles_per_block);
  const long long vox = (long long)D * H * W;
  const int grp = threadIdx.x % GROUPS, lane = threadIdx.x / GROUPS;
  const int oct = grp % (COUT / 8), trow = grp / (COUT / 8);  // output-channel octet, (dd,dh) tap row
  const int tdd = trow / 3, tdh = trow % 3;
  const bool active = lane < LANES;
  for (int ci = 0; ci < Cin; ++ci) {
    float acc[3][8];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[a][j] = 0.f;
    for (int t = t_begin; t < t_end; ++t) {
      const int w0 = (t % tW) * SW_TW;
      const int r = t / tW;
      const int h0 = (r % tH) * SW_TH, d0 = (r / tH) * SW_TD;
      __syncthreads();
      for (int i = threadIdx.x; i < SW_HALO; i += 256) {
        int hx = i % SW_HW, rr = i / SW_HW, hy = rr % SW_HH, hz = rr / SW_HH;
        int gz = d0 + hz - 1, gy = h0 + hy - 1, gx = w0 + hx - 1;
        float v = 0.f;
        if (gz >= 0 && gz < D && gy >= 0 && gy < H && gx >= 0 && gx < W)
          v = __ldg(x + ((size_t)n * vox + ((size_t)gz * H + gy) * W + gx) * Cin + ci);
        xs[i] = v;
      }
      for (int i = threadIdx.x; i < SW_VOX * (COUT / 8); i += 256) {
        int o8 = i % (COUT / 8), v = i / (COUT / 8);
        int vx = v % SW_TW, vy = (v / SW_TW) % SW_TH, vz = v / (SW_TW * SW_TH);
        int gz = d0 + vz, gy = h0 + vy, gx = w0 + vx;
        float f[8];
        if (gz < D && gy < H && gx < W) {
          unpack8(*reinterpret_cast<const bf16x8*>(dz + ((size_t)n * vox + ((size_t)gz * H + gy) * W + gx) * COUT + o8 * 8), f);
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] = 0.f;
        }
        float4* dst = reinterpret_cast<float4*>(dzs + (size_t)v * COUT + o8 * 8);
        dst[0] = make_float4(f[0], f[1], f[2], f[3]);
        dst[1] = make_float4(f[4], f[5], f[6], f[7]);
      }
      __syncthreads();
      if (active) {
        for (int v = lane; v < SW_VOX; v += LANES) {
          const int vx = v % SW_TW, vy = (v / SW_TW) % SW_TH, vz = v / (SW_TW * SW_TH);
          const float* xp = xs + ((vz + tdd) * SW_HH + (vy + tdh)) * SW_HW + vx;
          const float x0 = xp[0], x1 = xp[1], x2 = xp[2];
          const float4* dp = reinterpret_cast<const float4*>(dzs + (size_t)v * COUT + oct * 8);
          const float4 g0 = dp[0], g1 = dp[1];
          const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            acc[0][j] = fmaf(x0, g[j], acc[0][j]);
            acc[1][j] = fmaf(x1, g[j], acc[1][j]);
            acc[2][j] = fmaf(x2, g[j], acc[2][j]);
          }
        }
      }
    }
    // block reduction over the voxel lanes, then one atomic per output
    __syncthreads();
    float* red = dzs;  // [LANES][GROUPS][24]
    if (active) {
      float* rp = red + ((size_t)lane * GROUPS + grp) * 24;
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int j = 0; j < 8; ++j) rp[a * 8 + j] = acc[a][j];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < GROUPS * 24; i += 256) {
      float sum = 0.f;
      for (int l = 0; l < LANES; ++l) sum += red[(size_t)l * GROUPS * 24 + i];
      const int g = i / 24, a = (i % 24) / 8, j = i % 8;
      const int o8 = g % (COUT / 8), tr = g / (COUT / 8);
      const int tap = tr * 3 + a;
      // one split slot per block (G [N][S = gridDim.x][27][Cin][COUT]): plain stores, summed in a fixed order by b200_wgrad_finalize
      G[((((size_t)n * gridDim.x + blockIdx.x) * 27 + tap) * Cin + ci) * COUT + o8 * 8 + j] = sum;
    }
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
    if (stem_mma_supported(Cin, Cout) && !getenv("B200UNET_STEM_FMA"))  // warp-level MMA, x split into hi + lo halves
      return stem_mma_fwd(xf, (const bf16*)wf, n_w, biascls, n_b, act, slope, N, D, H, W, Cout, P, (bf16*)y, pmode, partials, ST(s));
    if (Cin == 1 && W % 4 == 0 && (Cout == 8 || Cout == 16)) {  // register-tiled: 4 voxels per thread
      if (Cout == 8)
        stem_conv_fwd_tiled_kernel<8, 4><<<grid, 256, sm2, ST(s)>>>(xf, (const bf16*)wf, n_w, biascls, n_b, act, slope, D, H, W, P, (bf16*)y, pmode, partials);
      else
        stem_conv_fwd_tiled_kernel<16, 4><<<grid, 256, sm2, ST(s)>>>(xf, (const bf16*)wf, n_w, biascls, n_b, act, slope, D, H, W, P, (bf16*)y, pmode, partials);
      B200_CHECK_LAUNCH("stem_conv_fwd_tiled");
      return 0;
    }
    if (Cin == 1 && W % 2 == 0 && Cout == 32) {
      stem_conv_fwd_tiled_kernel<32, 2><<<grid, 256, sm2, ST(s)>>>(xf, (const bf16*)wf, n_w, biascls, n_b, act, slope, D, H, W, P, (bf16*)y, pmode, partials);
      B200_CHECK_LAUNCH("stem_conv_fwd_tiled");
      return 0;
    }
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

static bool stem_wgrad_applies(int x_is_f32, int Cin, int Cout) { return x_is_f32 && Cin <= 4 && (Cout == 8 || Cout == 16 || Cout == 32); }
static void stem_wgrad_grid(int N, int D, int H, int W, int* blocks, int* tpb_out) {
  int ntiles = ceil_div(D, SW_TD) * ceil_div(H, SW_TH) * ceil_div(W, SW_TW);
  const int sms4 = 4 * sm_count();
  int tpb = ceil_div(ntiles, sms4 / (N > 0 ? N : 1) > 0 ? sms4 / N : 1);  // ~4 blocks per SM in total
  if (tpb < 1) tpb = 1;
  *blocks = ceil_div(ntiles, tpb);
  *tpb_out = tpb;
}
// split slots of G the direct wgrad writes: one per block of the fp32-input stem kernel, else 1
int b200_conv3_direct_wgrad_splits(int N, int D, int H, int W, int Cin, int Cout, int x_is_f32) {
  if (!stem_wgrad_applies(x_is_f32, Cin, Cout)) return 1;
  int blocks, tpb;
  if (stem_mma_supported(Cin, Cout) && !getenv("B200UNET_STEM_FMA"))
    stem_mma_wgrad_grid(N, D, H, W, &blocks, &tpb);
  else
    stem_wgrad_grid(N, D, H, W, &blocks, &tpb);
  return blocks;
}

int b200_conv3_direct_wgrad(const void* x, int x_is_f32, const void* dz, int N, int D, int H, int W, int Cin, int Cout, float* G,
                            b200_stream_t s) {
  long long vox = (long long)D * H * W;
  int total = 27 * Cin * Cout;
  if (x_is_f32 && stem_mma_supported(Cin, Cout) && !getenv("B200UNET_STEM_FMA"))
    return stem_mma_wgrad((const float*)x, (const bf16*)dz, N, D, H, W, Cout, G, ST(s));
  if (stem_wgrad_applies(x_is_f32, Cin, Cout)) {
    int blocks, tpb;
    stem_wgrad_grid(N, D, H, W, &blocks, &tpb);
    dim3 g2(blocks, N);
    size_t red_floats = (size_t)(256 / (9 * (Cout / 8))) * 9 * (Cout / 8) * 24;
    size_t dz_floats = (size_t)SW_VOX * Cout;
    size_t sm2 = (SW_HALO + (dz_floats > red_floats ? dz_floats : red_floats)) * sizeof(float);
    if (Cout == 8) {
      cudaFuncSetAttribute(stem_wgrad_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2);
      stem_wgrad_kernel<8><<<g2, 256, sm2, ST(s)>>>((const float*)x, (const bf16*)dz, D, H, W, Cin, tpb, G);
    } else if (Cout == 16) {
      cudaFuncSetAttribute(stem_wgrad_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2);
      stem_wgrad_kernel<16><<<g2, 256, sm2, ST(s)>>>((const float*)x, (const bf16*)dz, D, H, W, Cin, tpb, G);
    } else {
      cudaFuncSetAttribute(stem_wgrad_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2);
      stem_wgrad_kernel<32><<<g2, 256, sm2, ST(s)>>>((const float*)x, (const bf16*)dz, D, H, W, Cin, tpb, G);
    }
    B200_CHECK_LAUNCH("stem_wgrad");
    return 0;
  }
  size_t bytes = (size_t)N * 27 * Cin * Cout * sizeof(float);
  cudaError_t e = cudaMemsetAsync(G, 0, bytes, ST(s));
  B200_CHECK_ARG(e == cudaSuccess, "conv3_direct_wgrad: memset failed: %s", cudaGetErrorString(e));
  dim3 grid(ceil_div(total, 256), N, ceil_div(vox, WG_CHUNK));  // outputs on x (no 65535 limit), voxel chunks on z
  B200_CHECK_ARG(grid.z <= 65535, "conv3_direct_wgrad: volume too large for the fallback kernel (%lld voxels)", vox);
  if (x_is_f32)
    conv3_direct_wgrad_kernel<float><<<grid, 256, 0, ST(s)>>>((const float*)x, (const bf16*)dz, D, H, W, Cin, Cout, G);
  else
    conv3_direct_wgrad_kernel<bf16><<<grid, 256, 0, ST(s)>>>((const bf16*)x, (const bf16*)dz, D, H, W, Cin, Cout, G);
  B200_CHECK_LAUNCH("conv3_direct_wgrad");
  return 0;
}

}  // extern "C"
