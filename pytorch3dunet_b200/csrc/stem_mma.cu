// The network's first convolution (C_in == 1, fp32 input volume) on the warp-level tensor-core path.
//
// Reference op: SingleConv 'conv' of the first encoder, nn.Conv3d(1, C_out, 3, padding=1) forward and its weight gradient
// (pytorch3dunet/unet3d/buildingblocks.py:36-113 create_conv; the GroupNorm in front of it is folded into per-sample weights and border
// bias classes by the engine, as for every other conv).
//
// With K = 27 the layer is 3.6 GFLOP against 150 MB of HBM traffic at 2x128^3 -> 16 channels: it is bound by the output write (fprop)
// and the dz read (wgrad), not by math, and a tcgen05 pipeline (TMA cannot build a 1-channel im2col tile: the innermost box would be
// 2 bytes) buys nothing.  What the CUDA-core kernels it replaces could not do is stay out of the way of that traffic: 432 FMA per
// voxel kept them at 0.18 ms (fprop) and 0.24 ms (wgrad).  Here the 27-tap dot product is ONE warp-level MMA chain
// (mma.sync m16n8k16, fp32 accumulate), fed from a shared-memory halo tile of x:
//   * x stays fp32-accurate: every value is split into hi + lo 16-bit halves (x = hi + lo to 2^-17) that occupy two adjacent K
//     slots of the A fragment and meet the same weight in B, so the result matches an fp32-x product to well below the 16-bit
//     rounding of the output;
//   * fprop: M = 16 consecutive voxels of a line, K = 32 taps x {hi,lo} = 64, N = C_out; epilogue = border-class bias, activation,
//     16-bit rounding, GroupNorm partial sums of the rounded output (deterministic: fixed lane / warp order), 8..16-byte stores;
//   * wgrad: M = 32 taps x {hi,lo}, N = C_out, K = voxels; B fragments come from the dz tile with ldmatrix.trans; every block
//     writes its OWN split slot of G (plain stores, no atomics), summed in a fixed order by b200_wgrad_finalize.
#include "common.cuh"
#include "ew.cuh"
#include "conv_common.cuh"

namespace b200 {

constexpr int SM_TD = 2, SM_TH = 8, SM_TW = 32, SM_VOX = SM_TD * SM_TH * SM_TW;  // voxel tile
constexpr int SM_HD = SM_TD + 2, SM_HH = SM_TH + 2, SM_HW = SM_TW + 2;           // halo tile
// halo pitches chosen so that the (up to two) tap rows one LDS touches fall in disjoint bank ranges: row pitch 44 (bank +12),
// plane pitch 452 (452 - 2*44 = 364 = bank +12 after the third row)
constexpr int SM_PITCH = 44, SM_PLANE = 452, SM_HALO = SM_HD * SM_PLANE;

#ifdef B200_ACT_F16
#define B200_MMA_TYPES "f16.f16"
#else
#define B200_MMA_TYPES "bf16.bf16"
#endif
__device__ __forceinline__ void mma16816(float c[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32." B200_MMA_TYPES ".f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// x = hi + lo: hi = the upper 16 bits of the fp32 pattern (exact in bf16), lo = 16-bit rounding of the remainder (error 2^-17 |x|).
// fp16 build: hi = fp16 rounding of x.
__device__ __forceinline__ float split_hi(float x) {
#ifdef B200_ACT_F16
  return bf16_round(x);
#else
  return __uint_as_float(__float_as_uint(x) & 0xffff0000u);
#endif
}
__device__ __forceinline__ uint32_t split_hi_lo(float x) {  // low half = hi, high half = lo
  const float hi = split_hi(x);
  return pack2(hi, x - hi);
}
__device__ __forceinline__ void split_pair(float v0, float v1, uint32_t& hi, uint32_t& lo) {  // hi = {hi(v0), hi(v1)}, lo likewise
  const float h0 = split_hi(v0), h1 = split_hi(v1);
  hi = pack2(h0, h1);
  lo = pack2(v0 - h0, v1 - h1);
}
__device__ __forceinline__ void unpack2(uint32_t p, float& a, float& b) {  // the two 16-bit values of p as floats
#ifdef B200_ACT_F16
  const float2 t = __half22float2(*reinterpret_cast<const __half2*>(&p));
  a = t.x;
  b = t.y;
#else
  a = __uint_as_float(p << 16);
  b = __uint_as_float(p & 0xffff0000u);
#endif
}
__device__ __forceinline__ int tap_offset(int tap) { return (tap / 9) * SM_PLANE + ((tap / 3) % 3) * SM_PITCH + tap % 3; }

// The fp32 halo of a tile (origin -1 on every axis, zero outside the volume) goes global -> registers -> shared memory, so that the
// loads of tile i+1 are in flight while tile i is computed.  Warp `warp` owns rows warp, warp+8, ... (5 of the 40), lane = column
// (lanes 0,1 also carry columns 32,33).
struct HaloRegs {
  float v[5], e[5];
};
__device__ __forceinline__ void halo_fetch(HaloRegs& r, const float* __restrict__ xn, int d0, int h0, int w0, int D, int H, int W, int warp,
                                           int lane) {
  const int gx = w0 + lane - 1, gx2 = w0 + 31 + lane;
  const bool okx = gx >= 0 && gx < W, okx2 = lane < SM_HW - 32 && gx2 < W;
#pragma unroll
  for (int i = 0; i < 5; ++i) {
    const int row = warp + 8 * i;
    const int hz = row / SM_HH, hy = row - hz * SM_HH;
    const int gz = d0 + hz - 1, gy = h0 + hy - 1;
    const bool ok = gz >= 0 && gz < D && gy >= 0 && gy < H;
    const float* src = xn + ((long long)gz * H + gy) * W;
    r.v[i] = (ok && okx) ? __ldg(src + gx) : 0.f;
    r.e[i] = (ok && okx2) ? __ldg(src + gx2) : 0.f;
  }
}
__device__ __forceinline__ void halo_commit(const HaloRegs& r, float* xs, int warp, int lane) {
#pragma unroll
  for (int i = 0; i < 5; ++i) {
    const int row = warp + 8 * i;
    const int hz = row / SM_HH, hy = row - hz * SM_HH;
    float* dst = xs + hz * SM_PLANE + hy * SM_PITCH;
    dst[lane] = r.v[i];
    if (lane < SM_HW - 32) dst[32 + lane] = r.e[i];
  }
}
__device__ __forceinline__ void tile_origin(int tile, int tW, int tH, int& d0, int& h0, int& w0) {
  w0 = (tile % tW) * SM_TW;
  const int r = tile / tW;
  h0 = (r % tH) * SM_TH;
  d0 = (r / tH) * SM_TD;
}

template <int COUT>
__global__ void __launch_bounds__(256, 2) stem_mma_fwd_kernel(const float* __restrict__ x, const bf16* __restrict__ wf, int n_w,
                                                            const float* __restrict__ biascls, int n_b, int act, float slope, int D, int H,
                                                            int W, int P, bf16* __restrict__ y, int pmode, float* __restrict__ partials) {
  constexpr int NT = COUT / 8;  // n8 tiles
  constexpr int CPT = 2 * NT;   // channels per thread (contiguous): channel = tig*CPT + t*2 + j
  __shared__ float xs[SM_HALO];
  __shared__ float bsm[64 * COUT];
  __shared__ float red[8 * COUT * 2];
  const int p = blockIdx.x, n = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tig = lane & 3;
  const bf16* wn = wf + (size_t)(n_w > 1 ? n : 0) * 27 * COUT;
  if (n_b)
    for (int i = threadIdx.x; i < 64 * COUT; i += 256) bsm[i] = biascls[(size_t)(n_b > 1 ? n : 0) * 64 * COUT + i];
  // B fragments (weights) for all four k-steps: k rows (tig*2, tig*2+1) = (hi, lo) slots of tap ks*8 + tig, rows +8 = tap ks*8+4+tig
  uint32_t bfr[4][NT][2];
  int offs[4][2];
#pragma unroll
  for (int ks = 0; ks < 4; ++ks)
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      const int tap = ks * 8 + hh * 4 + tig;
      const bool valid = tap < 27;
      offs[ks][hh] = valid ? tap_offset(tap) : 0;
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        const int ch = (g >> 1) * CPT + t * 2 + (g & 1);
        const uint32_t wb = valid ? (uint32_t) * reinterpret_cast<const unsigned short*>(wn + tap * COUT + ch) : 0u;
        bfr[ks][t][hh] = wb | (wb << 16);
      }
    }
  // relu / leaky / identity as max(z,0) + ns*min(z,0) (bit-identical to act_fwd); ELU takes the generic route
  const bool elu = act == B200_ACT_ELU;
  const float ns = act == B200_ACT_RELU ? 0.f : (act == B200_ACT_LEAKY ? slope : 1.f);
  const long long vox = (long long)D * H * W;
  const float* xn = x + (size_t)n * vox;
  bf16* yn = y + (size_t)n * vox * COUT + tig * CPT;
  const int tD = (D + SM_TD - 1) / SM_TD, tH = (H + SM_TH - 1) / SM_TH, tW = (W + SM_TW - 1) / SM_TW;
  int t0, t1;
  ew_range_i(tD * tH * tW, p, P, t0, t1);
  float s[CPT], q[CPT];
#pragma unroll
  for (int i = 0; i < CPT; ++i) s[i] = q[i] = 0.f;
  HaloRegs hr;
  int d0, h0, w0;
  if (t0 < t1) {
    tile_origin(t0, tW, tH, d0, h0, w0);
    halo_fetch(hr, xn, d0, h0, w0, D, H, W, warp, lane);
  }
  for (int tile = t0; tile < t1; ++tile) {
    __syncthreads();  // everyone is done reading the previous tile (and bsm is complete)
    halo_commit(hr, xs, warp, lane);
    __syncthreads();
    const int cd0 = d0, ch0 = h0, cw0 = w0;
    if (tile + 1 < t1) {
      tile_origin(tile + 1, tW, tH, d0, h0, w0);
      halo_fetch(hr, xn, d0, h0, w0, D, H, W, warp, lane);
    }
#pragma unroll 1
    for (int mi = 0; mi < 4; ++mi) {
      const int line = warp * 2 + (mi >> 1), mx = mi & 1;  // 32 m-tiles of 16 voxels: line (vz, vy), half mx
      const int vz = line >> 3, vy = line & 7;
      const int gz = cd0 + vz, gy = ch0 + vy, gx0 = cw0 + mx * 16 + g, gx1 = gx0 + 8;
      if (gz >= D || gy >= H || cw0 + mx * 16 >= W) continue;  // warp-uniform
      const float* xb = xs + vz * SM_PLANE + vy * SM_PITCH + mx * 16 + g;
      float acc[NT][4];
      if (n_b) {
        const int cdh = (axis_cls(gz, D) << 4) | (axis_cls(gy, H) << 2);
        const float* b0 = bsm + (cdh | axis_cls(gx0, W)) * COUT + tig * CPT;
        const float* b1 = bsm + (cdh | axis_cls(gx1, W)) * COUT + tig * CPT;
#pragma unroll
        for (int t = 0; t < NT; ++t) {
          const float2 u0 = *reinterpret_cast<const float2*>(b0 + t * 2), u1 = *reinterpret_cast<const float2*>(b1 + t * 2);
          acc[t][0] = u0.x;
          acc[t][1] = u0.y;
          acc[t][2] = u1.x;
          acc[t][3] = u1.y;
        }
      } else {
#pragma unroll
        for (int t = 0; t < NT; ++t) acc[t][0] = acc[t][1] = acc[t][2] = acc[t][3] = 0.f;
      }
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const uint32_t a0 = split_hi_lo(xb[offs[ks][0]]);
        const uint32_t a1 = split_hi_lo(xb[offs[ks][0] + 8]);
        const uint32_t a2 = split_hi_lo(xb[offs[ks][1]]);
        const uint32_t a3 = split_hi_lo(xb[offs[ks][1] + 8]);
#pragma unroll
        for (int t = 0; t < NT; ++t) mma16816(acc[t], a0, a1, a2, a3, bfr[ks][t][0], bfr[ks][t][1]);
      }
      // epilogue: activation, 16-bit rounding, statistics of the rounded values, one vector store per voxel
      if (elu) {
#pragma unroll
        for (int t = 0; t < NT; ++t)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[t][j] = acc[t][j] > 0.f ? acc[t][j] : expm1f(acc[t][j]);
      } else {
#pragma unroll
        for (int t = 0; t < NT; ++t)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[t][j] = fmaxf(acc[t][j], 0.f) + ns * fminf(acc[t][j], 0.f);
      }
      bf16* yrow = yn + (((long long)gz * H + gy) * W) * COUT;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int gx = half ? gx1 : gx0;
        uint32_t pk[NT];
#pragma unroll
        for (int t = 0; t < NT; ++t) pk[t] = pack2(acc[t][half * 2], acc[t][half * 2 + 1]);
        if (gx < W) {
#pragma unroll
          for (int t = 0; t < NT; ++t) {
            float r0, r1;
            unpack2(pk[t], r0, r1);
            s[t * 2] += r0;
            s[t * 2 + 1] += r1;
            q[t * 2] = fmaf(r0, r0, q[t * 2]);
            q[t * 2 + 1] = fmaf(r1, r1, q[t * 2 + 1]);
          }
          bf16* dst = yrow + (long long)gx * COUT;
          if (NT == 1) {
            *reinterpret_cast<uint32_t*>(dst) = pk[0];
          } else if (NT == 2) {
            *reinterpret_cast<uint2*>(dst) = make_uint2(pk[0], pk[1]);
          } else {
#pragma unroll
            for (int i = 0; i < NT / 4; ++i) *reinterpret_cast<uint4*>(dst + i * 8) = make_uint4(pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
          }
        }
      }
    }
  }
  if (pmode) {
#pragma unroll
    for (int i = 0; i < CPT; ++i) {
      float a = s[i], b = q[i];
      for (int o = 4; o < 32; o <<= 1) {  // over the 8 voxel rows g (lanes with the same tig)
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
      }
      if (g == 0) {
        red[(warp * COUT + tig * CPT + i) * 2] = a;
        red[(warp * COUT + tig * CPT + i) * 2 + 1] = b;
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

// ---- weight gradient ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void ldmatrix_x2_trans(uint32_t& r0, uint32_t& r1, const void* p) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(a));
}

template <int COUT>
__global__ void __launch_bounds__(256, (COUT <= 16 ? 2 : 1)) stem_mma_wgrad_kernel(const float* __restrict__ x, const bf16* __restrict__ dz, int D,
                                                                                  int H, int W, int tiles_per_block, float* __restrict__ G) {
  constexpr int NT = COUT / 8;
  constexpr int SWZ_SHIFT = NT == 4 ? 1 : (NT == 2 ? 2 : 3);  // chunk swizzle of the dz tile: phys = chunk ^ ((v >> SHIFT) & (NT-1))
  extern __shared__ __align__(16) unsigned char smraw[];
  float* xs = reinterpret_cast<float*>(smraw);                           // [SM_HALO]
  bf16* dzs = reinterpret_cast<bf16*>(smraw + SM_HALO * sizeof(float));  // [SM_VOX][COUT], 16-byte chunks swizzled; reused as red
  const int n = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tig = lane & 3;
  const int tD = (D + SM_TD - 1) / SM_TD, tH = (H + SM_TH - 1) / SM_TH, tW = (W + SM_TW - 1) / SM_TW;
  const int ntiles = tD * tH * tW;
  const int t_begin = blockIdx.x * tiles_per_block;
  const int t_end = min(ntiles, t_begin + tiles_per_block);
  const long long vox = (long long)D * H * W;
  const float* xn = x + (size_t)n * vox;
  const bf16* dzn = dz + (size_t)n * vox * COUT;
  // A rows owned by this thread: taps mtap*16 + g and +8 (mtap = 0,1).  Rows of taps >= 27 read a valid address and are never
  // written out.
  int off[2][2];
#pragma unroll
  for (int mtap = 0; mtap < 2; ++mtap)
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      const int tap = mtap * 16 + hh * 8 + g;
      off[mtap][hh] = tap < 27 ? tap_offset(tap) : 0;
    }
  float acc[4][NT][4];  // [hi taps 0-15, hi taps 16-31, lo taps 0-15, lo taps 16-31]
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int t = 0; t < NT; ++t) acc[a][t][0] = acc[a][t][1] = acc[a][t][2] = acc[a][t][3] = 0.f;

  // the dz tile: warp `warp` stages (and later consumes) lines 2*warp, 2*warp+1 of the 16; a line = 32 voxels = 32*NT 16-byte chunks,
  // lane l carries chunks l, l+32, ... of each
  HaloRegs hr;
  uint4 zr[2][NT];
  int d0, h0, w0;
  auto dz_fetch = [&](int fd0, int fh0, int fw0) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int line = warp * 2 + i;
      const int gz = fd0 + (line >> 3), gy = fh0 + (line & 7);
      const bool ok = gz < D && gy < H;
      const bf16* src = dzn + (((long long)gz * H + gy) * W + fw0) * COUT;
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        const int ci = lane + 32 * j;
        zr[i][j] = (ok && fw0 + ci / NT < W) ? __ldg(reinterpret_cast<const uint4*>(src) + ci) : make_uint4(0u, 0u, 0u, 0u);
      }
    }
  };
  auto dz_commit = [&]() {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int line = warp * 2 + i;
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        const int ci = lane + 32 * j;
        const int v = line * SM_TW + ci / NT, chunk = ci % NT;
        const int phys = chunk ^ ((v >> SWZ_SHIFT) & (NT - 1));
        *reinterpret_cast<uint4*>(dzs + v * COUT + phys * 8) = zr[i][j];
      }
    }
  };
  if (t_begin < t_end) {
    tile_origin(t_begin, tW, tH, d0, h0, w0);
    halo_fetch(hr, xn, d0, h0, w0, D, H, W, warp, lane);
    dz_fetch(d0, h0, w0);
  }
  for (int tile = t_begin; tile < t_end; ++tile) {
    __syncthreads();
    halo_commit(hr, xs, warp, lane);
    dz_commit();
    __syncthreads();
    const int cd0 = d0, ch0 = h0, cw0 = w0;
    if (tile + 1 < t_end) {
      tile_origin(tile + 1, tW, tH, d0, h0, w0);
      halo_fetch(hr, xn, d0, h0, w0, D, H, W, warp, lane);
      dz_fetch(d0, h0, w0);
    }
#pragma unroll 1
    for (int ki = 0; ki < 4; ++ki) {
      const int line = warp * 2 + (ki >> 1), kx = ki & 1;  // 32 k-steps of 16 voxels: line (vz, vy), half kx
      const int vz = line >> 3, vy = line & 7;
      if (cd0 + vz >= D || ch0 + vy >= H || cw0 + kx * 16 >= W) continue;  // warp-uniform; dz there is zero anyway
      const int v0 = line * SM_TW + kx * 16;  // first voxel of the k-step inside the tile
      const float* xb = xs + vz * SM_PLANE + vy * SM_PITCH + kx * 16 + tig * 2;
      // B fragments: n-tile t, k rows = voxels.  Lane l addresses row (l & 7) of matrix (l >> 3) & 1 (voxels +0 / +8).
      uint32_t bq[NT][2];
      {
        const int vrow = v0 + ((lane >> 3) & 1) * 8 + (lane & 7);
        const int sw = (vrow >> SWZ_SHIFT) & (NT - 1);
#pragma unroll
        for (int t = 0; t < NT; ++t) ldmatrix_x2_trans(bq[t][0], bq[t][1], dzs + vrow * COUT + ((t ^ sw) * 8));
      }
#pragma unroll
      for (int mtap = 0; mtap < 2; ++mtap) {
        uint32_t hi[4], lo[4];
        const float* pa = xb + off[mtap][0];
        const float* pb = xb + off[mtap][1];
        split_pair(pa[0], pa[1], hi[0], lo[0]);  // row g,   k = tig*2, +1
        split_pair(pb[0], pb[1], hi[1], lo[1]);  // row g+8, k = tig*2, +1
        split_pair(pa[8], pa[9], hi[2], lo[2]);  // row g,   k = tig*2+8, +9
        split_pair(pb[8], pb[9], hi[3], lo[3]);  // row g+8, k = tig*2+8, +9
#pragma unroll
        for (int t = 0; t < NT; ++t) {
          mma16816(acc[mtap][t], hi[0], hi[1], hi[2], hi[3], bq[t][0], bq[t][1]);
          mma16816(acc[2 + mtap][t], lo[0], lo[1], lo[2], lo[3], bq[t][0], bq[t][1]);
        }
      }
    }
  }
  // hi + lo, then a fixed-order sum over the 8 warps; the block's slot gets every [tap][co] entry (zeros if it had no tile)
  __syncthreads();
  float* red = reinterpret_cast<float*>(dzs);  // [8][32][COUT]
#pragma unroll
  for (int mtap = 0; mtap < 2; ++mtap)
#pragma unroll
    for (int t = 0; t < NT; ++t) {
      float* rp = red + ((size_t)warp * 32 + mtap * 16 + g) * COUT + t * 8 + tig * 2;
      rp[0] = acc[mtap][t][0] + acc[2 + mtap][t][0];
      rp[1] = acc[mtap][t][1] + acc[2 + mtap][t][1];
      rp[8 * COUT] = acc[mtap][t][2] + acc[2 + mtap][t][2];
      rp[8 * COUT + 1] = acc[mtap][t][3] + acc[2 + mtap][t][3];
    }
  __syncthreads();
  float* Gb = G + ((size_t)n * gridDim.x + blockIdx.x) * 27 * COUT;
  for (int i = threadIdx.x; i < 27 * COUT; i += 256) {
    float sum = 0.f;
    for (int wv = 0; wv < 8; ++wv) sum += red[(size_t)wv * 32 * COUT + i];
    Gb[i] = sum;
  }
}

bool stem_mma_supported(int Cin, int Cout) { return Cin == 1 && (Cout == 8 || Cout == 16 || Cout == 32); }

int stem_mma_fwd(const float* x, const bf16* wf, int n_w, const float* biascls, int n_b, int act, float slope, int N, int D, int H, int W,
                 int Cout, int P, bf16* y, int pmode, float* partials, cudaStream_t s) {
  dim3 grid(P, N);
  if (Cout == 8)
    stem_mma_fwd_kernel<8><<<grid, 256, 0, s>>>(x, wf, n_w, biascls, n_b, act, slope, D, H, W, P, y, pmode, partials);
  else if (Cout == 16)
    stem_mma_fwd_kernel<16><<<grid, 256, 0, s>>>(x, wf, n_w, biascls, n_b, act, slope, D, H, W, P, y, pmode, partials);
  else
    stem_mma_fwd_kernel<32><<<grid, 256, 0, s>>>(x, wf, n_w, biascls, n_b, act, slope, D, H, W, P, y, pmode, partials);
  B200_CHECK_LAUNCH("stem_mma_fwd");
  return 0;
}

void stem_mma_wgrad_grid(int N, int D, int H, int W, int* blocks, int* tiles_per_block) {
  const int ntiles = ceil_div(D, SM_TD) * ceil_div(H, SM_TH) * ceil_div(W, SM_TW);
  int want = 4 * sm_count() / (N > 0 ? N : 1);  // ~4 blocks per SM in total
  if (want < 1) want = 1;
  int tpb = ceil_div(ntiles, want);
  if (tpb < 1) tpb = 1;
  *blocks = ceil_div(ntiles, tpb);
  *tiles_per_block = tpb;
}

int stem_mma_wgrad(const float* x, const bf16* dz, int N, int D, int H, int W, int Cout, float* G, cudaStream_t s) {
  int blocks, tpb;
  stem_mma_wgrad_grid(N, D, H, W, &blocks, &tpb);
  dim3 grid(blocks, N);
  size_t red_bytes = (size_t)8 * 32 * Cout * sizeof(float), dz_bytes = (size_t)SM_VOX * Cout * sizeof(bf16);
  size_t smem = SM_HALO * sizeof(float) + (red_bytes > dz_bytes ? red_bytes : dz_bytes);
  if (Cout == 8) {
    stem_mma_wgrad_kernel<8><<<grid, 256, smem, s>>>(x, dz, D, H, W, tpb, G);
  } else if (Cout == 16) {
    stem_mma_wgrad_kernel<16><<<grid, 256, smem, s>>>(x, dz, D, H, W, tpb, G);
  } else {
    cudaFuncSetAttribute(stem_mma_wgrad_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    stem_mma_wgrad_kernel<32><<<grid, 256, smem, s>>>(x, dz, D, H, W, tpb, G);
  }
  B200_CHECK_LAUNCH("stem_mma_wgrad");
  return 0;
}

}  // namespace b200
