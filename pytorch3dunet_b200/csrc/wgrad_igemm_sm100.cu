// Weight gradient of the 3x3x3 convolution on the sm_100a tensor cores.
//
//   G[n][split][tap][ci][co] = sum over the voxels v of the split of  x[n, v+tap-1, ci] * dz[n, v, co]
//
// i.e. 27 GEMMs  D_tap[ci, co] = X_tap^T [ci, voxels] * dZ [voxels, co]  that share the dZ operand; the reduction
// (GEMM-K) dimension is the voxel index.  Both operands are read exactly as they sit in HBM (NDHWC: one row per
// voxel, channels contiguous), so they are "MN-major" UMMA operands: TMA box loads -> swizzled smem tiles
// [128 voxels][AW channels], tcgen05.mma with a_major = b_major = MN, K = 16 voxels per instruction.
// The zero padding of x is again the TMA out-of-bounds fill.  The GroupNorm scale/shift that the forward pass
// folded into the weights is undone analytically in b200_wgrad_finalize (a[n,ci] * G + b[n,ci] * T).
//
// One CTA: (voxel split, sample, group of TG taps, 128-channel slice of C_in).  TMEM holds TG accumulators of
// [128 x C_out] fp32 (TG * C_out <= 512 columns).  Warp 0 = TMA producer, warp 1 = MMA issuer, warps 2..5 = epilogue.
#include <string.h>

#include "common.cuh"
#include "sm100_ptx.cuh"

namespace b200 {

int make_act_tmap(CUtensorMap* tm, const void* ptr, int N, int D, int H, int W, int C, int kc, int bd, int bh, int bw);
int make_act_tmap_stride2(CUtensorMap* tm, const void* ptr, int N, int D, int H, int W, int C, int kc, int bd, int bh, int bw);
int choose_box(int D, int H, int W, int* bd, int* bh, int* bw);
int wgrad_halo_splits(int N, int D, int H, int W, int Cin, int Cout);
int wgrad_halo_run(const void* x, const void* dz, int N, int D, int H, int W, int Cin, int Cout, int co0, int CoutTotal, float* G, cudaStream_t s);
int wgrad_up_splits(int N, int d, int h, int w, int Cout, int C1);
int wgrad_up_run(const void* dz, const void* low, int N, int d, int h, int w, int Cout, int C1, float* Q, cudaStream_t s);

constexpr int WG_THREADS = 192;
constexpr int WG_MAX_A_STAGES = 8;
constexpr int WG_B_STAGES = 2;
constexpr int WG_A_FULL_BYTES = 128 * 128 * 2;  // an M=128 MMA reads up to 128 voxels x 128 channels past a stage base

struct WgradParams {
  int N, D, H, W, Cin, Cout;
  int BD, BH, BW, tilesD, tilesH, tilesW, tiles;
  int S, tiles_per_split;
  int TG, ngroups, mchunks;
  int AWa, AWb;  // channels per smem atom tile (64/32/16) on the x side and the dz side
  int a_stages, a_stage_bytes, b_stage_bytes;
  int tmem_cols;
  int NTAPS;           // 27; 1 for the 1x1x1 conv; 64 for conv3 o nearest-upsample (G is [n][S][NTAPS][Cin][Cout])
  signed char toff[64 * 3];  // tap t reads the x-operand at a_mul * (tile origin) + toff[t]
  int a_mul;
  // C_in <= 64: `tps` = 128 / C_in taps are stacked along M of ONE MMA (their x tiles sit side by side as M atoms), NST = stacks
  int tps, NST;
  int co0, CoutTotal;  // this launch covers output channels [co0, co0 + Cout) of CoutTotal (C_out > 256 is processed in slices)
  float* G;
};

__global__ void __launch_bounds__(WG_THREADS)
conv3_wgrad_igemm_kernel(const __grid_constant__ CUtensorMap tmapX, const __grid_constant__ CUtensorMap tmapZ, const WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t a_full[WG_MAX_A_STAGES], a_empty[WG_MAX_A_STAGES];
  __shared__ __align__(8) uint64_t b_full[WG_B_STAGES], b_empty[WG_B_STAGES];
  __shared__ __align__(8) uint64_t tmem_full_bar;
  __shared__ uint32_t tmem_slot;

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smemB = smem;                                           // WG_B_STAGES * b_stage_bytes
  uint8_t* smemA = smem + (size_t)WG_B_STAGES * p.b_stage_bytes;   // a_stages * WG_A_STAGE_BYTES
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;  // provably warp-uniform

  const int split = blockIdx.x % p.S;
  const int n = blockIdx.x / p.S;
  const int grp = blockIdx.y;
  const int m0 = blockIdx.z * 128;
  const int st0 = grp * p.TG;                  // first tap stack of this group
  const int ntaps = min(p.TG, p.NST - st0);    // tap stacks (= accumulators) of this CTA
  const int t0 = split * p.tiles_per_split;
  const int t1 = min(p.tiles, t0 + p.tiles_per_split);
  const int m_real = min(128, p.Cin - m0);
  const int apt = m_real / p.AWa;  // M atoms per tap
  const int natoms_b = p.Cout / p.AWb;
  const int a_atom_bytes = 128 * p.AWa * 2, b_atom_bytes = 128 * p.AWb * 2;

  if (threadIdx.x == 0) {
    for (int i = 0; i < p.a_stages; ++i) {
      mbar_init(&a_full[i], 1);
      mbar_init(&a_empty[i], 1);
    }
    for (int i = 0; i < WG_B_STAGES; ++i) {
      mbar_init(&b_full[i], 1);
      mbar_init(&b_empty[i], 1);
    }
    mbar_init(&tmem_full_bar, 1);
    fence_mbar_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmapX);
    tma_prefetch_desc(&tmapZ);
  }
  if (warp == 1) tmem_alloc(&tmem_slot, (uint32_t)p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int ai = 0;  // running A-stage counter
      for (int t = t0, bi = 0; t < t1; ++t, ++bi) {
        int tt = t;
        const int w0 = (tt % p.tilesW) * p.BW;
        tt /= p.tilesW;
        const int h0 = (tt % p.tilesH) * p.BH;
        const int d0 = (tt / p.tilesH) * p.BD;
        {
          const int bs = bi % WG_B_STAGES;
          mbar_wait(&b_empty[bs], ((uint32_t)(bi / WG_B_STAGES) & 1u) ^ 1u);
          mbar_arrive_expect_tx(&b_full[bs], (uint32_t)(natoms_b * b_atom_bytes));
          for (int j = 0; j < natoms_b; ++j)
            tma_load_5d(smemB + (size_t)bs * p.b_stage_bytes + (size_t)j * b_atom_bytes, &tmapZ, &b_full[bs], p.co0 + j * p.AWb, w0, h0, d0, n);
        }
        for (int tp = 0; tp < ntaps; ++tp, ++ai) {
          const int tapb = (st0 + tp) * p.tps;
          const int rt = min(p.tps, p.NTAPS - tapb);  // taps really present in this stack (the last one may be ragged)
          const int as = ai % p.a_stages;
          mbar_wait(&a_empty[as], ((uint32_t)(ai / p.a_stages) & 1u) ^ 1u);
          mbar_arrive_expect_tx(&a_full[as], (uint32_t)(rt * apt * a_atom_bytes));
          for (int jt = 0; jt < rt; ++jt) {
            const int tap = tapb + jt;
            const int od = p.toff[3 * tap], oh = p.toff[3 * tap + 1], ow = p.toff[3 * tap + 2];
            for (int j = 0; j < apt; ++j)
              tma_load_5d(smemA + (size_t)as * p.a_stage_bytes + (size_t)(jt * apt + j) * a_atom_bytes, &tmapX, &a_full[as],
                          m0 + j * p.AWa, p.a_mul * w0 + ow, p.a_mul * h0 + oh, p.a_mul * d0 + od, n);
          }
        }
      }
    }
  } else if (warp == 1) {
    // whole warp converged, one elected lane issues; descriptors = constant high word + low word advanced by adds (the loop
    // must stay under the ~54-cycle dispatch time of a tcgen05.mma)
    {
      const uint32_t idesc = umma_idesc_bf16(128, p.Cout, 1, 1);
      const int rba = p.AWa * 2, rbb = p.AWb * 2;
      const uint64_t hiA = umma_smem_desc(0, 16u, (uint32_t)(8 * rba), umma_layout_for_row_bytes(rba)) & 0xFFFFFFFF00000000ull;
      const uint64_t hiB = umma_smem_desc(0, 16u, (uint32_t)(8 * rbb), umma_layout_for_row_bytes(rbb)) & 0xFFFFFFFF00000000ull;
      const uint32_t lboA = (((uint32_t)a_atom_bytes >> 4) & 0x3FFFu) << 16, lboB = (((uint32_t)b_atom_bytes >> 4) & 0x3FFFu) << 16;
      const uint32_t ka = (uint32_t)rba, kb = (uint32_t)rbb;  // 16 voxel rows of K, in 16-byte units
      int ai = 0;
      uint32_t accum = 0;
      for (int t = t0, bi = 0; t < t1; ++t, ++bi) {
        const int bs = bi % WG_B_STAGES;
        mbar_wait(&b_full[bs], (uint32_t)(bi / WG_B_STAGES) & 1u);
        const uint32_t b_lo = ((smem_u32(smemB + (size_t)bs * p.b_stage_bytes) >> 4) & 0x3FFFu) | lboB;
        for (int tp = 0; tp < ntaps; ++tp, ++ai) {
          const int as = ai % p.a_stages;
          mbar_wait(&a_full[as], (uint32_t)(ai / p.a_stages) & 1u);
          tc_fence_after();
          const uint32_t a_lo = ((smem_u32(smemA + (size_t)as * p.a_stage_bytes) >> 4) & 0x3FFFu) | lboA;
          const uint32_t tacc = tmem_base + (uint32_t)(tp * p.Cout);
          umma_bf16_elect(tacc, hiA | (uint64_t)a_lo, hiB | (uint64_t)b_lo, idesc, accum);
#pragma unroll
          for (uint32_t k = 1; k < 8; ++k)  // 128 voxels = 8 x K16
            umma_bf16_elect(tacc, hiA | (uint64_t)(a_lo + k * ka), hiB | (uint64_t)(b_lo + k * kb), idesc, 1u);
          umma_commit_elect(&a_empty[as]);
        }
        accum = 1u;
        umma_commit_elect(&b_empty[bs]);
      }
      umma_commit_elect(&tmem_full_bar);
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int jt = row / m_real;            // which tap of the stack this accumulator row belongs to
    const int ci = m0 + row - jt * m_real;
    mbar_wait(&tmem_full_bar, 0);
    __syncwarp();
    tc_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    const bool have_work = t1 > t0;
    for (int tp = 0; tp < ntaps; ++tp) {
      const int tap = (st0 + tp) * p.tps + jt;
      const bool valid = jt < p.tps && tap < p.NTAPS;
      float* grow = p.G + ((((size_t)n * p.S + split) * p.NTAPS + (valid ? tap : 0)) * p.Cin + (valid ? ci : 0)) * p.CoutTotal + p.co0;
      for (int c0 = 0; c0 < p.Cout; c0 += 16) {
        uint32_t raw[16];
        tmem_ld_32x32b_x16(taddr + (uint32_t)(tp * p.Cout + c0), raw);
        tmem_ld_wait();
        if (valid) {
          float4* o = reinterpret_cast<float4*>(grow + c0);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float4 f;
            f.x = have_work ? __uint_as_float(raw[4 * i]) : 0.f;
            f.y = have_work ? __uint_as_float(raw[4 * i + 1]) : 0.f;
            f.z = have_work ? __uint_as_float(raw[4 * i + 2]) : 0.f;
            f.w = have_work ? __uint_as_float(raw[4 * i + 3]) : 0.f;
            o[i] = f;
          }
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

static int atom_width(int C) { return (C % 64 == 0) ? 64 : (C % 32 == 0 ? 32 : 16); }

static int cout_slice(int Cout) {  // C_out per launch (N dimension of the MMA <= 256)
  if (Cout <= 256) return Cout;
  return Cout % 256 == 0 ? 256 : (Cout % 128 == 0 ? 128 : (Cout % 64 == 0 ? 64 : (Cout % 32 == 0 ? 32 : 16)));
}
static bool wgrad_supported(int N, int D, int H, int W, int Cin, int Cout) {
  (void)N;
  int bd, bh, bw;
  if (Cin % 16 != 0 || Cout % 16 != 0) return false;
  if (choose_box(D, H, W, &bd, &bh, &bw)) return false;
  return true;
}

// NT = 27: 3x3x3 taps at offsets -1..1;  64: 4x4x4 at -1..2 over a stride-2 operand;  1: no shift;
// NT_DECONV: 3x3x3 at offsets 0..2 over a stride-2 operand (transposed-conv weight gradient)
constexpr int NT_DECONV = -27;
static void wgrad_plan(int N, int D, int H, int W, int Cin, int Cout, WgradParams& p, int NT = 27) {
  memset(&p, 0, sizeof(p));
  p.N = N; p.D = D; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout;
  const bool deconv = NT == NT_DECONV;
  if (deconv) NT = 27;
  p.NTAPS = NT;
  p.a_mul = (NT == 64 || deconv) ? 2 : 1;
  for (int t = 0; t < NT; ++t) {
    if (deconv) {
      p.toff[3 * t] = (signed char)(t / 9);
      p.toff[3 * t + 1] = (signed char)((t / 3) % 3);
      p.toff[3 * t + 2] = (signed char)(t % 3);
    } else if (NT == 27) {
      p.toff[3 * t] = (signed char)(t / 9 - 1);
      p.toff[3 * t + 1] = (signed char)((t / 3) % 3 - 1);
      p.toff[3 * t + 2] = (signed char)(t % 3 - 1);
    } else if (NT == 64) {
      p.toff[3 * t] = (signed char)((t >> 4) - 1);
      p.toff[3 * t + 1] = (signed char)(((t >> 2) & 3) - 1);
      p.toff[3 * t + 2] = (signed char)((t & 3) - 1);
    } else {
      p.toff[3 * t] = p.toff[3 * t + 1] = p.toff[3 * t + 2] = 0;
    }
  }
  choose_box(D, H, W, &p.BD, &p.BH, &p.BW);
  p.tilesD = (D + p.BD - 1) / p.BD;
  p.tilesH = (H + p.BH - 1) / p.BH;
  p.tilesW = (W + p.BW - 1) / p.BW;
  p.tiles = p.tilesD * p.tilesH * p.tilesW;
  p.tps = (Cin <= 64 && NT > 1) ? 128 / Cin : 1;
  if (p.tps > NT) p.tps = NT;
  p.NST = (NT + p.tps - 1) / p.tps;
  int tgmax = 512 / Cout;
  if (tgmax > p.NST) tgmax = p.NST;
  p.ngroups = (p.NST + tgmax - 1) / tgmax;
  p.TG = (p.NST + p.ngroups - 1) / p.ngroups;
  p.ngroups = (p.NST + p.TG - 1) / p.TG;
  p.mchunks = (Cin + 127) / 128;
  // one CTA per SM (TMEM: TG*Cout columns, smem: deep A pipeline) -> size the split count for a single wave
  int ctas_per_split = N * p.ngroups * p.mchunks;
  int want = sm_count() / ctas_per_split;
  if (want < 1) want = 1;
  if (want > p.tiles) want = p.tiles;
  p.tiles_per_split = (p.tiles + want - 1) / want;
  p.S = (p.tiles + p.tiles_per_split - 1) / p.tiles_per_split;
  p.AWa = atom_width(Cin);
  p.AWb = atom_width(Cout);
  p.b_stage_bytes = 128 * Cout * 2;
  // an A stage only holds the channels that exist (garbage M rows of the MMA read past it, into the next stage or the tail pad)
  p.a_stage_bytes = 128 * (p.tps * Cin < 128 ? p.tps * Cin : 128) * 2;
  int budget = 190 * 1024 - WG_B_STAGES * p.b_stage_bytes - WG_A_FULL_BYTES;
  p.a_stages = budget / p.a_stage_bytes;
  if (p.a_stages > WG_MAX_A_STAGES) p.a_stages = WG_MAX_A_STAGES;
  if (p.a_stages < 2) p.a_stages = 2;
  int cols = 32;
  while (cols < p.TG * Cout) cols <<= 1;
  p.tmem_cols = cols;
}

static int wgrad_plain_launch(const void* x, const void* dz, int N, int D, int H, int W, int Cin, int Cout, float* G, int NT,
                              cudaStream_t s) {
  const int cs = cout_slice(Cout);
  WgradParams p;
  wgrad_plan(N, D, H, W, Cin, cs, p, NT);
  p.G = G;
  p.CoutTotal = Cout;
  CUtensorMap tmX, tmZ;
  int rc = p.a_mul == 2 ? make_act_tmap_stride2(&tmX, x, N, 2 * D, 2 * H, 2 * W, Cin, p.AWa, p.BD, p.BH, p.BW)
                        : make_act_tmap(&tmX, x, N, D, H, W, Cin, p.AWa, p.BD, p.BH, p.BW);
  if (rc) return rc;
  rc = make_act_tmap(&tmZ, dz, N, D, H, W, Cout, p.AWb, p.BD, p.BH, p.BW);
  if (rc) return rc;
  size_t smem = (size_t)WG_B_STAGES * p.b_stage_bytes + (size_t)p.a_stages * p.a_stage_bytes + WG_A_FULL_BYTES + 1024;
  cudaError_t e = cudaFuncSetAttribute(conv3_wgrad_igemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  B200_CHECK_ARG(e == cudaSuccess, "conv3_wgrad_igemm: cudaFuncSetAttribute(%zu) failed: %s", smem, cudaGetErrorString(e));
  dim3 grid((unsigned)(N * p.S), (unsigned)p.ngroups, (unsigned)p.mchunks);
  for (int co0 = 0; co0 < Cout; co0 += cs) {
    p.co0 = co0;
    conv3_wgrad_igemm_kernel<<<grid, WG_THREADS, smem, s>>>(tmX, tmZ, p);
    B200_CHECK_LAUNCH("conv3_wgrad_igemm");
  }
  return 0;
}
}  // namespace b200

using namespace b200;

extern "C" {

int b200_conv3_wgrad_igemm_supported(int N, int D, int H, int W, int Cin, int Cout) {
  return wgrad_supported(N, D, H, W, Cin, Cout) ? 1 : 0;
}

int b200_conv3_wgrad_igemm_splits(int N, int D, int H, int W, int Cin, int Cout) {
  if (!wgrad_supported(N, D, H, W, Cin, Cout)) return 0;
  const int cs = cout_slice(Cout);
  int hs = wgrad_halo_splits(N, D, H, W, Cin, cs);
  if (hs > 0) return hs;
  WgradParams p;
  wgrad_plan(N, D, H, W, Cin, cs, p);
  return p.S;
}

int b200_conv3_wgrad_igemm(const void* x, const void* dz, int N, int D, int H, int W, int Cin, int Cout, float* G, b200_stream_t s) {
  B200_CHECK_ARG(wgrad_supported(N, D, H, W, Cin, Cout), "conv3_wgrad_igemm: unsupported shape N=%d D=%d H=%d W=%d Cin=%d Cout=%d", N,
                 D, H, W, Cin, Cout);
  const int cs = cout_slice(Cout);
  if (wgrad_halo_splits(N, D, H, W, Cin, cs) > 0) {
    for (int co0 = 0; co0 < Cout; co0 += cs) {
      int rc = wgrad_halo_run(x, dz, N, D, H, W, Cin, cs, co0, Cout, G, (cudaStream_t)s);
      if (rc) return rc;
    }
    return 0;
  }
  return wgrad_plain_launch(x, dz, N, D, H, W, Cin, Cout, G, 27, (cudaStream_t)s);
}

// ---- weight gradient of the 1x1x1 conv: G[n][split][ci][co] = sum_v x[n,v,ci] * dy[n,v,co] (flat voxel list, no shift)
int b200_pointwise_tc_wgrad_splits(int N, long long vox, int Cin, int Cout) {
  if (Cin % 16 != 0 || Cout % 16 != 0 || vox < 1 || vox >= (1ll << 31)) return 0;
  WgradParams p;
  wgrad_plan(N, 1, 1, (int)vox, Cin, cout_slice(Cout), p, 1);
  return p.S;
}
int b200_pointwise_tc_wgrad(const void* x, const void* dy, int N, long long vox, int Cin, int Cout, float* G, b200_stream_t s) {
  B200_CHECK_ARG(b200_pointwise_tc_wgrad_splits(N, vox, Cin, Cout) > 0, "pointwise_tc_wgrad: unsupported N=%d vox=%lld Cin=%d Cout=%d", N,
                 vox, Cin, Cout);
  return wgrad_plain_launch(x, dy, N, 1, 1, (int)vox, Cin, Cout, G, 1, (cudaStream_t)s);
}

// ---- weight gradient of conv3 o nearest-upsample w.r.t. the 64 (offset e) x channel blocks:
//   Q[n][split][e][co][c1] = sum_u dz[n, 2u+e, co] * b[n, u, c1],  e in {-1..2}^3
// (dW[t] = sum over the parities r of Q[r - t], assembled by b200_upcat_assemble_wgrad).  The strided, shifted operand is dz.
int b200_conv3_up_wgrad_splits(int N, int d, int h, int w, int Cout, int C1) {
  if (const int S = wgrad_up_splits(N, d, h, w, Cout, C1)) return S;  // halo + stacked-offset kernel (wgrad_up_sm100.cu)
  if (!wgrad_supported(N, d, h, w, Cout, C1)) return 0;
  WgradParams p;
  wgrad_plan(N, d, h, w, Cout, cout_slice(C1), p, 64);
  return p.S;
}
// ---- weight gradient of the transposed conv by phases: Q[n][split][e][co][ci] = sum_u gp[n, 2u+e, co] * x[n, u, ci], e in {0,1,2}^3
int b200_deconv_phase_wgrad_splits(int N, int d, int h, int w, int Cout, int Cin) {
  if (!wgrad_supported(N, d, h, w, Cout, Cin)) return 0;
  WgradParams p;
  wgrad_plan(N, d, h, w, Cout, cout_slice(Cin), p, NT_DECONV);
  return p.S;
}
int b200_deconv_phase_wgrad(const void* gp, const void* x, int N, int d, int h, int w, int Cout, int Cin, float* Q, b200_stream_t s) {
  B200_CHECK_ARG(b200_deconv_phase_wgrad_splits(N, d, h, w, Cout, Cin) > 0, "deconv_phase_wgrad: unsupported N=%d %dx%dx%d Cout=%d Cin=%d", N, d, h,
                 w, Cout, Cin);
  return wgrad_plain_launch(gp, x, N, d, h, w, Cout, Cin, Q, NT_DECONV, (cudaStream_t)s);
}
int b200_conv3_up_wgrad(const void* dz, const void* b, int N, int d, int h, int w, int Cout, int C1, float* Q, b200_stream_t s) {
  B200_CHECK_ARG(b200_conv3_up_wgrad_splits(N, d, h, w, Cout, C1) > 0, "conv3_up_wgrad: unsupported N=%d %dx%dx%d Cout=%d C1=%d", N, d, h, w,
                 Cout, C1);
  if (wgrad_up_splits(N, d, h, w, Cout, C1)) return wgrad_up_run(dz, b, N, d, h, w, Cout, C1, Q, (cudaStream_t)s);
  return wgrad_plain_launch(dz, b, N, d, h, w, Cout, C1, Q, 64, (cudaStream_t)s);
}

}  // extern "C"
