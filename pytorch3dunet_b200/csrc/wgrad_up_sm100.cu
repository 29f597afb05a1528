// Weight gradient of conv3x3x3 o nearest-upsample(x2) with respect to the low-resolution half of a virtual concat, halo + stacked version.
//
//   Q[n][split][t][co][c1] = sum over the coarse voxels u of   dz[n, 2u + t - 1, co] * low[n, u, c1],     t in {0..3}^3
//
// (reference op: Decoder's InterpolateUpsampling + torch.cat + SingleConv weight gradient, pytorch3dunet/unet3d/buildingblocks.py
// :310-392, :480-507; the 64 offset blocks are folded to the 27 taps by b200_upcat_assemble_wgrad.)
//
// The tap-loop kernel (wgrad_igemm_sm100.cu, NTAPS = 64) reloads a 128-voxel dz tile per STACK of offsets: 512 KB of TMA traffic per
// 128 coarse voxels, L2-bandwidth bound (0.40 ms for 64 -> 32 @ 2x128^3).  Here, as in wgrad_hs_kernel, every offset is a VIEW:
//   * A (M side) = low, MN-major, 64-channel slice, box of 2 planes x 16 lines x 10 voxels; the M dimension stacks two w-shifted views
//     (atom stride = ONE row): rows (m, c1), m = 0,1;
//   * B (N side) = ONE FINE PLANE of dz sampled with element stride 2 along w (TMA element strides), 34 consecutive fine lines x 8
//     samples; the N dimension stacks the four h offsets as line-shifted views (atom stride = one fine line, K-group stride = TWO
//     fine lines because consecutive coarse lines are two fine lines apart): columns (t_h, co);
//   * per axis the offset is fixed by which lattice / plane of dz meets which view / plane of low:
//       w:  dz samples at fine w = 2(w0+j) + q,  low view shifted by m rows (box starts at w0-1):   t_w = q + 3 - 2m,  m in {q, q+1}
//       d:  dz fine plane 2 d0 + c,              low plane d0 - 1 + e:                              t_d = c + 3 - 2e,  e in {c, c+1}
//     so one CTA (fixed c; q = 0,1 or fixed) keeps 4 (2) accumulators [128 x 4*C_out] in TMEM and issues, per tile of 128 coarse
//     voxels, 32 (16) MMAs of M = 128, N = 4*C_out (128 / 256), K = 16 -- every row and column of every instruction is useful.
//   Every (coarse voxel, offset) pair of the volume is produced exactly once: pairs are enumerated by the dz SAMPLE (tile row), the
//   low voxel follows from the view shift; pairs whose low voxel or dz sample lies outside the volume are zero-filled by TMA.
// warp 0: TMA producer, warp 1: TMEM alloc + MMA issue (warp-converged, elected lane), warps 2..5: final read-out.
#include <stdlib.h>

#include "conv_common.cuh"

namespace b200 {

int make_act_tmap(CUtensorMap* tm, const void* ptr, int N, int D, int H, int W, int C, int kc, int bd, int bh, int bw);
int make_act_tmap_stride2_w(CUtensorMap* tm, const void* ptr, int N, int D, int H, int W, int C, int kc, int bh, int bw);

constexpr int WU_THREADS = 192;
constexpr int WU_BH = 16, WU_BW = 8;                 // coarse tile: 1 plane x 16 lines x 8 voxels
constexpr int WU_XW = WU_BW + 2;                     // low box: 2 planes x 16 lines x 10 voxels
constexpr int WU_XROWS = 2 * WU_BH * WU_XW;          // 320
constexpr int WU_ZH = 2 * WU_BH + 2;                 // dz box: 34 fine lines x 8 samples
constexpr int WU_ZROWS = WU_ZH * WU_BW;              // 272
constexpr int WU_CA = 64;                            // channels of low per CTA slice (one 128-byte swizzle row)
constexpr int WU_MAX_A = 4, WU_MAX_B = 6;

struct UpWgradParams {
  int N, d, h, w, C1, Cout, CoutTotal;
  int tilesH, tilesW, tiles, S;
  int nslices;   // C1 / 64
  int ncos;      // CoutTotal / Cout
  int NQ;        // dz w-lattices per CTA: 2 (4 accumulators) or 1 (2 accumulators)
  int a_stages, b_stages, a_bytes, b_bytes, tmem_cols;
  float* Q;
};

__global__ void __launch_bounds__(WU_THREADS, 1)
wgrad_up_kernel(const __grid_constant__ CUtensorMap tmapX, const __grid_constant__ CUtensorMap tmapZ, const UpWgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t a_full[WU_MAX_A], a_empty[WU_MAX_A], b_full[WU_MAX_B], b_empty[WU_MAX_B], done_bar;
  __shared__ uint32_t tmem_slot;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smemA = smem;
  uint8_t* smemB = smem + (size_t)p.a_stages * p.a_bytes;
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;  // provably warp-uniform
  constexpr int rbA = WU_CA * 2;
  const int rbB = p.Cout * 2;
  const int NC4 = 4 * p.Cout;  // columns of one accumulator

  const int split = blockIdx.x % p.S, n = blockIdx.x / p.S;
  const int slice = blockIdx.y;
  // blockIdx.z = ((cos * 2 + c) * (2 / NQ) + qg)
  int zz = blockIdx.z;
  const int qg = zz % (2 / p.NQ);
  zz /= (2 / p.NQ);
  const int c = zz & 1;
  const int co0 = (zz >> 1) * p.Cout;
  const int q0 = p.NQ == 2 ? 0 : qg;  // first w-lattice of this CTA

  if (threadIdx.x == 0) {
    for (int i = 0; i < p.a_stages; ++i) {
      mbar_init(&a_full[i], 1);
      mbar_init(&a_empty[i], 1);
    }
    for (int i = 0; i < p.b_stages; ++i) {
      mbar_init(&b_full[i], 1);
      mbar_init(&b_empty[i], 1);
    }
    mbar_init(&done_bar, 1);
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
  const int ntiles_mine = (p.tiles - split + p.S - 1) / p.S;

  if (warp == 0) {
    if (lane == 0) {
      int it = 0, ib = 0;
      for (int t = split; t < p.tiles; t += p.S, ++it) {
        const int tw_i = t % p.tilesW;
        const int r = t / p.tilesW;
        const int h0 = (r % p.tilesH) * WU_BH, d0 = r / p.tilesH, w0 = tw_i * WU_BW;
        const int as = it % p.a_stages;
        mbar_wait(&a_empty[as], ((uint32_t)(it / p.a_stages) & 1u) ^ 1u);
        mbar_arrive_expect_tx(&a_full[as], (uint32_t)(WU_XROWS * rbA));
        tma_load_5d(smemA + (size_t)as * p.a_bytes, &tmapX, &a_full[as], slice * WU_CA, w0 - 1, h0, d0 - 1 + c, n);
        for (int qi = 0; qi < p.NQ; ++qi, ++ib) {
          const int bs = ib % p.b_stages;
          mbar_wait(&b_empty[bs], ((uint32_t)(ib / p.b_stages) & 1u) ^ 1u);
          mbar_arrive_expect_tx(&b_full[bs], (uint32_t)(WU_ZROWS * rbB));
          tma_load_5d(smemB + (size_t)bs * p.b_bytes, &tmapZ, &b_full[bs], co0, 2 * w0 + q0 + qi, 2 * h0 - 1, 2 * d0 + c, n);
        }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = umma_idesc_bf16(128, NC4, 1, 1);
    // A: MN-major, atom stride (LBO) = ONE row -> atom m = view shifted by m voxels in w; K-group stride (SBO) = one low line (10 rows).
    // B: MN-major, atom stride (LBO) = ONE fine line (8 rows) -> atom i = view shifted by i fine lines; K-group stride = TWO fine lines.
    const uint64_t hiA = umma_smem_desc(0, 0, (uint32_t)(WU_XW * rbA), umma_layout_for_row_bytes(rbA)) & 0xFFFFFFFF00000000ull;
    const uint64_t hiB = umma_smem_desc(0, 0, (uint32_t)(16 * rbB), umma_layout_for_row_bytes(rbB)) & 0xFFFFFFFF00000000ull;
    const uint32_t lboA = ((uint32_t)rbA >> 4) << 16, lboB = (((uint32_t)(8 * rbB) >> 4) & 0x3FFFu) << 16;
    constexpr uint32_t A_LINE = (uint32_t)(WU_XW * rbA) >> 4;  // one low line, 16-byte units
    constexpr uint32_t A_ROW = (uint32_t)rbA >> 4;
    const uint32_t b_k = (uint32_t)(32 * rbB) >> 4;  // K16 = two coarse lines = four fine lines of dz
    int it = 0, ib = 0;
    for (int t = split; t < p.tiles; t += p.S, ++it) {
      const int as = it % p.a_stages;
      mbar_wait(&a_full[as], (uint32_t)(it / p.a_stages) & 1u);
      const uint32_t a_lo = ((smem_u32(smemA + (size_t)as * p.a_bytes) >> 4) & 0x3FFFu) | lboA;
      const uint32_t accum = it != 0 ? 1u : 0u;
      for (int qi = 0; qi < p.NQ; ++qi, ++ib) {
        const int bs = ib % p.b_stages;
        mbar_wait(&b_full[bs], (uint32_t)(ib / p.b_stages) & 1u);
        tc_fence_after();
        const uint32_t b_lo = ((smem_u32(smemB + (size_t)bs * p.b_bytes) >> 4) & 0x3FFFu) | lboB;
        const uint32_t a_q = a_lo + (uint32_t)(q0 + qi) * A_ROW;  // views m = q, q+1
#pragma unroll 1
        for (int ei = 0; ei < 2; ++ei) {
          const uint32_t a_g = a_q + (uint32_t)(ei * WU_BH) * A_LINE;
          const uint32_t tacc = tmem_base + (uint32_t)((qi * 2 + ei) * NC4);
#pragma unroll
          for (int k = 0; k < 8; ++k)
            umma_bf16_elect(tacc, hiA | (uint64_t)(a_g + (uint32_t)(2 * k) * A_LINE), hiB | (uint64_t)(b_lo + (uint32_t)k * b_k), idesc,
                            (k != 0) ? 1u : accum);
        }
        umma_commit_elect(&b_empty[bs]);
      }
      umma_commit_elect(&a_empty[as]);
    }
    umma_commit_elect(&done_bar);
  } else {
    // ================= final read-out (warps 2..5): row = (m, c1), columns (t_h, co) =================
    const int qd = warp & 3;
    const int row = qd * 32 + lane;
    const int m = row >> 6, c1 = slice * WU_CA + (row & 63);
    mbar_wait(&done_bar, 0);
    __syncwarp();
    tc_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(qd * 32) << 16);
    for (int qi = 0; qi < p.NQ; ++qi)
      for (int ei = 0; ei < 2; ++ei) {
        const int tw = 3 - (q0 + qi) - 2 * m;  // q + 3 - 2 (q + m)
        const int td = 3 - c - 2 * ei;         // c + 3 - 2 (c + ei)
        for (int th = 0; th < 4; ++th) {
          const int tap = (td * 4 + th) * 4 + tw;
          float* qrow = p.Q + ((((size_t)n * p.S + split) * 64 + tap) * p.CoutTotal + co0) * p.C1 + c1;
          for (int cc = 0; cc < p.Cout; cc += 16) {
            uint32_t raw[16];
            tmem_ld_32x32b_x16(taddr + (uint32_t)((qi * 2 + ei) * NC4 + th * p.Cout + cc), raw);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 16; ++e) qrow[(size_t)(cc + e) * p.C1] = ntiles_mine ? __uint_as_float(raw[e]) : 0.f;
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

// one plane of an activation sampled with element stride 2 along w only: box of bh lines x bw SAMPLES
int make_act_tmap_stride2_w(CUtensorMap* tm, const void* ptr, int N, int D, int H, int W, int C, int kc, int bh, int bw);

bool wgrad_up_plan(int N, int d, int h, int w, int Cout, int C1, UpWgradParams* pp) {
  UpWgradParams& p = *pp;
  memset(&p, 0, sizeof(p));
  const char* dis = getenv("B200UNET_UP_WGRAD_HS");
  if (dis && dis[0] == '0') return false;
  if (C1 % WU_CA != 0 || Cout % 16 != 0) return false;
  if (d < 2 || h < WU_BH + 1 || w < WU_XW) return false;  // every TMA box fits inside its tensor (2h >= 34)
  int nc = Cout;
  if (nc > 64) nc = 64;
  if (Cout % nc != 0 || !(nc == 16 || nc == 32 || nc == 64)) return false;
  p.N = N; p.d = d; p.h = h; p.w = w; p.C1 = C1; p.Cout = nc; p.CoutTotal = Cout;
  p.tilesH = (h + WU_BH - 1) / WU_BH;
  p.tilesW = (w + WU_BW - 1) / WU_BW;
  p.tiles = d * p.tilesH * p.tilesW;
  p.nslices = C1 / WU_CA;
  p.ncos = Cout / nc;
  p.NQ = (16 * nc <= 512) ? 2 : 1;
  p.a_bytes = (WU_XROWS * WU_CA * 2 + 1023) & ~1023;
  p.b_bytes = (WU_ZROWS * nc * 2 + 1023) & ~1023;
  p.a_stages = p.NQ == 2 ? 3 : 2;
  p.b_stages = (196 * 1024 - p.a_stages * p.a_bytes) / p.b_bytes;
  if (p.b_stages > WU_MAX_B) p.b_stages = WU_MAX_B;
  if (p.b_stages < 2) return false;
  int cols = 32;
  while (cols < p.NQ * 2 * 4 * nc) cols <<= 1;
  p.tmem_cols = cols;
  const int ctas_per_split = N * p.nslices * p.ncos * 2 * (2 / p.NQ);
  int want = sm_count() / ctas_per_split;
  if (const char* e = getenv("B200UNET_WGRAD_SPLITS")) {  // tests: few splits => many tiles accumulated per CTA
    const int v = atoi(e);
    if (v >= 1) want = v;
  }
  if (want < 1) want = 1;
  if (want > p.tiles) want = p.tiles;
  p.S = want;
  return true;
}

int wgrad_up_splits(int N, int d, int h, int w, int Cout, int C1) {
  UpWgradParams p;
  return wgrad_up_plan(N, d, h, w, Cout, C1, &p) ? p.S : 0;
}

int wgrad_up_run(const void* dz, const void* low, int N, int d, int h, int w, int Cout, int C1, float* Q, cudaStream_t s) {
  UpWgradParams p;
  if (!wgrad_up_plan(N, d, h, w, Cout, C1, &p)) return -1;
  p.Q = Q;
  CUtensorMap tmX, tmZ;
  int rc = make_act_tmap(&tmX, low, N, d, h, w, C1, WU_CA, 2, WU_BH, WU_XW);
  if (rc) return rc;
  rc = make_act_tmap_stride2_w(&tmZ, dz, N, 2 * d, 2 * h, 2 * w, Cout, p.Cout, WU_ZH, WU_BW);
  if (rc) return rc;
  size_t smem = (size_t)p.a_stages * p.a_bytes + (size_t)p.b_stages * p.b_bytes + 1024;
  cudaError_t e = cudaFuncSetAttribute(wgrad_up_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  B200_CHECK_ARG(e == cudaSuccess, "wgrad_up: cudaFuncSetAttribute(%zu) failed: %s", smem, cudaGetErrorString(e));
  dim3 grid((unsigned)(N * p.S), (unsigned)p.nslices, (unsigned)(p.ncos * 2 * (2 / p.NQ)));
  wgrad_up_kernel<<<grid, WU_THREADS, smem, s>>>(tmX, tmZ, p);
  B200_CHECK_LAUNCH("wgrad_up");
  return 0;
}

}  // namespace b200
