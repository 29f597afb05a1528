// Weight gradient of the 3x3x3 convolution, halo + stacked-tap version (large volumes, D>=3, H>=18, W>=10).
//
//   G[n][split][tap][ci][co] = sum over the voxels v of the split of  x[n, v+tap-1, ci] * dz[n, v, co]
//
// The plain kernel (wgrad_igemm_sm100.cu) issues, per 128-voxel tile and per TAP, a TMA load of the shifted x tile and
// 8 MMAs that use only C_in of the 128 M rows -- for the C_in <= 32 full-resolution layers that is TMA-latency bound and
// wastes 3/4 of every instruction.  Here, per tile (1x16x8 voxels) and per 32-channel (16-channel) slice of C_in:
//   * ONE TMA box load of the (3 x 18 x 10)-voxel halo of x; every tap is a row-shifted view of it (as in the fprop halo
//     kernel; the views work because the swizzle is a function of the shared-memory address);
//   * the M dimension of each MMA stacks the three dw taps: the MN-major A operand is described with a leading-dimension
//     byte offset (atom stride) of ONE ROW, so atom j is the same view shifted by j voxels in w.  One instruction
//     (M=128, N=C_out, K=16 voxels) therefore produces D[(dw, ci), co] for dw = 0,1,2 (+ ignored garbage for dw = 3);
//     72 MMAs per tile instead of 216;
//   * 9 accumulators (one per (dd,dh)) of [128 x C_out] fp32 live in TMEM for the whole CTA lifetime (C_out <= 56);
//     for wider C_out the (dd,dh) pairs are split over blockIdx.z.
// Both operands are read as they sit in HBM (NDHWC rows): MN-major A (x) and B (dz); the reduction dim is the voxel index.
// warp 0: TMA producer, warp 1: TMEM alloc + MMA issue (warp-converged, elected lane), warps 2..5: final read-out.
#include <stdlib.h>

#include "conv_common.cuh"

namespace b200 {

constexpr int WH_BH = 16, WH_BW = 8, WH_HD = 3, WH_HH = WH_BH + 2, WH_HW = WH_BW + 2;
constexpr int WH_ROWS = WH_HD * WH_HH * WH_HW;  // 540
constexpr int WH_THREADS = 224;  // warp 0 TMA producer, warps 1 and 6 MMA issuers, warps 2..5 read-out
constexpr int WH_ISSUERS = 2;
constexpr int WH_MAX_A = 6, WH_MAX_B = 4;

struct WgradHaloParams {
  int N, D, H, W, Cin, Cout;
  int tilesH, tilesW, tiles;
  int S;        // voxel splits per sample (tiles are dealt round-robin to the splits)
  int CA;       // channels of x per CTA slice (32 or 16)
  int nslices;  // Cin / CA
  int PG;       // (dd,dh) pairs per CTA: 9, 3 (one dd) or 1
  int AWb;      // dz atom width (channels)
  int a_stages, b_stages, a_bytes, b_bytes;
  int tmem_cols;
  int co0, CoutTotal;  // slice [co0, co0 + Cout) of a wider C_out
  float* G;
  int hs;       // 1: h-stacked variant (wgrad_hs_kernel): the three dh taps stacked along N; PG then counts dd planes per CTA (3 or 1)
};

template <int CA>
__global__ void __launch_bounds__(WH_THREADS, 1)
wgrad_halo_kernel(const __grid_constant__ CUtensorMap tmapX, const __grid_constant__ CUtensorMap tmapZ, const WgradHaloParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t a_full[WH_MAX_A], a_empty[WH_MAX_A], b_full[WH_MAX_B], b_empty[WH_MAX_B], done_bar;
  __shared__ uint32_t tmem_slot;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smemA = smem;
  uint8_t* smemB = smem + (size_t)p.a_stages * p.a_bytes;
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;  // provably warp-uniform
  constexpr int rbA = CA * 2;

  const int split = blockIdx.x % p.S, n = blockIdx.x / p.S;
  const int slice = blockIdx.y;          // which CA-channel slice of C_in
  const int pg0 = blockIdx.z * p.PG;     // first (dd,dh) pair of this CTA
  const int natoms_b = p.Cout / p.AWb;
  const int b_atom_bytes = 128 * p.AWb * 2;

  if (threadIdx.x == 0) {
    for (int i = 0; i < p.a_stages; ++i) {
      mbar_init(&a_full[i], 1);
      mbar_init(&a_empty[i], WH_ISSUERS);  // every issuer commits each stage it has read
    }
    for (int i = 0; i < p.b_stages; ++i) {
      mbar_init(&b_full[i], 1);
      mbar_init(&b_empty[i], WH_ISSUERS);
    }
    mbar_init(&done_bar, WH_ISSUERS);
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
      int it = 0;
      for (int t = split; t < p.tiles; t += p.S, ++it) {
        const int tw_i = t % p.tilesW;
        const int r = t / p.tilesW;
        const int h0 = (r % p.tilesH) * WH_BH, d0 = r / p.tilesH, w0 = tw_i * WH_BW;
        const int bs = it % p.b_stages, as = it % p.a_stages;
        mbar_wait(&b_empty[bs], ((uint32_t)(it / p.b_stages) & 1u) ^ 1u);
        mbar_arrive_expect_tx(&b_full[bs], (uint32_t)(natoms_b * b_atom_bytes));
        for (int j = 0; j < natoms_b; ++j)
          tma_load_5d(smemB + (size_t)bs * p.b_bytes + (size_t)j * b_atom_bytes, &tmapZ, &b_full[bs], p.co0 + j * p.AWb, w0, h0, d0, n);
        mbar_wait(&a_empty[as], ((uint32_t)(it / p.a_stages) & 1u) ^ 1u);
        mbar_arrive_expect_tx(&a_full[as], (uint32_t)(WH_ROWS * rbA));
        tma_load_5d(smemA + (size_t)as * p.a_bytes, &tmapX, &a_full[as], slice * CA, w0 - 1, h0 - 1, d0 - 1, n);
      }
    }
  } else if (warp == 1 || warp == 6) {
    // two MMA issuers (warps 1 and 6) share every tile: each owns half of the (dd,dh) accumulators.  One warp can issue an
    // N <= 64 MMA every ~54 cycles, two warps interleave to ~40 (profiles/probes_r01.md, probe 4).  Both wait on the same
    // full barriers in lockstep (the producer cannot refill a stage before BOTH have committed it) -> no phase aliasing.
    // Whole warp converged; one elected lane issues (see sm100_ptx.cuh)
    const int issuer = warp == 1 ? 0 : 1;
    const int g_half = (p.PG + 1) / 2;
    const int g_begin = issuer * g_half, g_end = issuer == 0 ? g_half : p.PG;
    const int rbB = p.AWb * 2;
    const uint32_t idesc = umma_idesc_bf16(128, p.Cout, 1, 1);
    // A: MN-major, atom stride (LBO) = ONE halo row -> atom j = view shifted by j voxels in w; 8-row K group stride (SBO) = one
    // halo line (10 rows).  B: MN-major, atom stride = one [128 x AWb] tile, K group stride = 8 rows.
    const uint64_t hiA = umma_smem_desc(0, 0, (uint32_t)(WH_HW * rbA), umma_layout_for_row_bytes(rbA)) & 0xFFFFFFFF00000000ull;
    const uint64_t hiB = umma_smem_desc(0, 0, (uint32_t)(8 * rbB), umma_layout_for_row_bytes(rbB)) & 0xFFFFFFFF00000000ull;
    const uint32_t lboA = ((uint32_t)rbA >> 4) << 16, lboB = (((uint32_t)b_atom_bytes >> 4) & 0x3FFFu) << 16;
    constexpr uint32_t A_LINE = (uint32_t)(WH_HW * rbA) >> 4;  // one halo line, 16-byte units
    const uint32_t b_k = (uint32_t)(16 * rbB) >> 4;           // 16 voxel rows of dz
    int it = 0;
    for (int t = split; t < p.tiles; t += p.S, ++it) {
      const int bs = it % p.b_stages, as = it % p.a_stages;
      mbar_wait(&b_full[bs], (uint32_t)(it / p.b_stages) & 1u);
      mbar_wait(&a_full[as], (uint32_t)(it / p.a_stages) & 1u);
      tc_fence_after();
      const uint32_t a_lo = ((smem_u32(smemA + (size_t)as * p.a_bytes) >> 4) & 0x3FFFu) | lboA;
      const uint32_t b_lo = ((smem_u32(smemB + (size_t)bs * p.b_bytes) >> 4) & 0x3FFFu) | lboB;
      const uint32_t accum = it != 0 ? 1u : 0u;
#pragma unroll 1
      for (int g = g_begin; g < g_end; ++g) {
        const int pair = pg0 + g;  // dd*3 + dh
        const uint32_t a_g = a_lo + (uint32_t)((pair / 3) * WH_HH + pair % 3) * A_LINE;
        const uint32_t tacc = tmem_base + (uint32_t)(g * p.Cout);
#pragma unroll
        for (int k = 0; k < 8; ++k)  // 128 voxels = 8 x K16 = lines (2k, 2k+1) of the tile
          umma_bf16_elect(tacc, hiA | (uint64_t)(a_g + (uint32_t)(2 * k) * A_LINE), hiB | (uint64_t)(b_lo + (uint32_t)k * b_k), idesc,
                          (k != 0) ? 1u : accum);
      }
      umma_commit_elect(&a_empty[as]);
      umma_commit_elect(&b_empty[bs]);
    }
    umma_commit_elect(&done_bar);
  } else {
    // ================= final read-out (warps 2..5) =================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int dw = row / CA, ci = slice * CA + row % CA;
    const bool valid = dw < 3;
    mbar_wait(&done_bar, 0);
    __syncwarp();
    tc_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    for (int g = 0; g < p.PG; ++g) {
      const int tap = (pg0 + g) * 3 + (valid ? dw : 0);
      float* grow = p.G + ((((size_t)n * p.S + split) * 27 + tap) * p.Cin + ci) * p.CoutTotal + p.co0;
      for (int c0 = 0; c0 < p.Cout; c0 += 16) {
        uint32_t raw[16];
        tmem_ld_32x32b_x16(taddr + (uint32_t)(g * p.Cout + c0), raw);
        tmem_ld_wait();
        if (valid) {
          float4* o = reinterpret_cast<float4*>(grow + c0);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float4 f;
            f.x = ntiles_mine ? __uint_as_float(raw[4 * i]) : 0.f;
            f.y = ntiles_mine ? __uint_as_float(raw[4 * i + 1]) : 0.f;
            f.z = ntiles_mine ? __uint_as_float(raw[4 * i + 2]) : 0.f;
            f.w = ntiles_mine ? __uint_as_float(raw[4 * i + 3]) : 0.f;
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

// ------------------------------------------------------------------------------------------------------------------------------
// h-stacked variant.  The instruction above (M = 4 dw-atoms x C_in slice, N = C_out, K = 16 voxels) fetches (128 + C_out) * 32 bytes
// of operands per 2 * 128 * C_out * 16 flops: at C_out = 32 the tensor pipe waits for shared memory 60 % of the time (an MMA is
// paced by max(N/2, 32 + N/4) cycles, see conv_zs_sm100.cu).  Here the N dimension additionally stacks the three dh taps: the B
// operand (dz, MN-major) is described with an atom stride of ONE TILE LINE (8 voxels), so atom i is the dz tile shifted by i-1 lines
// in h (the tile is loaded with one halo line above and below; lines outside the volume are zero-filled by TMA):
//     D[(dw, ci), (i, co)] = sum_{u in tile} x[u + (dd, 0, dw)] * dz[u + (0, i-1, 0)]   =  G[(dd, dh = 1 - i, dw)][ci][co]
// (substitute v = u + (0, i-1, 0): every (v, tap) pair of the volume is covered exactly once over all tiles, the pairs whose dz
// line falls outside the volume are zero).  x needs no h halo any more: per tile ONE x box of 3 x 16 x 10 voxels and ONE dz box of
// 1 x 18 x 8; 3 accumulators [128 x 3*C_out] (one per dd); 24 instructions of N = 3*C_out per tile instead of 72 of N = C_out.
// Needs C_out = one swizzle row of dz (16 / 32 / 64 channels).  warp 0: TMA producer, warp 1: TMEM alloc + MMA issue, warps 2..5: read-out.
constexpr int WS_THREADS = 192;
constexpr int WS_XH = WH_BH, WS_XW = WH_BW + 2;             // x box: 3 planes x 16 lines x 10 voxels
constexpr int WS_XROWS = WH_HD * WS_XH * WS_XW;             // 480
constexpr int WS_ZROWS = (WH_BH + 2) * WH_BW;               // dz box: 18 lines x 8 voxels = 144 rows

template <int CA>
__global__ void __launch_bounds__(WS_THREADS, 3)
wgrad_hs_kernel(const __grid_constant__ CUtensorMap tmapX, const __grid_constant__ CUtensorMap tmapZ, const WgradHaloParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t a_full[WH_MAX_A], a_empty[WH_MAX_A], b_full[WH_MAX_B], b_empty[WH_MAX_B], done_bar;
  __shared__ uint32_t tmem_slot;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smemA = smem;
  uint8_t* smemB = smem + (size_t)p.a_stages * p.a_bytes;
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;  // provably warp-uniform
  constexpr int rbA = CA * 2;
  const int rbB = p.Cout * 2;

  const int split = blockIdx.x % p.S, n = blockIdx.x / p.S;
  const int slice = blockIdx.y;          // which CA-channel slice of C_in
  const int dd0 = blockIdx.z * p.PG;     // first depth tap of this CTA

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
      int it = 0;
      for (int t = split; t < p.tiles; t += p.S, ++it) {
        const int tw_i = t % p.tilesW;
        const int r = t / p.tilesW;
        const int h0 = (r % p.tilesH) * WH_BH, d0 = r / p.tilesH, w0 = tw_i * WH_BW;
        const int bs = it % p.b_stages, as = it % p.a_stages;
        mbar_wait(&b_empty[bs], ((uint32_t)(it / p.b_stages) & 1u) ^ 1u);
        mbar_arrive_expect_tx(&b_full[bs], (uint32_t)(WS_ZROWS * rbB));
        tma_load_5d(smemB + (size_t)bs * p.b_bytes, &tmapZ, &b_full[bs], p.co0, w0, h0 - 1, d0, n);
        mbar_wait(&a_empty[as], ((uint32_t)(it / p.a_stages) & 1u) ^ 1u);
        mbar_arrive_expect_tx(&a_full[as], (uint32_t)(p.PG * WS_XH * WS_XW * rbA));
        tma_load_5d(smemA + (size_t)as * p.a_bytes, &tmapX, &a_full[as], slice * CA, w0 - 1, h0, d0 - 1 + dd0, n);  // PG planes from dd0
      }
    }
  } else if (warp == 1) {
    // MMA issuer (whole warp converged, one elected lane issues)
    const uint32_t idesc = umma_idesc_bf16(128, 3 * p.Cout, 1, 1);
    // A: MN-major, atom stride (LBO) = ONE row -> atom j = view shifted by j voxels in w; 8-row K group stride (SBO) = one x line (10 rows).
    // B: MN-major, atom stride (LBO) = ONE tile line (8 rows) -> atom i = view shifted by i lines in h; K group stride = one line too.
    const uint64_t hiA = umma_smem_desc(0, 0, (uint32_t)(WS_XW * rbA), umma_layout_for_row_bytes(rbA)) & 0xFFFFFFFF00000000ull;
    const uint64_t hiB = umma_smem_desc(0, 0, (uint32_t)(8 * rbB), umma_layout_for_row_bytes(rbB)) & 0xFFFFFFFF00000000ull;
    const uint32_t lboA = ((uint32_t)rbA >> 4) << 16, lboB = (((uint32_t)(8 * rbB) >> 4) & 0x3FFFu) << 16;
    constexpr uint32_t A_LINE = (uint32_t)(WS_XW * rbA) >> 4;  // one x line, 16-byte units
    const uint32_t b_k = (uint32_t)(16 * rbB) >> 4;           // 16 voxel rows (two lines) of dz
    int it = 0;
    for (int t = split; t < p.tiles; t += p.S, ++it) {
      const int bs = it % p.b_stages, as = it % p.a_stages;
      mbar_wait(&b_full[bs], (uint32_t)(it / p.b_stages) & 1u);
      mbar_wait(&a_full[as], (uint32_t)(it / p.a_stages) & 1u);
      tc_fence_after();
      const uint32_t a_lo = ((smem_u32(smemA + (size_t)as * p.a_bytes) >> 4) & 0x3FFFu) | lboA;
      const uint32_t b_lo = ((smem_u32(smemB + (size_t)bs * p.b_bytes) >> 4) & 0x3FFFu) | lboB;
      const uint32_t accum = it != 0 ? 1u : 0u;
#pragma unroll 1
      for (int g = 0; g < p.PG; ++g) {
        const uint32_t a_g = a_lo + (uint32_t)(g * WS_XH) * A_LINE;
        const uint32_t tacc = tmem_base + (uint32_t)(g * 3 * p.Cout);
#pragma unroll
        for (int k = 0; k < 8; ++k)  // 128 voxels = 8 x K16 = lines (2k, 2k+1) of the tile; dz atoms start at halo lines 2k, 2k+1, 2k+2
          umma_bf16_elect(tacc, hiA | (uint64_t)(a_g + (uint32_t)(2 * k) * A_LINE), hiB | (uint64_t)(b_lo + (uint32_t)k * b_k), idesc,
                          (k != 0) ? 1u : accum);
      }
      umma_commit_elect(&a_empty[as]);
      umma_commit_elect(&b_empty[bs]);
    }
    umma_commit_elect(&done_bar);
  } else {
    // ================= final read-out (warps 2..5) =================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int dw = row / CA, ci = slice * CA + row % CA;
    const bool valid = dw < 3;
    mbar_wait(&done_bar, 0);
    __syncwarp();
    tc_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    for (int g = 0; g < p.PG; ++g)
      for (int i = 0; i < 3; ++i) {  // dz shift i-1  <->  dh = 1 - i  <->  tap index dh + 1 = 2 - i
        const int tap = ((dd0 + g) * 3 + (2 - i)) * 3 + (valid ? dw : 0);
        float* grow = p.G + ((((size_t)n * p.S + split) * 27 + tap) * p.Cin + ci) * p.CoutTotal + p.co0;
        for (int c0 = 0; c0 < p.Cout; c0 += 16) {
          uint32_t raw[16];
          tmem_ld_32x32b_x16(taddr + (uint32_t)((g * 3 + i) * p.Cout + c0), raw);
          tmem_ld_wait();
          if (valid) {
            float4* o = reinterpret_cast<float4*>(grow + c0);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float4 f;
              f.x = ntiles_mine ? __uint_as_float(raw[4 * e]) : 0.f;
              f.y = ntiles_mine ? __uint_as_float(raw[4 * e + 1]) : 0.f;
              f.z = ntiles_mine ? __uint_as_float(raw[4 * e + 2]) : 0.f;
              f.w = ntiles_mine ? __uint_as_float(raw[4 * e + 3]) : 0.f;
              o[e] = f;
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

bool wgrad_halo_plan(int N, int D, int H, int W, int Cin, int Cout, WgradHaloParams* pp) {
  WgradHaloParams& p = *pp;
  memset(&p, 0, sizeof(p));
  if (Cin % 16 != 0 || Cout % 16 != 0 || Cout > 256) return false;
  if (D < WH_HD || H < WH_HH || W < WH_HW) return false;
  const char* dis = getenv("B200UNET_NO_HALO");
  if (dis && dis[0] == '1') return false;
  p.N = N; p.D = D; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout;
  p.tilesH = (H + WH_BH - 1) / WH_BH;
  p.tilesW = (W + WH_BW - 1) / WH_BW;
  p.tiles = D * p.tilesH * p.tilesW;
  p.CA = (Cin % 32 == 0) ? 32 : 16;
  p.nslices = Cin / p.CA;
  const char* nohs = getenv("B200UNET_WGRAD_HS");
  if ((Cout == 16 || Cout == 32 || Cout == 64) && !(nohs && nohs[0] == '0')) {
    // h-stacked variant: N = 3 * C_out per instruction, one accumulator per depth tap
    p.hs = 1;
    p.PG = (9 * Cout <= 512) ? 3 : 1;
    // C_out <= 32: ONE depth tap per CTA and THREE CTAs per SM (72 KB of shared memory and <= 128 TMEM columns each).  A single issuing
    // warp needs ~80 cycles per instruction where the pipe wants one per 56 (N = 96); three co-resident CTAs give the SM three
    // issuers, each on its own accumulator (deterministic), for the price of reading the dz tile three times from L2.
    const char* pg1 = getenv("B200UNET_WGRAD_HS_PG1");
    int per_sm = 1;
    if (Cout <= 32 && !(pg1 && pg1[0] == '0')) {
      p.PG = 1;
      per_sm = 3;
    }
    p.AWb = Cout;
    p.a_bytes = (p.PG * WS_XH * WS_XW * p.CA * 2 + 1023) & ~1023;
    p.b_bytes = (WS_ZROWS * Cout * 2 + 1023) & ~1023;
    p.b_stages = 3;
    const int budget = (per_sm == 3 ? 72 : 200) * 1024;
    p.a_stages = (budget - p.b_stages * p.b_bytes) / p.a_bytes;
    if (p.a_stages > WH_MAX_A) p.a_stages = WH_MAX_A;
    if (p.a_stages < 2) return false;
    if (p.b_stages > p.a_stages) p.b_stages = p.a_stages;
    int cols = 32;
    while (cols < p.PG * 3 * Cout) cols <<= 1;
    p.tmem_cols = cols;
    int ctas_per_split = N * p.nslices * (3 / p.PG);
    int want = per_sm * sm_count() / ctas_per_split;
    if (const char* e = getenv("B200UNET_WGRAD_SPLITS")) {  // tests: few splits => many tiles accumulated per CTA
      const int v = atoi(e);
      if (v >= 1) want = v;
    }
    if (want < 1) want = 1;
    if (want > p.tiles) want = p.tiles;
    p.S = want;
    return true;
  }
  p.PG = (9 * Cout <= 512) ? 9 : ((3 * Cout <= 512) ? 3 : 1);
  p.AWb = (Cout % 64 == 0) ? 64 : (Cout % 32 == 0 ? 32 : 16);
  p.a_bytes = (WH_ROWS * p.CA * 2 + 1023) & ~1023;
  p.b_bytes = 128 * Cout * 2;
  p.b_stages = 2;
  int budget = 200 * 1024 - p.b_stages * p.b_bytes;
  p.a_stages = budget / p.a_bytes;
  if (p.a_stages > WH_MAX_A) p.a_stages = WH_MAX_A;
  if (p.a_stages < 2) return false;
  if (p.b_stages < p.a_stages && (budget - p.a_stages * p.a_bytes) >= 2 * p.b_bytes) p.b_stages = 4;
  int cols = 32;
  while (cols < p.PG * Cout) cols <<= 1;
  p.tmem_cols = cols;
  int ctas_per_split = N * p.nslices * (9 / p.PG);
  int want = sm_count() / ctas_per_split;
  if (want < 1) want = 1;
  if (want > p.tiles) want = p.tiles;
  p.S = want;
  return true;
}

int wgrad_halo_launch(const void* x, const void* dz, WgradHaloParams& p, cudaStream_t s) {
  CUtensorMap tmX, tmZ;
  if (p.hs) {
    int rc = make_act_tmap(&tmX, x, p.N, p.D, p.H, p.W, p.Cin, p.CA, p.PG, WS_XH, WS_XW);
    if (rc) return rc;
    rc = make_act_tmap(&tmZ, dz, p.N, p.D, p.H, p.W, p.CoutTotal, p.Cout, 1, WH_BH + 2, WH_BW);
    if (rc) return rc;
    size_t smem = (size_t)p.a_stages * p.a_bytes + (size_t)p.b_stages * p.b_bytes + 1024;
    auto kern = p.CA == 32 ? wgrad_hs_kernel<32> : wgrad_hs_kernel<16>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    B200_CHECK_ARG(e == cudaSuccess, "wgrad_hs: cudaFuncSetAttribute(%zu) failed: %s", smem, cudaGetErrorString(e));
    dim3 grid((unsigned)(p.N * p.S), (unsigned)p.nslices, (unsigned)(3 / p.PG));
    kern<<<grid, WS_THREADS, smem, s>>>(tmX, tmZ, p);
    B200_CHECK_LAUNCH("wgrad_hs");
    return 0;
  }
  int rc = make_act_tmap(&tmX, x, p.N, p.D, p.H, p.W, p.Cin, p.CA, WH_HD, WH_HH, WH_HW);
  if (rc) return rc;
  rc = make_act_tmap(&tmZ, dz, p.N, p.D, p.H, p.W, p.CoutTotal, p.AWb, 1, WH_BH, WH_BW);
  if (rc) return rc;
  size_t smem = (size_t)p.a_stages * p.a_bytes + (size_t)p.b_stages * p.b_bytes + 1024;
  auto kern = p.CA == 32 ? wgrad_halo_kernel<32> : wgrad_halo_kernel<16>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  B200_CHECK_ARG(e == cudaSuccess, "wgrad_halo: cudaFuncSetAttribute(%zu) failed: %s", smem, cudaGetErrorString(e));
  dim3 grid((unsigned)(p.N * p.S), (unsigned)p.nslices, (unsigned)(9 / p.PG));
  kern<<<grid, WH_THREADS, smem, s>>>(tmX, tmZ, p);
  B200_CHECK_LAUNCH("wgrad_halo");
  return 0;
}

int wgrad_halo_splits(int N, int D, int H, int W, int Cin, int Cout) {
  WgradHaloParams p;
  return wgrad_halo_plan(N, D, H, W, Cin, Cout, &p) ? p.S : 0;
}

int wgrad_halo_run(const void* x, const void* dz, int N, int D, int H, int W, int Cin, int Cout, int co0, int CoutTotal, float* G,
                   cudaStream_t s) {
  WgradHaloParams p;
  if (!wgrad_halo_plan(N, D, H, W, Cin, Cout, &p)) return -1;
  p.G = G;
  p.co0 = co0;
  p.CoutTotal = CoutTotal;
  return wgrad_halo_launch(x, dz, p, s);
}

}  // namespace b200
