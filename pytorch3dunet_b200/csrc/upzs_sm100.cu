// The upsampled half of the virtual-concat decoder convolution, z-stacked:
//
//   R[2u + p] = sum_{j in {0,1}^3} Wp[p][j] * low[u + off(p, j)]        (per axis: p = 0: offsets {-1, 0};  p = 1: offsets {0, +1})
//
// = conv3(nearest_up2x(low)) per output parity phase p (upcat_conv.cu), 8/27 of the MACs of the materialised form.  The tap-loop
// kernel (b200_conv3_up_phase_fwd in conv_igemm_sm100.cu) issues it as 64 (phase, tap) products of N = C_out columns, re-fetching the
// 128-voxel input tile for each: N = 32 instructions are paced by their operand fetch (40 cycles for 16 of math) and the tiles come
// 19x over the L2 (profiles/ncu_r01_full_summary.md).  Here, as in conv_zs_sm100.cu, the DEPTH direction is stacked along N:
//
//   one CTA = one in-plane phase (ph, pw) (grid.z), walking columns of 16x8 LOW-RES voxels along the depth axis;
//   input plane z (one 18x10 halo tile, fetched once) x the weights of the four (depth phase, depth tap) pairs it feeds
//        = contributions to the FULL-RES output planes  2z-1 | 2z | 2z+1 | 2z+2   (pd,jd) = (1,1) (0,1) (1,0) (0,0)
//   which are consecutive blocks of the lane's TMEM ring: N = 4*C_out per instruction, 4 in-plane views (jh, jw) per input plane.
//
// Each input plane opens two output planes (2z+1, 2z+2) and completes two (2z-1, 2z).  Same lanes / ring / first-step handling as the
// 3x3x3 kernel (zs_common.cuh).  The epilogue only rounds and stores (R is the `residual` input of the encoder-channel convolution,
// which applies bias / activation / statistics): row (xh, xw) of the tile goes to voxel (J, 2xh+ph, 2xw+pw) of the (2d,2h,2w) volume.
#include <stdlib.h>

#include "conv_common.cuh"
#include "zs_common.cuh"

namespace b200 {

struct UpzsParams {
  int N, d, h, w, C1, Cout;   // low-res dims, channels
  int tilesH, tilesW;
  int n_w;
  int NT;                     // output channels per CTA
  int KC, kchunks;
  int a_stages, a_bytes, b_total_bytes;
  int tmem_bufs;              // ring blocks (both lanes)
  int ctas_per_sample;
  bf16* R;
};

// all 4 in-plane views x KC/16 k-steps of one halo chunk.  View (jh, jw) of in-plane phase (ph, pw) starts at halo row
// (ph + jh) * 10 + (pw + jw).  b_lo points at [view 0][block 0] of this chunk; one view advances 4 blocks.
template <int KC, bool ONE_RUN>
__device__ __forceinline__ void upzs_issue_chunk(const ZsRun (&rr)[3], uint32_t a_lo, uint32_t b_lo, uint32_t b_view, uint64_t hiA, uint64_t hiB,
                                                 int ph, int pw, bool skip_first) {
  constexpr uint32_t RB16 = KC * 2 / 16;
  if (ONE_RUN) b_lo += rr[0].boff;
#pragma unroll
  for (int v = 0; v < 4; ++v) {
    const uint32_t offA = (uint32_t)((ph + (v >> 1)) * ZS_HW + pw + (v & 1)) * RB16;
#pragma unroll
    for (int k = 0; k < KC / 16; ++k) {
      const uint64_t adesc = hiA | (uint64_t)(a_lo + offA + 2u * k);
      const uint64_t bdesc = hiB | (uint64_t)(b_lo + 2u * k);
      if (v == 0 && k == 0) {
        if (!skip_first) {
          if (ONE_RUN) umma_bf16_elect(rr[0].tacc, adesc, bdesc, rr[0].idesc, 1u);
          else zs_issue(rr, adesc, bdesc, false);
        }
      } else {
        if (ONE_RUN) umma_bf16_elect(rr[0].tacc, adesc, bdesc, rr[0].idesc, 1u);
        else zs_issue(rr, adesc, bdesc, false);
      }
    }
    b_lo += b_view;
  }
}

template <int KC>
__global__ void __launch_bounds__(ZS_THREADS, 1)
conv3_upzs_kernel(const __grid_constant__ CUtensorMap tmapA, const __grid_constant__ CUtensorMap tmapB, const UpzsParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t a_full_[ZS_MAX_STAGES], a_empty_[ZS_MAX_STAGES];
  __shared__ __align__(8) uint64_t b_full, tmem_full_[ZS_MAX_SLOTS], tmem_empty_[ZS_MAX_SLOTS];
  __shared__ uint32_t tmem_slot;

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smemB = smem;
  const int b_region = (p.b_total_bytes + 1023) & ~1023;
  uint8_t* smemA = smem + b_region;
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;  // provably warp-uniform
  const int n = blockIdx.y, cta = blockIdx.x, cps = gridDim.x;
  const int nslices = p.Cout / p.NT;
  const int ipz = blockIdx.z / nslices;             // in-plane phase (ph, pw)
  const int n0 = (blockIdx.z - ipz * nslices) * p.NT;
  const int ph = ipz >> 1, pw = ipz & 1;
  const int nchunks = p.C1 / KC;
  constexpr int rb = KC * 2;
  const int R = p.tmem_bufs / ZS_LANES;
  const int S = p.a_stages / ZS_LANES;
  const int D = p.d;  // the walk is over LOW-RES planes
  const int lane_id = warp < 8 ? (warp >> 2) : (warp & 1);
  ZsWalk walk;
  {
    const long long T = (long long)p.tilesH * p.tilesW * D;
    const int vc = cta * ZS_LANES + lane_id, vn = cps * ZS_LANES;
    walk.L = (int)(T * vc / vn);
    walk.L1 = (int)(T * (vc + 1) / vn);
    walk.D = D;
  }
  uint64_t* a_full = a_full_ + lane_id * S;
  uint64_t* a_empty = a_empty_ + lane_id * S;
  uint64_t* tmem_full = tmem_full_ + lane_id * R;
  uint64_t* tmem_empty = tmem_empty_ + lane_id * R;

  if (threadIdx.x == 0) {
    for (int i = 0; i < ZS_LANES * S; ++i) {
      mbar_init(&a_full_[i], 1);
      mbar_init(&a_empty_[i], 1);
    }
    mbar_init(&b_full, 1);
    for (int i = 0; i < ZS_LANES * R; ++i) {
      mbar_init(&tmem_full_[i], 1);
      mbar_init(&tmem_empty_[i], 4);
    }
    fence_mbar_init();
  }
  if (warp == ZS_WARP_PRODUCER && lane == 0) {
    tma_prefetch_desc(&tmapA);
    tma_prefetch_desc(&tmapB);
  }
  if (warp == ZS_WARP_MMA) tmem_alloc(&tmem_slot, 512u);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot + (uint32_t)(lane_id * R * p.NT);
  uint8_t* smemA_lane = smemA + (size_t)lane_id * S * p.a_bytes;

  if (warp >= ZS_WARP_PRODUCER && warp < ZS_WARP_MMA) {
    // ================= TMA producer =================
    if (lane == 0) {
      const int wsample = p.n_w > 1 ? n : 0;
      if (lane_id == 0) {
        mbar_arrive_expect_tx(&b_full, (uint32_t)p.b_total_bytes);
        // smem layout [chunk][view = (jh,jw)][block b][NT][KC]; block b <-> (pd, jd) = (1,1) (0,1) (1,0) (0,0): ascending output plane
        for (int cb = 0; cb < nchunks; ++cb)
          for (int v = 0; v < 4; ++v)
            for (int b = 0; b < 4; ++b) {
              const int pd = (b == 0 || b == 2) ? 1 : 0, jd = b < 2 ? 1 : 0;
              const int phase = (pd << 2) | (ph << 1) | pw, j = (jd << 2) | v;  // v = (jh << 1) | jw
              tma_load_3d(smemB + ((size_t)((cb * 4 + v) * 4 + b)) * p.NT * rb, &tmapB, &b_full, cb * KC, n0, wsample * 64 + phase * 8 + j);
            }
      }
      ZsRing st = {0, 0u};
      ZsSeg sg;
      while (walk.next(sg)) {
        const int th_i = sg.col / p.tilesW, tw_i = sg.col - th_i * p.tilesW;
        const int h0 = th_i * ZS_BH, w0 = tw_i * ZS_BW;
        const int zin0 = sg.z0 > 0 ? sg.z0 - 1 : 0, zin1 = sg.z1 < D ? sg.z1 : D - 1;
        for (int zin = zin0; zin <= zin1; ++zin)
          for (int j = 0; j < nchunks; ++j) {
            mbar_wait(&a_empty[st.idx], st.ph ^ 1u);
            mbar_arrive_expect_tx(&a_full[st.idx], (uint32_t)(ZS_ROWS * rb));
            tma_load_5d(smemA_lane + (size_t)st.idx * p.a_bytes, &tmapA, &a_full[st.idx], j * KC, w0 - 1, h0 - 1, zin, n);
            st.step(S);
          }
      }
    }
  } else if (warp >= ZS_WARP_MMA) {
    // ================= MMA issuer of this lane =================
    const uint32_t lay = umma_layout_for_row_bytes(rb);
    const uint64_t hiA = umma_smem_desc(0, 16u, (uint32_t)(ZS_HW * rb), lay) & 0xFFFFFFFF00000000ull;
    const uint64_t hiB = umma_smem_desc(0, 16u, (uint32_t)(8 * rb), lay) & 0xFFFFFFFF00000000ull;
    const uint32_t lo_lbo = 1u << 16;
    const uint32_t sB0 = smem_u32(smemB);
    const uint32_t blk16 = (uint32_t)(p.NT * rb) >> 4;
    const uint32_t b_view = 4u * blk16;
    const uint32_t chunkB16 = 16u * blk16;
    uint32_t idesc_n[5];
    idesc_n[0] = 0u;
    for (int m = 1; m <= 4; ++m) idesc_n[m] = umma_idesc_bf16(128, m * p.NT, 0, 0);
    mbar_wait(&b_full, 0);
    tc_fence_after();
    ZsRing st = {0, 0u};
    ZsRing open = {0, 0u};   // block of the next output plane to be opened
    ZsRing done = {0, 0u};   // block of the next output plane to complete
    ZsSeg sg;
    while (walk.next(sg)) {
      const int z0 = sg.z0, z1 = sg.z1;
      const int J0 = 2 * z0, J1 = 2 * z1;  // full-res output planes of the segment
      const int zin0 = z0 > 0 ? z0 - 1 : 0, zin1 = z1 < D ? z1 : D - 1;
      int slot_a = open.idx;   // block of plane `a_prev`
      int a_prev = J0;
      int opened = J0;          // planes [J0, opened) have been opened
      int completed = J0;       // planes [J0, completed) have been committed
      for (int zin = zin0; zin <= zin1; ++zin) {
        // output planes this input plane contributes to: [2zin-1, 2zin+2] clipped to the segment
        const int a = 2 * zin - 1 > J0 ? 2 * zin - 1 : J0;
        const int b = 2 * zin + 2 < J1 - 1 ? 2 * zin + 2 : J1 - 1;
        const int m = b - a + 1;
        while (a_prev < a) {  // a advances by up to two planes per input plane
          if (++slot_a == R) slot_a = 0;
          ++a_prev;
        }
        const int f = opened - a;  // index of the first fresh block inside [a, b] (m = none)
        for (int J = opened; J <= b; ++J) {
          mbar_wait(&tmem_empty[open.idx], open.ph ^ 1u);
          open.step(R);
        }
        opened = b + 1;
        tc_fence_after();
        int w = R - slot_a;  // first block index at which the ring wraps
        if (w > m) w = m;
        ZsRun rf[3], rr[3];
        {
          const int c1 = w < f ? w : f, c2 = w < f ? f : w;
          const int bf[4] = {0, c1, c2, m}, br[4] = {0, w, m, m};
          const int tdr_a = a - (2 * zin - 1);  // weight block of plane a
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            const int s0 = bf[k], len = bf[k + 1] - bf[k];
            int sl = slot_a + s0;
            if (sl >= R) sl -= R;
            rf[k].tacc = tmem_base + (uint32_t)(sl * p.NT);
            rf[k].boff = (uint32_t)(tdr_a + s0) * blk16;
            rf[k].idesc = len > 0 ? idesc_n[len] : 0u;
            rf[k].accum = s0 >= f ? 0u : 1u;
            const int s1 = br[k], len1 = br[k + 1] - br[k];
            int sl1 = slot_a + s1;
            if (sl1 >= R) sl1 -= R;
            rr[k].tacc = tmem_base + (uint32_t)(sl1 * p.NT);
            rr[k].boff = (uint32_t)(tdr_a + s1) * blk16;
            rr[k].idesc = len1 > 0 ? idesc_n[len1] : 0u;
            rr[k].accum = 1u;
          }
        }
        const bool one_run = w >= m;
        uint32_t b_lo = ((sB0 >> 4) & 0x3FFFu) | lo_lbo;
        for (int j = 0; j < nchunks; ++j) {
          mbar_wait(&a_full[st.idx], st.ph);
          tc_fence_after();
          const uint32_t a_lo = ((smem_u32(smemA_lane + (size_t)st.idx * p.a_bytes) >> 4) & 0x3FFFu) | lo_lbo;
          if (j == 0) {  // first (view, k) step of the plane: per-block accumulate flags
            constexpr uint32_t RB16 = KC * 2 / 16;
            const uint32_t offA = (uint32_t)(ph * ZS_HW + pw) * RB16;
            zs_issue(rf, hiA | (uint64_t)(a_lo + offA), hiB | (uint64_t)b_lo, true);
          }
          if (one_run) upzs_issue_chunk<KC, true>(rr, a_lo, b_lo, b_view, hiA, hiB, ph, pw, j == 0);
          else upzs_issue_chunk<KC, false>(rr, a_lo, b_lo, b_view, hiA, hiB, ph, pw, j == 0);
          umma_commit_elect(&a_empty[st.idx]);
          st.step(S);
          b_lo += chunkB16;
        }
        // planes that received their last contribution: everything up to 2*zin (2*zin + 1 too when this is the last input plane)
        int last = 2 * zin;
        if (zin == D - 1) last = 2 * zin + 1;
        if (last > J1 - 1) last = J1 - 1;
        for (; completed <= last; ++completed) {
          umma_commit_elect(&tmem_full[done.idx]);
          done.step(R);
        }
      }
    }
  } else {
    // ================= epilogue: round + store =================
    const int qd = warp & 3;
    const int row = qd * 32 + lane;
    const int bx = row % ZS_BW, by = row / ZS_BW;
    const int NT = p.NT;
    ZsRing cur = {0, 0u};
    const int H2 = 2 * p.h, W2 = 2 * p.w;
    ZsSeg sg;
    while (walk.next(sg)) {
      const int th_i = sg.col / p.tilesW, tw_i = sg.col - th_i * p.tilesW;
      const int xh = th_i * ZS_BH + by, xw = tw_i * ZS_BW + bx;
      const bool valid = xh < p.h && xw < p.w;
      const size_t vox_hw = (size_t)n * (2 * D) * H2 * W2 + (size_t)(2 * xh + ph) * W2 + (2 * xw + pw);
      for (int J = 2 * sg.z0; J < 2 * sg.z1; ++J, cur.step(R)) {
        const int slot = cur.idx;
        mbar_wait(&tmem_full[slot], cur.ph);
        __syncwarp();
        tc_fence_after();
        const uint32_t taddr = tmem_base + (uint32_t)(slot * NT) + ((uint32_t)(qd * 32) << 16);
        bf16* orow = p.R + (vox_hw + (size_t)J * H2 * W2) * p.Cout + n0;
        for (int c0 = 0; c0 < NT; c0 += 32) {
          const bool wide = c0 + 32 <= NT;
          uint32_t raw[32];
          if (wide) tmem_ld_32x32b_x32(taddr + c0, raw);
          else tmem_ld_32x32b_x16(taddr + c0, raw);
          tmem_ld_wait();
          if (valid) {
            const int cw = wide ? 32 : 16;
            float v[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = i < cw ? __uint_as_float(raw[i]) : 0.f;
            bf16x8* op = reinterpret_cast<bf16x8*>(orow + c0);
#pragma unroll
            for (int i = 0; i < 4; ++i)
              if (8 * i < cw) op[i] = pack8(&v[8 * i]);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty[slot]);
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == ZS_WARP_MMA) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_slot, 512u);
  }
}

static bool upzs_plan(int N, int d, int h, int w, int C1, int Cout, UpzsParams* pp) {
  UpzsParams& p = *pp;
  memset(&p, 0, sizeof(p));
  const char* e = getenv("B200UNET_UPZS");
  if (e && e[0] == '0') return false;
  if (C1 % 16 != 0 || Cout % 16 != 0) return false;
  if (h < ZS_HH || w < ZS_HW || d < 1) return false;
  const int budget = 222 * 1024;
  const int kc = (C1 % 64 == 0) ? 64 : (C1 % 32 == 0 ? 32 : 16);
  const int a_bytes = (ZS_ROWS * kc * 2 + 1023) & ~1023;
  int NT = 0, stages = 0, b_total = 0;
  for (int nt : {Cout, 32}) {
    // NT <= 32: the lane's ring of 8 blocks holds four accumulating planes + the ones being drained; at most two slices
    if (nt > Cout || Cout % nt != 0 || nt > 32 || (nt != Cout && 2 * nt < Cout)) continue;
    const int bt = 16 * nt * C1 * 2;
    int st = (budget - ((bt + 1023) & ~1023) - 1024) / a_bytes;
    if (st > ZS_MAX_STAGES) st = ZS_MAX_STAGES;
    st &= ~1;
    if (st < 4) continue;
    NT = nt; stages = st; b_total = bt;
    break;
  }
  if (!NT) return false;
  int slots = 512 / NT;
  if (slots > ZS_MAX_SLOTS) slots = ZS_MAX_SLOTS;
  slots &= ~1;
  if (slots < 16) return false;
  p.N = N; p.d = d; p.h = h; p.w = w; p.C1 = C1; p.Cout = Cout;
  p.tilesH = (h + ZS_BH - 1) / ZS_BH;
  p.tilesW = (w + ZS_BW - 1) / ZS_BW;
  p.NT = NT;
  p.KC = kc;
  p.kchunks = C1 / kc;
  p.a_stages = stages;
  p.a_bytes = a_bytes;
  p.b_total_bytes = b_total;
  p.tmem_bufs = slots;
  const long long T = (long long)p.tilesH * p.tilesW * d;
  if (T >= (1ll << 30)) return false;
  int cps = sm_count() / (N * 4 * (Cout / NT));
  if (cps < 1) cps = 1;
  if (const char* c = getenv("B200UNET_ZS_CTAS")) {
    const int v = atoi(c);
    if (v >= 1) cps = v;
  }
  if ((long long)cps * ZS_LANES > T) cps = (int)((T + ZS_LANES - 1) / ZS_LANES);
  p.ctas_per_sample = cps;
  return true;
}

// returns -1 when the shape is not taken (the caller falls back to the tap-loop kernel)
int conv3_upzs_run(const void* low, const void* wp, int n_w, int N, int d, int h, int w, int C1, int Cout, void* R, cudaStream_t s) {
  UpzsParams p;
  if (!upzs_plan(N, d, h, w, C1, Cout, &p)) return -1;
  p.n_w = n_w;
  p.R = (bf16*)R;
  CUtensorMap tmA, tmB;
  int rc = make_act_tmap(&tmA, low, N, d, h, w, C1, p.KC, 1, ZS_HH, ZS_HW);
  if (rc) return rc;
  rc = make_w_tmap(&tmB, wp, 64 * n_w, Cout, C1, p.KC, p.NT, 1);
  if (rc) return rc;
  size_t smem = (size_t)((p.b_total_bytes + 1023) & ~1023) + (size_t)p.a_stages * p.a_bytes + 1024;
  auto kern = p.KC == 64 ? conv3_upzs_kernel<64> : (p.KC == 32 ? conv3_upzs_kernel<32> : conv3_upzs_kernel<16>);
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  B200_CHECK_ARG(e == cudaSuccess, "conv3_upzs: cudaFuncSetAttribute(%zu) failed: %s", smem, cudaGetErrorString(e));
  dim3 grid((unsigned)p.ctas_per_sample, (unsigned)N, (unsigned)(4 * (Cout / p.NT)));
  kern<<<grid, ZS_THREADS, smem, s>>>(tmA, tmB, p);
  B200_CHECK_LAUNCH("conv3_upzs");
  return 0;
}

}  // namespace b200
