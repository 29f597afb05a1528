// Gradient of the virtual-concat phase conv w.r.t. the LOW-RES tensor, z-stacked (the transpose of upzs_sm100.cu):
//
//   dlow[u] = sum_{e in {-1,0,1,2}^3} Wd[e] * dz[2u + e]          (a 4x4x4 stride-2 convolution of the full-res output gradient)
//
// The tap-loop version (b200_conv3_up_dgrad in conv_igemm_sm100.cu) issues 64 taps of N = C1 columns over stride-2 TMA tiles that are
// re-fetched per tap.  Here one CTA owns one IN-PLANE PARITY (rh, rw) of the dz lattice (grid.z): its input rows are the samples
// dz[J][2i+rh][2j+rw] (one element-stride-2 TMA box per full-res plane J, 18x10 samples), its in-plane taps are the 2x2 shifted views
// (rh = 0: e_h in {0, 2} = shifts {0, +1};  rh = 1: e_h in {-1, +1} = shifts {-1, 0}), and the DEPTH direction is stacked along N:
// input plane J feeds the two low-res output planes  t-1 | t  (t = ceil(J/2); e_d = +1 | -1 for odd J, 2 | 0 for even J), adjacent blocks
// of the lane's TMEM ring: N = 2*C1 per instruction.  Output plane u opens at J = 2u-1 and completes after J = 2u+2.  The four parity
// CTAs produce four PARTIAL gradients (bf16 [4][N][d][h][w][C1]); b200_sum_parts adds them (the partials are 1/8 the size of dz).
#include <stdlib.h>

#include "conv_common.cuh"
#include "zs_common.cuh"

namespace b200 {

struct UpdzsParams {
  int N, d, h, w, C1, Cout;   // low-res dims; C1 = channels of the low-res tensor (N side of the MMA), Cout = channels of dz (K side)
  int tilesH, tilesW;
  int NT;                     // output channels (of C1) per CTA
  int KC, kchunks;
  int a_stages, a_bytes, b_total_bytes;
  int tmem_bufs;              // ring blocks (both lanes)
  int ctas_per_sample;
  bf16* parts;                // [4][N][d][h][w][C1]
};

// all 4 in-plane views x KC/16 k-steps of one halo chunk.  View (a_h, a_w) of lattice parity (rh, rw) starts at halo row
// (1 - rh + a_h) * 10 + (1 - rw + a_w).  b_lo points at [view 0][block 0] of this chunk and depth parity; one view advances 2 blocks.
template <int KC, bool ONE_RUN>
__device__ __forceinline__ void updzs_issue_chunk(const ZsRun (&rr)[3], uint32_t a_lo, uint32_t b_lo, uint32_t b_view, uint64_t hiA, uint64_t hiB,
                                                  int rh, int rw, bool skip_first) {
  constexpr uint32_t RB16 = KC * 2 / 16;
  if (ONE_RUN) b_lo += rr[0].boff;
#pragma unroll
  for (int v = 0; v < 4; ++v) {
    const uint32_t offA = (uint32_t)((1 - rh + (v >> 1)) * ZS_HW + 1 - rw + (v & 1)) * RB16;
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
conv3_updzs_kernel(const __grid_constant__ CUtensorMap tmapA, const __grid_constant__ CUtensorMap tmapB, const UpdzsParams p) {
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
  const int nslices = p.C1 / p.NT;
  const int ipz = blockIdx.z / nslices;             // in-plane lattice parity (rh, rw)
  const int n0 = (blockIdx.z - ipz * nslices) * p.NT;
  const int rh = ipz >> 1, rw = ipz & 1;
  const int nchunks = p.Cout / KC;
  constexpr int rb = KC * 2;
  const int R = p.tmem_bufs / ZS_LANES;
  const int S = p.a_stages / ZS_LANES;
  const int D = p.d;          // the walk is over LOW-RES output planes
  const int JD = 2 * p.d;     // full-res depth
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
      if (lane_id == 0) {
        mbar_arrive_expect_tx(&b_full, (uint32_t)p.b_total_bytes);
        // smem layout [chunk][rd = J & 1][view = (a_h, a_w)][block b][NT][KC]; block b <-> output plane t-1+b:
        //   rd = 1 (J = 2t-1): e_d = +1 | -1;   rd = 0 (J = 2t): e_d = 2 | 0;   e_h = rh ? 2*a_h - 1 : 2*a_h  (same for w)
        for (int cb = 0; cb < nchunks; ++cb)
          for (int rd = 0; rd < 2; ++rd)
            for (int v = 0; v < 4; ++v)
              for (int b = 0; b < 2; ++b) {
                const int ed = rd ? (b == 0 ? 1 : -1) : (b == 0 ? 2 : 0);
                const int ah = v >> 1, aw = v & 1;
                const int eh = rh ? 2 * ah - 1 : 2 * ah, ew = rw ? 2 * aw - 1 : 2 * aw;
                const int e = ((ed + 1) << 4) | ((eh + 1) << 2) | (ew + 1);
                tma_load_3d(smemB + ((size_t)(((cb * 2 + rd) * 4 + v) * 2 + b)) * p.NT * rb, &tmapB, &b_full, cb * KC, n0, e);
              }
      }
      ZsRing st = {0, 0u};
      ZsSeg sg;
      while (walk.next(sg)) {
        const int th_i = sg.col / p.tilesW, tw_i = sg.col - th_i * p.tilesW;
        const int h0 = th_i * ZS_BH, w0 = tw_i * ZS_BW;
        const int Jb = 2 * sg.z0 - 1 > 0 ? 2 * sg.z0 - 1 : 0, Je = 2 * sg.z1 < JD - 1 ? 2 * sg.z1 : JD - 1;
        for (int J = Jb; J <= Je; ++J)
          for (int j = 0; j < nchunks; ++j) {
            mbar_wait(&a_empty[st.idx], st.ph ^ 1u);
            mbar_arrive_expect_tx(&a_full[st.idx], (uint32_t)(ZS_ROWS * rb));
            // element-stride-2 box: sample (i, jx) of the tile is dz[J][2*(h0-1+i)+rh][2*(w0-1+jx)+rw]
            tma_load_5d(smemA_lane + (size_t)st.idx * p.a_bytes, &tmapA, &a_full[st.idx], j * KC, 2 * (w0 - 1) + rw, 2 * (h0 - 1) + rh, J, n);
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
    const uint32_t b_view = 2u * blk16;
    const uint32_t b_rd = 8u * blk16;
    const uint32_t chunkB16 = 16u * blk16;
    const uint32_t idesc1 = umma_idesc_bf16(128, p.NT, 0, 0), idesc2 = umma_idesc_bf16(128, 2 * p.NT, 0, 0);
    mbar_wait(&b_full, 0);
    tc_fence_after();
    ZsRing st = {0, 0u};
    ZsRing open = {0, 0u};   // block of the next output plane to be opened
    ZsRing done = {0, 0u};   // block of the next output plane to complete
    ZsSeg sg;
    while (walk.next(sg)) {
      const int z0 = sg.z0, z1 = sg.z1;
      const int Jb = 2 * z0 - 1 > 0 ? 2 * z0 - 1 : 0, Je = 2 * z1 < JD - 1 ? 2 * z1 : JD - 1;
      int slot_a = open.idx;   // block of plane `a_prev`
      int a_prev = z0;
      int opened = z0;          // output planes [z0, opened) have been opened
      int completed = z0;       // output planes [z0, completed) have been committed
      for (int J = Jb; J <= Je; ++J) {
        const int t = (J + 1) >> 1;  // J = 2t-1 or 2t feeds output planes t-1 and t
        const int a = t - 1 > z0 ? t - 1 : z0;
        const int b = t < z1 - 1 ? t : z1 - 1;
        const int m = b - a + 1;     // 1 or 2
        while (a_prev < a) {
          if (++slot_a == R) slot_a = 0;
          ++a_prev;
        }
        const int f = opened - a;    // index of the first fresh block inside [a, b] (m = none)
        for (int u = opened; u <= b; ++u) {
          mbar_wait(&tmem_empty[open.idx], open.ph ^ 1u);
          open.step(R);
        }
        if (opened < b + 1) opened = b + 1;
        tc_fence_after();
        int w = R - slot_a;          // first block index at which the ring wraps
        if (w > m) w = m;
        ZsRun rf[3], rr[3];
        {
          const int c1 = w < f ? w : f, c2 = w < f ? f : w;
          const int bf[4] = {0, c1, c2, m}, br[4] = {0, w, m, m};
          const int tdr_a = a - (t - 1);  // weight block of plane a
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            const int s0 = bf[k], len = bf[k + 1] - bf[k];
            int sl = slot_a + s0;
            if (sl >= R) sl -= R;
            rf[k].tacc = tmem_base + (uint32_t)(sl * p.NT);
            rf[k].boff = (uint32_t)(tdr_a + s0) * blk16;
            rf[k].idesc = len <= 0 ? 0u : (len == 1 ? idesc1 : idesc2);
            rf[k].accum = s0 >= f ? 0u : 1u;
            const int s1 = br[k], len1 = br[k + 1] - br[k];
            int sl1 = slot_a + s1;
            if (sl1 >= R) sl1 -= R;
            rr[k].tacc = tmem_base + (uint32_t)(sl1 * p.NT);
            rr[k].boff = (uint32_t)(tdr_a + s1) * blk16;
            rr[k].idesc = len1 <= 0 ? 0u : (len1 == 1 ? idesc1 : idesc2);
            rr[k].accum = 1u;
          }
        }
        const bool one_run = w >= m;
        uint32_t b_lo = (((sB0 >> 4) & 0x3FFFu) | lo_lbo) + (uint32_t)(J & 1) * b_rd;
        for (int j = 0; j < nchunks; ++j) {
          mbar_wait(&a_full[st.idx], st.ph);
          tc_fence_after();
          const uint32_t a_lo = ((smem_u32(smemA_lane + (size_t)st.idx * p.a_bytes) >> 4) & 0x3FFFu) | lo_lbo;
          if (j == 0) {  // first (view, k) step of the plane: per-block accumulate flags
            constexpr uint32_t RB16 = KC * 2 / 16;
            const uint32_t offA = (uint32_t)((1 - rh) * ZS_HW + 1 - rw) * RB16;
            zs_issue(rf, hiA | (uint64_t)(a_lo + offA), hiB | (uint64_t)b_lo, true);
          }
          if (one_run) updzs_issue_chunk<KC, true>(rr, a_lo, b_lo, b_view, hiA, hiB, rh, rw, j == 0);
          else updzs_issue_chunk<KC, false>(rr, a_lo, b_lo, b_view, hiA, hiB, rh, rw, j == 0);
          umma_commit_elect(&a_empty[st.idx]);
          st.step(S);
          b_lo += chunkB16;
        }
        // output plane u is complete after input plane 2u+2 (or after the segment's / volume's last input plane)
        int last = J >= 2 ? (J - 2) >> 1 : -1;
        if (J == Je) last = z1 - 1;
        if (last > z1 - 1) last = z1 - 1;
        for (; completed <= last; ++completed) {
          umma_commit_elect(&tmem_full[done.idx]);
          done.step(R);
        }
      }
    }
  } else {
    // ================= epilogue: round + store this parity's partial gradient =================
    const int qd = warp & 3;
    const int row = qd * 32 + lane;
    const int bx = row % ZS_BW, by = row / ZS_BW;
    const int NT = p.NT;
    ZsRing cur = {0, 0u};
    const size_t HW = (size_t)p.h * p.w;
    bf16* part = p.parts + (size_t)ipz * p.N * D * HW * p.C1;
    ZsSeg sg;
    while (walk.next(sg)) {
      const int th_i = sg.col / p.tilesW, tw_i = sg.col - th_i * p.tilesW;
      const int xh = th_i * ZS_BH + by, xw = tw_i * ZS_BW + bx;
      const bool valid = xh < p.h && xw < p.w;
      const size_t vox_hw = (size_t)n * D * HW + (size_t)xh * p.w + xw;
      for (int u = sg.z0; u < sg.z1; ++u, cur.step(R)) {
        const int slot = cur.idx;
        mbar_wait(&tmem_full[slot], cur.ph);
        __syncwarp();
        tc_fence_after();
        const uint32_t taddr = tmem_base + (uint32_t)(slot * NT) + ((uint32_t)(qd * 32) << 16);
        bf16* orow = part + (vox_hw + (size_t)u * HW) * p.C1 + n0;
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

// out = sum of `nparts` tensors of `elems` 16-bit elements each (fp32 accumulation)
__global__ void sum_parts_kernel(const bf16* __restrict__ parts, int nparts, size_t elems8, bf16* __restrict__ out) {
  const bf16x8* pp = reinterpret_cast<const bf16x8*>(parts);
  bf16x8* op = reinterpret_cast<bf16x8*>(out);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < elems8; i += (size_t)gridDim.x * blockDim.x) {
    float acc[8] = {0};
    for (int q = 0; q < nparts; ++q) {
      float f[8];
      unpack8(pp[(size_t)q * elems8 + i], f);
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] += f[k];
    }
    op[i] = pack8(acc);
  }
}

int make_act_tmap_stride2_hw(CUtensorMap* tm, const void* ptr, int N, int D, int H, int W, int C, int kc, int bh, int bw);  // conv_igemm_sm100.cu

static bool updzs_plan(int N, int d, int h, int w, int C1, int Cout, UpdzsParams* pp) {
  UpdzsParams& p = *pp;
  memset(&p, 0, sizeof(p));
  const char* e = getenv("B200UNET_UPZS");
  if (e && e[0] == '0') return false;
  if (C1 % 16 != 0 || Cout % 16 != 0) return false;
  if (h < ZS_HH || w < ZS_HW || d < 1) return false;
  const int budget = 222 * 1024;
  const int kc = (Cout % 64 == 0) ? 64 : (Cout % 32 == 0 ? 32 : 16);
  const int a_bytes = (ZS_ROWS * kc * 2 + 1023) & ~1023;
  int NT = 0, stages = 0, b_total = 0;
  for (int nt : {C1, 64}) {
    // N = 2*NT <= 128: the lane's ring (>= 4 blocks) holds two accumulating planes + the ones being drained; at most two slices
    if (nt > C1 || C1 % nt != 0 || nt > 64 || (nt != C1 && 2 * nt < C1)) continue;
    const int bt = 16 * nt * Cout * 2;
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
  if (slots < 8) return false;
  p.N = N; p.d = d; p.h = h; p.w = w; p.C1 = C1; p.Cout = Cout;
  p.tilesH = (h + ZS_BH - 1) / ZS_BH;
  p.tilesW = (w + ZS_BW - 1) / ZS_BW;
  p.NT = NT;
  p.KC = kc;
  p.kchunks = Cout / kc;
  p.a_stages = stages;
  p.a_bytes = a_bytes;
  p.b_total_bytes = b_total;
  p.tmem_bufs = slots;
  const long long T = (long long)p.tilesH * p.tilesW * d;
  if (T >= (1ll << 30)) return false;
  int cps = sm_count() / (N * 4 * (C1 / NT));
  if (cps < 1) cps = 1;
  if (const char* c = getenv("B200UNET_ZS_CTAS")) {
    const int v = atoi(c);
    if (v >= 1) cps = v;
  }
  if ((long long)cps * ZS_LANES > T) cps = (int)((T + ZS_LANES - 1) / ZS_LANES);
  p.ctas_per_sample = cps;
  return true;
}

bool conv3_updzs_supported(int N, int d, int h, int w, int Cout, int C1) {
  UpdzsParams p;
  return updzs_plan(N, d, h, w, C1, Cout, &p);
}

// parts: bf16 [4][N][d][h][w][C1] scratch; dxb = their sum.  Returns -1 when the shape is not taken.
int conv3_updzs_run(const void* dz, const void* wd, int N, int d, int h, int w, int Cout, int C1, void* parts, void* dxb, cudaStream_t s) {
  UpdzsParams p;
  if (!updzs_plan(N, d, h, w, C1, Cout, &p)) return -1;
  p.parts = (bf16*)parts;
  CUtensorMap tmA, tmB;
  int rc = make_act_tmap_stride2_hw(&tmA, dz, N, 2 * d, 2 * h, 2 * w, Cout, p.KC, ZS_HH, ZS_HW);
  if (rc) return rc;
  rc = make_w_tmap(&tmB, wd, 64, C1, Cout, p.KC, p.NT, 1);
  if (rc) return rc;
  size_t smem = (size_t)((p.b_total_bytes + 1023) & ~1023) + (size_t)p.a_stages * p.a_bytes + 1024;
  auto kern = p.KC == 64 ? conv3_updzs_kernel<64> : (p.KC == 32 ? conv3_updzs_kernel<32> : conv3_updzs_kernel<16>);
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  B200_CHECK_ARG(e == cudaSuccess, "conv3_updzs: cudaFuncSetAttribute(%zu) failed: %s", smem, cudaGetErrorString(e));
  dim3 grid((unsigned)p.ctas_per_sample, (unsigned)N, (unsigned)(4 * (C1 / p.NT)));
  kern<<<grid, ZS_THREADS, smem, s>>>(tmA, tmB, p);
  B200_CHECK_LAUNCH("conv3_updzs");
  const size_t elems8 = (size_t)N * d * h * w * C1 / 8;
  size_t blocks = (elems8 + 255) / 256;
  const size_t cap = (size_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  sum_parts_kernel<<<(unsigned)blocks, 256, 0, s>>>((const bf16*)parts, 4, elems8, (bf16*)dxb);
  B200_CHECK_LAUNCH("sum_parts");
  return 0;
}

}  // namespace b200
