// 3x3x3 convolution with the three DEPTH taps stacked along the MMA's N dimension ("z-stacked" kernel): the fprop / dgrad kernel of the
// small-channel, large-volume layers (C_out <= 128, 27*C_in*C_out*2 bytes of weights resident in shared memory).
//
// Why.  A tcgen05.mma with both operands in shared memory (M=128, K=16, bf16) is paced by max(math, operand fetch):
//   math   = N/2 cycles,   operand fetch = (128 + N) * 32 B / (128 B/cycle) = 32 + N/4 cycles
// (fits every probe in profiles/probes_r01.md: 40 / 48 / 64 cycles at N = 32 / 64 / 128 with two issuers).  With N = C_out = 32 an
// instruction does 16 cycles of math in 40: the layers that hold most of the net's FLOPs ran at 40 % of the tensor pipe whatever the
// issue loop did.  Stacking the weights of the three depth taps along N (N = 3*C_out = 96: 48 cycles of math in 56) makes ONE
// instruction do the work of three:
//
//   input plane z (one 18x10-voxel halo tile, K = C_in)  x  [W(dd=+1) | W(dd=0) | W(dd=-1)] (N = 3*C_out)
//        = contributions to the output planes  z-1 | z | z+1,   which live in ADJACENT column blocks of TMEM.
//
// Each persistent CTA walks a column of tiles along the depth axis; output plane z accumulates in TMEM while the input planes z-1, z,
// z+1 stream through, then the epilogue warps drain it while the next planes accumulate.  Every input plane tile is fetched once
// (1.4x halo overhead instead of 4.2x), the MMA count per output tile drops from 27*C_in/16 to 9*C_in/16.
//
//   TMEM: per lane a ring of R = min(16, 512 / C_out) / 2 column blocks of C_out columns; output plane number q (in the lane's own order)
//         lives in block q mod R.  An input plane targets up to three consecutive blocks = one MMA, or two when the ring wraps.
//   first touch of a block uses accumulate = 0: the very first (tap, k) step of an input plane is issued as separate MMAs for the
//         already-open blocks (accumulate) and the newly opened one (overwrite); all other steps are single instructions.
//   warps: two independent lanes per CTA (see ZS_LANES): 0..3 / 4..7 epilogue warpgroups, 8 / 9 TMA producers, 10 / 11 MMA issuers.
//   GroupNorm statistics of the output: per-thread register accumulators across all planes of the CTA (C_out <= 32) or per-warp
//         shared-memory accumulators (wider), reduced ONCE at the end: partials [N][P = CTAs per sample][C_out][2].
//
// Same contract as conv3_halo_kernel (conv_halo_sm100.cu): y = act(conv(x, wf) + bias[cls] (+ residual)), bf16 NDHWC.
#include <stdlib.h>

#include "conv_common.cuh"
#include "zs_common.cuh"

namespace b200 {

// EW = epilogue warpgroups per lane.  One warpgroup (a single warp per SM sub-partition) drains a 128 x 32 plane block in ~3 300 cycles
// (tcgen05.ld + ~250 dependent ALU instructions + stores, measured with the MMAs switched off: profiles/zs_debug_r02_v4_two_lanes.log),
// more than the ~1 500 cycles its MMAs take: with EW = 2 the two warpgroups of a lane split every block's COLUMNS in 16-wide slabs
// (warps w and w+4 may read the same TMEM lane quarter), which halves the per-thread work and the register footprint.
template <int KC, int EW>
__global__ void __launch_bounds__((8 * EW + 4) * 32, 1)
conv3_zs_kernel(const __grid_constant__ CUtensorMap tmapA, const __grid_constant__ CUtensorMap tmapB, const ConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t a_full_[ZS_MAX_STAGES], a_empty_[ZS_MAX_STAGES];   // lane l owns entries [l*S, (l+1)*S)
  __shared__ __align__(8) uint64_t b_full, tmem_full_[ZS_MAX_SLOTS], tmem_empty_[ZS_MAX_SLOTS];  // lane l owns [l*R, (l+1)*R)
  __shared__ uint32_t tmem_slot;

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smemB = smem;
  const int b_region = (p.b_total_bytes + 1023) & ~1023;
  uint8_t* smemA = smem + b_region;
  constexpr int NEW = 8 * EW;                      // epilogue warps
  constexpr int THREADS = (NEW + 4) * 32;
  constexpr int WARP_PRODUCER = NEW, WARP_MMA = NEW + 2;  // two producers, two issuers (lane = (warp - NEW) & 1)
  float* stat_acc = reinterpret_cast<float*>(smemA + (size_t)p.a_stages * p.a_bytes);  // [NEW epilogue warps][NT][2]
  float* bias_interior = stat_acc + NEW * p.NT * 2;                                      // [8 parity variants][NT]
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;  // provably warp-uniform: the role branches and descriptor math below stay on the uniform datapath
  const int n = blockIdx.y, cta = blockIdx.x, cps = gridDim.x;
  const int n0 = blockIdx.z * p.NT;  // output-channel slice of this CTA (C_out > NT: the resident weights of NT channels fit, those of C_out do not)
  const int nchunks = p.Cin / KC;
  constexpr int rb = KC * 2;
  const int R = p.tmem_bufs / ZS_LANES;  // ring slots per lane
  const int S = p.a_stages / ZS_LANES;   // halo stages per lane
  const int D = p.D;
  // lane of this warp: epilogue warps 0..3 -> 0, 4..7 -> 1; producer / issuer warps alternate
  const int lane_id = warp < NEW ? warp / (4 * EW) : ((warp - NEW) & 1);
  ZsWalk walk = zs_walk(p, cta * ZS_LANES + lane_id, cps * ZS_LANES);
  uint64_t* a_full = a_full_ + lane_id * S;
  uint64_t* a_empty = a_empty_ + lane_id * S;
  uint64_t* tmem_full = tmem_full_ + lane_id * R;
  uint64_t* tmem_empty = tmem_empty_ + lane_id * R;
#ifdef B200_DEBUG
  const int planes_mine = walk.L1 - walk.L;
#endif

  if (threadIdx.x == 0) {
    for (int i = 0; i < ZS_LANES * S; ++i) {
      mbar_init(&a_full_[i], 1);
      mbar_init(&a_empty_[i], 1);
    }
    mbar_init(&b_full, 1);
    for (int i = 0; i < ZS_LANES * R; ++i) {
      mbar_init(&tmem_full_[i], 1);
      mbar_init(&tmem_empty_[i], 4 * (EW == 2 && p.NT >= 32 ? 2 : 1));  // one arrival per epilogue warp that drains a slab of the block
    }
    fence_mbar_init();
  }
  if (warp == WARP_PRODUCER && lane == 0) {
    tma_prefetch_desc(&tmapA);
    tma_prefetch_desc(&tmapB);
  }
  if (warp == WARP_MMA) tmem_alloc(&tmem_slot, (uint32_t)p.tmem_cols);
  if (p.n_b)
    for (int i = threadIdx.x; i < 8 * p.NT; i += THREADS) {
      const int v = i / p.NT, c = i - v * p.NT;
      const int cls = ((v & 4 ? 3 : 1) << 4) | ((v & 2 ? 3 : 1) << 2) | (v & 1 ? 3 : 1);
      bias_interior[i] = p.biascls[((size_t)(p.n_b > 1 ? n : 0) * 64 + cls) * p.Cout + n0 + c];
    }
  for (int i = threadIdx.x; i < NEW * p.NT * 2; i += THREADS) stat_acc[i] = 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot + (uint32_t)(lane_id * R * p.NT);   // this lane's half of the accumulator ring
  uint8_t* smemA_lane = smemA + (size_t)lane_id * S * p.a_bytes;

  if (warp >= WARP_PRODUCER && warp < WARP_MMA) {
    // ================= TMA producer: resident weights once, then one halo tile per (input plane, channel chunk) =================
    if (lane == 0) {
      const int wsample = p.n_w > 1 ? n : 0;
      if (lane_id == 0) mbar_arrive_expect_tx(&b_full, (uint32_t)p.b_total_bytes);
      // smem layout [chunk][t9][tdr = 2 - td][C_out][KC]: the three depth taps of one in-plane tap are adjacent row blocks, in the
      // order of ascending OUTPUT plane (z-1 <- dd=+1, z <- dd=0, z+1 <- dd=-1)
      for (int cb = 0; cb < nchunks && lane_id == 0; ++cb)
        for (int t9 = 0; t9 < 9; ++t9)
          for (int tdr = 0; tdr < 3; ++tdr)
            tma_load_3d(smemB + ((size_t)((cb * 9 + t9) * 3 + tdr)) * p.NT * rb, &tmapB, &b_full, cb * KC, n0,
                        wsample * 27 + (2 - tdr) * 9 + t9);
      ZsRing st = {0, 0u};
      long long w_prod = 0, t_begin = dbg_clock();
      ZsSeg sg;
      while (walk.next(sg)) {
        const int th_i = sg.col / p.tilesW, tw_i = sg.col - th_i * p.tilesW;
        const int h0 = th_i * ZS_BH, w0 = tw_i * ZS_BW;
        const int zin0 = sg.z0 > 0 ? sg.z0 - 1 : 0, zin1 = sg.z1 < D ? sg.z1 : D - 1;
        for (int zin = zin0; zin <= zin1; ++zin)
          for (int j = 0; j < nchunks; ++j) {
            const long long c0 = dbg_clock();
            mbar_wait(&a_empty[st.idx], st.ph ^ 1u);
            w_prod += dbg_clock() - c0;
            mbar_arrive_expect_tx(&a_full[st.idx], (uint32_t)(ZS_ROWS * rb));
            tma_load_5d(smemA_lane + (size_t)st.idx * p.a_bytes, &tmapA, &a_full[st.idx], j * KC, w0 - 1, h0 - 1, zin, n);
            st.step(S);
          }
      }
#ifdef B200_DEBUG
      if (p.dbg && lane_id == 0) {
        long long* o = p.dbg + (size_t)(blockIdx.y * gridDim.x + blockIdx.x) * 16;
        o[0] = w_prod;
        o[1] = dbg_clock() - t_begin;
      }
#else
      (void)w_prod; (void)t_begin;
#endif
    }
  } else if (warp >= WARP_MMA) {
    // ================= MMA issuer of this lane (whole warp converged, one elected lane issues) =================
    const uint32_t lay = umma_layout_for_row_bytes(rb);
    const uint64_t hiA = umma_smem_desc(0, 16u, (uint32_t)(ZS_HW * rb), lay) & 0xFFFFFFFF00000000ull;
    const uint64_t hiB = umma_smem_desc(0, 16u, (uint32_t)(8 * rb), lay) & 0xFFFFFFFF00000000ull;
    const uint32_t lo_lbo = 1u << 16;
    const uint32_t sB0 = smem_u32(smemB);
    const uint32_t blk16 = (uint32_t)(p.NT * rb) >> 4;  // one weight block [C_out][KC], 16-byte units
    const uint32_t b_t9 = 3u * blk16;
    const uint32_t chunkB16 = 27u * blk16;
    const uint32_t idesc1 = umma_idesc_bf16(128, p.NT, 0, 0), idesc2 = umma_idesc_bf16(128, 2 * p.NT, 0, 0),
                   idesc3 = umma_idesc_bf16(128, 3 * p.NT, 0, 0);
    mbar_wait(&b_full, 0);
    tc_fence_after();
    ZsRing st = {0, 0u};      // halo stage being consumed
    ZsRing open = {0, 0u};    // TMEM block of the NEXT output plane to be opened (planes are opened and completed in order)
    ZsRing done = {0, 0u};    // TMEM block of the next output plane to complete
    int slot_a = 0;           // TMEM block of plane `a` (the oldest plane the current input plane touches)
    long long w_afull = 0, w_tempty = 0, t_begin = dbg_clock();
    ZsSeg sg;
    while (walk.next(sg)) {
      const int z0 = sg.z0, z1 = sg.z1;
      const int zin0 = z0 > 0 ? z0 - 1 : 0, zin1 = z1 < D ? z1 : D - 1;
      slot_a = open.idx;  // the segment's first output plane is the next one to be opened
      int a_prev = z0;
      int opened = z0;     // output planes [z0, opened) have been opened
      for (int zin = zin0; zin <= zin1; ++zin) {
        // output planes this input plane contributes to: [a, b] = [zin-1, zin+1] clipped to the segment
        const int a = zin - 1 > z0 ? zin - 1 : z0;
        const int b = zin + 1 < z1 - 1 ? zin + 1 : z1 - 1;
        const int m = b - a + 1;
        if (a != a_prev) {  // a advances by at most one plane per input plane
          if (++slot_a == R) slot_a = 0;
          a_prev = a;
        }
        // fresh blocks = planes not opened yet: [opened, b]; f = index of the first one inside [a, b] (m = none)
        const int f = opened - a;  // opened >= a always (plane a was opened by an earlier input plane or is opened now)
        {
          const long long c0 = dbg_clock();
          for (int zo = opened; zo <= b; ++zo) {  // the block must have been drained by the epilogue of the plane that used it R planes ago
            mbar_wait(&tmem_empty[open.idx], open.ph ^ 1u);
            open.step(R);
          }
          opened = b + 1;
          w_tempty += dbg_clock() - c0;
        }
        tc_fence_after();
        ZsRun rf[3], rr[3];
        bool one_run;
        if (m == 3 && f == 2 && slot_a + 3 <= R) {
          // steady state: blocks z-1, z (open) and z+1 (fresh) are contiguous: ONE instruction of N = 3*C_out per step; the first
          // step is an N = 2*C_out accumulate plus an N = C_out overwrite
          const uint32_t t0 = tmem_base + (uint32_t)(slot_a * p.NT);
          rr[0].tacc = t0; rr[0].boff = 0u; rr[0].idesc = idesc3; rr[0].accum = 1u;
          rr[1].idesc = 0u; rr[2].idesc = 0u; rr[1].tacc = rr[2].tacc = t0; rr[1].boff = rr[2].boff = 0u; rr[1].accum = rr[2].accum = 1u;
          rf[0].tacc = t0; rf[0].boff = 0u; rf[0].idesc = idesc2; rf[0].accum = 1u;
          rf[1].tacc = t0 + (uint32_t)(2 * p.NT); rf[1].boff = 2u * blk16; rf[1].idesc = idesc1; rf[1].accum = 0u;
          rf[2].idesc = 0u; rf[2].tacc = t0; rf[2].boff = 0u; rf[2].accum = 1u;
          one_run = true;
        } else {
          // w: first block index at which the ring wraps (m = no wrap inside this range).  Runs = maximal ranges of blocks that one
          // MMA can cover: cut at the ring wrap; the FIRST (tap, k) step is also cut where the accumulate flag changes (open blocks
          // accumulate, fresh blocks are overwritten)
          int w = R - slot_a;
          if (w > m) w = m;
          const int c1 = w < f ? w : f, c2 = w < f ? f : w;   // sorted cuts (m = none)
          const int bf[4] = {0, c1, c2, m}, br[4] = {0, w, m, m};
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            const int s0 = bf[k], len = bf[k + 1] - bf[k];
            int sl = slot_a + s0;
            if (sl >= R) sl -= R;
            rf[k].tacc = tmem_base + (uint32_t)(sl * p.NT);
            rf[k].boff = (uint32_t)(a - zin + 1 + s0) * blk16;   // tdr of block i: (a + i) - zin + 1
            rf[k].idesc = len <= 0 ? 0u : (len == 1 ? idesc1 : (len == 2 ? idesc2 : idesc3));
            rf[k].accum = s0 >= f ? 0u : 1u;
            const int s1 = br[k], len1 = br[k + 1] - br[k];
            int sl1 = slot_a + s1;
            if (sl1 >= R) sl1 -= R;
            rr[k].tacc = tmem_base + (uint32_t)(sl1 * p.NT);
            rr[k].boff = (uint32_t)(a - zin + 1 + s1) * blk16;
            rr[k].idesc = len1 <= 0 ? 0u : (len1 == 1 ? idesc1 : (len1 == 2 ? idesc2 : idesc3));
            rr[k].accum = 1u;
          }
          one_run = w >= m;
        }
        uint32_t b_lo = ((sB0 >> 4) & 0x3FFFu) | lo_lbo;
        for (int j = 0; j < nchunks; ++j) {
          const long long c1 = dbg_clock();
          mbar_wait(&a_full[st.idx], st.ph);
          w_afull += dbg_clock() - c1;
          tc_fence_after();
          const uint32_t a_lo = ((smem_u32(smemA_lane + (size_t)st.idx * p.a_bytes) >> 4) & 0x3FFFu) | lo_lbo;
          if (!DBG_FLAG(p, 8)) {
            if (j == 0)  // first (tap, k) step of the plane: per-block accumulate flags
              zs_issue(rf, hiA | (uint64_t)a_lo, hiB | (uint64_t)b_lo, true);
            if (one_run) zs_issue_chunk<KC, true>(rr, a_lo, b_lo, b_t9, hiA, hiB, j == 0);
            else zs_issue_chunk<KC, false>(rr, a_lo, b_lo, b_t9, hiA, hiB, j == 0);
          }
          umma_commit_elect(&a_empty[st.idx]);
          st.step(S);
          b_lo += chunkB16;
        }
        // output planes that received their last contribution: zin-1 always (if in the segment); plane D-1 when zin == D-1
        if (zin - 1 >= z0) {
          umma_commit_elect(&tmem_full[done.idx]);
          done.step(R);
        }
        if (zin == D - 1 && D - 1 < z1) {
          umma_commit_elect(&tmem_full[done.idx]);
          done.step(R);
        }
      }
    }
#ifdef B200_DEBUG
    if (p.dbg && lane == 0 && lane_id == 0) {
      long long* o = p.dbg + (size_t)(blockIdx.y * gridDim.x + blockIdx.x) * 16;
      o[2] = w_afull;
      o[3] = w_tempty;
      o[4] = dbg_clock() - t_begin;
      o[7] = planes_mine * ZS_LANES;
    }
#else
    (void)w_afull; (void)w_tempty; (void)t_begin;
#endif
  } else {
    // ================= epilogue: the first 4*EW warps drain lane 0's planes, the next 4*EW lane 1's =================
    constexpr int SW = EW == 2 ? 16 : 32;  // slab width: columns one warp handles at a time
    const int wg = (warp >> 2) % EW;        // warpgroup of the lane: takes slabs wg, wg + EW, ...
    const int qd = warp & 3;                // TMEM lane quarter
    const int row = qd * 32 + lane;
    const int bx = row % ZS_BW, by = row / ZS_BW;
    const int NT = p.NT;
    const bool drains = wg * SW < NT;       // (NT = 16 with EW = 2: the second warpgroup has no slab and never arrives)
    const bool reg_stats = p.pmode != 0 && NT <= 32;   // then a warp meets exactly one slab per plane
    float rs[SW], rq[SW];
#pragma unroll
    for (int i = 0; i < SW; ++i) rs[i] = rq[i] = 0.f;
    float* my_acc = stat_acc + (size_t)warp * NT * 2;
    ZsRing cur = {0, 0u};   // TMEM block of the plane being drained
    long long w_tfull = 0, t_ld = 0, t_begin = dbg_clock();
    const size_t HW = (size_t)p.H * p.W;
    ZsSeg sg;
    while (drains && walk.next(sg)) {
      const int th_i = sg.col / p.tilesW, tw_i = sg.col - th_i * p.tilesW;
      const int xh = th_i * ZS_BH + by, xw = tw_i * ZS_BW + bx;
      const bool valid = xh < p.H && xw < p.W;
      const size_t vox_hw = (size_t)n * D * HW + (size_t)xh * p.W + xw;
      for (int zo = sg.z0; zo < sg.z1; ++zo, cur.step(R)) {
        const int slot = cur.idx;
        const size_t vox_off = vox_hw + (size_t)zo * HW;
        const float* bias_row = nullptr;
        if (p.n_b && valid) {
          const int cls = conv_bias_cls(p.cls_mode, zo, xh, xw, D, p.H, p.W);
          const bool interior = (cls & 0x15) == 0x15;  // every axis class is 1 or 3
          bias_row = interior ? bias_interior + (((cls >> 3) & 4) | ((cls >> 2) & 2) | ((cls >> 1) & 1)) * NT
                              : p.biascls + ((size_t)(p.n_b > 1 ? n : 0) * 64 + cls) * p.Cout + n0;
        }
        const long long cw0 = dbg_clock();
        mbar_wait(&tmem_full[slot], cur.ph);
        w_tfull += dbg_clock() - cw0;
        __syncwarp();
        tc_fence_after();
        const uint32_t taddr = tmem_base + (uint32_t)(slot * NT) + ((uint32_t)(qd * 32) << 16);
        for (int c0 = wg * SW; c0 < NT; c0 += EW * SW) {
          const bool wide = SW == 32 && c0 + 32 <= NT;  // else 16 columns (a slab of the two-warpgroup split, or the tail of NT = 48)
          uint32_t raw[SW];
          const long long cl0 = dbg_clock();
          if constexpr (SW == 32) {
            if (wide) tmem_ld_32x32b_x32(taddr + c0, raw);
            else tmem_ld_32x32b_x16(taddr + c0, raw);
          } else {
            tmem_ld_32x32b_x16(taddr + c0, raw);
          }
          tmem_ld_wait();
          t_ld += dbg_clock() - cl0;
          if (DBG_FLAG(p, 4)) continue;
          const int cw = wide ? 32 : 16;
          float v[SW];
#pragma unroll
          for (int i = 0; i < SW; ++i) v[i] = i < cw ? __uint_as_float(raw[i]) : 0.f;
          const size_t goff = vox_off * p.Cout + n0 + c0;
          if (valid) {
            if (bias_row) {
              const float4* bp = reinterpret_cast<const float4*>(bias_row + c0);
#pragma unroll
              for (int i = 0; i < SW / 4; ++i)
                if (4 * i < cw) {
                  const float4 bb = bp[i];
                  v[4 * i] += bb.x; v[4 * i + 1] += bb.y; v[4 * i + 2] += bb.z; v[4 * i + 3] += bb.w;
                }
            }
            if (p.residual) {
              const bf16x8* rp = reinterpret_cast<const bf16x8*>(p.residual + goff);
#pragma unroll
              for (int i = 0; i < SW / 8; ++i)
                if (8 * i < cw) {
                  float f[8];
                  unpack8(rp[i], f);
#pragma unroll
                  for (int j = 0; j < 8; ++j) v[8 * i + j] += f[j];
                }
            }
            if (p.act == B200_ACT_RELU) {
#pragma unroll
              for (int i = 0; i < SW; ++i) v[i] = fmaxf(v[i], 0.f);
            } else if (p.act == B200_ACT_LEAKY) {
#pragma unroll
              for (int i = 0; i < SW; ++i) v[i] = v[i] > 0.f ? v[i] : v[i] * p.slope;
            } else if (p.act == B200_ACT_ELU) {
#pragma unroll
              for (int i = 0; i < SW; ++i) v[i] = v[i] > 0.f ? v[i] : expm1f(v[i]);
            }
#pragma unroll
            for (int i = 0; i < SW; ++i) v[i] = bf16_round(v[i]);
            bf16x8* op = reinterpret_cast<bf16x8*>(p.y + goff);
            if (!DBG_FLAG(p, 1)) {
#pragma unroll
              for (int i = 0; i < SW / 8; ++i)
                if (8 * i < cw) op[i] = pack8(&v[8 * i]);
            }
          } else {
#pragma unroll
            for (int i = 0; i < SW; ++i) v[i] = 0.f;
          }
          if (p.pmode) {
            if (reg_stats) {  // C_out <= 32: one slab per warp; per-thread accumulators across all planes of the CTA
#pragma unroll
              for (int i = 0; i < SW; ++i) {
                rs[i] += v[i];
                rq[i] += v[i] * v[i];
              }
            } else {          // wider: reduce over the warp's 32 rows now, accumulate per warp in shared memory
              float w[SW];
#pragma unroll
              for (int i = 0; i < SW; ++i) w[i] = v[i] * v[i];
              const float sum = warp_reduce_scatter<SW>(v, lane);
              const float qq = warp_reduce_scatter<SW>(w, lane);
              if (lane < cw) {
                my_acc[(c0 + lane) * 2] += sum;
                my_acc[(c0 + lane) * 2 + 1] += qq;
              }
            }
          }
        }
        // block drained -> hand it back to the MMA warp
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty[slot]);
      }
    }
#ifdef B200_DEBUG
    if (p.dbg && threadIdx.x == 0) {
      long long* o = p.dbg + (size_t)(blockIdx.y * gridDim.x + blockIdx.x) * 16;
      o[5] = w_tfull;
      o[6] = dbg_clock() - t_begin;
      o[8] = t_ld;
    }
#else
    (void)w_tfull; (void)t_ld; (void)t_begin;
#endif
    if (reg_stats && drains) {
      const int c0 = wg * SW;  // the warp's only slab
      const float sum = warp_reduce_scatter<SW>(rs, lane);
      const float qq = warp_reduce_scatter<SW>(rq, lane);
      if (lane < SW && c0 + lane < NT) {
        my_acc[(c0 + lane) * 2] = sum;
        my_acc[(c0 + lane) * 2 + 1] = qq;
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (p.pmode) {  // fixed-order sum over the epilogue warps -> this CTA's partial row
    float* out = p.partials + (((size_t)n * cps + cta) * p.Cout + n0) * 2;
    for (int i = threadIdx.x; i < p.NT * 2; i += THREADS) {
      float acc = 0.f;
#pragma unroll
      for (int w = 0; w < NEW; ++w) acc += stat_acc[(size_t)w * p.NT * 2 + i];
      out[i] = acc;
    }
  }
  if (warp == WARP_MMA) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

static bool zs_enabled() {  // B200UNET_ZS=0 falls back to conv3_halo_kernel / the tap-loop kernel (read per call: tests toggle it)
  const char* e = getenv("B200UNET_ZS");
  return !(e && e[0] == '0');
}

// decides whether the z-stacked kernel takes this layer and fills the plan
bool conv_zs_plan(int N, int D, int H, int W, int Cin, int Cout, ConvParams* pp) {
  ConvParams& p = *pp;
  memset(&p, 0, sizeof(p));
  if (!zs_enabled()) return false;
  if (Cin % 16 != 0 || Cout % 16 != 0) return false;
  if (H < ZS_HH || W < ZS_HW || D < 1) return false;  // keep the TMA box inside the tensor extent
  const int budget = 222 * 1024;
  const int kc = (Cin % 64 == 0) ? 64 : (Cin % 32 == 0 ? 32 : 16);
  const int a_bytes = (ZS_ROWS * kc * 2 + 1023) & ~1023;
  // NT = output channels per CTA: all of them when their resident weights fit and one MMA can cover three NT-wide blocks (N <= 256,
  // TMEM ring of >= 8 blocks); else HALF of them (>= 32) when that fits -- the slices re-read the input from L2, which
  // still beats the tap-loop kernel by ~2x on these shapes (64->64 at 64^3 / 160^3).  Wider layers stay on the tap-loop kernel
  // (N = C_out >= 128 is math-bound there).
  int NT = 0, stages = 0, b_total = 0;
  for (int nt : {Cout, 64, 48, 32}) {
    if (nt > Cout || Cout % nt != 0 || nt > 64 || (nt != Cout && (nt < 32 || 2 * nt < Cout))) continue;  // at most two slices
    const int bt = 27 * nt * Cin * 2;
    const int scratch = (16 * nt * 2 + 8 * nt) * (int)sizeof(float);  // per-warp statistics rows (up to 16 epilogue warps) + interior bias variants
    int st = (budget - ((bt + 1023) & ~1023) - scratch - 1024) / a_bytes;
    if (st > ZS_MAX_STAGES) st = ZS_MAX_STAGES;
    st &= ~1;  // two lanes, half of the stages each
    if (st < 4) continue;
    NT = nt; stages = st; b_total = bt;
    break;
  }
  if (!NT) return false;
  p.N = N; p.D = D; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout;
  p.BD = 1; p.BH = ZS_BH; p.BW = ZS_BW;
  p.tilesD = D;
  p.tilesH = (H + ZS_BH - 1) / ZS_BH;
  p.tilesW = (W + ZS_BW - 1) / ZS_BW;
  p.NT = NT;
  p.KC = kc;
  p.KCb = kc;
  p.kchunks = Cin / kc;
  p.a_stages = stages;
  p.a_bytes = a_bytes;
  p.b_total_bytes = b_total;
  int slots = 512 / NT;
  if (slots > ZS_MAX_SLOTS) slots = ZS_MAX_SLOTS;
  slots &= ~1;  // two lanes, half of the ring each: three blocks accumulating + one being drained
  if (slots < 8) return false;
  p.tmem_bufs = slots;
  p.tmem_cols = 512;
  long long T = (long long)p.tilesH * p.tilesW * D;
  if (T >= (1ll << 30)) return false;  // the kernel's plane-tile counters are 32-bit
  int cps = sm_count() / (N * (Cout / NT));  // one persistent CTA per SM over (sample, channel slice)
  if (cps < 1) cps = 1;
  if (const char* e = getenv("B200UNET_ZS_CTAS")) {  // tests: few CTAs per sample => long depth walks (ring wrap, mid-column segment cuts)
    const int v = atoi(e);
    if (v >= 1) cps = v;
  }
  if ((long long)cps * ZS_LANES > T) cps = (int)((T + ZS_LANES - 1) / ZS_LANES);
  p.ctas_per_sample = cps;
  return true;
}

#ifdef B200_DEBUG
long long* get_debug_buffer();
#endif

int conv_zs_launch(const void* x, const void* wf, ConvParams& p, cudaStream_t s) {
#ifdef B200_DEBUG
  p.dbg = get_debug_buffer();
  const char* fl = getenv("B200UNET_DBG_FLAGS");
  p.dbg_flags = fl ? atoi(fl) : 0;
#endif
  CUtensorMap tmA, tmB;
  int rc = make_act_tmap(&tmA, x, p.N, p.D, p.H, p.W, p.Cin, p.KC, 1, ZS_HH, ZS_HW);
  if (rc) return rc;
  rc = make_w_tmap(&tmB, wf, 27 * p.n_w, p.Cout, p.Cin, p.KC, p.NT, 1);
  if (rc) return rc;
  size_t smem = (size_t)((p.b_total_bytes + 1023) & ~1023) + (size_t)p.a_stages * p.a_bytes + (size_t)(16 * p.NT * 2 + 8 * p.NT) * sizeof(float) + 1024;
  const char* e1 = getenv("B200UNET_ZS_EPI");  // "1": one epilogue warpgroup per lane (the round-2 first version)
  const int ew = (e1 && e1[0] == '1') ? 1 : 2;
  auto kern = ew == 2 ? (p.KC == 64 ? conv3_zs_kernel<64, 2> : (p.KC == 32 ? conv3_zs_kernel<32, 2> : conv3_zs_kernel<16, 2>))
                      : (p.KC == 64 ? conv3_zs_kernel<64, 1> : (p.KC == 32 ? conv3_zs_kernel<32, 1> : conv3_zs_kernel<16, 1>));
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  B200_CHECK_ARG(e == cudaSuccess, "conv3_zs: cudaFuncSetAttribute(%zu) failed: %s", smem, cudaGetErrorString(e));
  dim3 grid((unsigned)p.ctas_per_sample, (unsigned)p.N, (unsigned)(p.Cout / p.NT));
  kern<<<grid, (8 * ew + 4) * 32, smem, s>>>(tmA, tmB, p);
  B200_CHECK_LAUNCH("conv3_zs");
  return 0;
}

}  // namespace b200
