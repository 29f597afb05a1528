// Shared pieces of the z-stacked kernels (conv_zs_sm100.cu: 3x3x3 conv; upzs_sm100.cu: the virtual-concat phase conv): tile geometry,
// run descriptors of the issue loop, the plane-tile walk and the incremental ring bookkeeping.
#pragma once
#include "conv_common.cuh"

namespace b200 {

constexpr int ZS_BH = 16, ZS_BW = 8;
constexpr int ZS_HH = ZS_BH + 2, ZS_HW = ZS_BW + 2;
constexpr int ZS_ROWS = ZS_HH * ZS_HW;  // 180 rows of one input-plane halo tile
// two independent LANES per CTA, each = one epilogue warpgroup + one TMA producer warp + one MMA issuer warp + half of the halo stages + half of
// the TMEM ring, walking its own half of the CTA's plane-tiles.  (One issuing warp sustains an MMA per ~55-80 cycles; two warps
// issuing into DIFFERENT accumulators keep the tensor pipe fed -- and, unlike two warps sharing an accumulator, leave the fp32
// accumulation order, hence the result bits, independent of warp timing.)
constexpr int ZS_LANES = 2;
constexpr int ZS_THREADS = 2 * 128 + 4 * 32;
constexpr int ZS_WARP_PRODUCER = 8, ZS_WARP_MMA = 10;  // warps 8,9 producers; 10,11 issuers (lane = warp & 1)
constexpr int ZS_MAX_STAGES = 8;
constexpr int ZS_MAX_SLOTS = 16;

struct ZsRun {
  uint32_t tacc;   // TMEM column address of the first block
  uint32_t boff;   // offset of the first weight block, 16-byte units
  uint32_t idesc;  // instruction descriptor for N = blocks * C_out
  uint32_t accum;  // accumulate flag of the FIRST (tap, k) step of the input plane (always 1 afterwards)
};

// one (tap, k) step = one MMA per non-empty run (idesc == 0 marks an empty run) -- general path (ring wrap, first step of a plane)
__device__ __forceinline__ void zs_issue(const ZsRun (&r)[3], uint64_t adesc, uint64_t bdesc, bool first) {
#pragma unroll
  for (int k = 0; k < 3; ++k)
    if (r[k].idesc) umma_bf16_elect(r[k].tacc, adesc, bdesc + r[k].boff, r[k].idesc, first ? r[k].accum : 1u);
}

// all 9 in-plane taps x KC/16 k-steps of one halo chunk.  b_lo points at [t9 = 0][tdr = 0] of this chunk; one t9 advances 3 blocks.
// `skip_first`: the (t9 = 0, k = 0) step of the plane's first chunk is issued separately (its accumulate flags differ per block).
// ONE_RUN: the plane's blocks are contiguous in TMEM (no ring wrap): one instruction per step, descriptors advance by immediates.
template <int KC, bool ONE_RUN>
__device__ __forceinline__ void zs_issue_chunk(const ZsRun (&rr)[3], uint32_t a_lo, uint32_t b_lo, uint32_t b_t9, uint64_t hiA, uint64_t hiB,
                                               bool skip_first) {
  constexpr uint32_t RB16 = KC * 2 / 16;  // one halo row in 16-byte units
  if (ONE_RUN) b_lo += rr[0].boff;
#pragma unroll
  for (int t9 = 0; t9 < 9; ++t9) {
    const uint32_t offA = (uint32_t)((t9 / 3) * ZS_HW + t9 % 3) * RB16;
#pragma unroll
    for (int k = 0; k < KC / 16; ++k) {
      const uint64_t adesc = hiA | (uint64_t)(a_lo + offA + 2u * k);
      const uint64_t bdesc = hiB | (uint64_t)(b_lo + 2u * k);
      if (t9 == 0 && k == 0) {
        if (!skip_first) {
          if (ONE_RUN) umma_bf16_elect(rr[0].tacc, adesc, bdesc, rr[0].idesc, 1u);
          else zs_issue(rr, adesc, bdesc, false);
        }
      } else {
        if (ONE_RUN) umma_bf16_elect(rr[0].tacc, adesc, bdesc, rr[0].idesc, 1u);
        else zs_issue(rr, adesc, bdesc, false);
      }
    }
    b_lo += b_t9;
  }
}

// The CTA's share of the sample: plane-tiles [L0, L1) in the order (column = th*tilesW + tw, then depth), cut into segments that stay
// inside one column.  All per-plane bookkeeping below is 32-bit and incremental (ring indices and mbarrier phase bits are carried,
// never recomputed with divisions: a 64-bit division costs hundreds of cycles and the issuing warp has ~56 per instruction).
struct ZsSeg {
  int col, z0, z1;
};
struct ZsWalk {
  int L, L1, D;
  __device__ __forceinline__ bool next(ZsSeg& s) {
    if (L >= L1) return false;
    s.col = L / D;
    s.z0 = L - s.col * D;
    const int rest = L1 - L;
    s.z1 = rest < D - s.z0 ? s.z0 + rest : D;
    L += s.z1 - s.z0;
    return true;
  }
};
__device__ __forceinline__ ZsWalk zs_walk(const ConvParams& p, int cta, int cps) {
  const long long T = (long long)p.tilesH * p.tilesW * p.D;  // < 2^31 (checked by the plan)
  ZsWalk w;
  w.L = (int)(T * cta / cps);
  w.L1 = (int)(T * (cta + 1) / cps);
  w.D = p.D;
  return w;
}
// ring position: index + phase bit of its current use, advanced one step at a time
struct ZsRing {
  int idx;
  uint32_t ph;
  __device__ __forceinline__ void step(int n) {
    if (++idx == n) {
      idx = 0;
      ph ^= 1u;
    }
  }
};

}  // namespace b200
