// Thread mapping + partial-sum helpers shared by the HBM-bound vector kernels and the direct conv kernels.
#pragma once
#include "common.cuh"

namespace b200 {

// ------------------------------------------------------------------------------------------------
// thread mapping shared by the vector kernels: a block of 256 threads = CG channel groups (8 ch each)
// x VL voxel lanes; consecutive threads touch consecutive 16-byte chunks of one voxel (coalesced).
// ------------------------------------------------------------------------------------------------
constexpr int EW_THREADS = 256;
struct EwMap {
  int CG, VL, cg, vl;
  bool active;
};
__device__ __forceinline__ EwMap ew_map(int C) {
  EwMap m;
  m.CG = C >> 3;
  m.VL = EW_THREADS / m.CG;
  m.cg = threadIdx.x % m.CG;
  m.vl = threadIdx.x / m.CG;
  m.active = m.vl < m.VL;
  return m;
}
// blocks per sample for a pass over `voxels` voxels of C channels
static inline int ew_blocks(long long voxels, int C) {
  int CG = C / 8;
  int VL = EW_THREADS / CG;
  if (VL < 1) VL = 1;
  long long per_block = (long long)VL * 16;  // >= 16 voxels per lane
  long long p = (voxels + per_block - 1) / per_block;
  if (p > 1024) p = 1024;
  if (p < 1) p = 1;
  return (int)p;
}
// [v0,v1) voxel range of block p out of P
__device__ __forceinline__ void ew_range(long long voxels, int p, int P, long long& v0, long long& v1) {
  long long per = (voxels + P - 1) / P;
  v0 = (long long)p * per;
  v1 = v0 + per;
  if (v1 > voxels) v1 = voxels;
  if (v0 > voxels) v0 = voxels;
}
// line-structured iteration for kernels that need (d,h,w): the voxel lanes of a block are split into LPB groups of `lpl` lanes;
// a group walks one (d,h) line at a time (32-bit index math, no per-voxel division), lane lw covering w = lw, lw+lpl, ...
struct LineMap {
  int lpl, LPB, ls, lw;
  bool active;
};
__device__ __forceinline__ LineMap line_map(const EwMap& m, int W) {
  LineMap lm;
  lm.lpl = m.VL < W ? m.VL : W;
  if (lm.lpl < 1) lm.lpl = 1;
  lm.LPB = m.VL / lm.lpl;
  if (lm.LPB < 1) lm.LPB = 1;
  lm.ls = m.vl / lm.lpl;
  lm.lw = m.vl - lm.ls * lm.lpl;
  lm.active = m.active && lm.ls < lm.LPB;
  return lm;
}
__device__ __forceinline__ void ew_range_i(int items, int p, int P, int& i0, int& i1) {
  int per = (items + P - 1) / P;
  i0 = p * per;
  i1 = i0 + per;
  if (i1 > items) i1 = items;
  if (i0 > items) i0 = items;
}
// blocks per sample for a pass WITHOUT partial sums (no reason to keep >= 16 voxels per lane): ~4 items per lane
static inline int ew_blocks_dense(long long items, int C) {
  int CG = C / 8;
  int VL = EW_THREADS / CG;
  if (VL < 1) VL = 1;
  long long p = (items + (long long)VL * 4 - 1) / ((long long)VL * 4);
  if (p > 16384) p = 16384;
  if (p < 1) p = 1;
  return (int)p;
}
// reduce per-thread (s[8], q[8]) over the voxel lanes of the block and write [C][2] partials
__device__ __forceinline__ void ew_write_partials(const float s[8], const float q[8], const EwMap& m, float* out /*[C][2]*/,
                                                  float* red /* smem EW_THREADS*16 */) {
  if (m.active) {
    float* r = red + (size_t)(m.vl * m.CG + m.cg) * 16;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      r[i] = s[i];
      r[8 + i] = q[i];
    }
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < m.CG * 16; idx += EW_THREADS) {
    int cg = idx >> 4, i = idx & 15;
    float acc = 0.f;
    for (int vl = 0; vl < m.VL; ++vl) acc += red[(size_t)(vl * m.CG + cg) * 16 + i];
    int c = cg * 8 + (i & 7), k = i >> 3;
    out[c * 2 + k] = acc;
  }
}


}  // namespace b200
