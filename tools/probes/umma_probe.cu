// Hardware probe (test tooling, not on the product path): does tcgen05.mma accept a K-major SWIZZLE_128B A operand
// whose start address is shifted by whole 128-byte rows (not a multiple of the 1024-byte swizzle pattern) and whose
// 8-row groups sit at a stride that is not a multiple of 1024 bytes?  If the swizzle XOR is a pure function of the
// shared-memory address bits (as it is for TMA writes), shifted views of ONE halo tile can serve all 27 filter taps.
//   D[r][n] = sum_k A[row(r)][k] * B[n][k],  row(r) = shift + (r / 8) * group_rows + (r % 8)
#include <string.h>

#include "common.cuh"
#include "sm100_ptx.cuh"
#include "b200probe.h"

namespace b200 {

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* t, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                   smem_u32(dst)),
               "l"(reinterpret_cast<uint64_t>(t)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
               : "memory");
}

__global__ void __launch_bounds__(128) umma_probe_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                         int rows, int shift, int group_rows, float* __restrict__ D) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar, done_bar;
  __shared__ uint32_t tmem_slot;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;                                   // rows x 128 B
  uint8_t* sB = smem + (((size_t)rows * 128 + 1023) & ~(size_t)1023);  // 16 x 128 B
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&full_bar, 1);
    mbar_init(&done_bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&tmem_slot, 32);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(&full_bar, (uint32_t)(rows * 128 + 16 * 128));
    // TMA boxes are limited to 256 rows
    tma_load_2d(sA, &tmA, &full_bar, 0, 0);
    tma_load_2d(sB, &tmB, &full_bar, 0, 0);
    mbar_wait(&full_bar, 0);
    tc_fence_after();
    const uint32_t idesc = umma_idesc_bf16(128, 16, 0, 0);
    for (int k = 0; k < 4; ++k) {
      const uint64_t adesc = umma_smem_desc(smem_u32(sA) + (uint32_t)(shift * 128 + k * 32), 16u, (uint32_t)(group_rows * 128), UMMA_LAYOUT_SW128);
      const uint64_t bdesc = umma_smem_desc(smem_u32(sB) + (uint32_t)(k * 32), 16u, 1024u, UMMA_LAYOUT_SW128);
      umma_bf16(tmem_base, adesc, bdesc, idesc, k != 0 ? 1u : 0u);
    }
    umma_commit(&done_bar);
  }
  mbar_wait(&done_bar, 0);
  __syncwarp();
  tc_fence_after();
  uint32_t raw[16];
  tmem_ld_32x32b_x16(tmem_base + ((uint32_t)(warp * 32) << 16), raw);
  tmem_ld_wait();
  for (int i = 0; i < 16; ++i) D[(size_t)(warp * 32 + lane) * 16 + i] = __uint_as_float(raw[i]);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, 32);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode_tiled();

}  // namespace b200

using namespace b200;

extern "C" int b200_probe_umma_rowshift(const void* A, int rows, const void* B, int shift, int group_rows, float* D, b200_stream_t s) {
  B200_CHECK_ARG(rows <= 256 && shift + 15 * group_rows + 8 <= rows, "probe: rows=%d too small for shift=%d group_rows=%d", rows, shift,
                 group_rows);
  EncodeTiledFn enc = get_encode_tiled();
  B200_CHECK_ARG(enc, "cuTensorMapEncodeTiled entry point not available");
  CUtensorMap tmA, tmB;
  cuuint64_t dA[2] = {64, (cuuint64_t)rows}, dB[2] = {64, 16};
  cuuint64_t st[1] = {128};
  cuuint32_t bA[2] = {64, (cuuint32_t)rows}, bB[2] = {64, 16}, es[2] = {1, 1};
  CUresult r1 = enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(A), dA, st, bA, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CUresult r2 = enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(B), dB, st, bB, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  B200_CHECK_ARG(r1 == CUDA_SUCCESS && r2 == CUDA_SUCCESS, "probe: tensor map encode failed %d %d", (int)r1, (int)r2);
  size_t smem = (((size_t)rows * 128 + 1023) & ~(size_t)1023) + 2048 + 1024;
  cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  umma_probe_kernel<<<1, 128, smem, (cudaStream_t)s>>>(tmA, tmB, rows, shift, group_rows, D);
  B200_CHECK_LAUNCH("umma_probe");
  return 0;
}

// ---- probe 2: tcgen05.mma issue/accumulate throughput -----------------------------------------------------------
// cycles per MMA (M=128, N, K=16, bf16) when `iters*4` MMAs are spread round-robin over `n_acc` independent
// accumulators (TMEM column ranges).  n_acc=1 exposes the latency of the accumulate dependency chain.
namespace b200 {
__global__ void __launch_bounds__(128) umma_issue_probe_kernel(int N, int n_acc, int iters, int rb, int group_rows, int shift, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t done_bar;
  __shared__ uint32_t tmem_slot;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;               // up to 256 rows x 128 B
  uint8_t* sB = smem + 32768;       // N x 128 B
  for (int i = threadIdx.x; i < (32768 + N * 128) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    mbar_init(&done_bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = umma_idesc_bf16(128, N, 0, 0);
    uint64_t ad[4], bd[4];
    const uint32_t lay = umma_layout_for_row_bytes(rb);
    const int ksl = rb / 32;  // K=16 slices available per row
    for (int k = 0; k < 4; ++k) {
      ad[k] = umma_smem_desc(smem_u32(sA) + shift * rb + (k % ksl) * 32, 16u, (uint32_t)(group_rows * rb), lay);
      bd[k] = umma_smem_desc(smem_u32(sB) + (k % ksl) * 32, 16u, (uint32_t)(8 * rb), lay);
    }
    long long t0 = clock64();
    int a = 0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        umma_bf16(tmem_base + (uint32_t)(a * N), ad[k], bd[k], idesc, 1u);
        a = (a + 1 == n_acc) ? 0 : a + 1;
      }
    }
    long long t1 = clock64();
    umma_commit(&done_bar);
    mbar_wait(&done_bar, 0);
    long long t2 = clock64();
    out[0] = t1 - t0;  // issue time
    out[1] = t2 - t0;  // until all complete
  }
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}
}  // namespace b200

extern "C" int b200_probe_umma_issue(int N, int n_acc, int iters, int rb, int group_rows, int shift, long long* out, b200_stream_t s) {
  B200_CHECK_ARG(N % 16 == 0 && N >= 16 && N <= 256 && n_acc >= 1 && n_acc * N <= 512, "probe: bad N=%d n_acc=%d", N, n_acc);
  B200_CHECK_ARG((rb == 32 || rb == 64 || rb == 128) && (shift + 15 * group_rows + 8) * rb <= 32768, "probe: bad view");
  size_t smem = 32768 + (size_t)N * 128 + 2048;
  cudaFuncSetAttribute(umma_issue_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  umma_issue_probe_kernel<<<1, 128, smem, (cudaStream_t)s>>>(N, n_acc, iters, rb, group_rows, shift, out);
  B200_CHECK_LAUNCH("umma_issue_probe");
  return 0;
}

// ---- probe 3: does a stream of tcgen05.mma starve tcgen05.ld of another TMEM column range? ----------------------
// warp 0 (one lane) issues `mma_iters*4` MMAs (N columns at column 0); warps 4..7 time `ld_iters` x (tcgen05.ld 32x32b.x32 +
// wait) of columns [256, 288).  out[0] = cycles per ld with the MMA stream running (mma_iters > 0) or idle (mma_iters == 0).
namespace b200 {
__global__ void __launch_bounds__(256) tmem_ld_contention_kernel(int N, int mma_iters, int ld_iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t done_bar;
  __shared__ uint32_t tmem_slot;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + 16384;
  for (int i = threadIdx.x; i < (16384 + N * 128) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&done_bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  if (warp == 0) {
    if (lane == 0 && mma_iters > 0) {
      const uint32_t idesc = umma_idesc_bf16(128, N, 0, 0);
      uint64_t ad[4], bd[4];
      for (int k = 0; k < 4; ++k) {
        ad[k] = umma_smem_desc(smem_u32(sA) + k * 32, 16u, 1024u, UMMA_LAYOUT_SW128);
        bd[k] = umma_smem_desc(smem_u32(sB) + k * 32, 16u, 1024u, UMMA_LAYOUT_SW128);
      }
      long long t0 = clock64();
      for (int it = 0; it < mma_iters; ++it) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tmem_base, ad[k], bd[k], idesc, 1u);
      }
      umma_commit(&done_bar);
      mbar_wait(&done_bar, 0);
      out[1] = (clock64() - t0);
    }
  } else if (warp >= 4) {
    const uint32_t taddr = tmem_base + 256 + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t raw[32];
    float acc = 0.f;
    __syncwarp();
    long long t0 = clock64();
    for (int i = 0; i < ld_iters; ++i) {
      tmem_ld_32x32b_x32(taddr, raw);
      tmem_ld_wait();
      acc += __uint_as_float(raw[i & 31]);
    }
    long long t1 = clock64();
    if (threadIdx.x == 128) out[0] = t1 - t0;
    if (acc == 123.456f) out[2] = 1;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}
}  // namespace b200

extern "C" int b200_probe_tmem_ld_contention(int N, int mma_iters, int ld_iters, long long* out, b200_stream_t s) {
  size_t smem = 16384 + (size_t)N * 128 + 2048;
  cudaFuncSetAttribute(b200::tmem_ld_contention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  b200::tmem_ld_contention_kernel<<<1, 256, smem, (cudaStream_t)s>>>(N, mma_iters, ld_iters, out);
  B200_CHECK_LAUNCH("tmem_ld_contention");
  return 0;
}

// ---- probe 4: is the per-MMA dispatch floor per issuing warp / per CTA, or per SM? --------------------------------
// `n_issuers` warps of one CTA each issue iters*4 MMAs (M=128, N, K=16) into their own accumulator; `grid` CTAs run at once
// (grid = 2 x #SMs with 256 TMEM columns each puts two CTAs on every SM).  out[cta*4 + w] = cycles until warp w's MMAs completed.
namespace b200 {
__global__ void __launch_bounds__(128) umma_multi_issue_probe_kernel(int N, int n_issuers, int iters, int tmem_cols, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t done_bar[4];
  __shared__ uint32_t tmem_slot;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;          // 128 rows x 128 B
  uint8_t* sB = smem + 16384;  // N x 128 B
  for (int i = threadIdx.x; i < (16384 + N * 128) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(&done_bar[i], 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&tmem_slot, (uint32_t)tmem_cols);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  if (warp < n_issuers && lane == 0) {
    const uint32_t idesc = umma_idesc_bf16(128, N, 0, 0);
    uint64_t ad[4], bd[4];
    for (int k = 0; k < 4; ++k) {
      ad[k] = umma_smem_desc(smem_u32(sA) + k * 32, 16u, 1024u, UMMA_LAYOUT_SW128);
      bd[k] = umma_smem_desc(smem_u32(sB) + k * 32, 16u, 1024u, UMMA_LAYOUT_SW128);
    }
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_bf16(tmem_base + (uint32_t)(warp * N), ad[k], bd[k], idesc, 1u);
    }
    umma_commit(&done_bar[warp]);
    mbar_wait(&done_bar[warp], 0);
    out[blockIdx.x * 4 + warp] = clock64() - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)tmem_cols);
  }
}
}  // namespace b200

extern "C" int b200_probe_umma_multi_issue(int N, int n_issuers, int iters, int grid, long long* out, b200_stream_t s) {
  B200_CHECK_ARG(N % 16 == 0 && N >= 16 && N <= 256 && n_issuers >= 1 && n_issuers <= 4 && n_issuers * N <= 256 && grid >= 1,
                 "probe: bad N=%d n_issuers=%d", N, n_issuers);
  size_t smem = 16384 + (size_t)N * 128 + 2048;
  cudaFuncSetAttribute(b200::umma_multi_issue_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  b200::umma_multi_issue_probe_kernel<<<grid, 128, smem, (cudaStream_t)s>>>(N, n_issuers, iters, 256, out);
  B200_CHECK_LAUNCH("umma_multi_issue_probe");
  return 0;
}
