/* Hardware probes (TEST TOOLING, not part of the product library): built into tools/probes/libb200probe.so, which links
 * against libb200unet.so for the TMA-descriptor / error helpers.  Results: profiles/probes_r01.md. */
#ifndef B200PROBE_H_
#define B200PROBE_H_
#include "../../include/b200unet.h"
#ifdef __cplusplus
extern "C" {
#endif
/* hardware probe (test tooling): tcgen05.mma on a row-shifted / odd-strided view of a SWIZZLE_128B tile.
 * A: [rows][64] bf16, B: [16][64] bf16, D: [128][16] f32 with D[r][n] = sum_k A[shift + (r/8)*group_rows + r%8][k] * B[n][k] */
int b200_probe_umma_rowshift(const void* A, int rows, const void* B, int shift, int group_rows, float* D, b200_stream_t s);
/* hardware probe: cycles to issue / complete iters*4 tcgen05.mma (M=128,N,K=16) spread over n_acc accumulators;
 * out[0] = issue cycles, out[1] = cycles until all completed */
/* hardware probe: cycles for ld_iters tcgen05.ld (4 warps, 32 columns each) while mma_iters*4 MMAs (N columns) run; out[0] ld cycles, out[1] mma cycles */
int b200_probe_tmem_ld_contention(int N, int mma_iters, int ld_iters, long long* out, b200_stream_t s);
/* out[cta*4 + w] = cycles until issuing warp w's iters*4 MMAs (own accumulator) completed; grid CTAs, 256 TMEM columns each */
int b200_probe_umma_multi_issue(int N, int n_issuers, int iters, int grid, long long* out, b200_stream_t s);
int b200_probe_umma_issue(int N, int n_acc, int iters, int rb, int group_rows, int shift, long long* out, b200_stream_t s);

#ifdef __cplusplus
}
#endif
#endif
