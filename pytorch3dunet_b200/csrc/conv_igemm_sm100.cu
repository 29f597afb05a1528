// 3x3x3 convolution (padding 1) as an implicit GEMM on the sm_100a tensor cores.
//
//   D[128 voxels, NT out-channels] = sum over (tap, 64-channel chunk) of  A_tap[128, KC] * B_tap[NT, KC]^T
//
//   A_tap : a BDxBHxBW box of the NDHWC activation, shifted by the filter tap, fetched by ONE 5-D TMA tile load
//           whose out-of-bounds elements are zero-filled by the hardware (= the conv's zero padding); the im2col
//           matrix never exists in memory.  Rows land K-major in shared memory with the TMA 128/64/32-byte swizzle.
//   B_tap : [NT, KC] slice of the (GroupNorm-folded, per-sample) weights wf[n][tap][co][ci], one 3-D TMA tile load.
//   MMA   : tcgen05.mma.cta_group::1.kind::f16, M=128, N=NT, K=16, issued by one thread; fp32 accumulators in TMEM.
//   Epilogue (4 warps): tcgen05.ld -> + border-class bias (the GroupNorm shift through zero padding) -> + residual
//           -> activation -> bf16 -> global; per-channel (sum, sum*w) partials for the next GroupNorm / GN backward.
//
// The same kernel is the dgrad (input gradient): x := dz, weights := tap-flipped transposed weights.
// Warp roles: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..5 = epilogue.
#include "conv_common.cuh"

namespace b200 {

__global__ void __launch_bounds__(CONV_THREADS)
conv3_igemm_kernel(const __grid_constant__ CUtensorMap tmapA, const __grid_constant__ CUtensorMap tmapB, const ConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[CONV_MAX_STAGES];
  __shared__ __align__(8) uint64_t empty_bar[CONV_MAX_STAGES];
  __shared__ __align__(8) uint64_t tmem_full_bar[8];  // one per accumulator
  __shared__ uint32_t tmem_slot;

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;  // provably warp-uniform
  const int stage_bytes = p.a_bytes + p.b_bytes;

  // tile coordinates
  int t = blockIdx.x;
  const int tw_i = t % p.tilesW;
  t /= p.tilesW;
  const int th_i = t % p.tilesH;
  t /= p.tilesH;
  const int td_i = t % p.tilesD;
  const int n = t / p.tilesD;
  const int tile_in_sample = blockIdx.x - n * (p.tilesD * p.tilesH * p.tilesW);
  const int d0 = td_i * p.BD, h0 = th_i * p.BH, w0 = tw_i * p.BW;
  const int n0 = blockIdx.y * p.NT;
  const int numK = p.ntaps * p.kchunks;

  if (threadIdx.x == 0) {
    for (int i = 0; i < p.stages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 8; ++i) mbar_init(&tmem_full_bar[i], 1);
    fence_mbar_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmapA);
    tma_prefetch_desc(&tmapB);
  }
  if (warp == 1) tmem_alloc(&tmem_slot, (uint32_t)p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      const int wsample = p.n_w > 1 ? n : 0;
      for (int kb = 0; kb < numK; ++kb) {
        const int stage = kb % p.stages;
        const uint32_t phase = (uint32_t)(kb / p.stages) & 1u;
        mbar_wait(&empty_bar[stage], phase ^ 1u);
        const int tap = kb / p.kchunks, kc = kb - tap * p.kchunks;
        const int od = p.toff[3 * tap], oh = p.toff[3 * tap + 1], ow = p.toff[3 * tap + 2];
        uint8_t* sa = smem + (size_t)stage * stage_bytes;
        uint8_t* sb = sa + p.a_bytes;
        mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)(128 * p.KC * 2 + p.NT * p.KC * 2));
        tma_load_5d(sa, &tmapA, &full_bar[stage], kc * p.KC, p.in_mul * w0 + ow, p.in_mul * h0 + oh, p.in_mul * d0 + od, n);
        tma_load_3d(sb, &tmapB, &full_bar[stage], kc * p.KC, n0, wsample * p.w_rows + p.w_row0 + tap);
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (whole warp converged, one elected lane issues) =================
    {
      const uint32_t idesc = umma_idesc_bf16(128, p.NT, 0, 0);
      const int rb = p.KC * 2;
      const uint64_t hi = umma_smem_desc(0, 16u, 8u * (uint32_t)rb, umma_layout_for_row_bytes(rb)) & 0xFFFFFFFF00000000ull;
      const int ksteps = p.KC / 16;
      int kb = 0;
      for (int acc = 0; acc < p.nacc; ++acc) {
        const uint32_t tacc = tmem_base + (uint32_t)(acc * p.NT);
        const int kb_per_acc = (p.acc_ntaps[acc] ? p.acc_ntaps[acc] : p.ntaps / p.nacc) * p.kchunks;
        uint32_t accum = 0;
        for (int i = 0; i < kb_per_acc; ++i, ++kb) {
          const int stage = kb % p.stages;
          const uint32_t phase = (uint32_t)(kb / p.stages) & 1u;
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
          const uint32_t a_lo = ((sa >> 4) & 0x3FFFu) | (1u << 16);
          const uint32_t b_lo = (((sa + (uint32_t)p.a_bytes) >> 4) & 0x3FFFu) | (1u << 16);
          for (int k = 0; k < ksteps; ++k) {
            umma_bf16_elect(tacc, hi | (uint64_t)(a_lo + 2u * k), hi | (uint64_t)(b_lo + 2u * k), idesc, accum);
            accum = 1u;
          }
          umma_commit_elect(&empty_bar[stage]);  // frees the smem slot once these MMAs have read it
        }
        umma_commit_elect(&tmem_full_bar[acc]);  // this accumulator is complete
      }
    }
  } else {
    // ================= epilogue (warps 2..5) =================
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;
    const int bx = row % p.BW, by = (row / p.BW) % p.BH, bz = row / (p.BW * p.BH);
    const int xd = d0 + bz, xh = h0 + by, xw = w0 + bx;
    const bool valid = xd < p.D && xh < p.H && xw < p.W;
    const float* bias_row = nullptr;
    if (p.n_b && valid) {
      const int cls = conv_bias_cls(p.cls_mode, xd, xh, xw, p.D, p.H, p.W);
      bias_row = p.biascls + ((size_t)(p.n_b > 1 ? n : 0) * 64 + cls) * p.Cout;
    }
    // stats scratch [4][NT][2]: stage 0 of the smem ring, free once the (single) accumulator is complete (pmode needs nacc == 1)
    float* scratch = reinterpret_cast<float*>(smem) + (size_t)q * p.NT * 2;
    for (int acc = 0; acc < p.nacc; ++acc) {
      const int o0 = p.nacc > 1 ? p.acc_off[3 * acc] : p.out_off[0], o1 = p.nacc > 1 ? p.acc_off[3 * acc + 1] : p.out_off[1],
                o2 = p.nacc > 1 ? p.acc_off[3 * acc + 2] : p.out_off[2];
      const size_t vox_off = (size_t)n * p.OD * p.OH * p.OW + ((size_t)(p.out_mul * xd + o0) * p.OH + (p.out_mul * xh + o1)) * p.OW +
                             (p.out_mul * xw + o2);
      mbar_wait(&tmem_full_bar[acc], 0);
      __syncwarp();
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.NT);
      int c0 = 0;
      for (; c0 + 32 <= p.NT; c0 += 32) conv_epilogue_slab<32>(p, taddr, c0, n0, valid, vox_off, bias_row, lane, scratch);
      if (c0 < p.NT) conv_epilogue_slab<16>(p, taddr, c0, n0, valid, vox_off, bias_row, lane, scratch);
    }
    if (p.pmode) {
      asm volatile("bar.sync 1, 128;" ::: "memory");  // the 4 epilogue warps only
      const int et = threadIdx.x - 64;
      const float* sc = reinterpret_cast<const float*>(smem);
      float* out = p.partials + (((size_t)n * (p.tilesD * p.tilesH * p.tilesW) + tile_in_sample) * p.Cout + n0) * 2;
      for (int i = et; i < p.NT * 2; i += 128) out[i] = sc[i] + sc[p.NT * 2 + i] + sc[p.NT * 4 + i] + sc[p.NT * 6 + i];
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

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

static CUtensorMapSwizzle swizzle_for_row_bytes(int rb) {
  return rb == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (rb == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
}

// NDHWC bf16 activation viewed as a rank-5 tensor {C, W, H, D, N}; box {kc, bw, bh, bd, 1}
int make_act_tmap(CUtensorMap* tm, const void* ptr, int N, int D, int H, int W, int C, int kc, int bd, int bh, int bw) {
  EncodeTiledFn enc = get_encode_tiled();
  B200_CHECK_ARG(enc, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N};
  cuuint64_t strides[4] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2, (cuuint64_t)D * H * W * C * 2};
  cuuint32_t box[5] = {(cuuint32_t)kc, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bd, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(tm, B200_TMAP_DTYPE, 5, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for_row_bytes(kc * 2), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  B200_CHECK_ARG(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(activation %dx%dx%dx%dx%d, box %d,%d,%d,%d) failed: %d", N, D, H, W, C, kc,
                 bw, bh, bd, (int)r);
  return 0;
}
// weights [rows2][rows1][C] bf16 viewed as rank-3 {C, rows1, rows2}; box {kc, nt, 1}
int make_w_tmap(CUtensorMap* tm, const void* ptr, int rows2, int rows1, int C, int kc, int nt, int ntaps_box) {
  EncodeTiledFn enc = get_encode_tiled();
  B200_CHECK_ARG(enc, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)rows1, (cuuint64_t)rows2};
  cuuint64_t strides[2] = {(cuuint64_t)C * 2, (cuuint64_t)rows1 * C * 2};
  cuuint32_t box[3] = {(cuuint32_t)kc, (cuuint32_t)nt, (cuuint32_t)ntaps_box};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(tm, B200_TMAP_DTYPE, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for_row_bytes(kc * 2), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  B200_CHECK_ARG(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(weights %dx%dx%d, box %d,%d) failed: %d", rows2, rows1, C, kc, nt, (int)r);
  return 0;
}

// pick the 128-voxel box that wastes the fewest padded voxels; 0 on success
int choose_box(int D, int H, int W, int* bd, int* bh, int* bw) {
  static const int cand[][3] = {{2, 8, 8},  {1, 8, 16}, {4, 4, 8},  {1, 16, 8}, {2, 4, 16}, {4, 8, 4}, {8, 4, 4},  {1, 4, 32},
                                {2, 2, 32}, {1, 2, 64}, {1, 1, 128}, {8, 8, 2}, {16, 8, 1}, {4, 2, 16}, {2, 16, 4}, {8, 2, 8},
                                {16, 4, 2}, {32, 4, 1}, {128, 1, 1}, {1, 128, 1}, {64, 2, 1}, {1, 32, 4}, {1, 64, 2}, {16, 2, 4},
                                {32, 2, 2}, {2, 32, 2}, {2, 64, 1}, {4, 16, 2}, {4, 32, 1}, {8, 16, 1}, {16, 1, 8}, {8, 1, 16},
                                {4, 1, 32}, {2, 1, 64}, {32, 1, 4}, {64, 1, 2}};
  long long best = -1;
  // pass 0: boxes that fit inside the volume; pass 1 (tiny volumes, e.g. the 6^3 bottom level of a 5-level net): any box --
  // the part of a TMA box that lies outside the tensor is zero-filled on load and masked on store like any ragged border
  for (int pass = 0; pass < 2 && best < 0; ++pass) {
    for (auto& c : cand) {
      if (pass == 0 && (c[0] > D || c[1] > H || c[2] > W)) continue;
      long long padded = (long long)((D + c[0] - 1) / c[0]) * c[0] * ((H + c[1] - 1) / c[1]) * c[1] * ((W + c[2] - 1) / c[2]) * c[2];
      if (best < 0 || padded < best) {
        best = padded;
        *bd = c[0];
        *bh = c[1];
        *bw = c[2];
      }
    }
  }
  return best < 0 ? 1 : 0;
}

bool conv_igemm_supported(int N, int D, int H, int W, int Cin, int Cout) {
  (void)N;
  int bd, bh, bw;
  if (Cin % 16 != 0 || Cout % 16 != 0) return false;
  if (choose_box(D, H, W, &bd, &bh, &bw)) return false;
  return true;
}

static int pick_nt(int Cout) {
  if (Cout <= 256) return Cout;
  if (Cout % 256 == 0) return 256;
  if (Cout % 128 == 0) return 128;
  if (Cout % 64 == 0) return 64;
  if (Cout % 32 == 0) return 32;
  return 16;
}

}  // namespace b200

namespace b200 {
// Geometry of a plain (tap-loop) launch.  The defaults describe the 3x3x3 / pad 1 convolution.
struct PlainGeom {
  int ntaps = 27;
  signed char toff[64 * 3];
  int in_mul = 1;                   // 2: the input map subsamples a (2D,2H,2W) tensor with element stride 2
  int out_mul = 1, out_off[3] = {0, 0, 0};  // 2: outputs are written to one parity phase of a (2D,2H,2W) volume
  int w_rows = 27, w_row0 = 0;      // weight rows per sample in wf, first row of this launch
  int cls_mode = 0;
  int nacc = 1;                     // accumulators per CTA (taps split evenly), each with its own output offset acc_off[a]
  signed char acc_off[8 * 3] = {0};
  signed char acc_ntaps[8] = {0};   // uneven split: taps per accumulator (all 0 = even)
  PlainGeom() {
    for (int t = 0; t < 27; ++t) {
      toff[3 * t] = (signed char)(t / 9 - 1);
      toff[3 * t + 1] = (signed char)((t / 3) % 3 - 1);
      toff[3 * t + 2] = (signed char)(t % 3 - 1);
    }
  }
};

// activation map whose boxes subsample with element stride 2 (box of bd x bh x bw ELEMENTS LOADED)
int make_act_tmap_stride2(CUtensorMap* tm, const void* ptr, int N, int D, int H, int W, int C, int kc, int bd, int bh, int bw) {
  EncodeTiledFn enc = get_encode_tiled();
  B200_CHECK_ARG(enc, "cuTensorMapEncodeTiled entry point not available");
  B200_CHECK_ARG(2 * bd <= 256 && 2 * bh <= 256 && 2 * bw <= 256, "stride-2 box %dx%dx%d too large", bd, bh, bw);
  cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N};
  cuuint64_t strides[4] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2, (cuuint64_t)D * H * W * C * 2};
  cuuint32_t box[5] = {(cuuint32_t)kc, (cuuint32_t)(2 * bw), (cuuint32_t)(2 * bh), (cuuint32_t)(2 * bd), 1};
  cuuint32_t estr[5] = {1, 2, 2, 2, 1};
  CUresult r = enc(tm, B200_TMAP_DTYPE, 5, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for_row_bytes(kc * 2), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  B200_CHECK_ARG(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(stride-2 activation %dx%dx%dx%dx%d) failed: %d", N, D, H, W, C, (int)r);
  return 0;
}

// one plane of an activation sampled with element stride 2 in h and w (depth stride 1): box of bh x bw SAMPLES (updzs_sm100.cu)
int make_act_tmap_stride2_hw(CUtensorMap* tm, const void* ptr, int N, int D, int H, int W, int C, int kc, int bh, int bw) {
  EncodeTiledFn enc = get_encode_tiled();
  B200_CHECK_ARG(enc, "cuTensorMapEncodeTiled entry point not available");
  B200_CHECK_ARG(2 * bh <= 256 && 2 * bw <= 256, "stride-2 box %dx%d too large", bh, bw);
  cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N};
  cuuint64_t strides[4] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2, (cuuint64_t)D * H * W * C * 2};
  cuuint32_t box[5] = {(cuuint32_t)kc, (cuuint32_t)(2 * bw), (cuuint32_t)(2 * bh), 1, 1};
  cuuint32_t estr[5] = {1, 2, 2, 1, 1};
  CUresult r = enc(tm, B200_TMAP_DTYPE, 5, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swizzle_for_row_bytes(kc * 2), CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  B200_CHECK_ARG(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(stride-2 plane %dx%dx%dx%dx%d) failed: %d", N, D, H, W, C, (int)r);
  return 0;
}

// one plane of an activation sampled with element stride 2 along w only: box of bh lines x bw SAMPLES (wgrad_up_sm100.cu)
int make_act_tmap_stride2_w(CUtensorMap* tm, const void* ptr, int N, int D, int H, int W, int C, int kc, int bh, int bw) {
  EncodeTiledFn enc = get_encode_tiled();
  B200_CHECK_ARG(enc, "cuTensorMapEncodeTiled entry point not available");
  B200_CHECK_ARG(bh <= 256 && 2 * bw <= 256, "stride-2 box %dx%d too large", bh, bw);
  cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N};
  cuuint64_t strides[4] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2, (cuuint64_t)D * H * W * C * 2};
  cuuint32_t box[5] = {(cuuint32_t)kc, (cuuint32_t)(2 * bw), (cuuint32_t)bh, 1, 1};
  cuuint32_t estr[5] = {1, 2, 1, 1, 1};
  CUresult r = enc(tm, B200_TMAP_DTYPE, 5, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swizzle_for_row_bytes(kc * 2), CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  B200_CHECK_ARG(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(w-stride-2 plane %dx%dx%dx%dx%d) failed: %d", N, D, H, W, C, (int)r);
  return 0;
}

// plain (tap-loop) kernel launch over the tile domain (D,H,W) = the OUTPUT lattice the CTAs enumerate
static int conv_igemm_plain_launch(const void* x, const void* wf, int n_w, const float* biascls, int n_b, const void* residual, int act,
                                   float slope, int N, int D, int H, int W, int Cin, int Cout, void* y, int pmode, const void* aux,
                                   float* partials, const PlainGeom& g, cudaStream_t s) {
  ConvParams p;
  memset(&p, 0, sizeof(p));
  p.N = N; p.D = D; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout;
  p.ntaps = g.ntaps;
  memcpy(p.toff, g.toff, sizeof(p.toff));
  p.in_mul = g.in_mul;
  p.out_mul = g.out_mul;
  for (int i = 0; i < 3; ++i) p.out_off[i] = g.out_off[i];
  p.OD = g.out_mul * D; p.OH = g.out_mul * H; p.OW = g.out_mul * W;
  p.w_rows = g.w_rows;
  p.w_row0 = g.w_row0;
  p.cls_mode = g.cls_mode;
  p.nacc = g.nacc;
  memcpy(p.acc_off, g.acc_off, sizeof(p.acc_off));
  memcpy(p.acc_ntaps, g.acc_ntaps, sizeof(p.acc_ntaps));
  B200_CHECK_ARG(g.nacc >= 1 && g.nacc <= 8 && (g.acc_ntaps[0] || g.ntaps % g.nacc == 0) && (g.nacc == 1 || pmode == 0),
                 "conv3_igemm: bad accumulator split");
  choose_box(D, H, W, &p.BD, &p.BH, &p.BW);
  p.tilesD = (D + p.BD - 1) / p.BD;
  p.tilesH = (H + p.BH - 1) / p.BH;
  p.tilesW = (W + p.BW - 1) / p.BW;
  p.n_w = n_w;
  p.n_b = biascls ? n_b : 0;
  p.NT = pick_nt(Cout);
  while (g.nacc * p.NT > 512 && p.NT % 32 == 0) p.NT /= 2;  // all accumulators of a CTA live in its 512 TMEM columns
  {
    // small volumes (16^3 and below: 64 tiles for 148 SMs): split the output channels over more CTAs while that at least doubles the
    // number of busy SMs; N = 128 -> 64 costs 48 instead of 32 cycles per 64 columns, twice the SMs more than pays for it
    static const bool split = [] { const char* e = getenv("B200UNET_IGEMM_NSPLIT"); return !(e && e[0] == '0'); }();
    const long long tiles = (long long)N * p.tilesD * p.tilesH * p.tilesW;
    const int sms = sm_count();
    while (split && tiles * (Cout / p.NT) * 2 <= sms && p.NT >= 128 && p.NT % 32 == 0) p.NT /= 2;
  }
  B200_CHECK_ARG(g.nacc * p.NT <= 512 && Cout % p.NT == 0, "conv3_igemm: %d accumulators of %d columns do not fit TMEM", g.nacc, p.NT);
  p.KC = (Cin % 64 == 0) ? 64 : (Cin % 32 == 0 ? 32 : 16);
  p.kchunks = Cin / p.KC;
  p.a_bytes = (128 * p.KC * 2 + 1023) & ~1023;
  p.b_bytes = (p.NT * p.KC * 2 + 1023) & ~1023;
  const int stage_bytes = p.a_bytes + p.b_bytes;
  const int numK = g.ntaps * p.kchunks;
  // one tile per CTA: MMA, epilogue and pipeline fill of a CTA are serial, so latency is hidden by CO-RESIDENT CTAs -- size the
  // ring to ~96 KB (two CTAs per SM); 72 / 48 KB measured no different (env B200UNET_IGEMM_SMEM_KB overrides, for experiments)
  static const int smem_kb = [] { const char* e = getenv("B200UNET_IGEMM_SMEM_KB"); return e ? atoi(e) : 96; }();
  int stages = (smem_kb * 1024) / stage_bytes;
  {
    // fewer CTAs than SMs: nothing to co-reside with, the K loop is a chain of TMA round trips -> spend the whole SM on pipeline depth
    const long long ctas = (long long)N * p.tilesD * p.tilesH * p.tilesW * (Cout / p.NT);
    if (ctas <= sm_count() && !getenv("B200UNET_IGEMM_SMEM_KB")) stages = (196 * 1024) / stage_bytes;
  }
  if (stages > numK) stages = numK;  // short K loops: keep the CTA small so several fit on an SM
  if (stages < 3 && numK >= 3) stages = 3;
  if (stages < 1) stages = 1;
  if (stages > CONV_MAX_STAGES) stages = CONV_MAX_STAGES;
  p.stages = stages;
  int cols = 32;
  while (cols < g.nacc * p.NT) cols <<= 1;
  p.tmem_cols = cols;
  p.act = act;
  p.slope = slope;
  p.pmode = pmode;
  p.biascls = biascls;
  p.residual = (const bf16*)residual;
  p.aux = (const bf16*)aux;
  p.y = (bf16*)y;
  p.partials = partials;

  CUtensorMap tmA, tmB;
  int rc = g.in_mul == 2 ? make_act_tmap_stride2(&tmA, x, N, 2 * D, 2 * H, 2 * W, Cin, p.KC, p.BD, p.BH, p.BW)
                         : make_act_tmap(&tmA, x, N, D, H, W, Cin, p.KC, p.BD, p.BH, p.BW);
  if (rc) return rc;
  rc = make_w_tmap(&tmB, wf, g.w_rows * n_w, Cout, Cin, p.KC, p.NT, 1);
  if (rc) return rc;

  size_t smem = (size_t)stages * stage_bytes + 1024;
  size_t scratch = (size_t)4 * p.NT * 2 * sizeof(float);
  if (smem < scratch + 1024) smem = scratch + 1024;
  cudaError_t e = cudaFuncSetAttribute(conv3_igemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  B200_CHECK_ARG(e == cudaSuccess, "conv3_igemm: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
  dim3 grid((unsigned)(N * p.tilesD * p.tilesH * p.tilesW), (unsigned)(Cout / p.NT));
  conv3_igemm_kernel<<<grid, CONV_THREADS, smem, s>>>(tmA, tmB, p);
  B200_CHECK_LAUNCH("conv3_igemm");
  return 0;
}
}  // namespace b200

namespace b200 {
void deconv_phase_table(signed char* k3 /*[27*3]: kernel index per axis*/, signed char* off /*[27*3]: low-res input offset*/,
                        signed char* ntaps /*[8]*/) {
  int r = 0;
  for (int phase = 0; phase < 8; ++phase) {
    const int pp[3] = {(phase >> 2) & 1, (phase >> 1) & 1, phase & 1};
    const int cnt[3] = {pp[0] ? 1 : 2, pp[1] ? 1 : 2, pp[2] ? 1 : 2};
    int n = 0;
    for (int a = 0; a < cnt[0]; ++a)
      for (int b = 0; b < cnt[1]; ++b)
        for (int c = 0; c < cnt[2]; ++c, ++r, ++n) {
          const int sel[3] = {a, b, c};
          for (int ax = 0; ax < 3; ++ax) {
            // p = 1: (k 1, offset 0);  p = 0: selection 0 -> (k 2, offset -1), selection 1 -> (k 0, offset 0)
            k3[3 * r + ax] = (signed char)(pp[ax] ? 1 : (sel[ax] == 0 ? 2 : 0));
            off[3 * r + ax] = (signed char)(pp[ax] ? 0 : (sel[ax] == 0 ? -1 : 0));
          }
        }
    ntaps[phase] = (signed char)n;
  }
}
}  // namespace b200

using namespace b200;

extern "C" {

int b200_conv3_igemm_supported(int N, int D, int H, int W, int Cin, int Cout) {
  return conv_igemm_supported(N, D, H, W, Cin, Cout) ? 1 : 0;
}

int b200_conv3_igemm_partials_count(int N, int D, int H, int W, int Cin, int Cout) {
  (void)N;
  (void)Cin;
  (void)Cout;
  ConvParams hp;
  if (conv_zs_plan(N, D, H, W, Cin, Cout, &hp)) return hp.ctas_per_sample;  // one partial row per persistent CTA
  if (conv_halo_plan(N, D, H, W, Cin, Cout, &hp)) return hp.tilesD * hp.tilesH * hp.tilesW;
  int bd, bh, bw;
  if (choose_box(D, H, W, &bd, &bh, &bw)) return 0;
  return ((D + bd - 1) / bd) * ((H + bh - 1) / bh) * ((W + bw - 1) / bw);
}

int b200_conv3_igemm_fwd(const void* x, const void* wf, int n_w, const float* biascls, int n_b, const void* residual, int act,
                         float slope, int N, int D, int H, int W, int Cin, int Cout, void* y, int pmode, const void* aux,
                         float* partials, b200_stream_t s) {
  B200_CHECK_ARG(conv_igemm_supported(N, D, H, W, Cin, Cout), "conv3_igemm: unsupported shape N=%d D=%d H=%d W=%d Cin=%d Cout=%d", N, D,
                 H, W, Cin, Cout);
  const int cls_mode = (pmode >> 8) & 1;  // B200_PMODE_PHASE_BIAS
  pmode &= 0xff;
  ConvParams p;
  memset(&p, 0, sizeof(p));
  B200_CHECK_ARG(pmode == 0 || partials, "conv3_igemm: pmode=%d needs a partials buffer", pmode);
  B200_CHECK_ARG(pmode != 2 || aux, "conv3_igemm: pmode=2 needs aux");
  const bool zs = conv_zs_plan(N, D, H, W, Cin, Cout, &p);
  B200_CHECK_ARG(!(zs && pmode == 2), "conv3_igemm: pmode=2 is not provided by the z-stacked kernel (B200UNET_ZS=0 selects the halo kernel)");
  if (zs || conv_halo_plan(N, D, H, W, Cin, Cout, &p)) {
    // small-channel / large-volume layers: halo tiles in shared memory feed all taps, weights stay resident
    p.n_w = n_w;
    p.n_b = biascls ? n_b : 0;
    p.act = act;
    p.slope = slope;
    p.pmode = pmode;
    p.cls_mode = cls_mode;
    p.biascls = biascls;
    p.residual = (const bf16*)residual;
    p.aux = (const bf16*)aux;
    p.y = (bf16*)y;
    p.partials = partials;
    return zs ? conv_zs_launch(x, wf, p, (cudaStream_t)s) : conv_halo_launch(x, wf, p, (cudaStream_t)s);
  }
  PlainGeom g;
  g.cls_mode = cls_mode;
  return conv_igemm_plain_launch(x, wf, n_w, biascls, n_b, residual, act, slope, N, D, H, W, Cin, Cout, y, pmode, aux, partials, g,
                                 (cudaStream_t)s);
}

// ---- conv3x3x3 over a nearest-2x-upsampled tensor WITHOUT materialising it (decoder concat path, buildingblocks.py:493 + :575).
// Output parity phase p (per axis) of conv3(up(b)) is a 2x2x2 convolution of the low-res b with phase-specific summed weights:
//   p = 0: low-res offsets {-1, 0} carry taps {-1}, {0,+1};   p = 1: offsets {0, +1} carry taps {-1,0}, {+1}
// (8/27 of the MACs).  One launch per phase writes R[2u+p] (bf16, no bias/activation); the caller adds R as the `residual`
// of the encoder-channel convolution.  Zero padding of the upsampled volume == TMA zero fill of the low-res volume.
int b200_conv3_up_supported(int N, int d, int h, int w, int C1, int Cout) {
  (void)N;
  int bd, bh, bw;
  if (C1 % 16 != 0 || Cout % 16 != 0 || d < 1 || h < 1 || w < 1) return 0;
  return choose_box(d, h, w, &bd, &bh, &bw) ? 0 : 1;
}
int b200_conv3_up_phase_fwd(const void* b, const void* wp, int n_w, int N, int d, int h, int w, int C1, int Cout, void* R, b200_stream_t s) {
  B200_CHECK_ARG(b200_conv3_up_supported(N, d, h, w, C1, Cout), "conv3_up_phase_fwd: unsupported N=%d %dx%dx%d C1=%d Cout=%d", N, d, h, w,
                 C1, Cout);
  {  // large low-res planes, resident weights: the z-stacked kernel (upzs_sm100.cu, N = 4*C_out per instruction)
    const int rc = conv3_upzs_run(b, wp, n_w, N, d, h, w, C1, Cout, R, (cudaStream_t)s);
    if (rc >= 0) return rc;
  }
  // one launch: every CTA owns a tile of 128 low-res voxels and all 8 phases (8 accumulators; the epilogue of phase k overlaps
  // the MMAs of phase k+1)
  PlainGeom g;
  g.ntaps = 64;
  g.nacc = 8;
  for (int phase = 0; phase < 8; ++phase) {
    const int pp[3] = {(phase >> 2) & 1, (phase >> 1) & 1, phase & 1};
    for (int j = 0; j < 8; ++j) {
      const int jj[3] = {(j >> 2) & 1, (j >> 1) & 1, j & 1};
      for (int a = 0; a < 3; ++a)
        g.toff[3 * (phase * 8 + j) + a] = (signed char)(pp[a] == 0 ? (jj[a] == 0 ? -1 : 0) : (jj[a] == 0 ? 0 : 1));
    }
    for (int a = 0; a < 3; ++a) g.acc_off[3 * phase + a] = (signed char)pp[a];
  }
  g.out_mul = 2;
  g.w_rows = 64;
  g.cls_mode = 2;
  return conv_igemm_plain_launch(b, wp, n_w, nullptr, 0, nullptr, B200_ACT_NONE, 0.f, N, d, h, w, C1, Cout, R, 0, nullptr, nullptr, g,
                                 (cudaStream_t)s);
}
// transpose of the above: d b[u] = sum over the 4x4x4 offsets e in {-1..2}^3 of Wd[e] dz[2u+e]  (stride-2 reads of dz through an
// element-stride-2 tensor map).  wd: bf16 [64][C1][Cout].
// z-stacked version (updzs_sm100.cu): needs a scratch of 4 partial gradients [4][N][d][h][w][C1] (16-bit)
int b200_conv3_up_dgrad_zs_supported(int N, int d, int h, int w, int Cout, int C1) {
  return (b200_device_is_sm100() && conv3_updzs_supported(N, d, h, w, Cout, C1)) ? 1 : 0;
}
int b200_conv3_up_dgrad_zs(const void* dz, const void* wd, int N, int d, int h, int w, int Cout, int C1, void* parts, void* dxb, b200_stream_t s) {
  const int rc = conv3_updzs_run(dz, wd, N, d, h, w, Cout, C1, parts, dxb, (cudaStream_t)s);
  B200_CHECK_ARG(rc >= 0, "conv3_up_dgrad_zs: shape not taken N=%d %dx%dx%d Cout=%d C1=%d", N, d, h, w, Cout, C1);
  return rc;
}
int b200_conv3_up_dgrad(const void* dz, const void* wd, int N, int d, int h, int w, int Cout, int C1, void* dxb, b200_stream_t s) {
  B200_CHECK_ARG(b200_conv3_up_supported(N, d, h, w, C1, Cout), "conv3_up_dgrad: unsupported N=%d %dx%dx%d C1=%d Cout=%d", N, d, h, w, C1,
                 Cout);
  PlainGeom g;
  g.ntaps = 64;
  for (int e = 0; e < 64; ++e) {
    g.toff[3 * e] = (signed char)((e >> 4) - 1);
    g.toff[3 * e + 1] = (signed char)(((e >> 2) & 3) - 1);
    g.toff[3 * e + 2] = (signed char)((e & 3) - 1);
  }
  g.in_mul = 2;
  g.w_rows = 64;
  g.cls_mode = 2;
  return conv_igemm_plain_launch(dz, wd, 1, nullptr, 0, nullptr, B200_ACT_NONE, 0.f, N, d, h, w, Cout, C1, dxb, 0, nullptr, nullptr, g,
                                 (cudaStream_t)s);
}

// ---- ConvTranspose3d(k3, s2, p1) by output parity phases (TransposeConvUpsampling, buildingblocks.py:617-664), on the LOW-RES lattice:
//   T[o] = sum_{i, k : 2i - 1 + k = o} Wt[k] x[i]  on the (2d-1)^3 grid;  written here as  P[j] = T[j - 1]  on a (2d)^3 grid (P[0] along
//   an axis is a don't-care: the nearest resize to the encoder size reads T[max(j-1, 0)] = P[max(j, 1)], see b200_shift_add_fwd).
//   Per axis, j = 2u + p:  p = 1 (T index 2u, even):   one tap  Wt[1] x[u]
//                          p = 0 (T index 2u-1, odd):  two taps Wt[2] x[u-1] + Wt[0] x[u]
// => 27 (phase, tap) products in total (1,2,2,4,2,4,4,8 per phase) instead of the 27 taps PER OUTPUT VOXEL of a convolution over the
// zero-inserted input: 8x fewer MACs.  One launch, 8 accumulators per CTA (one per phase, unequal tap counts).
// wq: bf16 [27][Cout][Cin] in the (phase, tap) order of deconv_phase_table(); b200_deconv_phase_weights writes it.
int b200_deconv_phase_supported(int N, int d, int h, int w, int Cin, int Cout) {
  (void)N;
  int bd, bh, bw;
  if (Cin % 16 != 0 || Cout % 16 != 0 || d < 1 || h < 1 || w < 1) return 0;
  return choose_box(d, h, w, &bd, &bh, &bw) ? 0 : 1;
}
int b200_deconv_phase_fwd(const void* x, const void* wq, int N, int d, int h, int w, int Cin, int Cout, void* P, b200_stream_t s) {
  B200_CHECK_ARG(b200_deconv_phase_supported(N, d, h, w, Cin, Cout), "deconv_phase_fwd: unsupported N=%d %dx%dx%d Cin=%d Cout=%d", N, d, h, w,
                 Cin, Cout);
  PlainGeom g;
  g.ntaps = 27;
  g.nacc = 8;
  signed char k3[27 * 3];
  deconv_phase_table(k3, g.toff, g.acc_ntaps);
  for (int phase = 0; phase < 8; ++phase) {
    g.acc_off[3 * phase] = (signed char)((phase >> 2) & 1);
    g.acc_off[3 * phase + 1] = (signed char)((phase >> 1) & 1);
    g.acc_off[3 * phase + 2] = (signed char)(phase & 1);
  }
  g.out_mul = 2;
  g.w_rows = 27;
  g.cls_mode = 2;
  return conv_igemm_plain_launch(x, wq, 1, nullptr, 0, nullptr, B200_ACT_NONE, 0.f, N, d, h, w, Cin, Cout, P, 0, nullptr, nullptr, g,
                                 (cudaStream_t)s);
}
// adjoint: dx[u] = sum_{e in {0,1,2}^3} Wt[e]^T gp[2u + e]  -- a 3x3x3 STRIDE-2 convolution of the (folded) output gradient, read through an
// element-stride-2 tensor map; indices past the (2d)^3 grid are zero-filled.  wd: bf16 [27][Cin][Cout] (b200_deconv_phase_weights).
int b200_deconv_phase_dgrad(const void* gp, const void* wd, int N, int d, int h, int w, int Cout, int Cin, void* dx, b200_stream_t s) {
  B200_CHECK_ARG(b200_deconv_phase_supported(N, d, h, w, Cin, Cout), "deconv_phase_dgrad: unsupported N=%d %dx%dx%d Cin=%d Cout=%d", N, d, h, w,
                 Cin, Cout);
  PlainGeom g;
  g.ntaps = 27;
  for (int e = 0; e < 27; ++e) {
    g.toff[3 * e] = (signed char)(e / 9);
    g.toff[3 * e + 1] = (signed char)((e / 3) % 3);
    g.toff[3 * e + 2] = (signed char)(e % 3);
  }
  g.in_mul = 2;
  g.w_rows = 27;
  g.cls_mode = 2;
  return conv_igemm_plain_launch(gp, wd, 1, nullptr, 0, nullptr, B200_ACT_NONE, 0.f, N, d, h, w, Cout, Cin, dx, 0, nullptr, nullptr, g,
                                 (cudaStream_t)s);
}

// ---- 1x1x1 convolution (ResNetBlock.conv1, buildingblocks.py:203) on the same kernel: the volume is a flat list of voxels
int b200_pointwise_tc_supported(int N, long long vox, int Cin, int Cout) {
  (void)N;
  return (Cin % 16 == 0 && Cout % 16 == 0 && vox >= 1 && vox < (1ll << 31)) ? 1 : 0;
}
int b200_pointwise_tc_partials_count(int N, long long vox) {
  (void)N;
  return (int)((vox + 127) / 128);
}
int b200_pointwise_tc_fwd(const void* x, const void* wq, const float* bias, int N, long long vox, int Cin, int Cout, void* y,
                          float* partials, b200_stream_t s) {
  B200_CHECK_ARG(b200_pointwise_tc_supported(N, vox, Cin, Cout), "pointwise_tc: unsupported N=%d vox=%lld Cin=%d Cout=%d", N, vox, Cin,
                 Cout);
  PlainGeom g;
  g.ntaps = 1;
  g.toff[0] = g.toff[1] = g.toff[2] = 0;
  g.w_rows = 1;
  g.cls_mode = 2;
  return conv_igemm_plain_launch(x, wq, 1, bias, bias ? 1 : 0, nullptr, B200_ACT_NONE, 0.f, N, 1, 1, (int)vox, Cin, Cout, y, partials ? 1 : 0,
                                 nullptr, partials, g, (cudaStream_t)s);
}

}  // extern "C"
