// Shared pieces of the tcgen05 convolution kernels: parameter block, the epilogue slab, host helpers.
#pragma once
#include <string.h>

#include "common.cuh"
#include "sm100_ptx.cuh"

namespace b200 {

#ifdef B200_DEBUG
#define DBG_FLAG(p, bit) ((p).dbg_flags & (bit))
#else
#define DBG_FLAG(p, bit) 0
#endif

constexpr int CONV_THREADS = 192;
constexpr int CONV_MAX_STAGES = 12;
constexpr int HALO_MAX_STAGES = 6;

struct ConvParams {
  int N, D, H, W, Cin, Cout;
  int BD, BH, BW;
  int tilesD, tilesH, tilesW;
  int n_w, n_b;
  // ---- plain igemm kernel only: generic tap table.  Tap i reads the input at  in_mul * (tile origin) + toff[i]  (per axis, TMA
  // coordinates of the input tensor map; in_mul = 2 with an element-stride-2 map subsamples the input) with weight row
  // wsample * w_rows + w_row0 + i.  Outputs go to voxel  out_mul * x + out_off  of a volume (OD, OH, OW).
  int ntaps;             // 27 (3x3x3), 1 (1x1x1), 8 (one phase of conv3 o nearest-upsample), 64 (its transpose, 4x4x4 stride 2)
  signed char toff[64 * 3];
  int in_mul, out_mul, out_off[3], OD, OH, OW;
  int w_rows, w_row0;
  signed char acc_ntaps[8];  // taps of accumulator a (0 = the even split ntaps / nacc): the transposed conv's phases have 1..8 taps
  int nacc;                 // accumulators per CTA (1; 8 = the eight parity phases): taps [a*ntaps/nacc, (a+1)*ntaps/nacc) feed
  signed char acc_off[8 * 3];  // accumulator a, whose outputs go to out_mul * x + acc_off[a]  (out_off when nacc == 1)
  int cls_mode;  // bias row: 0 = border class of the voxel, 1 = phase-aware border class (interior split by parity), 2 = row 0
  int NT;       // output channels per CTA
  int KC;       // channels per k-block (16/32/64)
  int kchunks;  // Cin / KC
  int stages;
  int a_bytes, b_bytes;  // per-stage tile sizes (1024-aligned)
  int tmem_cols;
  int act;
  float slope;
  int pmode;
  const float* biascls;
  const bf16* residual;
  const bf16* aux;
  bf16* y;
  float* partials;
  // halo kernel only
  int KCb;             // channels per resident-weight block (swizzle granule of B)
  int a_stages;        // halo stages of KC channels each
  int b_total_bytes;   // resident weights [Cin/KCb][27][NT][KCb]
  int ctas_per_sample;
  int tmem_bufs;       // accumulator buffers in TMEM (2 or 4)
  long long* dbg;      // optional per-CTA wait-cycle counters (b200_set_debug_buffer), NULL in production
  int dbg_flags;       // experiments only (env B200UNET_DBG_FLAGS): 1 = skip the global stores, 2 = skip the bias add
};

// row of the [64][Cout] bias table for output voxel (xd,xh,xw) of a (D,H,W) volume
__device__ __forceinline__ int conv_bias_cls(int cls_mode, int xd, int xh, int xw, int D, int H, int W) {
  if (cls_mode == 2) return 0;
  int cd = axis_cls(xd, D), ch = axis_cls(xh, H), cw = axis_cls(xw, W);
  if (cls_mode == 1) {  // even dims >= 2: class 3 ("both") cannot occur and is reused for "interior, odd coordinate"
    if (cd == 1 && (xd & 1)) cd = 3;
    if (ch == 1 && (xh & 1)) ch = 3;
    if (cw == 1 && (xw & 1)) cw = 3;
  }
  return (cd << 4) | (ch << 2) | cw;
}

// one 32-/16-column slab of the accumulator tile for one thread (= one output voxel row)
template <int CW>
__device__ __forceinline__ void conv_epilogue_slab(const ConvParams& p, uint32_t taddr, int c0 /*col in tile*/, int n0, bool valid,
                                                   size_t vox_off /* (n*vox+v) */, const float* bias_row /* or null */, int lane,
                                                   float* scratch /* [NT][2] for this warp */, long long* tim = nullptr) {
  uint32_t raw[CW];
  long long t0 = tim ? dbg_clock() : 0;
  if constexpr (CW == 32) tmem_ld_32x32b_x32(taddr + c0, raw);
  else tmem_ld_32x32b_x16(taddr + c0, raw);
  tmem_ld_wait();
  long long t1 = tim ? dbg_clock() : 0;
  float v[CW];
#pragma unroll
  for (int i = 0; i < CW; ++i) v[i] = __uint_as_float(raw[i]);
  const size_t goff = vox_off * p.Cout + n0 + c0;
  if (valid) {
    if (bias_row && !DBG_FLAG(p, 2)) {
      const float4* bp = reinterpret_cast<const float4*>(bias_row + n0 + c0);
#pragma unroll
      for (int i = 0; i < CW / 4; ++i) {
        float4 b = bp[i];  // generic load: bias_row may live in shared memory (halo kernel) or global
        v[4 * i] += b.x;
        v[4 * i + 1] += b.y;
        v[4 * i + 2] += b.z;
        v[4 * i + 3] += b.w;
      }
    }
    if (p.residual) {
      const bf16x8* rp = reinterpret_cast<const bf16x8*>(p.residual + goff);
#pragma unroll
      for (int i = 0; i < CW / 8; ++i) {
        float f[8];
        unpack8(rp[i], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[8 * i + j] += f[j];
      }
    }
    // activation: the switch is hoisted out of the unrolled loop (inside it, every element drags the inlined ELU expm1f
    // and a branch along: ~1600 instructions per slab instead of ~300)
    if (p.act == B200_ACT_RELU) {
#pragma unroll
      for (int i = 0; i < CW; ++i) v[i] = fmaxf(v[i], 0.f);
    } else if (p.act == B200_ACT_LEAKY) {
#pragma unroll
      for (int i = 0; i < CW; ++i) v[i] = v[i] > 0.f ? v[i] : v[i] * p.slope;
    } else if (p.act == B200_ACT_ELU) {
#pragma unroll
      for (int i = 0; i < CW; ++i) v[i] = v[i] > 0.f ? v[i] : expm1f(v[i]);
    }
#pragma unroll
    for (int i = 0; i < CW; ++i) v[i] = bf16_round(v[i]);
    bf16x8* op = reinterpret_cast<bf16x8*>(p.y + goff);
    if (!DBG_FLAG(p, 1)) {
#pragma unroll
      for (int i = 0; i < CW / 8; ++i) op[i] = pack8(&v[8 * i]);
    }
  } else {
#pragma unroll
    for (int i = 0; i < CW; ++i) v[i] = 0.f;
  }
  long long t2 = tim ? dbg_clock() : 0;
  if (p.pmode) {
    float w[CW];
    if (p.pmode == 1) {
#pragma unroll
      for (int i = 0; i < CW; ++i) w[i] = v[i] * v[i];
    } else {
      if (valid) {
        const bf16x8* ap = reinterpret_cast<const bf16x8*>(p.aux + goff);
#pragma unroll
        for (int i = 0; i < CW / 8; ++i) {
          float f[8];
          unpack8(ap[i], f);
#pragma unroll
          for (int j = 0; j < 8; ++j) w[8 * i + j] = v[8 * i + j] * f[j];
        }
      } else {
#pragma unroll
        for (int i = 0; i < CW; ++i) w[i] = 0.f;
      }
    }
    float s = warp_reduce_scatter<CW>(v, lane);
    float q = warp_reduce_scatter<CW>(w, lane);
    if (lane < CW) {
      scratch[(c0 + lane) * 2] = s;
      scratch[(c0 + lane) * 2 + 1] = q;
    }
  }
  if (tim) {
    long long t3 = dbg_clock();
    tim[0] += t1 - t0;
    tim[1] += t2 - t1;
    tim[2] += t3 - t2;
  }
}


int make_act_tmap(CUtensorMap* tm, const void* ptr, int N, int D, int H, int W, int C, int kc, int bd, int bh, int bw);
int make_w_tmap(CUtensorMap* tm, const void* ptr, int rows2, int rows1, int C, int kc, int nt, int ntaps_box = 1);
int choose_box(int D, int H, int W, int* bd, int* bh, int* bw);
bool conv_igemm_supported(int N, int D, int H, int W, int Cin, int Cout);
bool conv_halo_plan(int N, int D, int H, int W, int Cin, int Cout, ConvParams* p);
int conv_halo_launch(const void* x, const void* wf, ConvParams& p, cudaStream_t s);
// (phase, tap) enumeration of the transposed conv by output parity phases (conv_igemm_sm100.cu): k3[27*3] kernel index per axis,
// off[27*3] low-res input offset per axis, ntaps[8] taps per phase
void deconv_phase_table(signed char* k3, signed char* off, signed char* ntaps);
bool conv_zs_plan(int N, int D, int H, int W, int Cin, int Cout, ConvParams* p);   // conv_zs_sm100.cu: depth taps stacked along N
int conv_zs_launch(const void* x, const void* wf, ConvParams& p, cudaStream_t s);
// upzs_sm100.cu: z-stacked phase conv of the virtual concat; returns -1 when the shape is not taken
// updzs_sm100.cu: its transpose (gradient w.r.t. the low-res tensor)
bool conv3_updzs_supported(int N, int d, int h, int w, int Cout, int C1);
int conv3_updzs_run(const void* dz, const void* wd, int N, int d, int h, int w, int Cout, int C1, void* parts, void* dxb, cudaStream_t s);
int conv3_upzs_run(const void* low, const void* wp, int n_w, int N, int d, int h, int w, int C1, int Cout, void* R, cudaStream_t s);

// stem_mma.cu: first conv of the network (C_in == 1, fp32 x) on warp-level MMA
bool stem_mma_supported(int Cin, int Cout);
int stem_mma_fwd(const float* x, const bf16* wf, int n_w, const float* biascls, int n_b, int act, float slope, int N, int D, int H, int W,
                 int Cout, int P, bf16* y, int pmode, float* partials, cudaStream_t s);
void stem_mma_wgrad_grid(int N, int D, int H, int W, int* blocks, int* tiles_per_block);
int stem_mma_wgrad(const float* x, const bf16* dz, int N, int D, int H, int W, int Cout, float* G, cudaStream_t s);
}  // namespace b200
