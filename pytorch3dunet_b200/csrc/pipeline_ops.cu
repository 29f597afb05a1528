// Kernels either side of the model call, all HBM-bound streams:
//   * patch gather: a (C, pz, py, px) fp32 patch cut out of a DEVICE-RESIDENT (C,Z,Y,X) fp32 volume, with the reflect padding the
//     reference applies to the whole volume before slicing (datasets/utils.py:518-546 mirror_pad = np.pad(mode='reflect'),
//     hdf5.py:16-20 halo-extended indices) -- the padded volume is never materialised;
//   * patch scatter: halo crop (remove_padding, datasets/utils.py:549-565) + write-back into the device-resident (C_out,Z,Y,X)
//     output volume (StandardPredictor.__call__, predictor.py:148-193: `prediction_array[index] = pred`, later patches overwrite
//     earlier ones).  "Last writer wins" is evaluated analytically: patch (iz,iy,ix) of the z-outer / x-inner patch grid writes a
//     voxel only if it is the LAST patch covering it, i.e. owner_z[z] == iz && owner_y[y] == iy && owner_x[x] == ix, where
//     owner_a[c] = the highest patch index along axis a whose [start, stop) contains c (lexicographic max of a product set is the
//     tuple of per-axis maxima).  Every output voxel is then written exactly once, by one patch, on whichever GPU ran it: patch
//     sharding across GPUs needs no ordering between ranks and the shards' volumes are disjoint (merge = sum);
//   * fused Adam over a flat fp32 parameter buffer (torch.optim.Adam as created by create_optimizer, utils.py:246-316: L2 weight
//     decay added to the gradient, bias-corrected moments), optional gradient pre-scale (1/world after a sum-allreduce).
#include "common.cuh"

namespace b200 {

__device__ __forceinline__ int reflect_index(int i, int n) {
  // np.pad(mode='reflect'): ... 2 1 | 0 1 2 ... n-1 | n-2 n-3 ...   (valid for pad < n)
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return i;
}

// grid (ceil(px/128)?, py, C*pz) is wasteful for small rows; use a flat grid-stride loop with x fastest (coalesced rows).
__global__ void patch_gather_kernel(const float* __restrict__ vol, int C, int Z, int Y, int X, int z0, int y0, int x0, int pz, int py, int px,
                                    float* __restrict__ out) {
  const size_t total = (size_t)C * pz * py * px;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % px);
    size_t r = i / px;
    const int y = (int)(r % py);
    r /= py;
    const int z = (int)(r % pz);
    const int c = (int)(r / pz);
    const int sz = reflect_index(z0 + z, Z), sy = reflect_index(y0 + y, Y), sx = reflect_index(x0 + x, X);
    out[i] = vol[(((size_t)c * Z + sz) * Y + sy) * X + sx];
  }
}

// pred: [C][pz][py][px] (halo still attached); the patch's unpadded index is [z0, z0+pz-2hz) x ...; out: [C][Z][Y][X]
__global__ void patch_scatter_kernel(const float* __restrict__ pred, int C, int pz, int py, int px, int hz, int hy, int hx,
                                     float* __restrict__ out, int Z, int Y, int X, int z0, int y0, int x0, int iz, int iy, int ix,
                                     const int* __restrict__ owner_z, const int* __restrict__ owner_y, const int* __restrict__ owner_x) {
  const int cz = pz - 2 * hz, cy = py - 2 * hy, cx = px - 2 * hx;
  const size_t total = (size_t)C * cz * cy * cx;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % cx);
    size_t r = i / cx;
    const int y = (int)(r % cy);
    r /= cy;
    const int z = (int)(r % cz);
    const int c = (int)(r / cz);
    const int gz = z0 + z, gy = y0 + y, gx = x0 + x;
    if (owner_z[gz] != iz || owner_y[gy] != iy || owner_x[gx] != ix) continue;  // a later patch owns this voxel
    out[(((size_t)c * Z + gz) * Y + gy) * X + gx] = pred[(((size_t)c * pz + z + hz) * py + y + hy) * px + x + hx];
  }
}

// torch.optim.Adam (amsgrad=False, maximize=False), single tensor = the flat buffer.  bc1 = 1 - beta1^t, bc2 = 1 - beta2^t.
__global__ void adam_step_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, size_t n,
                                 float lr, float beta1, float beta2, float eps, float wd, float bc1, float bc2, float grad_scale) {
  const float step_size = lr / bc1;
  const float inv_sqrt_bc2 = rsqrtf(bc2);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float pi = p[i];
    const float gi = g[i] * grad_scale + wd * pi;
    const float mi = m[i] + (1.f - beta1) * (gi - m[i]);  // exp_avg.lerp_(grad, 1 - beta1)
    const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) * inv_sqrt_bc2 + eps;
    p[i] = pi - step_size * (mi / denom);
  }
}

}  // namespace b200

using namespace b200;
#define ST(s) ((cudaStream_t)(s))

static int stream_blocks(size_t total) {
  size_t b = (total + 255) / 256;
  const size_t cap = (size_t)sm_count() * 16;
  if (b > cap) b = cap;
  return b < 1 ? 1 : (int)b;
}

extern "C" {

int b200_patch_gather_f32(const float* vol, int C, int Z, int Y, int X, int z0, int y0, int x0, int pz, int py, int px, float* out,
                          b200_stream_t s) {
  B200_CHECK_ARG(C >= 1 && pz >= 1 && py >= 1 && px >= 1, "patch_gather: empty patch");
  // reflect padding needs pad < size on every axis (np.pad(mode='reflect') with a single reflection)
  B200_CHECK_ARG(-z0 < Z && -y0 < Y && -x0 < X && z0 + pz - Z < Z && y0 + py - Y < Y && x0 + px - X < X,
                 "patch_gather: patch (%d,%d,%d)+(%d,%d,%d) reaches more than one reflection outside the (%d,%d,%d) volume", z0, y0, x0, pz,
                 py, px, Z, Y, X);
  const size_t total = (size_t)C * pz * py * px;
  patch_gather_kernel<<<stream_blocks(total), 256, 0, ST(s)>>>(vol, C, Z, Y, X, z0, y0, x0, pz, py, px, out);
  B200_CHECK_LAUNCH("patch_gather");
  return 0;
}

int b200_patch_scatter_f32(const float* pred, int C, int pz, int py, int px, int hz, int hy, int hx, float* out, int Z, int Y, int X,
                           int z0, int y0, int x0, int iz, int iy, int ix, const int* owner_z, const int* owner_y, const int* owner_x,
                           b200_stream_t s) {
  B200_CHECK_ARG(pz > 2 * hz && py > 2 * hy && px > 2 * hx, "patch_scatter: halo (%d,%d,%d) swallows the patch (%d,%d,%d)", hz, hy, hx, pz,
                 py, px);
  B200_CHECK_ARG(z0 >= 0 && y0 >= 0 && x0 >= 0 && z0 + pz - 2 * hz <= Z && y0 + py - 2 * hy <= Y && x0 + px - 2 * hx <= X,
                 "patch_scatter: patch index outside the output volume");
  const size_t total = (size_t)C * (pz - 2 * hz) * (py - 2 * hy) * (px - 2 * hx);
  patch_scatter_kernel<<<stream_blocks(total), 256, 0, ST(s)>>>(pred, C, pz, py, px, hz, hy, hx, out, Z, Y, X, z0, y0, x0, iz, iy, ix,
                                                               owner_z, owner_y, owner_x);
  B200_CHECK_LAUNCH("patch_scatter");
  return 0;
}

int b200_adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2, float eps, float wd,
                   float bc1, float bc2, float grad_scale, b200_stream_t s) {
  B200_CHECK_ARG(n >= 0 && bc1 > 0.f && bc2 > 0.f, "adam_step: bad arguments");
  if (n == 0) return 0;
  adam_step_kernel<<<stream_blocks((size_t)n), 256, 0, ST(s)>>>(p, g, m, v, (size_t)n, lr, beta1, beta2, eps, wd, bc1, bc2, grad_scale);
  B200_CHECK_LAUNCH("adam_step");
  return 0;
}

}  // extern "C"
