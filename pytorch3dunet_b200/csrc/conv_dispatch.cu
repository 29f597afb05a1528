// Public conv entry points: choose between the tcgen05 implicit-GEMM kernels and the direct CUDA-core kernels.
// There is NO CPU path: both implementations are device kernels in this library.
#include "common.cuh"

extern "C" {

int b200_conv3_resolve_impl(int impl, int N, int D, int H, int W, int Cin, int Cout, int x_is_f32) {
  if (impl == B200_IMPL_DIRECT) return B200_IMPL_DIRECT;
  bool ok = !x_is_f32 && b200_device_is_sm100() && b200_conv3_igemm_supported(N, D, H, W, Cin, Cout);
  if (impl == B200_IMPL_TCGEN05) return ok ? B200_IMPL_TCGEN05 : -1;
  return ok ? B200_IMPL_TCGEN05 : B200_IMPL_DIRECT;
}

int b200_conv3_partials_count(int impl, int N, int D, int H, int W, int Cin, int Cout) {
  int r = b200_conv3_resolve_impl(impl, N, D, H, W, Cin, Cout, 0);
  if (r == B200_IMPL_TCGEN05) return b200_conv3_igemm_partials_count(N, D, H, W, Cin, Cout);
  return b200_conv3_direct_partials_count(N, D, H, W, Cout);
}

int b200_conv3_fwd(int impl, const void* x, int x_is_f32, const void* wf, int n_w, const float* biascls, int n_b, const void* residual,
                   int act, float slope, int N, int D, int H, int W, int Cin, int Cout, void* y, int pmode, const void* aux,
                   float* partials, b200_stream_t s) {
  int r = b200_conv3_resolve_impl(impl, N, D, H, W, Cin, Cout, x_is_f32);
  if (r < 0) {
    b200::set_error("conv3_fwd: tcgen05 implementation requested but unsupported for N=%d D=%d H=%d W=%d Cin=%d Cout=%d f32=%d", N, D, H,
                    W, Cin, Cout, x_is_f32);
    return 1;
  }
  if (r == B200_IMPL_TCGEN05)
    return b200_conv3_igemm_fwd(x, wf, n_w, biascls, n_b, residual, act, slope, N, D, H, W, Cin, Cout, y, pmode, aux, partials, s);
  B200_CHECK_ARG((pmode & 0x100) == 0, "conv3_fwd: phase-aware bias classes need the tcgen05 implementation");
  return b200_conv3_direct_fwd(x, x_is_f32, wf, n_w, biascls, n_b, residual, act, slope, N, D, H, W, Cin, Cout, y, pmode, aux, partials,
                               s);
}

int b200_conv3_wgrad_resolve_impl(int impl, int N, int D, int H, int W, int Cin, int Cout, int x_is_f32) {
  if (impl == B200_IMPL_DIRECT) return B200_IMPL_DIRECT;
  bool ok = !x_is_f32 && b200_device_is_sm100() && b200_conv3_wgrad_igemm_supported(N, D, H, W, Cin, Cout);
  if (impl == B200_IMPL_TCGEN05) return ok ? B200_IMPL_TCGEN05 : -1;
  return ok ? B200_IMPL_TCGEN05 : B200_IMPL_DIRECT;
}

int b200_conv3_wgrad_splits(int impl, int N, int D, int H, int W, int Cin, int Cout, int x_is_f32) {
  int r = b200_conv3_wgrad_resolve_impl(impl, N, D, H, W, Cin, Cout, x_is_f32);
  if (r == B200_IMPL_TCGEN05) return b200_conv3_wgrad_igemm_splits(N, D, H, W, Cin, Cout);
  return b200_conv3_direct_wgrad_splits(N, D, H, W, Cin, Cout, x_is_f32);
}

int b200_conv3_wgrad(int impl, const void* x, int x_is_f32, const void* dz, int N, int D, int H, int W, int Cin, int Cout, float* G,
                     b200_stream_t s) {
  int r = b200_conv3_wgrad_resolve_impl(impl, N, D, H, W, Cin, Cout, x_is_f32);
  if (r < 0) {
    b200::set_error("conv3_wgrad: tcgen05 implementation requested but unsupported for N=%d D=%d H=%d W=%d Cin=%d Cout=%d", N, D, H, W,
                    Cin, Cout);
    return 1;
  }
  if (r == B200_IMPL_TCGEN05) return b200_conv3_wgrad_igemm(x, dz, N, D, H, W, Cin, Cout, G, s);
  return b200_conv3_direct_wgrad(x, x_is_f32, dz, N, D, H, W, Cin, Cout, G, s);
}

}  // extern "C"
