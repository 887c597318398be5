// C ABI (include/plume_b200.h): argument checking, view construction, error reporting.
// Every entry point enqueues on the caller's stream and returns; errors never cross as exceptions.
#include "../../include/plume_b200.h"

#include "bandwidth.cuh"
#include "igemm.cuh"

#include <string>

namespace plume {
static thread_local std::string g_last_error;
void set_error(const std::string& msg) { g_last_error = msg; }
}  // namespace plume

using namespace plume;

namespace {
inline cudaStream_t S(plume_stream_t s) { return static_cast<cudaStream_t>(s); }

// Strided NHWC view of C channels with pixel stride ld (elements).
ActView view(const void* ptr, int ld, int N, int H, int W, int C) {
  ActView v;
  v.ptr = ptr;
  v.C = C; v.W = W; v.H = H; v.N = N;
  v.pix_stride = ld;
  v.row_stride = 1ll * W * ld;
  v.img_stride = 1ll * H * W * ld;
  return v;
}
// The (i,j) phase of a 2x-upsampled buffer [N][2H][2W][ld] seen as an N x H x W image.
ActView phase_view(const void* ptr, int ld, int N, int H, int W, int C, int i, int j) {
  ActView v;
  v.ptr = static_cast<const char*>(ptr) + (1ll * i * (2 * W) + j) * ld * 2;
  v.C = C; v.W = W; v.H = H; v.N = N;
  v.pix_stride = 2ll * ld;
  v.row_stride = 2ll * (2 * W) * ld;
  v.img_stride = 1ll * (2 * H) * (2 * W) * ld;
  return v;
}
bool bad_ld(int ld, int C) { return ld < C || (ld % 8) != 0; }
// bf16x3 views: the pixel stride covers the hi and the lo plane (ld / 2 elements each)
bool bad_ld(int ld, int C, int split) { return split ? (ld % 16 != 0 || ld / 2 < C) : bad_ld(ld, C); }
ActView with_plane(ActView v, int ld, int split) {
  v.plane = split ? ld / 2 : 0;
  return v;
}
#define PLUME_CHECK(cond, msg) \
  do {                         \
    if (!(cond)) {             \
      set_error(msg);          \
      return -1;               \
    }                          \
  } while (0)
}  // namespace

extern "C" {

const char* plume_version(void) { return "plume_b200 0.1 (sm_100a; tcgen05+TMA implicit GEMM)"; }
const char* plume_last_error(void) { return g_last_error.c_str(); }
int plume_debug_word(void) { return read_debug_word(); }
void plume_debug_set_prof(long long* buf) { set_prof_buffer(buf); }
void plume_set_sm_margin(int sms) { set_sm_margin(sms); }
int plume_get_sm_margin(void) { return get_sm_margin(); }
void plume_set_deterministic(int on) { set_deterministic(on); }
int plume_get_deterministic(void) { return deterministic() ? 1 : 0; }
int plume_num_sms(void) {
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  return n;
}

static int impl_conv3x3_fwd(const void* x, int ldx, const void* w, const float* scale, const float* shift,
                      int relu, void* y, int ldy, double* stat_sum, double* stat_sq, int N, int H, int W,
                      int Cin, int Cout, plume_stream_t stream, int split) {
  PLUME_CHECK(x && w && y, "conv3x3_fwd: null pointer");
  // ldx < Cin: x is dense with ldx channels per pixel; the weights' remaining input channels read as zero
  const int ldx1 = split ? ldx / 2 : ldx;   // channels per pixel of one plane
  const int Cx = ldx1 < Cin ? ldx1 : Cin;
  PLUME_CHECK(Cx > 0 && !bad_ld(ldx, Cx, split) && !bad_ld(ldy, Cout, split), "conv3x3_fwd: bad pixel stride");
  PLUME_CHECK((stat_sum == nullptr) == (stat_sq == nullptr), "conv3x3_fwd: stat_sum/stat_sq mismatch");
  FwdDesc d{};
  d.split = split;
  d.in[0] = with_plane(view(x, ldx, N, H, W, Cx), ldx, split);
  d.num_in_views = 1;
  d.num_taps = 9;
  d.Cin = Cin;
  d.wmat = w;
  d.out[0] = with_plane(view(y, ldy, N, H, W, Cout), ldy, split);
  d.num_out_views = 1;
  d.cout_per_view = Cout;
  d.scale = scale; d.shift = shift; d.relu = relu;
  d.stat_sum = stat_sum; d.stat_sq = stat_sq;
  d.stream = S(stream);
  return launch_igemm_fwd(d);
}
int plume_conv3x3_fwd(const void* x, int ldx, const void* w, const float* scale, const float* shift,
                      int relu, void* y, int ldy, double* stat_sum, double* stat_sq, int N, int H, int W,
                      int Cin, int Cout, plume_stream_t stream) {
  return impl_conv3x3_fwd(x, ldx, w, scale, shift, relu, y, ldy, stat_sum, stat_sq, N, H, W, Cin, Cout, stream, 0);
}
int plume_conv3x3_fwd_x3(const void* x, int ldx, const void* w, const float* scale, const float* shift,
                      int relu, void* y, int ldy, double* stat_sum, double* stat_sq, int N, int H, int W,
                      int Cin, int Cout, plume_stream_t stream) {
  return impl_conv3x3_fwd(x, ldx, w, scale, shift, relu, y, ldy, stat_sum, stat_sq, N, H, W, Cin, Cout, stream, 1);
}

static int impl_conv3x3_dgrad(const void* dy, int lddy, const void* w_dgrad, void* dx, int lddx, int N, int H,
                        int W, int Cin, int Cout, plume_stream_t stream, int split) {
  PLUME_CHECK(dy && w_dgrad && dx, "conv3x3_dgrad: null pointer");
  PLUME_CHECK(!bad_ld(lddy, Cout, split) && !bad_ld(lddx, Cin, split), "conv3x3_dgrad: bad pixel stride");
  FwdDesc d{};
  d.split = split;
  d.in[0] = with_plane(view(dy, lddy, N, H, W, Cout), lddy, split);
  d.num_in_views = 1;
  d.num_taps = 9;
  d.Cin = Cout;
  d.wmat = w_dgrad;
  d.out[0] = with_plane(view(dx, lddx, N, H, W, Cin), lddx, split);
  d.num_out_views = 1;
  d.cout_per_view = Cin;
  d.stream = S(stream);
  return launch_igemm_fwd(d);
}
int plume_conv3x3_dgrad(const void* dy, int lddy, const void* w_dgrad, void* dx, int lddx, int N, int H,
                        int W, int Cin, int Cout, plume_stream_t stream) {
  return impl_conv3x3_dgrad(dy, lddy, w_dgrad, dx, lddx, N, H, W, Cin, Cout, stream, 0);
}
int plume_conv3x3_dgrad_x3(const void* dy, int lddy, const void* w_dgrad, void* dx, int lddx, int N, int H,
                        int W, int Cin, int Cout, plume_stream_t stream) {
  return impl_conv3x3_dgrad(dy, lddy, w_dgrad, dx, lddx, N, H, W, Cin, Cout, stream, 1);
}

int plume_wgrad_splits(int N, int H, int W, int taps, int Cin, int Cout) {
  if (N <= 0 || H <= 0 || W <= 0 || Cin <= 0 || Cout <= 0 || (taps != 9 && taps != 4)) return 0;
  return wgrad_plan(N, H, W, taps, Cin, Cout);
}
size_t plume_wgrad_workspace_bytes(int N, int H, int W, int taps, int Cin, int Cout) {
  (void)N; (void)H; (void)W; (void)taps; (void)Cin; (void)Cout;
  return 0;  // split-K partial sums are reduced with fp32 atomics; no workspace is needed any more
}

static int impl_conv3x3_wgrad(const void* x, int ldx, const void* dy, int lddy, float* dw, int accumulate,
                        void* workspace, size_t workspace_bytes, int N, int H, int W, int Cin, int Cout,
                        plume_stream_t stream, int split) {
  (void)workspace; (void)workspace_bytes;
  PLUME_CHECK(x && dy && dw, "conv3x3_wgrad: null pointer");
  const int ldx1 = split ? ldx / 2 : ldx;
  const int Cx = ldx1 < Cin ? ldx1 : Cin;  // see plume_conv3x3_fwd
  PLUME_CHECK(Cx > 0 && !bad_ld(ldx, Cx, split) && !bad_ld(lddy, Cout, split), "conv3x3_wgrad: bad pixel stride");
  if (!accumulate) {
    cudaError_t e = cudaMemsetAsync(dw, 0, sizeof(float) * 9ull * Cin * Cout, S(stream));
    PLUME_CHECK(e == cudaSuccess, "conv3x3_wgrad: cudaMemsetAsync failed");
  }
  WgradDesc d{};
  d.split = split;
  d.x = with_plane(view(x, ldx, N, H, W, Cx), ldx, split);
  d.dy[0] = with_plane(view(dy, lddy, N, H, W, Cout), lddy, split);
  d.num_dy_views = 1;
  d.num_taps = 9;
  d.Cin = Cin; d.Cout = Cout;
  d.dw = dw;            // dw[co][t][ci]
  d.s_co = 9ll * Cin;
  d.s_t = Cin;
  d.stream = S(stream);
  return launch_igemm_wgrad(d);
}
int plume_conv3x3_wgrad(const void* x, int ldx, const void* dy, int lddy, float* dw, int accumulate,
                        void* workspace, size_t workspace_bytes, int N, int H, int W, int Cin, int Cout,
                        plume_stream_t stream) {
  return impl_conv3x3_wgrad(x, ldx, dy, lddy, dw, accumulate, workspace, workspace_bytes, N, H, W, Cin, Cout, stream, 0);
}
int plume_conv3x3_wgrad_x3(const void* x, int ldx, const void* dy, int lddy, float* dw, int accumulate,
                        void* workspace, size_t workspace_bytes, int N, int H, int W, int Cin, int Cout,
                        plume_stream_t stream) {
  return impl_conv3x3_wgrad(x, ldx, dy, lddy, dw, accumulate, workspace, workspace_bytes, N, H, W, Cin, Cout, stream, 1);
}

static int impl_convT2x2_concat_fwd(const void* x, int ldx, const void* w, const float* bias, void* u, int ldu,
                              int N, int H, int W, int Cin, int Cout, plume_stream_t stream, int split) {
  PLUME_CHECK(x && w && u, "convT2x2_fwd: null pointer");
  PLUME_CHECK(!bad_ld(ldx, Cin, split) && !bad_ld(ldu, Cout, split), "convT2x2_fwd: bad pixel stride");
  FwdDesc d{};
  d.split = split;
  d.in[0] = with_plane(view(x, ldx, N, H, W, Cin), ldx, split);
  d.num_in_views = 1;
  d.num_taps = 1;
  d.Cin = Cin;
  d.wmat = w;
  for (int ij = 0; ij < 4; ++ij)
    d.out[ij] = with_plane(phase_view(u, ldu, N, H, W, Cout, ij >> 1, ij & 1), ldu, split);
  d.num_out_views = 4;
  d.cout_per_view = Cout;
  d.shift = bias;
  d.stream = S(stream);
  return launch_igemm_fwd(d);
}
int plume_convT2x2_concat_fwd(const void* x, int ldx, const void* w, const float* bias, void* u, int ldu,
                              int N, int H, int W, int Cin, int Cout, plume_stream_t stream) {
  return impl_convT2x2_concat_fwd(x, ldx, w, bias, u, ldu, N, H, W, Cin, Cout, stream, 0);
}
int plume_convT2x2_concat_fwd_x3(const void* x, int ldx, const void* w, const float* bias, void* u, int ldu,
                              int N, int H, int W, int Cin, int Cout, plume_stream_t stream) {
  return impl_convT2x2_concat_fwd(x, ldx, w, bias, u, ldu, N, H, W, Cin, Cout, stream, 1);
}

static int impl_convT2x2_dgrad(const void* du, int lddu, const void* w_dgrad, void* dx, int lddx, int N, int H,
                         int W, int Cin, int Cout, plume_stream_t stream, int split) {
  PLUME_CHECK(du && w_dgrad && dx, "convT2x2_dgrad: null pointer");
  PLUME_CHECK(!bad_ld(lddu, Cout, split) && !bad_ld(lddx, Cin, split), "convT2x2_dgrad: bad pixel stride");
  FwdDesc d{};
  d.split = split;
  for (int ij = 0; ij < 4; ++ij)
    d.in[ij] = with_plane(phase_view(du, lddu, N, H, W, Cout, ij >> 1, ij & 1), lddu, split);
  d.num_in_views = 4;
  d.num_taps = 4;
  d.Cin = Cout;
  d.wmat = w_dgrad;
  d.out[0] = with_plane(view(dx, lddx, N, H, W, Cin), lddx, split);
  d.num_out_views = 1;
  d.cout_per_view = Cin;
  d.stream = S(stream);
  return launch_igemm_fwd(d);
}
int plume_convT2x2_dgrad(const void* du, int lddu, const void* w_dgrad, void* dx, int lddx, int N, int H,
                         int W, int Cin, int Cout, plume_stream_t stream) {
  return impl_convT2x2_dgrad(du, lddu, w_dgrad, dx, lddx, N, H, W, Cin, Cout, stream, 0);
}
int plume_convT2x2_dgrad_x3(const void* du, int lddu, const void* w_dgrad, void* dx, int lddx, int N, int H,
                         int W, int Cin, int Cout, plume_stream_t stream) {
  return impl_convT2x2_dgrad(du, lddu, w_dgrad, dx, lddx, N, H, W, Cin, Cout, stream, 1);
}

static int impl_convT2x2_wgrad(const void* x, int ldx, const void* du, int lddu, float* dw, int accumulate,
                         void* workspace, size_t workspace_bytes, int N, int H, int W, int Cin, int Cout,
                         plume_stream_t stream, int split) {
  (void)workspace; (void)workspace_bytes;
  PLUME_CHECK(x && du && dw, "convT2x2_wgrad: null pointer");
  PLUME_CHECK(!bad_ld(ldx, Cin, split) && !bad_ld(lddu, Cout, split), "convT2x2_wgrad: bad pixel stride");
  if (!accumulate) {
    cudaError_t e = cudaMemsetAsync(dw, 0, sizeof(float) * 4ull * Cin * Cout, S(stream));
    PLUME_CHECK(e == cudaSuccess, "convT2x2_wgrad: cudaMemsetAsync failed");
  }
  WgradDesc d{};
  d.split = split;
  d.x = with_plane(view(x, ldx, N, H, W, Cin), ldx, split);
  for (int ij = 0; ij < 4; ++ij)
    d.dy[ij] = with_plane(phase_view(du, lddu, N, H, W, Cout, ij >> 1, ij & 1), lddu, split);
  d.num_dy_views = 4;
  d.num_taps = 4;
  d.Cin = Cin; d.Cout = Cout;
  d.dw = dw;            // dw[ij][co][ci]
  d.s_co = Cin;
  d.s_t = 1ll * Cout * Cin;
  d.stream = S(stream);
  return launch_igemm_wgrad(d);
}
int plume_convT2x2_wgrad(const void* x, int ldx, const void* du, int lddu, float* dw, int accumulate,
                         void* workspace, size_t workspace_bytes, int N, int H, int W, int Cin, int Cout,
                         plume_stream_t stream) {
  return impl_convT2x2_wgrad(x, ldx, du, lddu, dw, accumulate, workspace, workspace_bytes, N, H, W, Cin, Cout, stream, 0);
}
int plume_convT2x2_wgrad_x3(const void* x, int ldx, const void* du, int lddu, float* dw, int accumulate,
                         void* workspace, size_t workspace_bytes, int N, int H, int W, int Cin, int Cout,
                         plume_stream_t stream) {
  return impl_convT2x2_wgrad(x, ldx, du, lddu, dw, accumulate, workspace, workspace_bytes, N, H, W, Cin, Cout, stream, 1);
}

int plume_pack_conv3x3(const float* w, void* wf, void* wd, int Cout, int Cin, plume_stream_t stream) {
  PLUME_CHECK(w, "pack_conv3x3: null pointer");
  return pack_conv3x3(w, wf, wd, Cout, Cin, S(stream));
}
int plume_pack_blocks(int kind, int Cout, int Cin) { return pack_blocks(kind, Cout, Cin); }
int plume_pack_batch(const plume_pack_desc* descs, int n, int total_blocks, plume_stream_t stream) {
  PLUME_CHECK(descs || n <= 0, "pack_batch: null descriptor table");
  return pack_batch(descs, n, total_blocks, S(stream));
}
int plume_pack_convT2x2(const float* w, void* wf, void* wd, int Cout, int Cin, plume_stream_t stream) {
  PLUME_CHECK(w, "pack_convT2x2: null pointer");
  return pack_convT2x2(w, wf, wd, Cout, Cin, S(stream));
}

static int impl_pad_channels(const void* in, int Cs, void* out, int Cd, long long pixels,
                       plume_stream_t stream, int dt) {
  PLUME_CHECK(in && out, "pad_channels: null pointer");
  return pad_channels(in, Cs, out, Cd, pixels, dt, S(stream));
}
int plume_pad_channels(const void* in, int Cs, void* out, int Cd, long long pixels,
                       plume_stream_t stream) {
  return impl_pad_channels(in, Cs, out, Cd, pixels, stream, 0);
}
int plume_pad_channels_x3(const void* in, int Cs, void* out, int Cd, long long pixels,
                       plume_stream_t stream) {
  return impl_pad_channels(in, Cs, out, Cd, pixels, stream, 1);
}

int plume_bn_finalize(const double* sum, const double* sq, long long count, const float* gamma,
                      const float* beta, float eps, float momentum, float* running_mean,
                      float* running_var, float* scale, float* shift, float* mean, float* invstd, int C,
                      plume_stream_t stream) {
  PLUME_CHECK(sum && sq && scale && shift, "bn_finalize: null pointer");
  return bn_finalize(sum, sq, count, gamma, beta, eps, momentum, running_mean, running_var, scale,
                     shift, mean, invstd, C, S(stream));
}
int plume_bn_fold_eval(const float* gamma, const float* beta, const float* running_mean,
                       const float* running_var, const float* conv_bias, float eps, float* scale,
                       float* shift, int C, plume_stream_t stream) {
  PLUME_CHECK(running_mean && running_var && scale && shift, "bn_fold_eval: null pointer");
  return bn_fold_eval(gamma, beta, running_mean, running_var, conv_bias, eps, scale, shift, C,
                      S(stream));
}

static int impl_scale_shift_act(const void* y, int ldy, const float* scale, const float* shift, int relu,
                          void* a, int lda, long long pixels, int C, plume_stream_t stream, int dt) {
  PLUME_CHECK(y && a && scale && shift, "scale_shift_act: null pointer");
  PLUME_CHECK(y != a, "scale_shift_act: in-place operation is not supported");
  return scale_shift_act(y, ldy, scale, shift, relu, a, lda, pixels, C, dt, S(stream));
}
int plume_scale_shift_act(const void* y, int ldy, const float* scale, const float* shift, int relu,
                          void* a, int lda, long long pixels, int C, plume_stream_t stream) {
  return impl_scale_shift_act(y, ldy, scale, shift, relu, a, lda, pixels, C, stream, 0);
}
int plume_scale_shift_act_x3(const void* y, int ldy, const float* scale, const float* shift, int relu,
                          void* a, int lda, long long pixels, int C, plume_stream_t stream) {
  return impl_scale_shift_act(y, ldy, scale, shift, relu, a, lda, pixels, C, stream, 1);
}
static int impl_scale_shift_act_pool(const void* y, int ldy, const float* scale, const float* shift, int relu,
                               void* skip, int ldskip, void* pooled, int ldpooled, uint8_t* argmax,
                               int N, int H, int W, int C, plume_stream_t stream, int dt) {
  PLUME_CHECK(y && pooled && argmax && scale && shift, "scale_shift_act_pool: null pointer");
  return scale_shift_act_pool(y, ldy, scale, shift, relu, skip, ldskip, pooled, ldpooled, argmax, N, H,
                              W, C, dt, S(stream));
}
int plume_scale_shift_act_pool(const void* y, int ldy, const float* scale, const float* shift, int relu,
                               void* skip, int ldskip, void* pooled, int ldpooled, uint8_t* argmax,
                               int N, int H, int W, int C, plume_stream_t stream) {
  return impl_scale_shift_act_pool(y, ldy, scale, shift, relu, skip, ldskip, pooled, ldpooled, argmax, N, H, W, C, stream, 0);
}
int plume_scale_shift_act_pool_x3(const void* y, int ldy, const float* scale, const float* shift, int relu,
                               void* skip, int ldskip, void* pooled, int ldpooled, uint8_t* argmax,
                               int N, int H, int W, int C, plume_stream_t stream) {
  return impl_scale_shift_act_pool(y, ldy, scale, shift, relu, skip, ldskip, pooled, ldpooled, argmax, N, H, W, C, stream, 1);
}
static int impl_maxpool2x2_fwd(const void* x, int ldx, void* y, int ldy, uint8_t* argmax, int N, int H, int W,
                         int C, plume_stream_t stream, int dt) {
  PLUME_CHECK(x && y && argmax, "maxpool2x2_fwd: null pointer");
  return maxpool2x2_fwd(x, ldx, y, ldy, argmax, N, H, W, C, dt, S(stream));
}
int plume_maxpool2x2_fwd(const void* x, int ldx, void* y, int ldy, uint8_t* argmax, int N, int H, int W,
                         int C, plume_stream_t stream) {
  return impl_maxpool2x2_fwd(x, ldx, y, ldy, argmax, N, H, W, C, stream, 0);
}
int plume_maxpool2x2_fwd_x3(const void* x, int ldx, void* y, int ldy, uint8_t* argmax, int N, int H, int W,
                         int C, plume_stream_t stream) {
  return impl_maxpool2x2_fwd(x, ldx, y, ldy, argmax, N, H, W, C, stream, 1);
}
static int impl_maxpool2x2_bwd(const void* dy, int lddy, const uint8_t* argmax, const void* dskip,
                         int lddskip, void* dx, int lddx, int N, int H, int W, int C,
                         plume_stream_t stream, int dt) {
  PLUME_CHECK(dy && argmax && dx, "maxpool2x2_bwd: null pointer");
  return maxpool2x2_bwd(dy, lddy, argmax, dskip, lddskip, dx, lddx, N, H, W, C, dt, S(stream));
}
int plume_maxpool2x2_bwd(const void* dy, int lddy, const uint8_t* argmax, const void* dskip,
                         int lddskip, void* dx, int lddx, int N, int H, int W, int C,
                         plume_stream_t stream) {
  return impl_maxpool2x2_bwd(dy, lddy, argmax, dskip, lddskip, dx, lddx, N, H, W, C, stream, 0);
}
int plume_maxpool2x2_bwd_x3(const void* dy, int lddy, const uint8_t* argmax, const void* dskip,
                         int lddskip, void* dx, int lddx, int N, int H, int W, int C,
                         plume_stream_t stream) {
  return impl_maxpool2x2_bwd(dy, lddy, argmax, dskip, lddskip, dx, lddx, N, H, W, C, stream, 1);
}

static int impl_maxpool2x2_bwd_bn(const void* dy, int lddy, const uint8_t* argmax, const void* dskip, int lddskip,
                                  void* dx, int lddx, const void* y, int ldy, const float* scale, const float* shift,
                                  const float* mean, const float* invstd, int relu, float* sum_g, float* sum_gx,
                                  int N, int H, int W, int C, plume_stream_t stream, int dt) {
  PLUME_CHECK(dy && argmax && dx, "maxpool2x2_bwd_bn: null pointer");
  BnReduceArgs bn{y, ldy, scale, shift, mean, invstd, relu, sum_g, sum_gx, nullptr};
  return maxpool2x2_bwd(dy, lddy, argmax, dskip, lddskip, dx, lddx, N, H, W, C, dt, S(stream), &bn);
}
int plume_maxpool2x2_bwd_bn(const void* dy, int lddy, const uint8_t* argmax, const void* dskip, int lddskip,
                            void* dx, int lddx, const void* y, int ldy, const float* scale, const float* shift,
                            const float* mean, const float* invstd, int relu, float* sum_g, float* sum_gx, int N,
                            int H, int W, int C, plume_stream_t stream) {
  return impl_maxpool2x2_bwd_bn(dy, lddy, argmax, dskip, lddskip, dx, lddx, y, ldy, scale, shift, mean, invstd, relu,
                                sum_g, sum_gx, N, H, W, C, stream, 0);
}
int plume_maxpool2x2_bwd_bn_x3(const void* dy, int lddy, const uint8_t* argmax, const void* dskip, int lddskip,
                               void* dx, int lddx, const void* y, int ldy, const float* scale, const float* shift,
                               const float* mean, const float* invstd, int relu, float* sum_g, float* sum_gx, int N,
                               int H, int W, int C, plume_stream_t stream) {
  return impl_maxpool2x2_bwd_bn(dy, lddy, argmax, dskip, lddskip, dx, lddx, y, ldy, scale, shift, mean, invstd, relu,
                                sum_g, sum_gx, N, H, W, C, stream, 1);
}
static int impl_head_bwd_bn(const void* feat, int ldf, const float* w, const float* logits, const uint8_t* target,
                            const float* sums, float bce_weight, float dice_weight, float dice_eps, float grad_scale,
                            void* dfeat, int lddf, float* dw, float* db, const void* y, int ldy, const float* scale,
                            const float* shift, const float* mean, const float* invstd, int relu, float* sum_g,
                            float* sum_gx, long long pixels, int C, plume_stream_t stream, int dt) {
  PLUME_CHECK(feat && w && logits && target && sums && dfeat && dw && db, "head_bwd_bn: null pointer");
  BnReduceArgs bn{y, ldy, scale, shift, mean, invstd, relu, sum_g, sum_gx, nullptr};
  return head_bwd(feat, ldf, w, logits, target, sums, bce_weight, dice_weight, dice_eps, grad_scale, dfeat, lddf, dw,
                  db, pixels, C, dt, S(stream), &bn);
}
int plume_head_bwd_bn(const void* feat, int ldf, const float* w, const float* logits, const uint8_t* target,
                      const float* sums, float bce_weight, float dice_weight, float dice_eps, float grad_scale,
                      void* dfeat, int lddf, float* dw, float* db, const void* y, int ldy, const float* scale,
                      const float* shift, const float* mean, const float* invstd, int relu, float* sum_g,
                      float* sum_gx, long long pixels, int C, plume_stream_t stream) {
  return impl_head_bwd_bn(feat, ldf, w, logits, target, sums, bce_weight, dice_weight, dice_eps, grad_scale, dfeat,
                          lddf, dw, db, y, ldy, scale, shift, mean, invstd, relu, sum_g, sum_gx, pixels, C, stream, 0);
}
int plume_head_bwd_bn_x3(const void* feat, int ldf, const float* w, const float* logits, const uint8_t* target,
                         const float* sums, float bce_weight, float dice_weight, float dice_eps, float grad_scale,
                         void* dfeat, int lddf, float* dw, float* db, const void* y, int ldy, const float* scale,
                         const float* shift, const float* mean, const float* invstd, int relu, float* sum_g,
                         float* sum_gx, long long pixels, int C, plume_stream_t stream) {
  return impl_head_bwd_bn(feat, ldf, w, logits, target, sums, bce_weight, dice_weight, dice_eps, grad_scale, dfeat,
                          lddf, dw, db, y, ldy, scale, shift, mean, invstd, relu, sum_g, sum_gx, pixels, C, stream, 1);
}

static int impl_bn_bwd_reduce(const void* da, int ldda, const void* y, int ldy, const float* scale,
                        const float* shift, const float* mean, const float* invstd, int relu,
                        float* sum_g, float* sum_gx, long long pixels, int C, plume_stream_t stream, int dt) {
  PLUME_CHECK(da && y && scale && shift && mean && invstd && sum_g && sum_gx,
              "bn_bwd_reduce: null pointer");
  return bn_bwd_reduce(da, ldda, y, ldy, scale, shift, mean, invstd, relu, sum_g, sum_gx, pixels, C,
                       dt, S(stream));
}
int plume_bn_bwd_reduce(const void* da, int ldda, const void* y, int ldy, const float* scale,
                        const float* shift, const float* mean, const float* invstd, int relu,
                        float* sum_g, float* sum_gx, long long pixels, int C, plume_stream_t stream) {
  return impl_bn_bwd_reduce(da, ldda, y, ldy, scale, shift, mean, invstd, relu, sum_g, sum_gx, pixels, C, stream, 0);
}
int plume_bn_bwd_reduce_x3(const void* da, int ldda, const void* y, int ldy, const float* scale,
                        const float* shift, const float* mean, const float* invstd, int relu,
                        float* sum_g, float* sum_gx, long long pixels, int C, plume_stream_t stream) {
  return impl_bn_bwd_reduce(da, ldda, y, ldy, scale, shift, mean, invstd, relu, sum_g, sum_gx, pixels, C, stream, 1);
}
static int impl_bn_bwd_apply(const void* da, int ldda, const void* y, int ldy, const float* scale,
                       const float* shift, const float* mean, const float* invstd, int relu,
                       const float* sum_g, const float* sum_gx, void* dy, int lddy, float* sum_dy,
                       float* dgamma, float* dbeta, int accumulate, long long pixels, int C,
                       plume_stream_t stream, int dt) {
  PLUME_CHECK(da && y && scale && shift && mean && invstd && sum_g && sum_gx && dy,
              "bn_bwd_apply: null pointer");
  return bn_bwd_apply(da, ldda, y, ldy, scale, shift, mean, invstd, relu, sum_g, sum_gx, dy, lddy,
                      sum_dy, dgamma, dbeta, accumulate, pixels, C, dt, S(stream));
}
int plume_bn_bwd_apply(const void* da, int ldda, const void* y, int ldy, const float* scale,
                       const float* shift, const float* mean, const float* invstd, int relu,
                       const float* sum_g, const float* sum_gx, void* dy, int lddy, float* sum_dy,
                       float* dgamma, float* dbeta, int accumulate, long long pixels, int C,
                       plume_stream_t stream) {
  return impl_bn_bwd_apply(da, ldda, y, ldy, scale, shift, mean, invstd, relu, sum_g, sum_gx, dy, lddy, sum_dy, dgamma, dbeta, accumulate, pixels, C, stream, 0);
}
int plume_bn_bwd_apply_x3(const void* da, int ldda, const void* y, int ldy, const float* scale,
                       const float* shift, const float* mean, const float* invstd, int relu,
                       const float* sum_g, const float* sum_gx, void* dy, int lddy, float* sum_dy,
                       float* dgamma, float* dbeta, int accumulate, long long pixels, int C,
                       plume_stream_t stream) {
  return impl_bn_bwd_apply(da, ldda, y, ldy, scale, shift, mean, invstd, relu, sum_g, sum_gx, dy, lddy, sum_dy, dgamma, dbeta, accumulate, pixels, C, stream, 1);
}
static int impl_relu_bwd(const void* da, int ldda, const void* a, int lda, void* dy, int lddy, float* sum_dy,
                   long long pixels, int C, plume_stream_t stream, int dt) {
  PLUME_CHECK(da && a && dy, "relu_bwd: null pointer");
  return relu_bwd(da, ldda, a, lda, dy, lddy, sum_dy, pixels, C, dt, S(stream));
}
int plume_relu_bwd(const void* da, int ldda, const void* a, int lda, void* dy, int lddy, float* sum_dy,
                   long long pixels, int C, plume_stream_t stream) {
  return impl_relu_bwd(da, ldda, a, lda, dy, lddy, sum_dy, pixels, C, stream, 0);
}
int plume_relu_bwd_x3(const void* da, int ldda, const void* a, int lda, void* dy, int lddy, float* sum_dy,
                   long long pixels, int C, plume_stream_t stream) {
  return impl_relu_bwd(da, ldda, a, lda, dy, lddy, sum_dy, pixels, C, stream, 1);
}
static int impl_channel_sum(const void* x, int ldx, float* out, long long pixels, int C,
                      plume_stream_t stream, int dt) {
  PLUME_CHECK(x && out, "channel_sum: null pointer");
  return channel_sum(x, ldx, out, pixels, C, dt, S(stream));
}
int plume_channel_sum(const void* x, int ldx, float* out, long long pixels, int C,
                      plume_stream_t stream) {
  return impl_channel_sum(x, ldx, out, pixels, C, stream, 0);
}
int plume_channel_sum_x3(const void* x, int ldx, float* out, long long pixels, int C,
                      plume_stream_t stream) {
  return impl_channel_sum(x, ldx, out, pixels, C, stream, 1);
}

static int impl_head_fwd(const void* feat, int ldf, const float* w, const float* b, const uint8_t* target,
                   float* logits, float* sums, long long pixels, int C, plume_stream_t stream, int dt) {
  PLUME_CHECK(feat && w && logits, "head_fwd: null pointer");
  PLUME_CHECK(!target || sums, "head_fwd: target given without sums");
  return head_fwd(feat, ldf, w, b, target, logits, sums, pixels, C, dt, S(stream));
}
int plume_head_fwd(const void* feat, int ldf, const float* w, const float* b, const uint8_t* target,
                   float* logits, float* sums, long long pixels, int C, plume_stream_t stream) {
  return impl_head_fwd(feat, ldf, w, b, target, logits, sums, pixels, C, stream, 0);
}
int plume_head_fwd_x3(const void* feat, int ldf, const float* w, const float* b, const uint8_t* target,
                   float* logits, float* sums, long long pixels, int C, plume_stream_t stream) {
  return impl_head_fwd(feat, ldf, w, b, target, logits, sums, pixels, C, stream, 1);
}
int plume_head_loss(const float* sums, long long pixels, float bce_weight, float dice_weight,
                    float dice_eps, float* loss_out, plume_stream_t stream) {
  PLUME_CHECK(sums && loss_out, "head_loss: null pointer");
  return head_loss(sums, pixels, bce_weight, dice_weight, dice_eps, loss_out, S(stream));
}
static int impl_head_bwd(const void* feat, int ldf, const float* w, const float* logits, const uint8_t* target,
                   const float* sums, float bce_weight, float dice_weight, float dice_eps,
                   float grad_scale, void* dfeat, int lddf, float* dw, float* db, long long pixels,
                   int C, plume_stream_t stream, int dt) {
  PLUME_CHECK(feat && w && logits && target && sums && dfeat && dw && db, "head_bwd: null pointer");
  return head_bwd(feat, ldf, w, logits, target, sums, bce_weight, dice_weight, dice_eps, grad_scale,
                  dfeat, lddf, dw, db, pixels, C, dt, S(stream));
}
int plume_head_bwd(const void* feat, int ldf, const float* w, const float* logits, const uint8_t* target,
                   const float* sums, float bce_weight, float dice_weight, float dice_eps,
                   float grad_scale, void* dfeat, int lddf, float* dw, float* db, long long pixels,
                   int C, plume_stream_t stream) {
  return impl_head_bwd(feat, ldf, w, logits, target, sums, bce_weight, dice_weight, dice_eps, grad_scale, dfeat, lddf, dw, db, pixels, C, stream, 0);
}
int plume_head_bwd_x3(const void* feat, int ldf, const float* w, const float* logits, const uint8_t* target,
                   const float* sums, float bce_weight, float dice_weight, float dice_eps,
                   float grad_scale, void* dfeat, int lddf, float* dw, float* db, long long pixels,
                   int C, plume_stream_t stream) {
  return impl_head_bwd(feat, ldf, w, logits, target, sums, bce_weight, dice_weight, dice_eps, grad_scale, dfeat, lddf, dw, db, pixels, C, stream, 1);
}

int plume_cast_f32_bf16(const float* in, void* out, long long n, plume_stream_t stream) {
  PLUME_CHECK(n <= 0 || (in && out), "cast_f32_bf16: null pointer");
  return cast_f32_bf16(in, out, n, S(stream));
}
int plume_cast_bf16_f32(const void* in, float* out, long long n, plume_stream_t stream) {
  PLUME_CHECK(n <= 0 || (in && out), "cast_bf16_f32: null pointer");
  return cast_bf16_f32(in, out, n, S(stream));
}

int plume_adam(float* param, const float* grad, float* m, float* v, long long n, double lr, double beta1,
               double beta2, double eps, int step, float grad_scale, plume_stream_t stream) {
  PLUME_CHECK(param && grad && m && v, "adam: null pointer");
  PLUME_CHECK((reinterpret_cast<uintptr_t>(param) | reinterpret_cast<uintptr_t>(grad) |
               reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v)) % 16 == 0,
              "adam: buffers must be 16-byte aligned");
  return adam(param, grad, m, v, n, lr, beta1, beta2, eps, step, grad_scale, S(stream));
}

int plume_adam_dev(float* param, const float* grad, float* m, float* v, long long n, const float* coef,
                   plume_stream_t stream) {
  PLUME_CHECK(param && grad && m && v && coef, "adam_dev: null pointer");
  PLUME_CHECK((reinterpret_cast<uintptr_t>(param) | reinterpret_cast<uintptr_t>(grad) |
               reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v)) % 16 == 0,
              "adam_dev: buffers must be 16-byte aligned");
  return adam_dev(param, grad, m, v, n, coef, S(stream));
}

static int impl_extract_tiles(const void* scene, int Hs, int Ws, int Cs, const int* ys, const int* xs,
                        int count, int T, void* tiles, int Cd, plume_stream_t stream, int dt) {
  PLUME_CHECK(scene && ys && xs && tiles, "extract_tiles: null pointer");
  return extract_tiles(scene, Hs, Ws, Cs, ys, xs, count, T, tiles, Cd, dt, S(stream));
}
int plume_extract_tiles(const void* scene, int Hs, int Ws, int Cs, const int* ys, const int* xs,
                        int count, int T, void* tiles, int Cd, plume_stream_t stream) {
  return impl_extract_tiles(scene, Hs, Ws, Cs, ys, xs, count, T, tiles, Cd, stream, 0);
}
int plume_extract_tiles_x3(const void* scene, int Hs, int Ws, int Cs, const int* ys, const int* xs,
                        int count, int T, void* tiles, int Cd, plume_stream_t stream) {
  return impl_extract_tiles(scene, Hs, Ws, Cs, ys, xs, count, T, tiles, Cd, stream, 1);
}
int plume_stitch_threshold(const float* logits, const int* ys, const int* xs, int count, int T,
                           int margin, float logit_threshold, uint8_t* mask, float* prob, int Hs, int Ws,
                           plume_stream_t stream) {
  PLUME_CHECK(logits && ys && xs && mask, "stitch_threshold: null pointer");
  return stitch_threshold(logits, ys, xs, count, T, margin, logit_threshold, mask, prob, Hs, Ws,
                          S(stream));
}

int plume_rasterize_hulls(const int* verts_xy, const int* poly_offsets, const int* bbox, int n_polys,
                          const int* ys, const int* xs, int count, int Hm, int Wm, uint8_t* masks,
                          plume_stream_t stream) {
  PLUME_CHECK(masks, "rasterize_hulls: null mask pointer");
  PLUME_CHECK(n_polys >= 0 && (n_polys == 0 || (verts_xy && poly_offsets && bbox)),
              "rasterize_hulls: null polygon arrays");
  PLUME_CHECK((reinterpret_cast<uintptr_t>(bbox) & 15) == 0, "rasterize_hulls: bbox must be 16-byte aligned");
  return rasterize_hulls(verts_xy, poly_offsets, bbox, n_polys, ys, xs, count, Hm, Wm, masks, S(stream));
}

size_t plume_locate_fires_workspace_bytes(int n_fires) { return locate_fires_workspace_bytes(n_fires); }
int plume_locate_fires(const double* lats, const double* lons, int H, int W, const double* fire_lat,
                       const double* fire_lon, int n_fires, double half_box_deg, void* workspace,
                       size_t workspace_bytes, int* out_row_col, plume_stream_t stream) {
  PLUME_CHECK(n_fires <= 0 || (lats && lons && fire_lat && fire_lon && workspace && out_row_col),
              "locate_fires: null pointer");
  return locate_fires(lats, lons, H, W, fire_lat, fire_lon, n_fires, half_box_deg, workspace, workspace_bytes,
                      out_row_col, S(stream));
}

int plume_threshold_masks(const float* aod, int H, int W, const double* thresholds, int T, uint8_t* masks,
                          plume_stream_t stream) {
  PLUME_CHECK(T <= 0 || (aod && thresholds && masks), "threshold_masks: null pointer");
  return threshold_masks(aod, 0, H, W, thresholds, T, masks, S(stream));
}
int plume_threshold_masks_f64(const double* aod, int H, int W, const double* thresholds, int T, uint8_t* masks,
                              plume_stream_t stream) {
  PLUME_CHECK(T <= 0 || (aod && thresholds && masks), "threshold_masks_f64: null pointer");
  return threshold_masks(aod, 1, H, W, thresholds, T, masks, S(stream));
}
int plume_label_components(const uint8_t* masks, int T, int H, int W, int* labels, int* sizes,
                           plume_stream_t stream) {
  PLUME_CHECK(T <= 0 || (masks && labels && sizes), "label_components: null pointer");
  return label_components(masks, T, H, W, labels, sizes, S(stream));
}
int plume_fire_extents(const int* labels, const int* sizes, int T, int H, int W, const int* fire_row_col,
                       int n_fires, int win, int* extents, plume_stream_t stream) {
  PLUME_CHECK(T <= 0 || n_fires <= 0 || (labels && sizes && fire_row_col && extents), "fire_extents: null pointer");
  return fire_extents(labels, sizes, T, H, W, fire_row_col, n_fires, win, extents, S(stream));
}
size_t plume_sweep_workspace_bytes(int H, int W, int T) { return sweep_workspace_bytes(H, W, T); }
int plume_threshold_mask_bits(const float* aod, int H, int W, const double* thresholds, int T, uint32_t* bits,
                              plume_stream_t stream) {
  PLUME_CHECK(T <= 0 || (aod && thresholds && bits), "threshold_mask_bits: null pointer");
  return threshold_mask_bits(aod, 0, H, W, thresholds, T, bits, S(stream));
}
int plume_threshold_mask_bits_f64(const double* aod, int H, int W, const double* thresholds, int T, uint32_t* bits,
                                  plume_stream_t stream) {
  PLUME_CHECK(T <= 0 || (aod && thresholds && bits), "threshold_mask_bits_f64: null pointer");
  return threshold_mask_bits(aod, 1, H, W, thresholds, T, bits, S(stream));
}
int plume_pack_mask_bits(const uint8_t* masks, int T, int H, int W, uint32_t* bits, plume_stream_t stream) {
  PLUME_CHECK(T <= 0 || (masks && bits), "pack_mask_bits: null pointer");
  return pack_mask_bits(masks, T, H, W, bits, S(stream));
}
int plume_bits_extents(const uint32_t* bits, int T, int H, int W, const int* fire_row_col, int n_fires, int win,
                       void* workspace, size_t workspace_bytes, int* extents, plume_stream_t stream) {
  PLUME_CHECK(T <= 0 || n_fires <= 0 || (bits && fire_row_col && workspace && extents), "bits_extents: null pointer");
  return bits_extents(bits, T, H, W, fire_row_col, n_fires, win, workspace, workspace_bytes, extents, S(stream));
}
int plume_fire_components(const uint32_t* bits, int T, int H, int W, const int* fire_row_col, const int* plane_of_fire,
                          int n_fires, int win, const void* workspace, size_t workspace_bytes, uint32_t* component_bits,
                          int* stats, plume_stream_t stream) {
  PLUME_CHECK(T <= 0 || n_fires <= 0 || (bits && fire_row_col && plane_of_fire && workspace && component_bits && stats),
              "fire_components: null pointer");
  return fire_components(bits, T, H, W, fire_row_col, plane_of_fire, n_fires, win, workspace, workspace_bytes,
                         component_bits, stats, S(stream));
}
int plume_sweep_extents(const float* aod, int H, int W, const double* thresholds, int T, const int* fire_row_col,
                        int n_fires, int win, void* workspace, size_t workspace_bytes, int* extents,
                        plume_stream_t stream) {
  PLUME_CHECK(T <= 0 || n_fires <= 0 || (aod && thresholds && fire_row_col && workspace && extents),
              "sweep_extents: null pointer");
  return sweep_extents(aod, 0, H, W, thresholds, T, fire_row_col, n_fires, win, workspace, workspace_bytes, extents,
                       S(stream));
}
int plume_sweep_extents_f64(const double* aod, int H, int W, const double* thresholds, int T, const int* fire_row_col,
                            int n_fires, int win, void* workspace, size_t workspace_bytes, int* extents,
                            plume_stream_t stream) {
  PLUME_CHECK(T <= 0 || n_fires <= 0 || (aod && thresholds && fire_row_col && workspace && extents),
              "sweep_extents_f64: null pointer");
  return sweep_extents(aod, 1, H, W, thresholds, T, fire_row_col, n_fires, win, workspace, workspace_bytes, extents,
                       S(stream));
}

size_t plume_fill_nearest_workspace_bytes(int H, int W) { return fill_nearest_workspace_bytes(H, W); }
int plume_fill_nearest(const float* aod, int H, int W, float null_value, void* workspace, size_t workspace_bytes,
                       float* out, plume_stream_t stream) {
  PLUME_CHECK(H <= 0 || W <= 0 || (aod && workspace && out), "fill_nearest: null pointer");
  return fill_nearest(aod, 0, H, W, null_value, workspace, workspace_bytes, out, S(stream));
}
int plume_fill_nearest_f64(const double* aod, int H, int W, double null_value, void* workspace, size_t workspace_bytes,
                           double* out, plume_stream_t stream) {
  PLUME_CHECK(H <= 0 || W <= 0 || (aod && workspace && out), "fill_nearest_f64: null pointer");
  return fill_nearest(aod, 1, H, W, null_value, workspace, workspace_bytes, out, S(stream));
}

int plume_utm_zone_histogram(const double* lon, long long n, int* hist64, plume_stream_t stream) {
  PLUME_CHECK(hist64 && (n <= 0 || lon), "utm_zone_histogram: null pointer");
  return utm_zone_histogram(lon, n, hist64, S(stream));
}
int plume_sinusoidal_grid_latlon(double x_start, double x_stop, double y_start, double y_stop, int ny, int nx,
                                 double radius, double* lat, double* lon, plume_stream_t stream) {
  PLUME_CHECK(ny <= 0 || nx <= 0 || (lat && lon), "sinusoidal_grid_latlon: null pointer");
  return sinusoidal_grid_latlon(x_start, x_stop, y_start, y_stop, ny, nx, radius, lat, lon, S(stream));
}
int plume_utm_forward(const double* lat, const double* lon, long long n, int zone, double* x, double* y,
                      plume_stream_t stream) {
  PLUME_CHECK(n <= 0 || (lat && lon && x && y), "utm_forward: null pointer");
  return utm_forward(lat, lon, n, zone, x, y, S(stream));
}
int plume_utm_inverse(const double* x, const double* y, long long n, int zone, double* lat, double* lon,
                      plume_stream_t stream) {
  PLUME_CHECK(n <= 0 || (lat && lon && x && y), "utm_inverse: null pointer");
  return utm_inverse(x, y, n, zone, lat, lon, S(stream));
}
size_t plume_resample_workspace_bytes(int n_src, double min_x, double min_y, double max_x, double max_y,
                                      double radius) {
  return resample_workspace_bytes(n_src, min_x, min_y, max_x, max_y, radius);
}
int plume_resample_nearest_index(const double* src_lat, const double* src_lon, int n_src, int zone, double min_x,
                                 double min_y, double max_x, double max_y, int x_size, int y_size, double radius,
                                 void* workspace, size_t workspace_bytes, int* out_idx, plume_stream_t stream) {
  PLUME_CHECK(out_idx && (n_src <= 0 || (src_lat && src_lon)), "resample_nearest_index: null pointer");
  return resample_nearest_index(src_lat, src_lon, n_src, zone, min_x, min_y, max_x, max_y, x_size, y_size, radius,
                                workspace, workspace_bytes, out_idx, S(stream));
}
int plume_gather_fill(const void* src, int elem_bytes, const int* idx, long long n, double fill_value, void* out,
                      plume_stream_t stream) {
  PLUME_CHECK(n <= 0 || (src && idx && out), "gather_fill: null pointer");
  return gather_fill(src, elem_bytes, idx, n, fill_value, out, S(stream));
}

}  // extern "C"
