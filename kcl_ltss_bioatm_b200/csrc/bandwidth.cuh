// Internal C++ interface of the bandwidth kernels (bandwidth.cu; `dt` = activation storage: 0 bf16, 1 fp32), the label-geometry kernels (geometry.cu)
// and the threshold-sweep kernels (sweep.cu).  Input and output buffers of one call must not alias: inputs
// are read through the non-coherent path.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <string>

#include "../../include/plume_b200.h"

namespace plume {

void set_error(const std::string& msg);

// Deterministic mode (PLUME_DETERMINISTIC=1 / plume_set_deterministic): reductions write per-block partial sums to
// a library-owned scratch (`which`: 0 = main / chain stream users, 1 = weight gradients on the side stream) and a
// second kernel adds them in block order instead of floating-point atomics.
bool deterministic();
void set_deterministic(int on);
void* det_scratch(int which, size_t bytes);
int ordered_sum(const float* part, int blocks, int n, float* out, cudaStream_t s);          // out[c] += sum_b part[b*n+c]
int ordered_sum_f64(const double* part, int blocks, int row, int n, double* out, cudaStream_t s);  // rows `row` apart

int pad_channels(const void* in, int Cs, void* out, int Cd, long long pixels, int dt, cudaStream_t s);
int bn_finalize(const double* sum, const double* sq, long long count, const float* gamma,
                const float* beta, float eps, float momentum, float* running_mean,
                float* running_var, float* scale, float* shift, float* mean, float* invstd, int C,
                cudaStream_t s);
int bn_fold_eval(const float* gamma, const float* beta, const float* rm, const float* rv,
                 const float* bias, float eps, float* scale, float* shift, int C, cudaStream_t s);
int scale_shift_act(const void* y, int ldy, const float* scale, const float* shift, int relu, void* a,
                    int lda, long long pixels, int C, int dt, cudaStream_t s);
int scale_shift_act_pool(const void* y, int ldy, const float* scale, const float* shift, int relu,
                         void* skip, int ldskip, void* pooled, int ldpooled, uint8_t* argmax, int N,
                         int H, int W, int C, int dt, cudaStream_t s);
int maxpool2x2_fwd(const void* x, int ldx, void* y, int ldy, uint8_t* argmax, int N, int H, int W,
                   int C, int dt, cudaStream_t s);
// Optional fused BatchNorm-backward reduction of the gradient a kernel produces (see bandwidth.cu)
struct BnReduceArgs {
  const void* y;            // raw conv output of the BatchNorm layer (activation storage format), pixel stride ldy
  long long ldy;
  const float *scale, *shift, *mean, *invstd;
  int relu;
  float *sum_g, *sum_gx;    // added to (zeroed per backward pass by the caller)
  float* det_part;          // internal: deterministic-mode scratch
};
int maxpool2x2_bwd(const void* dy, int lddy, const uint8_t* argmax, const void* dskip, int lddskip,
                   void* dx, int lddx, int N, int H, int W, int C, int dt, cudaStream_t s,
                   const BnReduceArgs* bn = nullptr);
int bn_bwd_reduce(const void* da, int ldda, const void* y, int ldy, const float* scale,
                  const float* shift, const float* mean, const float* invstd, int relu, float* sum_g,
                  float* sum_gx, long long pixels, int C, int dt, cudaStream_t s);
int bn_bwd_apply(const void* da, int ldda, const void* y, int ldy, const float* scale,
                 const float* shift, const float* mean, const float* invstd, int relu,
                 const float* sum_g, const float* sum_gx, void* dy, int lddy, float* sum_dy,
                 float* dgamma, float* dbeta, int accumulate, long long pixels, int C, int dt, cudaStream_t s);
int relu_bwd(const void* da, int ldda, const void* a, int lda, void* dy, int lddy, float* sum_dy,
             long long pixels, int C, int dt, cudaStream_t s);
int channel_sum(const void* x, int ldx, float* out, long long pixels, int C, int dt, cudaStream_t s);
int head_fwd(const void* feat, int ldf, const float* w, const float* b, const uint8_t* target,
             float* logits, float* sums, long long pixels, int C, int dt, cudaStream_t s);
int head_loss(const float* sums, long long pixels, float bce_w, float dice_w, float eps, float* out,
              cudaStream_t s);
int head_bwd(const void* feat, int ldf, const float* w, const float* logits, const uint8_t* target,
             const float* sums, float bce_w, float dice_w, float eps, float grad_scale, void* dfeat,
             int lddf, float* dw, float* db, long long pixels, int C, int dt, cudaStream_t s,
             const BnReduceArgs* bn = nullptr);
int cast_f32_bf16(const float* in, void* out, long long n, cudaStream_t s);
int cast_bf16_f32(const void* in, float* out, long long n, cudaStream_t s);
int adam(float* param, const float* grad, float* m, float* v, long long n, double lr, double beta1,
         double beta2, double eps, int step, float grad_scale, cudaStream_t s);
int adam_dev(float* param, const float* grad, float* m, float* v, long long n, const float* coef,
             cudaStream_t s);
int pack_conv3x3(const float* w, void* wf, void* wd, int Cout, int Cin, cudaStream_t s);
int pack_convT2x2(const float* w, void* wf, void* wd, int Cout, int Cin, cudaStream_t s);
int pack_blocks(int kind, int Cout, int Cin);
int pack_batch(const plume_pack_desc* descs, int n, int total_blocks, cudaStream_t s);
int extract_tiles(const void* scene, int Hs, int Ws, int Cs, const int* ys, const int* xs, int count,
                  int T, void* tiles, int Cd, int dt, cudaStream_t s);
int stitch_threshold(const float* logits, const int* ys, const int* xs, int count, int T, int margin,
                     float thr, uint8_t* mask, float* prob, int Hs, int Ws, cudaStream_t s);

int rasterize_hulls(const int* verts, const int* offs, const int* bbox, int n_polys, const int* ys,
                    const int* xs, int count, int Hm, int Wm, uint8_t* masks, cudaStream_t s);

size_t locate_fires_workspace_bytes(int n_fires);
int locate_fires(const double* lats, const double* lons, int H, int W, const double* fire_lat,
                 const double* fire_lon, int n_fires, double half_box, void* workspace, size_t workspace_bytes,
                 int* out_rc, cudaStream_t s);

// aod: float32 (f64 = 0) or float64 (f64 = 1) image
int threshold_masks(const void* aod, int f64, int H, int W, const double* thr, int T, uint8_t* masks, cudaStream_t s);
int label_components(const uint8_t* masks, int T, int H, int W, int* labels, int* sizes, cudaStream_t s);
int fire_extents(const int* labels, const int* sizes, int T, int H, int W, const int* fire_rc, int n_fires, int win,
                 int* extents, cudaStream_t s);
// bit-plane form (32 pixels per word; sweep_bits.cuh)
size_t sweep_workspace_bytes(int H, int W, int T);
int threshold_mask_bits(const void* aod, int f64, int H, int W, const double* thr, int T, uint32_t* bits,
                        cudaStream_t s);
int pack_mask_bits(const uint8_t* masks, int T, int H, int W, uint32_t* bits, cudaStream_t s);
int bits_extents(const uint32_t* bits, int T, int H, int W, const int* fire_rc, int n_fires, int win, void* workspace,
                 size_t workspace_bytes, int* extents, cudaStream_t s);
int fire_components(const uint32_t* bits, int T, int H, int W, const int* fire_rc, const int* plane_of_fire, int n_fires,
                    int win, const void* workspace, size_t workspace_bytes, uint32_t* comp, int* stats, cudaStream_t s);
int sweep_extents(const void* aod, int f64, int H, int W, const double* thr, int T, const int* fire_rc, int n_fires, int win,
                  void* workspace, size_t workspace_bytes, int* extents, cudaStream_t s);

// nearest-valid fill of the AOD grid (fill.cu)
size_t fill_nearest_workspace_bytes(int H, int W);
int fill_nearest(const void* aod, int f64, int H, int W, double null_value, void* workspace, size_t workspace_bytes,
                 void* out, cudaStream_t s);

// UTM projection + nearest-neighbour swath -> grid resampling (resample.cu)
int utm_zone_histogram(const double* lon, long long n, int* hist64, cudaStream_t s);
int utm_forward(const double* lat, const double* lon, long long n, int zone, double* x, double* y, cudaStream_t s);
int utm_inverse(const double* x, const double* y, long long n, int zone, double* lat, double* lon, cudaStream_t s);
int sinusoidal_grid_latlon(double x_start, double x_stop, double y_start, double y_stop, int ny, int nx, double radius,
                           double* lat, double* lon, cudaStream_t s);
size_t resample_workspace_bytes(int n_src, double min_x, double min_y, double max_x, double max_y, double radius);
int resample_nearest_index(const double* src_lat, const double* src_lon, int n_src, int zone, double min_x,
                           double min_y, double max_x, double max_y, int x_size, int y_size, double radius,
                           void* workspace, size_t workspace_bytes, int* out_idx, cudaStream_t s);
int gather_fill(const void* src, int elem_bytes, const int* idx, long long n, double fill, void* out,
                cudaStream_t s);

}  // namespace plume
