// Internal C++ interface of the tensor-core implicit-GEMM kernels (igemm.cu).
#pragma once
#include "tmap.cuh"

namespace plume {

// Forward-type implicit GEMM:  OUT[pixel, n] = epilogue( sum_{tap, c} IN_tap[pixel + shift(tap), c] * Wmat[n, tap*Cin + c] )
//   conv3x3 fwd   : 9 taps, shifts (-1..1)^2 on one input view, one output view
//   conv3x3 dgrad : same kernel on dY with the rotated/transposed weight matrix
//   convT2x2 fwd  : 1 tap, four strided output views (one per (i,j)), Wmat rows = (ij, co)
//   convT2x2 dgrad: 4 taps, each reading its own strided view of dU, no shift
struct FwdDesc {
  ActView in[4];
  int num_in_views;   // 1, or 4 when every tap reads its own view
  int num_taps;       // 9, 4 or 1
  int Cin;            // channels per tap (multiple of 64)
  const void* wmat;   // bf16 [Ntotal][num_taps*Cin]
  ActView out[4];
  int num_out_views;  // 1, or 4 (convT fwd)
  int cout_per_view;  // multiple of 64
  const float* scale; // per output channel (index = n % cout_per_view), may be null (=1)
  const float* shift; // per output channel, may be null (=0)
  int relu;
  double* stat_sum;   // optional per-channel sum / sum of squares of the bf16 outputs (fp64 accumulators)
  double* stat_sq;
  cudaStream_t stream;
  // bf16x3 high-precision mode: every view is a hi plane with its lo plane `view.plane` elements further, wmat is
  // the hi matrix followed by the lo matrix (same shape); three MMA passes (hi*hi + hi*lo + lo*hi)
  int split;
};
int launch_igemm_fwd(const FwdDesc& d);

// Weight-gradient implicit GEMM:
//   dw[co*s_co + tap*s_t + ci] += sum_{pixels} X_tap[pixel + shift(tap), ci] * DY_tap[pixel, co]   (fp32 atomics)
//   conv3x3 : X shifted per tap, one dY view
//   convT2x2: X unshifted, dY view per tap (the four strided views of dU)
struct WgradDesc {
  ActView x;
  ActView dy[4];
  int num_dy_views;   // 1 (conv3x3) or 4 (convT)
  int num_taps;       // 9 or 4
  int Cin;            // 64, or a multiple of 128
  int Cout;           // multiple of 64
  float* dw;          // fp32 gradient, accumulated into (caller zeroes it for a plain assignment)
  long long s_co, s_t;
  cudaStream_t stream;
  int split;          // bf16x3 mode (see FwdDesc): x and dy are hi/lo plane pairs
  float* det_part;    // set by launch_igemm_wgrad in deterministic mode: [splits][taps*Cin*Cout] partial sums
};
// Chooses the K split for a problem; returns the number of splits (>=1).
int wgrad_plan(int N, int H, int W, int num_taps, int Cin, int Cout);
int launch_igemm_wgrad(const WgradDesc& d);

// SMs to leave free for concurrent collective kernels (see igemm.cu)
void set_sm_margin(int k);
int get_sm_margin();

int read_debug_word();
// Diagnostics: when set, conv3 launches write per-CTA cycle counters ([grid][8] int64) there.
void set_prof_buffer(long long* buf);  // last watchdog tag written by a trapped kernel (0 if none)

}  // namespace plume
