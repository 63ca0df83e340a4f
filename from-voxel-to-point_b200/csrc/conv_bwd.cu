// Filter gradient of the sparse convolution (not on the inference hot path).
//
// Replaces, in indiceConvBackward (pcdet/ops/spconv/include/spconv/spconv_ops.h:365-457), the per-offset
// gather(features) + gather(outGrad) + torch::mm_out(filterGradSub, in^T, out) sequence and its host loop over the
// D2H-copied pair counts (:378, :399-436):
//
//     dW[k][ci][co] = sum over pairs (in, out) of offset k of  X[in][ci] * dY[out][co]
//
// One launch, no host synchronisation: the work is cut into (offset, slice of the output rows) items; a CTA compacts
// the rows of its slice that have a neighbour at its offset (ballot + prefix), stages 16 (X row, dY row) pairs at a
// time in shared memory and accumulates the Cin x Cout outer products in registers (thread (ty, tx) owns the entries
// ci = ty + 16 a, co = tx + 16 b); the slice's partial sum goes to dW with one atomicAdd per entry.  fp32 FMA
// arithmetic like the reference's sgemm; the order of the partial sums is not fixed, so results agree with the
// reference to rounding (1e-6 relative), not bitwise.  The input gradient needs no kernel of its own: it is the
// forward contraction on the transposed map with W^T (spconv/ops.py:indice_conv_backward).
#include "common.cuh"

namespace fv2p {
namespace {

constexpr int kStageRows = 16;
constexpr int kMaxC = 128;

__global__ void __launch_bounds__(kThreads)
grad_filters_kernel(const float *__restrict__ x, const float *__restrict__ dy, const int *__restrict__ nbr,
                    int64_t nbr_stride, int64_t n_out_cap, const int *__restrict__ n_out_dev, int kvol, int cin,
                    int cout, int slices, float *dw) {
  __shared__ float xs[kStageRows][kMaxC];
  __shared__ float ys[kStageRows][kMaxC];
  __shared__ int list_in[kThreads], list_out[kThreads];
  __shared__ int warp_cnt[kThreads / 32];
  int n_out = n_out_dev ? *n_out_dev : (int)n_out_cap;
  if (n_out > n_out_cap) n_out = (int)n_out_cap;
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int rows_per_slice = (n_out + slices - 1) / slices;
  for (int item = blockIdx.x; item < kvol * slices; item += gridDim.x) {
    const int k = item / slices, sl = item - k * slices;
    const int r0 = sl * rows_per_slice, r1 = min(n_out, r0 + rows_per_slice);
    float acc[kMaxC / 16][kMaxC / 16];
#pragma unroll
    for (int a = 0; a < kMaxC / 16; ++a)
#pragma unroll
      for (int b = 0; b < kMaxC / 16; ++b) acc[a][b] = 0.0f;
    bool any = false;
    for (int base = r0; base < r1; base += kThreads) {
      // compact the rows of this block of kThreads that have a neighbour through offset k
      const int i = base + threadIdx.x;
      const int src = i < r1 ? __ldg(&nbr[(size_t)k * nbr_stride + i]) : -1;
      const unsigned bal = __ballot_sync(0xFFFFFFFFu, src >= 0);
      if (lane == 0) warp_cnt[warp] = __popc(bal);
      __syncthreads();
      int before = 0, total = 0;
#pragma unroll
      for (int w = 0; w < kThreads / 32; ++w) {
        const int c = warp_cnt[w];
        if (w < warp) before += c;
        total += c;
      }
      if (src >= 0) {
        const int pos = before + __popc(bal & ((1u << lane) - 1u));
        list_in[pos] = src;
        list_out[pos] = i;
      }
      __syncthreads();
      any = any || total > 0;
      for (int g0 = 0; g0 < total; g0 += kStageRows) {
        const int rows = min(kStageRows, total - g0);
        // stage the pairs' rows: X[in] (cin floats) and dY[out] (cout floats)
        for (int e = threadIdx.x; e < rows * cin; e += kThreads) {
          const int r = e / cin, c = e - r * cin;
          xs[r][c] = __ldg(&x[(size_t)list_in[g0 + r] * cin + c]);
        }
        for (int e = threadIdx.x; e < rows * cout; e += kThreads) {
          const int r = e / cout, c = e - r * cout;
          ys[r][c] = __ldg(&dy[(size_t)list_out[g0 + r] * cout + c]);
        }
        __syncthreads();
        for (int r = 0; r < rows; ++r) {
          float xv[kMaxC / 16], yv[kMaxC / 16];
#pragma unroll
          for (int a = 0; a < kMaxC / 16; ++a) xv[a] = (ty + 16 * a < cin) ? xs[r][ty + 16 * a] : 0.0f;
#pragma unroll
          for (int b = 0; b < kMaxC / 16; ++b) yv[b] = (tx + 16 * b < cout) ? ys[r][tx + 16 * b] : 0.0f;
#pragma unroll
          for (int a = 0; a < kMaxC / 16; ++a)
#pragma unroll
            for (int b = 0; b < kMaxC / 16; ++b) acc[a][b] = fmaf(xv[a], yv[b], acc[a][b]);
        }
        __syncthreads();
      }
    }
    if (any) {
      float *dst = dw + (size_t)k * cin * cout;
#pragma unroll
      for (int a = 0; a < kMaxC / 16; ++a)
#pragma unroll
        for (int b = 0; b < kMaxC / 16; ++b) {
          const int ci = ty + 16 * a, co = tx + 16 * b;
          if (ci < cin && co < cout && acc[a][b] != 0.0f) atomicAdd(&dst[(size_t)ci * cout + co], acc[a][b]);
        }
    }
    __syncthreads();
  }
}

}  // namespace
}  // namespace fv2p

using namespace fv2p;

extern "C" int fv2p_conv_grad_filters(const float *features, const float *grad_out, const int32_t *nbr,
                                      int64_t nbr_stride, int kvol, int64_t n_out_cap, const int32_t *n_out_dev,
                                      int cin, int cout, float *grad_filters, fv2p_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FV2P_REQUIRE(kvol >= 1 && kvol <= FV2P_MAX_KVOL, "conv_grad_filters: kernel volume %d out of range", kvol);
  FV2P_REQUIRE(cin >= 1 && cin <= kMaxC && cout >= 1 && cout <= kMaxC,
               "conv_grad_filters: channel counts must be in [1,%d] (got %d->%d)", kMaxC, cin, cout);
  FV2P_REQUIRE(n_out_cap >= 0 && nbr_stride >= n_out_cap, "conv_grad_filters: bad sizes");
  FV2P_REQUIRE(grad_filters, "conv_grad_filters: null output");
  int st = cuda_status(cudaMemsetAsync(grad_filters, 0, sizeof(float) * (size_t)kvol * cin * cout, stream),
                       "conv_grad_filters");
  if (st) return st;
  if (n_out_cap == 0) return FV2P_OK;
  FV2P_REQUIRE(features && grad_out && nbr, "conv_grad_filters: null pointer argument");
  const int grid = persistent_grid(2);
  int slices = grid / kvol;
  if (slices < 1) slices = 1;
  const int64_t max_slices = (n_out_cap + kThreads - 1) / kThreads;  // no slice shorter than one block of rows
  if (slices > max_slices) slices = (int)(max_slices < 1 ? 1 : max_slices);
  grad_filters_kernel<<<grid, kThreads, 0, stream>>>(features, grad_out, nbr, nbr_stride, n_out_cap, n_out_dev, kvol,
                                                     cin, cout, slices, grad_filters);
  FV2P_LAUNCH_CHECK("conv_grad_filters");
  return FV2P_OK;
}
