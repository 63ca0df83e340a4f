// Training-mode BatchNorm1d over the rows of a sparse tensor's feature matrix (not on
// the inference hot path, where BatchNorm is folded into the conv epilogue).
//
// The reference's backbones wrap every conv in `nn.BatchNorm1d(eps=1e-3, momentum=0.01)` applied to `.features`
// (pcdet/models/backbones_3d/spconv_backbone.py:75, :193; SparseSequential hands `.features` to non-sparse
// modules, pcdet/ops/spconv/modules.py).  In training that is per-channel batch statistics over the N active rows:
//
//     mean[c] = sum_i x[i][c] / N          var[c] = sum_i (x[i][c] - mean[c])^2 / N        (biased, used to normalise)
//     y[i][c] = (x[i][c] - mean[c]) / sqrt(var[c] + eps) * weight[c] + bias[c]
//     running_mean = (1 - m) running_mean + m mean      running_var = (1 - m) running_var + m var N / (N - 1)
//
// and, backward,  dbias = sum dy,  dweight = sum dy xhat,  dx = weight invstd (dy - dbias / N - xhat dweight / N).
//
// Two launches each way, no host synchronisation, no library call: a column-sum kernel (thread = one column of a group
// of rows, fp64 partial sums, one atomicAdd(double) pair per column and CTA, the last CTA to finish turns the sums
// into the saved statistics) and an elementwise kernel.  The variance is taken about a per-column pivot (row 0) so
// that E[x^2] - E[x]^2 does not cancel.
#include "common.cuh"

namespace fv2p {
namespace {

constexpr int kBnMaxC = 1024;

// Sums over rows of a[i][c] - pivot_a[c] and of its square (kind 0: statistics of x) or of dy and dy * xhat (kind 1).
template <int kKind>
__global__ void __launch_bounds__(kThreads)
bn_reduce_kernel(const float *__restrict__ x, const float *__restrict__ dy, int64_t n, int c_total,
                 const float *__restrict__ mean, const float *__restrict__ invstd, double *sums, unsigned int *ticket,
                 // kind 0 epilogue
                 float eps, float momentum, float *running_mean, float *running_var, float *save_mean,
                 float *save_invstd,
                 // kind 1 epilogue
                 float *grad_weight, float *grad_bias) {
  __shared__ double red[2][kThreads];
  __shared__ bool last;
  const int rows_per_pass = c_total >= kThreads ? 1 : kThreads / c_total;
  for (int c_base = 0; c_base < c_total; c_base += kThreads) {
    const int c_here = min(c_total - c_base, kThreads);
    const int r = threadIdx.x / c_here, c = c_base + threadIdx.x % c_here;
    const bool active = r < rows_per_pass;
    double s0 = 0.0, s1 = 0.0;
    if (active) {
      const float pivot = kKind == 0 ? __ldg(&x[c]) : 0.0f;
      const float mu = kKind == 1 ? mean[c] : 0.0f, is = kKind == 1 ? invstd[c] : 0.0f;
      for (int64_t i = (int64_t)blockIdx.x * rows_per_pass + r; i < n; i += (int64_t)gridDim.x * rows_per_pass) {
        if (kKind == 0) {
          const float d = __ldg(&x[i * c_total + c]) - pivot;
          s0 += (double)d;
          s1 += (double)d * (double)d;
        } else {
          const float g = __ldg(&dy[i * c_total + c]);
          const float xh = (__ldg(&x[i * c_total + c]) - mu) * is;
          s0 += (double)g;
          s1 += (double)g * (double)xh;
        }
      }
    }
    red[0][threadIdx.x] = s0;
    red[1][threadIdx.x] = s1;
    __syncthreads();
    if (threadIdx.x < c_here) {  // r == 0: add the other row groups of this CTA, then one atomic pair per column
      for (int q = 1; q < rows_per_pass; ++q) {
        s0 += red[0][threadIdx.x + q * c_here];
        s1 += red[1][threadIdx.x + q * c_here];
      }
      atomicAdd(&sums[c], s0);
      atomicAdd(&sums[c_total + c], s1);
    }
    __syncthreads();
  }
  __threadfence();
  if (threadIdx.x == 0) last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!last) return;
  __threadfence();
  for (int c = threadIdx.x; c < c_total; c += kThreads) {
    const double s0 = __ldcg(&sums[c]), s1 = __ldcg(&sums[c_total + c]);
    if (kKind == 0) {
      const double dn = (double)n;
      const double m_rel = s0 / dn;                         // mean - pivot
      double var = s1 / dn - m_rel * m_rel;                 // biased
      if (var < 0.0) var = 0.0;
      const double mu = (double)__ldg(&x[c]) + m_rel;
      save_mean[c] = (float)mu;
      save_invstd[c] = (float)(1.0 / sqrt(var + (double)eps));
      if (running_mean) running_mean[c] = (float)((1.0 - (double)momentum) * (double)running_mean[c] + (double)momentum * mu);
      if (running_var) {
        const double unbiased = n > 1 ? var * dn / (dn - 1.0) : var;
        running_var[c] = (float)((1.0 - (double)momentum) * (double)running_var[c] + (double)momentum * unbiased);
      }
    } else {
      grad_bias[c] = (float)s0;
      grad_weight[c] = (float)s1;
    }
    sums[c] = 0.0;  // the workspace is handed back zeroed
    sums[c_total + c] = 0.0;
  }
  if (threadIdx.x == 0) *ticket = 0u;
}

__global__ void __launch_bounds__(kThreads)
bn_apply_kernel(const float *__restrict__ x, int64_t total, int c_total, const float *__restrict__ mean,
                const float *__restrict__ invstd, const float *__restrict__ weight, const float *__restrict__ bias,
                int relu, float *__restrict__ y) {
  for (int64_t e = (int64_t)blockIdx.x * kThreads + threadIdx.x; e < total; e += (int64_t)gridDim.x * kThreads) {
    const int c = (int)(e % c_total);
    const float w = weight ? __ldg(&weight[c]) : 1.0f, b = bias ? __ldg(&bias[c]) : 0.0f;
    float v = (x[e] - mean[c]) * invstd[c] * w + b;
    if (relu) v = fmaxf(v, 0.0f);
    y[e] = v;
  }
}

__global__ void __launch_bounds__(kThreads)
bn_grad_input_kernel(const float *__restrict__ x, const float *__restrict__ dy, int64_t total, int c_total, double inv_n,
                     const float *__restrict__ mean, const float *__restrict__ invstd,
                     const float *__restrict__ weight, const float *__restrict__ grad_weight,
                     const float *__restrict__ grad_bias, float *__restrict__ dx) {
  for (int64_t e = (int64_t)blockIdx.x * kThreads + threadIdx.x; e < total; e += (int64_t)gridDim.x * kThreads) {
    const int c = (int)(e % c_total);
    const float w = weight ? __ldg(&weight[c]) : 1.0f;
    const float is = invstd[c];
    const float xh = (x[e] - mean[c]) * is;
    const float t = dy[e] - (float)((double)grad_bias[c] * inv_n) - xh * (float)((double)grad_weight[c] * inv_n);
    dx[e] = w * is * t;
  }
}

int reduce_grid(int64_t n, int c_total) {
  const int rows_per_pass = c_total >= kThreads ? 1 : kThreads / c_total;
  const int64_t want = (n + (int64_t)rows_per_pass * 8 - 1) / ((int64_t)rows_per_pass * 8);  // >= 8 rows per thread
  const int cap = persistent_grid(4);
  return (int)(want < 1 ? 1 : (want > cap ? cap : want));
}

}  // namespace
}  // namespace fv2p

using namespace fv2p;

extern "C" size_t fv2p_batchnorm_workspace_bytes(int channels) {
  return sizeof(double) * 2 * (size_t)(channels > 0 ? channels : 0) + 16;
}

extern "C" int fv2p_batchnorm_train_fwd(const float *x, int64_t n, int channels, const float *weight, const float *bias,
                                        float eps, float momentum, float *running_mean, float *running_var, int relu,
                                        float *y, float *save_mean, float *save_invstd, void *workspace,
                                        size_t workspace_bytes, fv2p_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FV2P_REQUIRE(channels >= 1 && channels <= kBnMaxC, "batchnorm_train_fwd: channels must be in [1,%d] (got %d)", kBnMaxC,
               channels);
  // torch raises "Expected more than 1 value per channel when training" for n == 1 (torch/nn/functional.py)
  FV2P_REQUIRE(n >= 2, "batchnorm_train_fwd: expected more than 1 value per channel when training (got %lld rows)",
               (long long)n);
  FV2P_REQUIRE(x && y && save_mean && save_invstd && workspace, "batchnorm_train_fwd: null pointer argument");
  FV2P_REQUIRE(workspace_bytes >= fv2p_batchnorm_workspace_bytes(channels), "batchnorm_train_fwd: workspace too small");
  FV2P_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 7) == 0, "batchnorm_train_fwd: workspace must be 8-byte aligned");
  double *sums = static_cast<double *>(workspace);
  unsigned int *ticket = reinterpret_cast<unsigned int *>(sums + 2 * channels);
  bn_reduce_kernel<0><<<reduce_grid(n, channels), kThreads, 0, stream>>>(
      x, nullptr, n, channels, nullptr, nullptr, sums, ticket, eps, momentum, running_mean, running_var, save_mean,
      save_invstd, nullptr, nullptr);
  FV2P_LAUNCH_CHECK("batchnorm_train_fwd(statistics)");
  const int64_t total = n * channels;
  const int64_t blocks = (total + kThreads * 4 - 1) / (kThreads * 4);
  const int cap = persistent_grid(8);
  bn_apply_kernel<<<(int)(blocks > cap ? cap : blocks), kThreads, 0, stream>>>(x, total, channels, save_mean, save_invstd,
                                                                              weight, bias, relu, y);
  FV2P_LAUNCH_CHECK("batchnorm_train_fwd(apply)");
  return FV2P_OK;
}

extern "C" int fv2p_batchnorm_train_bwd(const float *x, const float *grad_out, int64_t n, int channels,
                                        const float *weight, const float *save_mean, const float *save_invstd,
                                        float *grad_input, float *grad_weight, float *grad_bias, void *workspace,
                                        size_t workspace_bytes, fv2p_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FV2P_REQUIRE(channels >= 1 && channels <= kBnMaxC, "batchnorm_train_bwd: channels must be in [1,%d] (got %d)", kBnMaxC,
               channels);
  FV2P_REQUIRE(n >= 1, "batchnorm_train_bwd: no rows");
  FV2P_REQUIRE(x && grad_out && save_mean && save_invstd && grad_input && grad_weight && grad_bias && workspace,
               "batchnorm_train_bwd: null pointer argument");
  FV2P_REQUIRE(workspace_bytes >= fv2p_batchnorm_workspace_bytes(channels), "batchnorm_train_bwd: workspace too small");
  FV2P_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 7) == 0, "batchnorm_train_bwd: workspace must be 8-byte aligned");
  double *sums = static_cast<double *>(workspace);
  unsigned int *ticket = reinterpret_cast<unsigned int *>(sums + 2 * channels);
  bn_reduce_kernel<1><<<reduce_grid(n, channels), kThreads, 0, stream>>>(
      x, grad_out, n, channels, save_mean, save_invstd, sums, ticket, 0.0f, 0.0f, nullptr, nullptr, nullptr, nullptr,
      grad_weight, grad_bias);
  FV2P_LAUNCH_CHECK("batchnorm_train_bwd(sums)");
  const int64_t total = n * channels;
  const int64_t blocks = (total + kThreads * 4 - 1) / (kThreads * 4);
  const int cap = persistent_grid(8);
  bn_grad_input_kernel<<<(int)(blocks > cap ? cap : blocks), kThreads, 0, stream>>>(
      x, grad_out, total, channels, 1.0 / (double)n, save_mean, save_invstd, weight, grad_weight, grad_bias, grad_input);
  FV2P_LAUNCH_CHECK("batchnorm_train_bwd(grad_input)");
  return FV2P_OK;
}
