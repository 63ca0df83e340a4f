// Error bookkeeping, device check and the small shared kernels of libfv2p_b200.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace fv2p {

static thread_local char g_error[512] = "";

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

int cuda_status(cudaError_t e, const char *what) {
  if (e == cudaSuccess) return 0;
  // mirrors TV_CHECK_CUDA_ERR (include/tensorview/tensorview.h:93-102): "cuda execution failed with error N"
  set_error("%s: cuda execution failed with error %d (%s)", what, (int)e, cudaGetErrorString(e));
  return (int)e;
}

static int g_geo_ctas = 0;
int geo_ctas_override() { return g_geo_ctas; }

constexpr int kMaxDevices = 64;

int current_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) dev = 0;
  return dev;
}

// Per device ordinal: one process may drive several GPUs.
int sm_count() {
  static int cached[kMaxDevices] = {0};
  const int dev = current_device();
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      cached[dev] = n;
    else
      cached[dev] = 148;
  }
  return cached[dev];
}

// Work that is off the caller's dependency chain moves to a second stream: `to` waits for what `from` has enqueued
// so far.  Returns the stream to continue on (`from` itself when there is no second stream).
cudaStream_t fork_stream(cudaStream_t from, void *to_) {
  cudaStream_t to = static_cast<cudaStream_t>(to_);
  if (!to_ || to == from) return from;
  cudaEvent_t ev;
  if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) return from;
  cudaEventRecord(ev, from);
  cudaStreamWaitEvent(to, ev, 0);
  cudaEventDestroy(ev);
  return to;
}

// One launch clears every range of the job (grid-stride over the concatenation of the ranges).
__global__ void __launch_bounds__(kThreads) fill_ranges_kernel(const FillJob job) {
  for (int r = 0; r < job.count; ++r) {
    uint4 *dst = static_cast<uint4 *>(job.r[r].ptr);
    const unsigned long long n = job.r[r].n16;
    const uint4 pat = job.r[r].pattern;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * blockDim.x)
      dst[i] = pat;
  }
}

int launch_fill(const FillJob &job, cudaStream_t stream) {
  if (job.count == 0) return 0;
  fill_ranges_kernel<<<persistent_grid(8), kThreads, 0, stream>>>(job);
  return cuda_status(cudaGetLastError(), "fill");
}

__global__ void set_scalar_kernel(int *dst, int value) { *dst = value; }

void launch_set_scalar(int *dst, int value, cudaStream_t stream) {
  set_scalar_kernel<<<1, 1, 0, stream>>>(dst, value);
}

// One CTA per segment: in-place exclusive scan of counts[seg][0..n_chunks), totals[seg] = sum.
__global__ void __launch_bounds__(kThreads) scan_chunk_counts_kernel(int *counts, int64_t seg_stride,
                                                                     int n_chunks, const int *n_dev, int *totals) {
  __shared__ int smem[kThreads / 32 + 1];
  if (n_dev) {
    int n = *n_dev;
    n = n < 0 ? 0 : n;
    const int live = live_chunks(n);
    n_chunks = live < n_chunks ? live : n_chunks;
  }
  int *row = counts + (size_t)blockIdx.x * seg_stride;
  int running = 0;
  for (int base = 0; base < n_chunks; base += kThreads) {
    int i = base + threadIdx.x;
    int v = i < n_chunks ? row[i] : 0;
    int total;
    int ex = block_exclusive_scan(v, smem, total);
    if (i < n_chunks) row[i] = running + ex;
    running += total;
  }
  if (threadIdx.x == 0 && totals) totals[blockIdx.x] = running;
}

void launch_scan_chunk_counts(int *counts, int segments, int64_t seg_stride, const int *n_dev, int64_t n_cap,
                              int *totals, cudaStream_t stream) {
  int n_chunks = (int)((n_cap + kChunk - 1) / kChunk);
  if (n_chunks < 1) n_chunks = 1;
  scan_chunk_counts_kernel<<<segments, kThreads, 0, stream>>>(counts, seg_stride, n_chunks, n_dev, totals);
}

}  // namespace fv2p

extern "C" {

int fv2p_abi_version(void) { return FV2P_ABI_VERSION; }

__attribute__((visibility("default"))) int fv2p_debug_geo_ctas(int v) {
  fv2p::g_geo_ctas = v > 0 ? v : 0;
  return 0;
}

const char *fv2p_last_error(void) { return fv2p::g_error; }

int fv2p_device_check(int *sm_count_out, int *cc_major, int *cc_minor) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    fv2p::set_error("no CUDA device: %s", cudaGetErrorString(e));
    return FV2P_ERR_DEVICE;
  }
  int major = 0, minor = 0, sms = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (sm_count_out) *sm_count_out = sms;
  if (cc_major) *cc_major = major;
  if (cc_minor) *cc_minor = minor;
  if (major != 10) {
    fv2p::set_error("libfv2p_b200 is built for sm_100a only; device %d is sm_%d%d", dev, major, minor);
    return FV2P_ERR_DEVICE;
  }
  return FV2P_OK;
}

}  // extern "C"
