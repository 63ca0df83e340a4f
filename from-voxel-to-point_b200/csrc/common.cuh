// Shared device/host helpers for libfv2p_b200 (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/fv2p_b200.h"

namespace fv2p {

// ---------------------------------------------------------------------------------------- errors
void set_error(const char *fmt, ...);
int cuda_status(cudaError_t e, const char *what);  // 0 or the cudaError_t, records the message
int sm_count();                                    // of the current device (cached per device ordinal)
int current_device();                              // ordinal, clamped to [0, 64)

#define FV2P_REQUIRE(cond, ...)       \
  do {                                \
    if (!(cond)) {                    \
      fv2p::set_error(__VA_ARGS__);   \
      return FV2P_ERR_INVALID;        \
    }                                 \
  } while (0)

#define FV2P_LAUNCH_CHECK(what)                                   \
  do {                                                            \
    int _st = fv2p::cuda_status(cudaGetLastError(), what);        \
    if (_st) return _st;                                          \
  } while (0)

// ---------------------------------------------------------------------------------- launch shape
// Every kernel is persistent/grid-stride over a row count that lives in device memory, so launch
// dimensions never depend on data: a few CTAs per SM, 148 SMs on B200.
constexpr int kThreads = 256;
constexpr int kItemsPerThread = 2;
constexpr int kChunk = kThreads * kItemsPerThread;  // rows per work item of the ordered passes

inline int persistent_grid(int ctas_per_sm = 4) { return sm_count() * ctas_per_sm; }

// workspace carving (256-byte aligned)
struct Carver {
  char *base;
  size_t used;
  explicit Carver(void *p) : base(static_cast<char *>(p)), used(0) {}
  template <typename T>
  T *take(size_t count) {
    size_t off = (used + 255) & ~size_t(255);
    used = off + count * sizeof(T);
    return base ? reinterpret_cast<T *>(base + off) : nullptr;
  }
};

// -------------------------------------------------------------------------------------- hashing
constexpr unsigned long long kEmptyKey = ~0ull;

__host__ __device__ inline uint32_t table_slots_for(int64_t n) {
  // power of two >= 2n, at least 1024
  uint64_t want = (uint64_t)(n < 512 ? 512 : n) * 2;
  uint64_t s = 1024;
  while (s < want) s <<= 1;
  return (uint32_t)s;
}

__device__ __forceinline__ uint32_t mix64(unsigned long long k) {
  k ^= k >> 33;
  k *= 0xff51afd7ed558ccdULL;
  k ^= k >> 33;
  k *= 0xc4ceb9fe1a85ec53ULL;
  k ^= k >> 33;
  return (uint32_t)k;
}

// Open addressing, linear probing.  Returns the slot holding `key` (inserting it if absent).
__device__ __forceinline__ uint32_t table_insert(unsigned long long *keys, uint32_t mask,
                                                 unsigned long long key) {
  uint32_t slot = mix64(key) & mask;
  while (true) {
    unsigned long long seen = keys[slot];
    if (seen == key) return slot;
    if (seen == kEmptyKey) {
      unsigned long long prev = atomicCAS(&keys[slot], kEmptyKey, key);
      if (prev == kEmptyKey || prev == key) return slot;
    }
    slot = (slot + 1) & mask;
  }
}

// Same for a table that may be too small for the keys offered (its size follows a caller-chosen capacity): gives up
// with 0xFFFFFFFF once every slot has been looked at instead of probing a full table forever.
__device__ __forceinline__ uint32_t table_insert_bounded(unsigned long long *keys, uint32_t mask,
                                                         unsigned long long key) {
  uint32_t slot = mix64(key) & mask;
  for (uint32_t probes = 0; probes <= mask; ++probes) {
    unsigned long long seen = keys[slot];
    if (seen == key) return slot;
    if (seen == kEmptyKey) {
      unsigned long long prev = atomicCAS(&keys[slot], kEmptyKey, key);
      if (prev == kEmptyKey || prev == key) return slot;
    }
    slot = (slot + 1) & mask;
  }
  return 0xFFFFFFFFu;
}

// Returns the slot of `key` or 0xFFFFFFFF.
__device__ __forceinline__ uint32_t table_find(const unsigned long long *__restrict__ keys,
                                               uint32_t mask, unsigned long long key) {
  uint32_t slot = mix64(key) & mask;
  while (true) {
    unsigned long long seen = __ldg(&keys[slot]);
    if (seen == key) return slot;
    if (seen == kEmptyKey) return 0xFFFFFFFFu;
    slot = (slot + 1) & mask;
  }
}

__device__ __forceinline__ unsigned long long voxel_key(int b, int z, int y, int x, int D, int H,
                                                        int W) {
  return (((unsigned long long)b * D + z) * H + y) * (unsigned long long)W + x;
}

// ------------------------------------------------------------------------------ block-level scan
// Exclusive scan of one value per thread over a kThreads block; returns the block total through
// `total`.  `smem` must hold kThreads/32 + 1 ints.  Contains two __syncthreads().
__device__ __forceinline__ int block_exclusive_scan(int v, int *smem, int &total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int up = __shfl_up_sync(0xFFFFFFFFu, incl, d);
    if (lane >= d) incl += up;
  }
  if (lane == 31) smem[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int w = lane < (kThreads / 32) ? smem[lane] : 0;
    int wi = w;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      int up = __shfl_up_sync(0xFFFFFFFFu, wi, d);
      if (lane >= d) wi += up;
    }
    if (lane < (kThreads / 32)) smem[lane] = wi - w;  // exclusive warp offsets
    if (lane == 31) smem[kThreads / 32] = wi;         // block total
  }
  __syncthreads();
  total = smem[kThreads / 32];
  int r = smem[warp] + incl - v;
  __syncthreads();  // smem may be reused by the next call
  return r;
}

// Exclusive scan of one 0/1 flag per thread (ballot + popc, two barriers); block total through `total`.
__device__ __forceinline__ int block_flag_scan(bool flag, int *smem, int &total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned bal = __ballot_sync(0xFFFFFFFFu, flag);
  if (lane == 0) smem[warp] = __popc(bal);
  __syncthreads();
  int base = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < kThreads / 32; ++w) {
    const int c = smem[w];
    if (w < warp) base += c;
    tot += c;
  }
  total = tot;
  __syncthreads();  // smem may be reused by the next call
  return base + __popc(bal & ((1u << lane) - 1u));
}

// Number of kChunk-row work items that hold live rows (device-side count).
__device__ __forceinline__ int live_chunks(int n) { return (n + kChunk - 1) / kChunk; }

// counts[segment][chunk] -> exclusive prefix in place, totals[segment] = sum.  One CTA per segment.  With n_dev
// the scan covers only the chunks that hold live rows (capacity-sized tails are never written nor read).
void launch_scan_chunk_counts(int *counts, int segments, int64_t seg_stride, const int *n_dev,
                              int64_t n_cap, int *totals, cudaStream_t stream);

__global__ void set_scalar_kernel(int *dst, int value);
cudaStream_t fork_stream(cudaStream_t from, void *to);
void launch_set_scalar(int *dst, int value, cudaStream_t stream);

}  // namespace fv2p
