// Shared device/host helpers for libfv2p_b200 (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <limits.h>
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

int geo_ctas_override();  // 0 = none (debug switch fv2p_debug_geo_ctas: CTAs per SM of the persistent geometry grids)
// Persistent grids of the geometry pass and the voxelizer: 8 CTAs per SM.  These kernels run next to a persistent conv
// CTA that leaves room for about one of theirs per SM, so the grid size mostly decides how much of the freed
// resources they pick up between conv layers: measured step time at 1 / 2 / 4 / 8 CTAs per SM on waymo_b4 fp32
// 3.08 / 2.75 / 2.66 / 2.65 ms, kitti_b8 1.129 / 1.075 / 1.069 / 1.051 (profiles/contention.py).
constexpr int kGeoCtasPerSm = 8;
inline int persistent_grid(int ctas_per_sm = 4) {
  const int o = geo_ctas_override();
  return sm_count() * (o > 0 ? o : ctas_per_sm);
}

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

// ------------------------------------------------------------------------- coordinate tables (v2)
// One 16-byte slot per key: the probe of a present key costs ONE 16-byte load (key and value arrive together; the
// first generation kept keys and values in two arrays = two dependent L2 round trips per hit).  `val` holds, in
// order: kValEmpty after the clear; for the output table of a strided rulebook the smallest bid (>= 0) while the
// candidates are being inserted; finally ~row (< 0) of the voxel stored under the key.  Tables are sized by the
// caller-stated row CAPACITY (host-known), so every kernel of a step agrees on the mask and the clear can run
// before the live row count exists.
struct __align__(16) Slot {
  unsigned long long key;
  int val;
  int aux;
};
constexpr int kValEmpty = INT_MAX;

__host__ __device__ inline uint32_t table_slots_cap(int64_t row_cap) { return table_slots_for(row_cap); }

// Returns the slot index of `key`, inserting it if absent; 0xFFFFFFFF when the table is full.
__device__ __forceinline__ uint32_t slot_insert(Slot *t, uint32_t mask, unsigned long long key) {
  uint32_t s = mix64(key) & mask;
  for (uint32_t probes = 0; probes <= mask; ++probes) {
    unsigned long long seen = *reinterpret_cast<volatile unsigned long long *>(&t[s].key);
    if (seen == key) return s;
    if (seen == kEmptyKey) {
      unsigned long long prev = atomicCAS(&t[s].key, kEmptyKey, key);
      if (prev == kEmptyKey || prev == key) return s;
    }
    s = (s + 1) & mask;
  }
  return 0xFFFFFFFFu;
}

// Value stored under `key` (read-only table, one 16-byte load per probe), or kValEmpty when the key is absent.
// The walk is bounded by the table size: a table that overflowed its caller-stated capacity may have no empty slot
// left, and an absent key must still terminate (the step is then flagged and run again with larger bounds).
__device__ __forceinline__ int slot_lookup(const Slot *__restrict__ t, uint32_t mask, unsigned long long key) {
  uint32_t s = mix64(key) & mask;
  for (uint32_t probes = 0; probes <= mask; ++probes) {
    const uint4 q = __ldg(reinterpret_cast<const uint4 *>(&t[s]));
    const unsigned long long seen = ((unsigned long long)q.y << 32) | q.x;
    if (seen == key) return (int)q.z;
    if (seen == kEmptyKey) return kValEmpty;
    s = (s + 1) & mask;
  }
  return kValEmpty;
}

// ------------------------------------------------------------------------- single-pass ordered scan
// Decoupled look-back over per-chunk totals (Merrill & Garland): chunk c publishes its total, then walks back over
// its predecessors' words until it meets one that already carries an inclusive prefix.  One 64-bit word per chunk:
// bits 62-63 = state (0 not ready, 1 total only, 2 inclusive prefix), low 32 bits = value; the words must be zero
// when the kernel starts.  Chunks are handed out through an atomic ticket, so every predecessor of a chunk has
// already been taken by a CTA that never waits on a later chunk: forward progress does not depend on the whole grid
// being resident (the geometry kernels share the SMs with persistent conv kernels).  Call with the whole CTA;
// returns the exclusive prefix of chunk `c` to every thread.  `smem_word` = one int of shared memory.
__device__ __forceinline__ int lookback_exclusive(unsigned long long *state, int c, int total, int *smem_word) {
  constexpr unsigned long long kAgg = 1ull << 62, kPre = 2ull << 62;
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    volatile unsigned long long *vs = state;
    int exclusive = 0;
    if (c == 0) {
      if (lane == 0) vs[0] = kPre | (unsigned int)total;
    } else {
      if (lane == 0) vs[c] = kAgg | (unsigned int)total;
      int idx = c - 1;
      while (true) {
        const int mine = idx - lane;
        unsigned long long w;
        do {
          w = mine >= 0 ? vs[mine] : kPre;  // the virtual chunk before chunk 0 has prefix 0
        } while (__any_sync(0xFFFFFFFFu, (w >> 62) == 0));
        const unsigned has_pre = __ballot_sync(0xFFFFFFFFu, (w >> 62) == 2);
        const int first = has_pre ? __ffs(has_pre) - 1 : 31;
        int v = lane <= first ? (int)(unsigned int)(w & 0xFFFFFFFFu) : 0;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, d);
        exclusive += v;
        if (has_pre) break;
        idx -= 32;
      }
      if (lane == 0) vs[c] = kPre | (unsigned int)(exclusive + total);
    }
    if (lane == 0) *smem_word = exclusive;
  }
  __syncthreads();
  const int r = *smem_word;
  __syncthreads();
  return r;
}

// Ranges cleared by ONE launch at the start of a geometry pass (tables, scan states, -1 fills).
struct FillRange {
  void *ptr;
  unsigned long long n16;  // 16-byte units
  uint4 pattern;
};
constexpr int kMaxFillRanges = 40;
struct FillJob {
  int count;
  FillRange r[kMaxFillRanges];
};
int launch_fill(const FillJob &job, cudaStream_t stream);
// region of the row-grouping workspace (sort.cu) that must be zero when fv2p_group_rows starts
void group_rows_zero_region(int64_t n_cap, size_t *off, size_t *bytes);
inline void add_fill(FillJob &job, void *ptr, size_t bytes, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  if (!ptr || bytes == 0 || job.count >= kMaxFillRanges) return;
  FillRange &r = job.r[job.count++];
  r.ptr = ptr;
  r.n16 = (bytes + 15) / 16;
  r.pattern = make_uint4(a, b, c, d);
}
inline void add_fill_table(FillJob &job, void *table, int64_t row_cap) {
  add_fill(job, table, (size_t)table_slots_cap(row_cap) * sizeof(Slot), 0xFFFFFFFFu, 0xFFFFFFFFu, (uint32_t)kValEmpty, 0u);
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
