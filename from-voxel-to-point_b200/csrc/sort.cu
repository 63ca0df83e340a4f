// Row grouping for the tensor-core conv: output rows with similar neighbour masks share a tile.
//
// The fused conv processes 128 output rows per tile and one kernel offset per pipeline stage; an offset is skipped
// only if NO row of the tile has that neighbour.  In point-cloud geometry a row has 4-13 of the 27 neighbours, so
// tiles cut from the rulebook's native row order need almost all 27 stages and 60-75 % of the gathered rows are
// zero fill.  The reference has no counterpart: its gather-GEMM-scatter (spconv_ops.h:308-357) works on compacted
// per-offset pair lists instead.
//
// Round 1 sorted the rows by their full 27-bit mask (stable LSD radix sort, 3 passes x 3 launches + 4 more).  Round
// 2 groups them by a 12-bit DIGEST of the mask in ONE counting pass (2 launches):
//
//     digest = [ which x-offsets occur anywhere | which (z,y) lines of the kernel have any neighbour ]
//
// (kx + kz*ky bits: 3 + 9 for 3x3x3, 1 + 3 for (3,1,1)).  Rows of a tile then agree on the lines they touch, which
// is what decides the tile's active offsets; on the synthetic KITTI / Waymo frames this needs FEWER stages per tile
// than the exact sort (e.g. 9.4 vs 10.7, 6.8 vs 8.2, 12.2 vs 15.2, 10.6 vs 14.7 at the first four rulebooks of a
// KITTI frame; unsorted: 18.8 / 23.1 / 25.6 / 26.4), because the exact sort splits rows that differ only in
// low-order bits across distant tiles.
//
// Results do not change: every output row still accumulates its offsets in ascending k, only the assignment of rows
// to tiles moves - so the position of a row INSIDE its group may come from an atomic cursor (no ordered scan, no
// stability needed).  Outputs: perm[t] = output row processed at position t, the neighbour map permuted the same way
// (nbr_sorted[k][t] = nbr[k][perm[t]]), and the tile list (tile, OR of its rows' masks) by descending number of
// active offsets - the conv's longest-processing-time-first schedule.
#include "common.cuh"

namespace fv2p {
namespace {

constexpr int kDigestBits = 12;
constexpr int kBins = 1 << kDigestBits;
constexpr int kConvTile = 128;

__device__ __forceinline__ int live_n(const int *n_dev, int64_t n_cap) {
  int n = n_dev ? *n_dev : (int)n_cap;
  return n < 0 ? 0 : (n > n_cap ? (int)n_cap : n);
}

// kx = extent of the fastest kernel axis, lines = kvol / kx.  kvol <= 12: the mask itself.
__device__ __forceinline__ uint32_t mask_digest(uint32_t m, int kvol, int kx, int lines) {
  if (kvol <= kDigestBits) return m;
  uint32_t line_bits = 0, x_bits = 0;
  const uint32_t xm = (1u << kx) - 1u;
  for (int l = 0; l < lines; ++l) {
    const uint32_t seg = (m >> (l * kx)) & xm;
    x_bits |= seg;
    line_bits |= (seg != 0u ? 1u : 0u) << l;
  }
  uint32_t d = (x_bits << lines) | line_bits;
  if (kx + lines > kDigestBits) d = (d ^ (d >> kDigestBits)) & (kBins - 1);  // exotic kernels: fold
  return d;
}

struct GroupWs {
  uint32_t *masks;
  uint16_t *digests;
  int *bin_base;             // [kBins] exclusive scan of the bins (written by the last CTA of the histogram)
  int *bins, *cursor;        // [kBins] each, zero on entry
  uint32_t *tile_masks;      // [tiles], zero on entry
  int *done;                 // [4], zero on entry: CTAs finished per kernel
  size_t zero_off, zero_bytes, bytes;
};

GroupWs carve(void *ws, int64_t n_cap) {
  GroupWs w;
  Carver c(ws);
  const size_t n = (size_t)(n_cap > 0 ? n_cap : 1);
  w.masks = c.take<uint32_t>(n);
  w.digests = c.take<uint16_t>(n);
  w.bin_base = c.take<int>(kBins);
  w.bins = c.take<int>(kBins);
  w.zero_off = (size_t)(reinterpret_cast<char *>(w.bins) - static_cast<char *>(ws));
  w.cursor = c.take<int>(kBins);
  w.tile_masks = c.take<uint32_t>(n / kConvTile + 2);
  w.done = c.take<int>(64);
  w.zero_bytes = c.used - w.zero_off;
  w.bytes = c.used + 256;
  return w;
}

// masks, digests and the digest histogram; the last CTA to finish turns the bins into their exclusive scan.  The
// histogram goes straight to the global bins, one atomic per distinct digest per warp: no shared-memory histogram, so
// the kernel fits next to a persistent conv CTA that has left it only a few KB of shared memory.
__global__ void __launch_bounds__(kThreads)
group_hist_kernel(const int *__restrict__ nbr, int64_t nbr_stride, int kvol, int kx, int lines, const int *n_dev,
                  int64_t n_cap, uint32_t *masks, uint16_t *digests, int *bins, int *bin_base, int *done) {
  __shared__ int scan_smem[kThreads / 32 + 1];
  __shared__ int s_last;
  const int n = live_n(n_dev, n_cap);
  const int lane = threadIdx.x & 31;
  const int n_round = (n + 31) & ~31;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += gridDim.x * blockDim.x) {
    const bool live = i < n;
    uint32_t m = 0;
    if (live) {
      // nine of the row's loads in flight at a time: the kernel is a pure stream over the map (more would cost the
      // registers that let it sit next to a conv CTA)
      for (int k0 = 0; k0 < kvol; k0 += 9) {
        int v[9];
#pragma unroll
        for (int q = 0; q < 9; ++q) v[q] = k0 + q < kvol ? __ldg(&nbr[(size_t)(k0 + q) * nbr_stride + i]) : -1;
#pragma unroll
        for (int q = 0; q < 9; ++q) m |= (v[q] >= 0 ? 1u : 0u) << (k0 + q);
      }
    }
    const int d = live ? (int)mask_digest(m, kvol, kx, lines) : kBins + lane;
    const unsigned peers = __match_any_sync(0xFFFFFFFFu, d);
    if (live) {
      masks[i] = m;
      digests[i] = (uint16_t)d;
      if (lane == __ffs(peers) - 1) atomicAdd(&bins[d], __popc(peers));
    }
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(&done[0], 1) == (int)gridDim.x - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  constexpr int kPer = kBins / kThreads;  // consecutive bins per thread
  int local[kPer];
  int sum = 0;
#pragma unroll
  for (int q = 0; q < kPer; ++q) {
    local[q] = *reinterpret_cast<volatile int *>(&bins[threadIdx.x * kPer + q]);
    sum += local[q];
  }
  int total;
  int run = block_exclusive_scan(sum, scan_smem, total);
#pragma unroll
  for (int q = 0; q < kPer; ++q) {
    bin_base[threadIdx.x * kPer + q] = run;
    run += local[q];
  }
}

// Scatter: position = start of the row's bin + a ticket from the bin's cursor (one atomic per distinct digest per
// warp).  Writes perm, the permuted neighbour map and ORs the row's mask into its tile's mask; the last CTA to
// finish ranks the tiles.
__global__ void __launch_bounds__(kThreads)
group_scatter_kernel(const int *__restrict__ nbr, int64_t nbr_stride, int kvol, const int *n_dev, int64_t n_cap,
                     const uint32_t *__restrict__ masks, const uint16_t *__restrict__ digests,
                     const int *__restrict__ bin_base, int *cursor, int *perm, int *nbr_sorted,
                     int64_t sorted_stride, uint32_t *tile_masks, int *done, int2 *tile_order) {
  __shared__ int s_last;
  __shared__ int hist[33], base[33];
  __shared__ int warp_cnt[kThreads / 32][33];
  const int n = live_n(n_dev, n_cap);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n_round = (n + 31) & ~31;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += gridDim.x * blockDim.x) {
    const bool live = i < n;
    const int d = live ? (int)digests[i] : kBins + lane;
    const unsigned peers = __match_any_sync(0xFFFFFFFFu, d);
    const int leader = __ffs(peers) - 1;
    int start = 0;
    if (live && lane == leader) start = __ldg(&bin_base[d]) + atomicAdd(&cursor[d], __popc(peers));
    start = __shfl_sync(0xFFFFFFFFu, start, leader);
    const int pos = start + __popc(peers & ((1u << lane) - 1u));
    const uint32_t m = live ? masks[i] : 0u;
    const int tile = live ? pos / kConvTile : -1 - lane;
    const unsigned tpeers = __match_any_sync(0xFFFFFFFFu, tile);
    const uint32_t tm = __reduce_or_sync(tpeers, m);
    if (live) {
      perm[pos] = i;
      if (nbr_sorted) {
        for (int k0 = 0; k0 < kvol; k0 += 9) {
          int v[9];
#pragma unroll
          for (int q = 0; q < 9; ++q) v[q] = k0 + q < kvol ? __ldg(&nbr[(size_t)(k0 + q) * nbr_stride + i]) : -1;
#pragma unroll
          for (int q = 0; q < 9; ++q)
            if (k0 + q < kvol) nbr_sorted[(size_t)(k0 + q) * sorted_stride + pos] = v[q];
        }
      }
      if (tile_order && lane == __ffs(tpeers) - 1) atomicOr(&tile_masks[tile], tm);
    }
  }
  if (!tile_order) return;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(&done[1], 1) == (int)gridDim.x - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // tile_order = (tile, mask) pairs by descending weight = popcount(mask) (counting sort over the 33 possible
  // weights), ties in tile order (per-warp ballots under a block-ordered cursor).
  const volatile uint32_t *tmv = tile_masks;
  const int n_tiles = (n + kConvTile - 1) / kConvTile;
  if (threadIdx.x < 33) hist[threadIdx.x] = 0;
  __syncthreads();
  for (int t = threadIdx.x; t < n_tiles; t += blockDim.x) atomicAdd(&hist[__popc(tmv[t])], 1);
  __syncthreads();
  if (threadIdx.x == 0) {
    int run = 0;
    for (int w = 32; w >= 0; --w) {
      base[w] = run;
      run += hist[w];
    }
  }
  __syncthreads();
  for (int t0 = 0; t0 < n_tiles; t0 += blockDim.x) {
    const int t = t0 + threadIdx.x;
    const uint32_t m = t < n_tiles ? tmv[t] : 0u;
    const int w = t < n_tiles ? __popc(m) : -1;
    const unsigned peers = __match_any_sync(0xFFFFFFFFu, w);
    for (int b = lane; b < 33; b += 32) warp_cnt[warp][b] = 0;
    __syncwarp();
    if (w >= 0 && lane == __ffs(peers) - 1) warp_cnt[warp][w] = __popc(peers);
    __syncthreads();
    if (w >= 0) {
      int before = 0;
      for (int q = 0; q < warp; ++q) before += warp_cnt[q][w];
      tile_order[base[w] + before + __popc(peers & ((1u << lane) - 1u))] = make_int2(t, (int)m);
    }
    __syncthreads();
    if (threadIdx.x < 33) {
      int add = 0;
      for (int q = 0; q < kThreads / 32; ++q) add += warp_cnt[q][threadIdx.x];
      base[threadIdx.x] += add;
    }
    __syncthreads();
  }
}

}  // namespace

void group_rows_zero_region(int64_t n_cap, size_t *off, size_t *bytes) {
  // offsets do not depend on the base pointer
  GroupWs w = carve(reinterpret_cast<void *>(uintptr_t(4096)), n_cap);
  *off = w.zero_off;
  *bytes = w.zero_bytes;
}

}  // namespace fv2p

using namespace fv2p;

extern "C" size_t fv2p_group_rows_workspace_bytes(int64_t n_cap) {
  if (n_cap < 0) return 0;
  return carve(nullptr, n_cap).bytes;
}

extern "C" size_t fv2p_sort_rows_workspace_bytes(int64_t n_cap) { return fv2p_group_rows_workspace_bytes(n_cap); }

extern "C" int fv2p_group_rows(const int32_t *nbr, int64_t nbr_stride, int kvol, int kx, int64_t n_cap,
                               const int32_t *n_dev, int32_t *perm, int32_t *nbr_sorted, int64_t sorted_stride,
                               int32_t *tile_order, void *workspace, size_t workspace_bytes, int flags,
                               fv2p_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FV2P_REQUIRE(kvol >= 1 && kvol <= FV2P_MAX_KVOL, "group_rows: kernel volume %d out of range", kvol);
  if (kx < 1 || kvol % kx != 0) kx = 1;
  FV2P_REQUIRE(n_cap >= 0 && n_cap < (1ll << 26) && nbr_stride >= n_cap, "group_rows: bad sizes");
  FV2P_REQUIRE(!nbr_sorted || sorted_stride >= n_cap, "group_rows: sorted_stride < row capacity");
  if (n_cap == 0) return FV2P_OK;
  FV2P_REQUIRE(nbr && perm, "group_rows: null pointer argument");
  GroupWs w = carve(workspace, n_cap);
  if (!workspace || workspace_bytes < w.bytes) {
    set_error("group_rows: workspace too small (%zu < %zu)", workspace_bytes, w.bytes);
    return FV2P_ERR_WORKSPACE;
  }
  if (!(flags & FV2P_FLAG_PREFILLED)) {
    int st = cuda_status(cudaMemsetAsync(static_cast<char *>(workspace) + w.zero_off, 0, w.zero_bytes, stream),
                         "group_rows");
    if (st) return st;
  }
  const int grid = persistent_grid(kGeoCtasPerSm);
  group_hist_kernel<<<grid, kThreads, 0, stream>>>(nbr, nbr_stride, kvol, kx, kvol / kx, n_dev, n_cap, w.masks,
                                                   w.digests, w.bins, w.bin_base, w.done);
  group_scatter_kernel<<<grid, kThreads, 0, stream>>>(nbr, nbr_stride, kvol, n_dev, n_cap, w.masks, w.digests,
                                                      w.bin_base, w.cursor, perm, nbr_sorted, sorted_stride,
                                                      w.tile_masks, w.done, reinterpret_cast<int2 *>(tile_order));
  FV2P_LAUNCH_CHECK("group_rows");
  return FV2P_OK;
}

// First-generation name and signature (3x3x3-style kernels: the fastest axis has extent 3 when kvol is a multiple
// of 9, else the mask is short enough to be its own digest).
extern "C" int fv2p_sort_rows_by_mask(const int32_t *nbr, int64_t nbr_stride, int kvol, int64_t n_cap,
                                      const int32_t *n_dev, int32_t *perm, int32_t *nbr_sorted,
                                      int64_t sorted_stride, int32_t *tile_order, void *workspace,
                                      size_t workspace_bytes, fv2p_stream_t stream_) {
  const int kx = kvol == 27 ? 3 : (kvol == 8 ? 2 : (kvol == 125 ? 5 : 1));
  return fv2p_group_rows(nbr, nbr_stride, kvol, kx, n_cap, n_dev, perm, nbr_sorted, sorted_stride, tile_order,
                         workspace, workspace_bytes, 0, stream_);
}
