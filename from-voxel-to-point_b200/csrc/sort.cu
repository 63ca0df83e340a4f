// Row ordering for the tensor-core conv: sort output rows by their neighbour mask.
//
// The fused conv processes 128 output rows per tile and one kernel offset per pipeline stage; an offset is skipped
// only if NO row of the tile has that neighbour.  In point-cloud geometry a row has 4-13 of the 27 neighbours, so
// tiles cut from the rulebook's native row order need almost all 27 stages and 60-75 % of the gathered rows are
// zero fill.  Grouping rows with equal / similar masks (the mask-sort of spconv v2's implicit GEMM) roughly halves
// the stages per tile (KITTI: 17.8->7.7, 25.9->12.9, 26.9->14.8, 26.9->15.7 at strides 1/2/4/8) and doubles the
// density of the rest.  The reference has no counterpart: its gather-GEMM-scatter (spconv_ops.h:308-357) works on
// compacted per-offset pair lists instead.
//
// Results do not change: every output row still accumulates its offsets in ascending k, only the assignment of
// rows to tiles moves.  The order is a stable LSD radix sort (9-bit digits, ascending mask, ties by row), hence
// deterministic.  Outputs: perm[t] = output row processed at sorted position t, and the neighbour map permuted the
// same way (nbr_sorted[k][t] = nbr[k][perm[t]]) so the conv reads it coalesced.
#include "common.cuh"

namespace fv2p {
namespace {

constexpr int kDigitBits = 9;
constexpr int kBins = 1 << kDigitBits;

__device__ __forceinline__ int live_n(const int *n_dev, int64_t n_cap) {
  int n = n_dev ? *n_dev : (int)n_cap;
  return n < 0 ? 0 : (n > n_cap ? (int)n_cap : n);
}

__global__ void __launch_bounds__(kThreads)
nbr_mask_kernel(const int *__restrict__ nbr, int64_t nbr_stride, int kvol, const int *n_dev, int64_t n_cap,
                uint32_t *keys, int *vals) {
  const int n = live_n(n_dev, n_cap);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    uint32_t m = 0;
    for (int k = 0; k < kvol; ++k) m |= (__ldg(&nbr[(size_t)k * nbr_stride + i]) >= 0 ? 1u : 0u) << k;
    keys[i] = m;
    vals[i] = i;
  }
}

// counts[bin][chunk] = items of the chunk whose digit is `bin`
__global__ void __launch_bounds__(kThreads)
sort_hist_kernel(const uint32_t *__restrict__ keys, const int *n_dev, int64_t n_cap, int shift, int *counts,
                 int n_chunks) {
  __shared__ int hist[kBins];
  const int n = live_n(n_dev, n_cap);
  for (int c = blockIdx.x; c < live_chunks(n); c += gridDim.x) {
    for (int b = threadIdx.x; b < kBins; b += kThreads) hist[b] = 0;
    __syncthreads();
    for (int p = 0; p < kItemsPerThread; ++p) {
      const int i = c * kChunk + p * kThreads + threadIdx.x;
      if (i < n) atomicAdd(&hist[(keys[i] >> shift) & (kBins - 1)], 1);
    }
    __syncthreads();
    for (int b = threadIdx.x; b < kBins; b += kThreads) counts[(size_t)b * n_chunks + c] = hist[b];
    __syncthreads();
  }
}

// Stable scatter: destination = (items in smaller bins) + (items of this bin in earlier chunks) + (earlier items of
// this bin inside the chunk).  `chunk_prefix` holds the per-bin exclusive scan over chunks, `bin_totals` the per-bin
// totals (both from launch_scan_chunk_counts).  Inside a chunk the 32-item slices ("virtual warps", in row order)
// publish their per-bin counts as bytes; an item's rank is the sum over earlier slices plus its position among the
// equal-digit lanes of its own slice (__match_any_sync), so a chunk costs two block barriers.
__global__ void __launch_bounds__(kThreads)
sort_scatter_kernel(const uint32_t *__restrict__ keys_in, const int *__restrict__ vals_in, const int *n_dev,
                    int64_t n_cap, int shift, const int *__restrict__ chunk_prefix,
                    const int *__restrict__ bin_totals, int n_chunks, uint32_t *keys_out, int *vals_out) {
  constexpr int kSlices = kChunk / 32;
  __shared__ int bin_base[kBins];
  __shared__ __align__(16) uint8_t slice_cnt[kSlices][kBins];
  __shared__ int scan_smem[kThreads / 32 + 1];
  const int n = live_n(n_dev, n_cap);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  {  // exclusive scan of the 512 bin totals (two per thread)
    const int a = bin_totals[2 * threadIdx.x], b = bin_totals[2 * threadIdx.x + 1];
    int total;
    const int ex = block_exclusive_scan(a + b, scan_smem, total);
    bin_base[2 * threadIdx.x] = ex;
    bin_base[2 * threadIdx.x + 1] = ex + a;
  }
  __syncthreads();
  for (int c = blockIdx.x; c < live_chunks(n); c += gridDim.x) {
    uint32_t *zero = reinterpret_cast<uint32_t *>(&slice_cnt[0][0]);
    for (int e = threadIdx.x; e < kSlices * kBins / 4; e += kThreads) zero[e] = 0u;
    __syncthreads();
    uint32_t key[kItemsPerThread];
    int val[kItemsPerThread], dig[kItemsPerThread], rank[kItemsPerThread];
#pragma unroll
    for (int p = 0; p < kItemsPerThread; ++p) {
      const int i = c * kChunk + p * kThreads + threadIdx.x;
      const bool live = i < n;
      key[p] = live ? keys_in[i] : 0u;
      val[p] = live ? vals_in[i] : 0;
      dig[p] = live ? (int)((key[p] >> shift) & (kBins - 1)) : -1;
      const unsigned peers = __match_any_sync(0xFFFFFFFFu, live ? dig[p] : kBins + lane);
      rank[p] = __popc(peers & ((1u << lane) - 1u));
      if (live && lane == __ffs(peers) - 1) slice_cnt[p * (kThreads / 32) + warp][dig[p]] = (uint8_t)__popc(peers);
    }
    __syncthreads();
#pragma unroll
    for (int p = 0; p < kItemsPerThread; ++p) {
      if (dig[p] < 0) continue;
      const int slice = p * (kThreads / 32) + warp;
      int before = 0;
      for (int q = 0; q < slice; ++q) before += slice_cnt[q][dig[p]];
      const int dst = bin_base[dig[p]] + chunk_prefix[(size_t)dig[p] * n_chunks + c] + before + rank[p];
      keys_out[dst] = key[p];
      vals_out[dst] = val[p];
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(kThreads)
permute_nbr_kernel(const int *__restrict__ nbr, int64_t nbr_stride, int kvol, const int *__restrict__ perm,
                   const int *n_dev, int64_t n_cap, int *nbr_sorted, int64_t sorted_stride) {
  const int n = live_n(n_dev, n_cap);
  const int64_t total = (int64_t)n * kvol;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(e / n), t = (int)(e - (int64_t)k * n);
    nbr_sorted[(size_t)k * sorted_stride + t] = __ldg(&nbr[(size_t)k * nbr_stride + __ldg(&perm[t])]);
  }
}

// Tile masks for the conv's tile scheduler: masks[t] = OR of the neighbour masks of sorted rows [128 t, 128 t + 128);
// its population count = pipeline stages (per 128-byte slice) the conv spends on the tile.  One warp per tile.
constexpr int kConvTile = 128;

__global__ void __launch_bounds__(kThreads)
tile_mask_kernel(const uint32_t *__restrict__ keys_sorted, const int *n_dev, int64_t n_cap, uint32_t *masks) {
  const int n = live_n(n_dev, n_cap);
  const int n_tiles = (n + kConvTile - 1) / kConvTile;
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; t < n_tiles; t += warps) {
    uint32_t m = 0;
#pragma unroll
    for (int q = 0; q < kConvTile / 32; ++q) {
      const int i = t * kConvTile + q * 32 + lane;
      if (i < n) m |= keys_sorted[i];
    }
    m = __reduce_or_sync(0xFFFFFFFFu, m);
    if (lane == 0) masks[t] = m;
  }
}

// tile_order = (tile, mask) pairs by descending weight = popcount(mask) (counting sort over the 33 possible weights,
// one CTA).  The conv hands tiles to its CTAs in this order, i.e. longest-processing-time-first list scheduling, and
// takes the tile's active offsets from the mask.  Ties are placed in tile order (per-warp ballots under a
// block-ordered cursor), so the order is deterministic.
__global__ void __launch_bounds__(1024)
tile_rank_kernel(const uint32_t *__restrict__ masks, const int *n_dev, int64_t n_cap, int2 *tile_order) {
  __shared__ int hist[33], base[33];
  __shared__ int warp_cnt[32][33];
  const int n = live_n(n_dev, n_cap);
  const int n_tiles = (n + kConvTile - 1) / kConvTile;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x < 33) hist[threadIdx.x] = 0;
  __syncthreads();
  for (int t = threadIdx.x; t < n_tiles; t += blockDim.x) atomicAdd(&hist[__popc(masks[t])], 1);
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
    const uint32_t m = t < n_tiles ? masks[t] : 0u;
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
      for (int q = 0; q < 32; ++q) add += warp_cnt[q][threadIdx.x];
      base[threadIdx.x] += add;
    }
    __syncthreads();
  }
}

struct SortWorkspace {
  uint32_t *keys_a, *keys_b;
  int *vals_b, *counts, *totals;
  uint32_t *tile_masks;
  int n_chunks;
  size_t bytes;
};

SortWorkspace carve(void *ws, int64_t n_cap) {
  SortWorkspace w;
  Carver c(ws);
  const size_t n = (size_t)(n_cap > 0 ? n_cap : 1);
  w.n_chunks = (int)((n_cap + kChunk - 1) / kChunk);
  if (w.n_chunks < 1) w.n_chunks = 1;
  w.keys_a = c.take<uint32_t>(n);
  w.keys_b = c.take<uint32_t>(n);
  w.vals_b = c.take<int>(n);
  w.counts = c.take<int>((size_t)kBins * w.n_chunks);
  w.totals = c.take<int>(kBins);
  w.tile_masks = c.take<uint32_t>(n / kConvTile + 1);
  w.bytes = c.used + 256;
  return w;
}

}  // namespace
}  // namespace fv2p

using namespace fv2p;

extern "C" size_t fv2p_sort_rows_workspace_bytes(int64_t n_cap) {
  if (n_cap < 0) return 0;
  return carve(nullptr, n_cap).bytes;
}

extern "C" int fv2p_sort_rows_by_mask(const int32_t *nbr, int64_t nbr_stride, int kvol, int64_t n_cap,
                                      const int32_t *n_dev, int32_t *perm, int32_t *nbr_sorted,
                                      int64_t sorted_stride, int32_t *tile_order, void *workspace,
                                      size_t workspace_bytes,
                                      fv2p_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FV2P_REQUIRE(kvol >= 1 && kvol <= FV2P_MAX_KVOL, "sort_rows: kernel volume %d out of range", kvol);
  FV2P_REQUIRE(n_cap >= 0 && n_cap < (1ll << 26) && nbr_stride >= n_cap, "sort_rows: bad sizes");
  FV2P_REQUIRE(!nbr_sorted || sorted_stride >= n_cap, "sort_rows: sorted_stride < row capacity");
  if (n_cap == 0) return FV2P_OK;
  FV2P_REQUIRE(nbr && perm, "sort_rows: null pointer argument");
  SortWorkspace w = carve(workspace, n_cap);
  if (!workspace || workspace_bytes < w.bytes) {
    set_error("sort_rows: workspace too small (%zu < %zu)", workspace_bytes, w.bytes);
    return FV2P_ERR_WORKSPACE;
  }
  const int grid = persistent_grid();
  const int passes = (kvol + kDigitBits - 1) / kDigitBits;
  // ping-pong (keys_a, perm) <-> (keys_b, vals_b), arranged so that the last pass lands in `perm`
  uint32_t *k_src = w.keys_a, *k_dst = w.keys_b;
  int *v_src = (passes % 2 == 0) ? perm : w.vals_b;
  int *v_dst = (passes % 2 == 0) ? w.vals_b : perm;
  nbr_mask_kernel<<<grid, kThreads, 0, stream>>>(nbr, nbr_stride, kvol, n_dev, n_cap, k_src, v_src);
  for (int p = 0; p < passes; ++p) {
    const int shift = p * kDigitBits;
    sort_hist_kernel<<<grid, kThreads, 0, stream>>>(k_src, n_dev, n_cap, shift, w.counts, w.n_chunks);
    launch_scan_chunk_counts(w.counts, kBins, w.n_chunks, n_dev, n_cap, w.totals, stream);
    sort_scatter_kernel<<<grid, kThreads, 0, stream>>>(k_src, v_src, n_dev, n_cap, shift, w.counts, w.totals,
                                                       w.n_chunks, k_dst, v_dst);
    uint32_t *tk = k_src;
    k_src = k_dst;
    k_dst = tk;
    int *tv = v_src;
    v_src = v_dst;
    v_dst = tv;
  }
  // v_src now holds the sorted rows and, by construction, is `perm`
  if (nbr_sorted)
    permute_nbr_kernel<<<grid, kThreads, 0, stream>>>(nbr, nbr_stride, kvol, perm, n_dev, n_cap, nbr_sorted,
                                                      sorted_stride);
  if (tile_order) {  // k_src holds the sorted masks
    tile_mask_kernel<<<grid, kThreads, 0, stream>>>(k_src, n_dev, n_cap, w.tile_masks);
    tile_rank_kernel<<<1, 1024, 0, stream>>>(w.tile_masks, n_dev, n_cap, reinterpret_cast<int2 *>(tile_order));
  }
  FV2P_LAUNCH_CHECK("sort_rows_by_mask");
  return FV2P_OK;
}
