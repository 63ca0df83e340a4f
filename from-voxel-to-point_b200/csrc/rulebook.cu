// Rulebook (indice pair) generation on sm_100a: hash build + probe + ordered compaction.
//
// Replaces getIndicePair<3> (pcdet/ops/spconv/include/spconv/spconv_ops.h:28-141), which memsets a
// dense int32 grid of batch*D*H*W cells (370 MB per KITTI/Waymo frame at stride 1) per call, and its
// functors (include/spconv/geometry.h:145-297 on CPU; include/spconv/indice.cu.h:24-203 on GPU).
// Here the grid is an O(N) open-addressing table that stays in L2, and the sequential orderings of the
// reference's CPU path are reproduced without atomics on the output order:
//
//   submanifold  out row i, offset k  ->  probe the input table at  in = out - pad + k*dil.
//                That gives the output-major neighbour map nbr[k][i] directly.  For odd kernels with
//                dilation 1 the pair list of offset k (ascending input row, geometry.h:281-295) is row
//                K-1-k of the same matrix compacted in row order; other geometries run an input-side
//                probe as well.
//   strided      every input enumerates its candidate outputs in getValidOutPos order (geometry.h:25-85),
//                inserts them into an output table and atomicMin's  key = in_row*E + enum_index  on the
//                slot.  The candidate that owns a slot's minimum is the one the serial loop would have
//                met first (geometry.h:181-187), so ranking the winners by key with an ordered scan gives
//                the reference's first-touch output rows.
//   compaction   per offset, flags along the input rows + an ordered block scan write the pair lists in
//                ascending input row, then the -1 tail, into the reference's [K,2,N] layout.
//
// All row counts live in device scalars; grids are persistent, so nothing here synchronises.
#include <limits.h>

#include "common.cuh"

namespace fv2p {
namespace {

struct Geom {
  int ksize[3], stride[3], pad[3], dil[3], out_shape[3];
  int kvol;
  int emax;  // bound on the raw candidate count per input
};

struct Candidates {
  int lo[3], hi[3], cnt[3], total;
};

// geometry.h:38-52 -- C integer division truncates toward zero on both host and device.
__device__ __forceinline__ Candidates candidate_range(const Geom &g, const int *p) {
  Candidates c;
  c.total = 1;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    c.lo[a] = (p[a] - (g.ksize[a] - 1) * g.dil[a] - 1 + g.stride[a] + g.pad[a]) / g.stride[a];
    c.hi[a] = (p[a] + g.pad[a]) / g.stride[a];
    c.cnt[a] = (c.hi[a] - c.lo[a]) / g.dil[a] + 1;
    c.total *= c.cnt[a];
  }
  if (c.total < 0) c.total = 0;
  return c;
}

// geometry.h:57-84 for raw enumeration index e (last axis fastest).  Returns validity.
__device__ __forceinline__ bool candidate_at(const Geom &g, const Candidates &c, const int *p, int e,
                                             int *o, int &offset) {
  bool ok = true;
  int mult = 1;
  offset = 0;
  int rest = e;
#pragma unroll
  for (int a = 2; a >= 0; --a) {
    int digit = rest % c.cnt[a];
    rest /= c.cnt[a];
    int v = c.hi[a] - digit * g.dil[a];
    o[a] = v;
    ok = ok && v >= 0 && v <= g.out_shape[a] - 1;
    offset += mult * (p[a] - v * g.stride[a] + g.pad[a]) / g.dil[a];
    mult *= g.ksize[a];
  }
  return ok;
}

__device__ __forceinline__ int live_count(const int *n_dev, int64_t n_cap) {
  int n = n_dev ? *n_dev : (int)n_cap;
  return n < 0 ? 0 : (n > n_cap ? (int)n_cap : n);
}

// ------------------------------------------------------------------------------------ table setup
__global__ void __launch_bounds__(kThreads)
table_clear_kernel(unsigned long long *keys, int *vals, const int *n_dev, int64_t n_cap, int64_t mult,
                   int64_t limit, int init) {
  int64_t want = (int64_t)live_count(n_dev, n_cap) * mult;
  if (want > limit) want = limit;
  const uint32_t slots = table_slots_for(want);
  for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < slots; s += gridDim.x * blockDim.x) {
    keys[s] = kEmptyKey;
    vals[s] = init;
  }
}

// geometry.h:276-280: grid[index] = j, later duplicates overwrite -> keep the largest row.
__global__ void __launch_bounds__(kThreads)
subm_insert_kernel(const int4 *__restrict__ indices, const int *n_dev, int64_t n_cap, int D, int H, int W,
                   unsigned long long *keys, int *vals) {
  const int n = live_count(n_dev, n_cap);
  const uint32_t mask = table_slots_for(n) - 1;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
    int4 c = __ldg(&indices[j]);
    uint32_t slot = table_insert(keys, mask, voxel_key(c.x, c.y, c.z, c.w, D, H, W));
    atomicMax(&vals[slot], j);
  }
}

// Output-side probe: mat[k][i] = input row at  in = out*stride - pad + k*dil  (stride 1 here), or -1.
// Also counts the hits per (matrix row, chunk) for the compaction.  Work item = (chunk, offset): K times more
// CTAs than a per-chunk split, and the eight probes of a thread are issued as independent loads before any
// of them is resolved (a probe is a dependent chain of L2 accesses; serialising 27 of them per thread was the
// single slowest kernel of the first profile).
__global__ void __launch_bounds__(kThreads)
subm_probe_kernel(const int4 *__restrict__ indices, const int *n_dev, int64_t n_cap, Geom g,
                  const unsigned long long *__restrict__ keys, const int *__restrict__ vals, int *mat,
                  int64_t mat_stride, int *counts, int n_chunks) {
  const int n = live_count(n_dev, n_cap);
  const uint32_t mask = table_slots_for(n) - 1;
  const int D = g.out_shape[0], H = g.out_shape[1], W = g.out_shape[2];
  const int work = live_chunks(n) * g.kvol;  // only chunks that hold live rows (capacity tails cost nothing)
  for (int w = blockIdx.x; w < work; w += gridDim.x) {
    const int c = w / g.kvol, k = w - c * g.kvol;
    const int base = c * kChunk;
    const int kx = k % g.ksize[2], ky = (k / g.ksize[2]) % g.ksize[1], kz = k / (g.ksize[2] * g.ksize[1]);
    const int dz = kz * g.dil[0] - g.pad[0], dy = ky * g.dil[1] - g.pad[1], dx = kx * g.dil[2] - g.pad[2];
    unsigned long long key[kItemsPerThread], seen[kItemsPerThread];
    uint32_t slot[kItemsPerThread];
#pragma unroll
    for (int p = 0; p < kItemsPerThread; ++p) {
      const int i = base + p * kThreads + threadIdx.x;
      key[p] = kEmptyKey;
      seen[p] = kEmptyKey;
      slot[p] = 0;
      if (i < n) {
        const int4 o = __ldg(&indices[i]);
        const int z = o.y + dz, y = o.z + dy, x = o.w + dx;
        if (z >= 0 && z < D && y >= 0 && y < H && x >= 0 && x < W) {
          key[p] = voxel_key(o.x, z, y, x, D, H, W);
          slot[p] = mix64(key[p]) & mask;
          seen[p] = __ldg(&keys[slot[p]]);
        }
      }
    }
    int found[kItemsPerThread];
#pragma unroll
    for (int p = 0; p < kItemsPerThread; ++p) {
      uint32_t s = slot[p];
      unsigned long long sk = seen[p];
      while (sk != key[p] && sk != kEmptyKey) {
        s = (s + 1) & mask;
        sk = __ldg(&keys[s]);
      }
      found[p] = (key[p] != kEmptyKey && sk == key[p]) ? (int)s : -1;
    }
#pragma unroll
    for (int p = 0; p < kItemsPerThread; ++p)
      if (found[p] >= 0) found[p] = __ldg(&vals[found[p]]);
    int total = 0;
#pragma unroll
    for (int p = 0; p < kItemsPerThread; ++p) {
      const int i = base + p * kThreads + threadIdx.x;
      if (i < n) mat[(size_t)k * mat_stride + i] = found[p];
      total += __syncthreads_count(found[p] >= 0);
    }
    if (threadIdx.x == 0) counts[(size_t)k * n_chunks + c] = total;
  }
}

// Input-side probe for submanifold geometries without mirror symmetry (even kernels, dilation > 1):
// min[k][j] = output row hit by input j through offset k, following geometry.h:281-295 literally.
__global__ void __launch_bounds__(kThreads)
subm_probe_in_kernel(const int4 *__restrict__ indices, const int *n_dev, int64_t n_cap, Geom g,
                     const unsigned long long *__restrict__ keys, const int *__restrict__ vals, int *mat,
                     int64_t mat_stride, int *counts, int n_chunks) {
  __shared__ int hits[FV2P_MAX_KVOL];
  const int n = live_count(n_dev, n_cap);
  const uint32_t mask = table_slots_for(n) - 1;
  const int D = g.out_shape[0], H = g.out_shape[1], W = g.out_shape[2];
  for (int c = blockIdx.x; c < live_chunks(n); c += gridDim.x) {
    if (threadIdx.x < FV2P_MAX_KVOL) hits[threadIdx.x] = 0;
    __syncthreads();
    const int base = c * kChunk;
    if (base < n) {
      for (int p = 0; p < kItemsPerThread; ++p) {
        const int j = base + p * kThreads + threadIdx.x;
        if (j >= n) continue;
        int4 q = __ldg(&indices[j]);
        for (int k = 0; k < g.kvol; ++k) mat[(size_t)k * mat_stride + j] = -1;
        const int pos[3] = {q.y, q.z, q.w};
        Candidates cs = candidate_range(g, pos);
        for (int e = 0; e < cs.total; ++e) {
          int o[3], k;
          if (!candidate_at(g, cs, pos, e, o, k)) continue;
          uint32_t slot = table_find(keys, mask, voxel_key(q.x, o[0], o[1], o[2], D, H, W));
          if (slot == 0xFFFFFFFFu) continue;
          mat[(size_t)k * mat_stride + j] = __ldg(&vals[slot]);
          atomicAdd(&hits[k], 1);
        }
      }
    }
    __syncthreads();
    if (threadIdx.x < g.kvol) counts[(size_t)threadIdx.x * n_chunks + c] = hits[threadIdx.x];
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------- strided: stage 1
// Each input inserts its candidate outputs and bids  j*E + e  for them.
__global__ void __launch_bounds__(kThreads)
conv_insert_kernel(const int4 *__restrict__ indices, const int *n_dev, int64_t n_cap, Geom g, int64_t out_cap,
                   unsigned long long *keys, int *vals, int *cand_slot, int *status) {
  const int n = live_count(n_dev, n_cap);
  int64_t want = (int64_t)n * g.emax;
  if (want > out_cap) want = out_cap;
  const uint32_t mask = table_slots_for(want) - 1;
  const int D = g.out_shape[0], H = g.out_shape[1], W = g.out_shape[2];
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
    int4 q = __ldg(&indices[j]);
    const int pos[3] = {q.y, q.z, q.w};
    Candidates cs = candidate_range(g, pos);
    for (int e = 0; e < g.emax; ++e) {
      int slot_out = -1;
      if (e < cs.total) {
        int o[3], k;
        if (candidate_at(g, cs, pos, e, o, k)) {
          // the table is sized by out_cap: with more distinct outputs than that it fills up - flag it like the
          // row overflow (the caller enlarges out_cap and runs the step again) instead of probing forever
          uint32_t slot = table_insert_bounded(keys, mask, voxel_key(q.x, o[0], o[1], o[2], D, H, W));
          if (slot != 0xFFFFFFFFu) {
            atomicMin(&vals[slot], j * g.emax + e);
            slot_out = (int)slot;
          } else if (status) {
            atomicOr(status, FV2P_STATUS_OUT_OVERFLOW);
          }
        }
      }
      cand_slot[(size_t)j * g.emax + e] = slot_out;
    }
  }
}

// wmask[j] = which of j's candidates own their output voxel; counts[chunk] = winners in the chunk.
__global__ void __launch_bounds__(kThreads)
conv_winner_kernel(const int *n_dev, int64_t n_cap, int emax, const int *__restrict__ vals,
                   const int *__restrict__ cand_slot, uint32_t *wmask, int *counts, int n_chunks) {
  __shared__ int smem[kThreads / 32 + 1];
  const int n = live_count(n_dev, n_cap);
  for (int c = blockIdx.x; c < live_chunks(n); c += gridDim.x) {
    const int base = c * kChunk;
    int mine = 0;
    if (base < n) {
      for (int p = 0; p < kItemsPerThread; ++p) {
        const int j = base + p * kThreads + threadIdx.x;
        if (j >= n) continue;
        uint32_t m = 0;
        for (int e = 0; e < emax; ++e) {
          int s = cand_slot[(size_t)j * emax + e];
          if (s >= 0 && vals[s] == j * emax + e) m |= 1u << e;
        }
        wmask[j] = m;
        mine += __popc(m);
      }
    }
    int total;
    block_exclusive_scan(mine, smem, total);
    if (threadIdx.x == 0) counts[c] = total;
  }
}

// Ordered scan of the winners -> output rows in first-touch order; vals[slot] becomes the row.
__global__ void __launch_bounds__(kThreads)
conv_assign_kernel(const int4 *__restrict__ indices, const int *n_dev, int64_t n_cap, Geom g, int64_t out_cap,
                   int *vals, const int *__restrict__ cand_slot, const uint32_t *__restrict__ wmask,
                   const int *__restrict__ chunk_prefix, const int *__restrict__ total_ptr, int n_chunks,
                   int4 *out_indices, int *n_out_dev, int *status) {
  __shared__ int smem[kThreads / 32 + 1];
  const int n = live_count(n_dev, n_cap);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    int total = *total_ptr;
    if (total > out_cap) {
      if (status) atomicOr(status, FV2P_STATUS_OUT_OVERFLOW);
      total = (int)out_cap;
    }
    *n_out_dev = total;
  }
  for (int c = blockIdx.x; c < live_chunks(n); c += gridDim.x) {
    const int base = c * kChunk;
    if (base >= n) continue;
    int running = chunk_prefix[c];
    for (int p = 0; p < kItemsPerThread; ++p) {
      const int j = base + p * kThreads + threadIdx.x;
      const uint32_t m = j < n ? wmask[j] : 0u;
      int total;
      int row = running + block_exclusive_scan(__popc(m), smem, total);
      running += total;
      if (m) {
        int4 q = __ldg(&indices[j]);
        const int pos[3] = {q.y, q.z, q.w};
        Candidates cs = candidate_range(g, pos);
        uint32_t rest = m;
        while (rest) {
          const int e = __ffs(rest) - 1;
          rest &= rest - 1;
          int o[3], k;
          candidate_at(g, cs, pos, e, o, k);
          const int s = cand_slot[(size_t)j * g.emax + e];
          if (row < out_cap) {
            out_indices[row] = make_int4(q.x, o[0], o[1], o[2]);
            vals[s] = row;
          } else {
            vals[s] = -1;
          }
          ++row;
        }
      }
    }
  }
}

__global__ void __launch_bounds__(kThreads)
nbr_fill_kernel(int *nbr, int64_t nbr_stride, int kvol, const int *n_out_dev, int64_t out_cap) {
  const int n = live_count(n_out_dev, out_cap);
  const int64_t total = (int64_t)kvol * n;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(t / n);
    const int i = (int)(t - (int64_t)k * n);
    nbr[(size_t)k * nbr_stride + i] = -1;
  }
}

// Emits both orientations: nbr[k][out_row] = j (output-major) and min[k][j] = out_row (input-major).
__global__ void __launch_bounds__(kThreads)
conv_pairs_kernel(const int4 *__restrict__ indices, const int *n_dev, int64_t n_cap, Geom g,
                  const int *__restrict__ vals, const int *__restrict__ cand_slot, int *nbr, int64_t nbr_stride,
                  int *mat, int64_t mat_stride, int *counts, int n_chunks) {
  __shared__ int hits[FV2P_MAX_KVOL];
  const int n = live_count(n_dev, n_cap);
  for (int c = blockIdx.x; c < live_chunks(n); c += gridDim.x) {
    if (threadIdx.x < FV2P_MAX_KVOL) hits[threadIdx.x] = 0;
    __syncthreads();
    const int base = c * kChunk;
    if (base < n) {
      for (int p = 0; p < kItemsPerThread; ++p) {
        const int j = base + p * kThreads + threadIdx.x;
        if (j >= n) continue;
        if (mat)
          for (int k = 0; k < g.kvol; ++k) mat[(size_t)k * mat_stride + j] = -1;
        int4 q = __ldg(&indices[j]);
        const int pos[3] = {q.y, q.z, q.w};
        Candidates cs = candidate_range(g, pos);
        for (int e = 0; e < cs.total && e < g.emax; ++e) {
          const int s = cand_slot[(size_t)j * g.emax + e];
          if (s < 0) continue;
          const int row = vals[s];
          if (row < 0) continue;
          int o[3], k;
          candidate_at(g, cs, pos, e, o, k);
          if (nbr) nbr[(size_t)k * nbr_stride + row] = j;
          if (mat) {
            mat[(size_t)k * mat_stride + j] = row;
            atomicAdd(&hits[k], 1);
          }
        }
      }
    }
    __syncthreads();
    if (mat && threadIdx.x < g.kvol) counts[(size_t)threadIdx.x * n_chunks + c] = hits[threadIdx.x];
    __syncthreads();
  }
}

// --------------------------------------------------------------------------------- compaction
// pairs[kk][0][t] = j, pairs[kk][1][t] = mat[src(kk)][j] for the t-th hit in ascending j; -1 tail up to n.
__global__ void __launch_bounds__(kThreads)
compact_pairs_kernel(const int *__restrict__ mat, int64_t mat_stride, const int *n_dev, int64_t n_cap, int kvol,
                     int mirror, const int *__restrict__ chunk_prefix, const int *__restrict__ row_totals,
                     int n_chunks, int *pairs, int64_t pair_stride, int *pair_num) {
  __shared__ int smem[kThreads / 32 + 1];
  const int n = live_count(n_dev, n_cap);
  const int lc = max(live_chunks(n), 1);  // chunk 0 always runs: it publishes pair_num even for an empty input
  const int work = kvol * lc;
  for (int w = blockIdx.x; w < work; w += gridDim.x) {
    const int kk = w / lc, c = w - kk * lc;
    const int src = mirror ? kvol - 1 - kk : kk;
    const int total = n > 0 ? row_totals[src] : 0;
    if (c == 0 && threadIdx.x == 0 && pair_num) pair_num[kk] = total;
    const int base = c * kChunk;
    if (base >= n || !pairs) continue;
    int running = chunk_prefix[(size_t)src * n_chunks + c];
    int *in_list = pairs + (size_t)(kk * 2 + 0) * pair_stride;
    int *out_list = pairs + (size_t)(kk * 2 + 1) * pair_stride;
    for (int p = 0; p < kItemsPerThread; ++p) {
      const int j = base + p * kThreads + threadIdx.x;
      const bool live = j < n;
      const int v = live ? mat[(size_t)src * mat_stride + j] : -1;
      int tot;
      const int pos = running + block_flag_scan(v >= 0, smem, tot);
      running += tot;
      if (live) {
        if (v >= 0) {
          in_list[pos] = j;
          out_list[pos] = v;
        } else {
          const int tail = total + (j - pos);  // misses before j = j - pos
          in_list[tail] = -1;
          out_list[tail] = -1;
        }
      }
    }
  }
}

// ------------------------------------------------------------------- pairs (reference layout) -> nbr
__global__ void __launch_bounds__(kThreads)
fill_i32_kernel(int *dst, int64_t count, int value) {
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < count; t += (int64_t)gridDim.x * blockDim.x)
    dst[t] = value;
}

__global__ void __launch_bounds__(kThreads)
pairs_to_nbr_kernel(const int *__restrict__ pairs, const int *__restrict__ pair_num, int kvol, int64_t pair_stride,
                    int inverse, int64_t n_out, int *nbr, int64_t nbr_stride) {
  for (int k = blockIdx.y; k < kvol; k += gridDim.y) {
    int hot = pair_num[k];
    if (hot > pair_stride) hot = (int)pair_stride;
    const int *src = pairs + (size_t)(k * 2 + (inverse ? 1 : 0)) * pair_stride;
    const int *dst = pairs + (size_t)(k * 2 + (inverse ? 0 : 1)) * pair_stride;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < hot; t += gridDim.x * blockDim.x) {
      const int o = dst[t];
      if (o >= 0 && o < n_out) nbr[(size_t)k * nbr_stride + o] = src[t];
    }
  }
}

// --------------------------------------------------------------------------------------- host side
struct RbWorkspace {
  int *scalars;  // [0] n (when the caller gave a host count), [1] n_out, [2..] spare
  unsigned long long *keys;
  int *vals, *cand_slot, *mat, *counts, *win_counts, *row_totals;
  uint32_t *wmask;
  int n_chunks;
  size_t bytes;
};

RbWorkspace carve(void *ws, int64_t n_in_cap, int64_t n_out_cap, int kvol) {
  RbWorkspace w;
  Carver c(ws);
  int64_t cand = n_in_cap * kvol;
  int64_t tab = cand < n_out_cap ? cand : n_out_cap;
  if (tab < n_in_cap) tab = n_in_cap;
  w.n_chunks = (int)((n_in_cap + kChunk - 1) / kChunk);
  if (w.n_chunks < 1) w.n_chunks = 1;
  w.scalars = c.take<int>(64);
  const uint32_t slots = table_slots_for(tab);
  w.keys = c.take<unsigned long long>(slots);
  w.vals = c.take<int>(slots);
  w.cand_slot = c.take<int>((size_t)(cand > 0 ? cand : 1));
  w.wmask = c.take<uint32_t>((size_t)(n_in_cap > 0 ? n_in_cap : 1));
  w.mat = c.take<int>((size_t)(cand > 0 ? cand : 1));
  w.counts = c.take<int>((size_t)kvol * w.n_chunks);
  w.win_counts = c.take<int>(w.n_chunks);
  w.row_totals = c.take<int>(FV2P_MAX_KVOL + 1);
  w.bytes = c.used + 256;
  return w;
}

int host_emax(const Geom &g) {
  // Largest raw candidate count per axis over every input position residue (brute force).
  int e = 1;
  for (int a = 0; a < 3; ++a) {
    int best = 0;
    const int span = 4 * g.stride[a] * g.ksize[a] * g.dil[a] + g.pad[a] + 8;
    for (int p = 0; p < span; ++p) {
      int lo = (p - (g.ksize[a] - 1) * g.dil[a] - 1 + g.stride[a] + g.pad[a]) / g.stride[a];
      int hi = (p + g.pad[a]) / g.stride[a];
      int cnt = (hi - lo) / g.dil[a] + 1;
      if (cnt > best) best = cnt;
    }
    e *= best;
  }
  return e;
}

int fill_geom(Geom &g, const int32_t *out_shape3, const int32_t *ksize3, const int32_t *stride3,
              const int32_t *pad3, const int32_t *dil3, const char *who) {
  g.kvol = 1;
  for (int a = 0; a < 3; ++a) {
    g.ksize[a] = ksize3[a];
    g.stride[a] = stride3 ? stride3[a] : 1;
    g.pad[a] = pad3 ? pad3[a] : ksize3[a] / 2;
    g.dil[a] = dil3 ? dil3[a] : 1;
    g.out_shape[a] = out_shape3[a];
    FV2P_REQUIRE(g.ksize[a] >= 1 && g.stride[a] >= 1 && g.dil[a] >= 1 && g.pad[a] >= 0 && g.out_shape[a] >= 1,
                 "%s: bad geometry on axis %d", who, a);
    // conv.py:79-80 / ops.py:72-73: "don't support this."
    FV2P_REQUIRE(g.stride[a] == 1 || g.dil[a] == 1, "%s: stride and dilation cannot both exceed 1", who);
    g.kvol *= g.ksize[a];
  }
  FV2P_REQUIRE(g.kvol <= FV2P_MAX_KVOL, "%s: kernel volume %d exceeds %d", who, g.kvol, FV2P_MAX_KVOL);
  g.emax = host_emax(g);
  FV2P_REQUIRE(g.emax >= 1 && g.emax <= 32, "%s: unsupported candidate fan-out %d", who, g.emax);
  return 0;
}

}  // namespace
}  // namespace fv2p

using namespace fv2p;

// The pair lists are only needed by consumers of the reference-layout tensors; with a second stream their compaction
// leaves the caller's dependency chain (the neighbour map is complete before it).
static cudaStream_t fork_for_pairs(cudaStream_t stream, fv2p_stream_t pairs_stream_) {
  return fork_stream(stream, pairs_stream_);
}

extern "C" size_t fv2p_rulebook_workspace_bytes(int64_t n_in_cap, int64_t n_out_cap, int kvol) {
  if (n_in_cap < 0 || n_out_cap < 0 || kvol < 1 || kvol > FV2P_MAX_KVOL) return 0;
  return carve(nullptr, n_in_cap, n_out_cap, kvol).bytes;
}

extern "C" int fv2p_rulebook_subm(const int32_t *indices, int64_t n_cap, const int32_t *n_dev, int batch,
                                  const int32_t *shape3, const int32_t *ksize3, const int32_t *dilation3,
                                  int32_t *pairs, int64_t pair_stride, int32_t *pair_num, int32_t *nbr,
                                  int64_t nbr_stride, void *workspace, size_t workspace_bytes,
                                  fv2p_stream_t stream_, fv2p_stream_t pairs_stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FV2P_REQUIRE(shape3 && ksize3, "rulebook_subm: null geometry");
  FV2P_REQUIRE(batch >= 1 && n_cap >= 0 && n_cap < (1ll << 26), "rulebook_subm: bad batch or row count");
  Geom g;
  // spconv_ops.h:76-80: submanifold forces stride 1 and padding ksize/2
  int st = fill_geom(g, shape3, ksize3, nullptr, nullptr, dilation3, "rulebook_subm");
  if (st) return st;
  FV2P_REQUIRE(!pairs || pair_stride >= n_cap, "rulebook_subm: pair_stride < row capacity");
  FV2P_REQUIRE(!nbr || nbr_stride >= n_cap, "rulebook_subm: nbr_stride < row capacity");
  if (n_cap == 0) {
    if (pair_num) cudaMemsetAsync(pair_num, 0, sizeof(int) * g.kvol, stream);
    return FV2P_OK;
  }
  FV2P_REQUIRE(indices, "rulebook_subm: null indices");
  RbWorkspace w = carve(workspace, n_cap, n_cap, g.kvol);
  if (!workspace || workspace_bytes < w.bytes) {
    set_error("rulebook_subm: workspace too small (%zu < %zu)", workspace_bytes, w.bytes);
    return FV2P_ERR_WORKSPACE;
  }
  const int *n_live = n_dev;
  if (!n_live) {  // host-known count: publish it once so the chunk scans can bound themselves the same way
    launch_set_scalar(w.scalars, (int)n_cap, stream);
    n_live = w.scalars;
  }
  bool mirror = true;
  for (int a = 0; a < 3; ++a) mirror = mirror && (g.ksize[a] % 2 == 1) && g.dil[a] == 1;
  const bool want_pairs = pairs || pair_num;
  const int grid = persistent_grid();
  const int4 *ind4 = reinterpret_cast<const int4 *>(indices);
  table_clear_kernel<<<grid, kThreads, 0, stream>>>(w.keys, w.vals, n_dev, n_cap, 1, n_cap, -1);
  subm_insert_kernel<<<grid, kThreads, 0, stream>>>(ind4, n_dev, n_cap, g.out_shape[0], g.out_shape[1],
                                                    g.out_shape[2], w.keys, w.vals);
  int *out_mat = nbr ? nbr : w.mat;
  int64_t out_stride = nbr ? nbr_stride : n_cap;
  if (nbr || mirror)
    subm_probe_kernel<<<grid, kThreads, 0, stream>>>(ind4, n_dev, n_cap, g, w.keys, w.vals, out_mat, out_stride,
                                                     w.counts, w.n_chunks);
  if (want_pairs) {
    const int *src_mat = out_mat;
    int64_t src_stride = out_stride;
    if (mirror) stream = fork_for_pairs(stream, pairs_stream_);  // (the other probe still needs the hash table)
    if (!mirror) {
      subm_probe_in_kernel<<<grid, kThreads, 0, stream>>>(ind4, n_dev, n_cap, g, w.keys, w.vals, w.mat, n_cap,
                                                          w.counts, w.n_chunks);
      src_mat = w.mat;
      src_stride = n_cap;
    }
    launch_scan_chunk_counts(w.counts, g.kvol, w.n_chunks, n_live, (int64_t)w.n_chunks * kChunk, w.row_totals,
                             stream);
    compact_pairs_kernel<<<grid, kThreads, 0, stream>>>(src_mat, src_stride, n_dev, n_cap, g.kvol, mirror ? 1 : 0,
                                                        w.counts, w.row_totals, w.n_chunks, pairs, pair_stride,
                                                        pair_num);
  }
  FV2P_LAUNCH_CHECK("rulebook_subm");
  return FV2P_OK;
}

extern "C" int fv2p_rulebook_conv(const int32_t *indices, int64_t n_cap, const int32_t *n_dev, int batch,
                                  const int32_t *out_shape3, const int32_t *ksize3, const int32_t *stride3,
                                  const int32_t *pad3, const int32_t *dilation3, int32_t *out_indices,
                                  int64_t out_cap, int32_t *n_out_dev, int32_t *pairs, int64_t pair_stride,
                                  int32_t *pair_num, int32_t *nbr, int64_t nbr_stride, int32_t *status_dev,
                                  void *workspace, size_t workspace_bytes, fv2p_stream_t stream_,
                                  fv2p_stream_t pairs_stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FV2P_REQUIRE(out_shape3 && ksize3 && stride3 && pad3, "rulebook_conv: null geometry");
  FV2P_REQUIRE(batch >= 1 && n_cap >= 0 && n_cap < (1ll << 26), "rulebook_conv: bad batch or row count");
  FV2P_REQUIRE(n_out_dev && out_indices, "rulebook_conv: null output pointer");
  Geom g;
  int st = fill_geom(g, out_shape3, ksize3, stride3, pad3, dilation3, "rulebook_conv");
  if (st) return st;
  FV2P_REQUIRE(n_cap * (int64_t)g.emax < (1ll << 31), "rulebook_conv: too many rows for 32-bit bids");
  FV2P_REQUIRE(!pairs || pair_stride >= n_cap, "rulebook_conv: pair_stride < row capacity");
  FV2P_REQUIRE(!nbr || nbr_stride >= out_cap, "rulebook_conv: nbr_stride < output capacity");
  if (n_cap == 0) {
    cudaMemsetAsync(n_out_dev, 0, sizeof(int), stream);
    if (pair_num) cudaMemsetAsync(pair_num, 0, sizeof(int) * g.kvol, stream);
    return FV2P_OK;
  }
  FV2P_REQUIRE(indices && out_cap >= 1, "rulebook_conv: null indices or zero output capacity");
  RbWorkspace w = carve(workspace, n_cap, out_cap, g.kvol);
  if (!workspace || workspace_bytes < w.bytes) {
    set_error("rulebook_conv: workspace too small (%zu < %zu)", workspace_bytes, w.bytes);
    return FV2P_ERR_WORKSPACE;
  }
  const bool want_pairs = pairs || pair_num;
  const int grid = persistent_grid();
  const int4 *ind4 = reinterpret_cast<const int4 *>(indices);
  const int *n_live = n_dev;
  if (!n_live) {
    launch_set_scalar(w.scalars, (int)n_cap, stream);
    n_live = w.scalars;
  }
  table_clear_kernel<<<grid, kThreads, 0, stream>>>(w.keys, w.vals, n_dev, n_cap, g.emax, out_cap, INT_MAX);
  conv_insert_kernel<<<grid, kThreads, 0, stream>>>(ind4, n_dev, n_cap, g, out_cap, w.keys, w.vals, w.cand_slot,
                                                    status_dev);
  conv_winner_kernel<<<grid, kThreads, 0, stream>>>(n_dev, n_cap, g.emax, w.vals, w.cand_slot, w.wmask,
                                                    w.win_counts, w.n_chunks);
  launch_scan_chunk_counts(w.win_counts, 1, w.n_chunks, n_live, (int64_t)w.n_chunks * kChunk,
                           w.row_totals + FV2P_MAX_KVOL, stream);
  conv_assign_kernel<<<grid, kThreads, 0, stream>>>(ind4, n_dev, n_cap, g, out_cap, w.vals, w.cand_slot, w.wmask,
                                                    w.win_counts, w.row_totals + FV2P_MAX_KVOL, w.n_chunks,
                                                    reinterpret_cast<int4 *>(out_indices), n_out_dev, status_dev);
  if (nbr) nbr_fill_kernel<<<grid, kThreads, 0, stream>>>(nbr, nbr_stride, g.kvol, n_out_dev, out_cap);
  if (nbr || want_pairs)
    conv_pairs_kernel<<<grid, kThreads, 0, stream>>>(ind4, n_dev, n_cap, g, w.vals, w.cand_slot, nbr, nbr_stride,
                                                     want_pairs ? w.mat : nullptr, n_cap, w.counts, w.n_chunks);
  if (want_pairs) {
    stream = fork_for_pairs(stream, pairs_stream_);
    launch_scan_chunk_counts(w.counts, g.kvol, w.n_chunks, n_live, (int64_t)w.n_chunks * kChunk, w.row_totals,
                             stream);
    compact_pairs_kernel<<<grid, kThreads, 0, stream>>>(w.mat, n_cap, n_dev, n_cap, g.kvol, 0, w.counts,
                                                        w.row_totals, w.n_chunks, pairs, pair_stride, pair_num);
  }
  FV2P_LAUNCH_CHECK("rulebook_conv");
  return FV2P_OK;
}

extern "C" int fv2p_get_indice_pairs_3d(const int32_t *indices, int64_t n, int batch, const int32_t *out_shape3,
                                        const int32_t *spatial_shape3, const int32_t *ksize3,
                                        const int32_t *stride3, const int32_t *pad3, const int32_t *dilation3,
                                        const int32_t *out_pad3, int subm, int transpose, int32_t *out_indices,
                                        int64_t out_cap, int32_t *pairs, int32_t *pair_num, int32_t *nbr,
                                        int64_t nbr_stride, int32_t *num_act_out_host, void *workspace,
                                        size_t workspace_bytes, fv2p_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  (void)out_pad3;
  FV2P_REQUIRE(num_act_out_host && out_shape3 && spatial_shape3 && ksize3, "get_indice_pairs_3d: null argument");
  if (transpose) {
    set_error("get_indice_pairs_3d: transposed convolution is outside the hot path (not built)");
    return FV2P_ERR_UNSUPPORTED;
  }
  if (subm) {
    for (int a = 0; a < 3; ++a)
      FV2P_REQUIRE(out_shape3[a] == spatial_shape3[a], "get_indice_pairs_3d: subm needs out_shape == spatial_shape");
    int st = fv2p_rulebook_subm(indices, n, nullptr, batch, spatial_shape3, ksize3, dilation3, pairs, n, pair_num,
                                nbr, nbr_stride, workspace, workspace_bytes, stream_, nullptr);
    if (st) return st;
    if (out_indices && out_indices != indices && n > 0) {
      // spconv_ops.h:104 returns the input tensor itself; a separate buffer gets a copy
      st = cuda_status(cudaMemcpyAsync(out_indices, indices, sizeof(int) * 4 * n, cudaMemcpyDeviceToDevice, stream),
                       "get_indice_pairs_3d");
      if (st) return st;
    }
    *num_act_out_host = (int32_t)n;
    return FV2P_OK;
  }
  FV2P_REQUIRE(workspace && workspace_bytes >= 256, "get_indice_pairs_3d: workspace too small");
  // the first 256 bytes carry the output count; the rest is the rulebook workspace
  int *n_out_dev = static_cast<int *>(workspace);
  int st = fv2p_rulebook_conv(indices, n, nullptr, batch, out_shape3, ksize3, stride3, pad3, dilation3, out_indices,
                              out_cap, n_out_dev, pairs, n, pair_num, nbr, nbr_stride, n_out_dev + 1,
                              static_cast<char *>(workspace) + 256, workspace_bytes - 256, stream_, nullptr);
  if (st) return st;
  int host[2] = {0, 0};
  if (n > 0) {
    st = cuda_status(cudaMemcpyAsync(host, n_out_dev, sizeof(int), cudaMemcpyDeviceToHost, stream),
                     "get_indice_pairs_3d");
    if (st) return st;
    st = cuda_status(cudaStreamSynchronize(stream), "get_indice_pairs_3d");
    if (st) return st;
  }
  *num_act_out_host = host[0];
  return FV2P_OK;
}

extern "C" int fv2p_pairs_to_nbr(const int32_t *pairs, const int32_t *pair_num, int kvol, int64_t pair_stride,
                                 int inverse, int64_t n_out, int32_t *nbr, int64_t nbr_stride,
                                 fv2p_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FV2P_REQUIRE(kvol >= 1 && kvol <= FV2P_MAX_KVOL, "pairs_to_nbr: kernel volume %d out of range", kvol);
  FV2P_REQUIRE(n_out >= 0 && nbr_stride >= n_out && pair_stride >= 0, "pairs_to_nbr: bad sizes");
  if (n_out == 0) return FV2P_OK;
  FV2P_REQUIRE(pairs || pair_stride == 0, "pairs_to_nbr: null pairs");
  FV2P_REQUIRE(pair_num && nbr, "pairs_to_nbr: null pointer argument");
  fill_i32_kernel<<<persistent_grid(), kThreads, 0, stream>>>(nbr, (int64_t)kvol * nbr_stride, -1);
  if (pair_stride > 0) {
    dim3 grid(sm_count(), kvol);
    pairs_to_nbr_kernel<<<grid, kThreads, 0, stream>>>(pairs, pair_num, kvol, pair_stride, inverse, n_out, nbr,
                                                       nbr_stride);
  }
  FV2P_LAUNCH_CHECK("pairs_to_nbr");
  return FV2P_OK;
}
