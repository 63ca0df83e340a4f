// Rulebook (indice pair) generation on sm_100a: hash build + probe, ordered by single-pass scans.
//
// Replaces getIndicePair<3> (pcdet/ops/spconv/include/spconv/spconv_ops.h:28-141), which memsets a
// dense int32 grid of batch*D*H*W cells (370 MB per KITTI/Waymo frame at stride 1) per call, and its
// functors (include/spconv/geometry.h:145-297 on CPU; include/spconv/indice.cu.h:24-203 on GPU).
// Here the grid is an O(N) open-addressing table of 16-byte slots that stays in L2, and the sequential
// orderings of the reference's CPU path are reproduced without atomics on the output order.
//
// What the convolution reads is the output-major neighbour map nbr[k][i] (input row feeding output row i through
// offset k, or -1).  Second generation of this file (round 2): a backbone's geometry pass is
//
//   ONE fill launch  (fv2p_geometry_prefill) clears every table / scan state / -1 region of every book of the step
//   level 0          table_insert (1 launch)
//   submanifold      subm_probe_sym (1 launch): for odd kernels with dilation 1 only offsets k <= K/2 are probed -
//                    a hit (i, k) -> j also IS the pair (j, K-1-k) -> i - and a hit costs one 16-byte load
//   strided          conv_insert -> conv_rank -> conv_nbr (3 launches).  Every input enumerates its candidate outputs
//                    in getValidOutPos order (geometry.h:25-85), inserts them into the output table and atomicMin's
//                    bid = in_row*E + enum_index on the slot.  The candidate that owns a slot's minimum is the one
//                    the serial loop would have met first (geometry.h:181-187); conv_rank counts winners per
//                    512-input chunk and turns the counts into first-touch output rows with ONE decoupled look-back
//                    scan (no separate count / scan / assign launches); the table then maps output coordinates to
//                    rows and doubles as the input table of the next level's submanifold rulebook.
//
// (Round 1 spent 5 launches per submanifold and 9 per strided rulebook, plus 13 for the row sort.)
//
// The reference-layout tensors pairs [K,2,N] / pair_num [K] are not on the path any more: fv2p_subm_pairs /
// fv2p_conv_pairs build them on demand from the neighbour map / the table (flags along the input rows + ordered
// block scans write the pair lists in ascending input row, then the -1 tail), bit-identical to the reference's CPU
// path.  The first-generation entry points (fv2p_rulebook_subm / _conv / fv2p_get_indice_pairs_3d) are kept on top
// of the same kernels.
//
// All row counts live in device scalars; grids are persistent, so nothing here synchronises.
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

// ------------------------------------------------------------------------------------ level-0 table
// geometry.h:276-280: grid[index] = j, later duplicates overwrite -> keep the largest row (smallest ~row).
__global__ void __launch_bounds__(kThreads)
table_insert_kernel(const int4 *__restrict__ indices, const int *n_dev, int64_t n_cap, int D, int H, int W,
                    Slot *table, uint32_t tmask, int *status) {
  const int n = live_count(n_dev, n_cap);
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
    int4 c = __ldg(&indices[j]);
    const uint32_t slot = slot_insert(table, tmask, voxel_key(c.x, c.y, c.z, c.w, D, H, W));
    if (slot != 0xFFFFFFFFu) atomicMin(&table[slot].val, ~j);
    else if (status) atomicOr(status, FV2P_STATUS_OUT_OVERFLOW);
  }
}

// ------------------------------------------------------------------------------------ submanifold
// Symmetric probe (odd kernel, dilation 1).  Work item = (512-row chunk, offset k <= K/2): the two probes of a
// thread are issued as independent loads before either is resolved.  nbr[k][i] is written for every row (coalesced);
// a hit also writes the mirrored entry nbr[K-1-k][j] = i, whose row must have been filled with -1 beforehand.
__global__ void __launch_bounds__(kThreads)
subm_probe_sym_kernel(const int4 *__restrict__ indices, const int *n_dev, int64_t n_cap, Geom g,
                      const Slot *__restrict__ table, uint32_t tmask, int *nbr, int64_t nbr_stride) {
  const int n = live_count(n_dev, n_cap);
  const int D = g.out_shape[0], H = g.out_shape[1], W = g.out_shape[2];
  const int half = g.kvol / 2;
  const int work = live_chunks(n) * (half + 1);
  for (int w = blockIdx.x; w < work; w += gridDim.x) {
    const int c = w / (half + 1), k = w - c * (half + 1);
    const int base = c * kChunk;
    const int kx = k % g.ksize[2], ky = (k / g.ksize[2]) % g.ksize[1], kz = k / (g.ksize[2] * g.ksize[1]);
    const int dz = kz - g.pad[0], dy = ky - g.pad[1], dx = kx - g.pad[2];
    unsigned long long key[kItemsPerThread];
    uint4 first[kItemsPerThread];
    uint32_t slot[kItemsPerThread];
#pragma unroll
    for (int p = 0; p < kItemsPerThread; ++p) {
      const int i = base + p * kThreads + threadIdx.x;
      key[p] = kEmptyKey;
      first[p] = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0u, 0u);
      slot[p] = 0;
      if (i < n) {
        const int4 o = __ldg(&indices[i]);
        const int z = o.y + dz, y = o.z + dy, x = o.w + dx;
        if (z >= 0 && z < D && y >= 0 && y < H && x >= 0 && x < W) {
          key[p] = voxel_key(o.x, z, y, x, D, H, W);
          slot[p] = mix64(key[p]) & tmask;
          first[p] = __ldg(reinterpret_cast<const uint4 *>(&table[slot[p]]));
        }
      }
    }
#pragma unroll
    for (int p = 0; p < kItemsPerThread; ++p) {
      const int i = base + p * kThreads + threadIdx.x;
      if (i >= n) continue;
      int found = -1;
      if (key[p] != kEmptyKey) {
        uint4 q = first[p];
        uint32_t s = slot[p];
        for (uint32_t probes = 0; probes <= tmask; ++probes) {  // bounded: an overflowed table may be full
          const unsigned long long seen = ((unsigned long long)q.y << 32) | q.x;
          if (seen == key[p]) {
            const int v = (int)q.z;
            found = v < 0 ? ~v : -1;
            break;
          }
          if (seen == kEmptyKey) break;
          s = (s + 1) & tmask;
          q = __ldg(reinterpret_cast<const uint4 *>(&table[s]));
        }
      }
      nbr[(size_t)k * nbr_stride + i] = found;
      if (k != half && found >= 0) nbr[(size_t)(g.kvol - 1 - k) * nbr_stride + found] = i;
    }
  }
}

// Output-side probe of EVERY offset (geometries without mirror symmetry: even kernels, dilation > 1):
// nbr[k][i] = input row at  in = out - pad + k*dil, or -1.
__global__ void __launch_bounds__(kThreads)
subm_probe_full_kernel(const int4 *__restrict__ indices, const int *n_dev, int64_t n_cap, Geom g,
                       const Slot *__restrict__ table, uint32_t tmask, int *nbr, int64_t nbr_stride) {
  const int n = live_count(n_dev, n_cap);
  const int D = g.out_shape[0], H = g.out_shape[1], W = g.out_shape[2];
  const int work = live_chunks(n) * g.kvol;
  for (int w = blockIdx.x; w < work; w += gridDim.x) {
    const int c = w / g.kvol, k = w - c * g.kvol;
    const int kx = k % g.ksize[2], ky = (k / g.ksize[2]) % g.ksize[1], kz = k / (g.ksize[2] * g.ksize[1]);
    const int dz = kz * g.dil[0] - g.pad[0], dy = ky * g.dil[1] - g.pad[1], dx = kx * g.dil[2] - g.pad[2];
    for (int p = 0; p < kItemsPerThread; ++p) {
      const int i = c * kChunk + p * kThreads + threadIdx.x;
      if (i >= n) continue;
      const int4 o = __ldg(&indices[i]);
      const int z = o.y + dz, y = o.z + dy, x = o.w + dx;
      int found = -1;
      if (z >= 0 && z < D && y >= 0 && y < H && x >= 0 && x < W) {
        const int v = slot_lookup(table, tmask, voxel_key(o.x, z, y, x, D, H, W));
        found = v < 0 ? ~v : -1;
      }
      nbr[(size_t)k * nbr_stride + i] = found;
    }
  }
}

// Input-side probe following geometry.h:281-295 literally: mat[k][j] = output row hit by input j through offset k.
// Used for the pair lists of non-mirror submanifold geometries and (with the output table) of strided rulebooks.
__global__ void __launch_bounds__(kThreads)
input_side_kernel(const int4 *__restrict__ indices, const int *n_dev, int64_t n_cap, Geom g,
                  const Slot *__restrict__ table, uint32_t tmask, int *mat, int64_t mat_stride) {
  const int n = live_count(n_dev, n_cap);
  const int D = g.out_shape[0], H = g.out_shape[1], W = g.out_shape[2];
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
    const int4 q = __ldg(&indices[j]);
    for (int k = 0; k < g.kvol; ++k) mat[(size_t)k * mat_stride + j] = -1;
    const int pos[3] = {q.y, q.z, q.w};
    const Candidates cs = candidate_range(g, pos);
    for (int e = 0; e < cs.total; ++e) {
      int o[3], k;
      if (!candidate_at(g, cs, pos, e, o, k)) continue;
      const int v = slot_lookup(table, tmask, voxel_key(q.x, o[0], o[1], o[2], D, H, W));
      if (v < 0) mat[(size_t)k * mat_stride + j] = ~v;
    }
  }
}

// ------------------------------------------------------------------------------- strided: stage 1
// Each input inserts its candidate outputs and bids  j*E + e  for them.
__global__ void __launch_bounds__(kThreads)
conv_insert_kernel(const int4 *__restrict__ indices, const int *n_dev, int64_t n_cap, Geom g, Slot *table,
                   uint32_t tmask, int *cand_slot, int *status) {
  const int n = live_count(n_dev, n_cap);
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
          // the table is sized by the caller's output capacity: with more distinct outputs than that it fills up -
          // flag it like the row overflow (the caller enlarges the capacity and runs the step again)
          const uint32_t slot = slot_insert(table, tmask, voxel_key(q.x, o[0], o[1], o[2], D, H, W));
          if (slot != 0xFFFFFFFFu) {
            atomicMin(&table[slot].val, j * g.emax + e);
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

// ------------------------------------------------------------------------------- strided: stage 2
// Winners -> first-touch output rows, in one pass.  A candidate owns its output voxel iff the slot still holds its
// bid (losers see a smaller bid, or the winner's ~row, which is negative).  Chunks of 512 inputs are taken through a
// ticket; the chunk's winner count goes through the decoupled look-back scan; the winners then write their output
// coordinates and replace the bid by ~row.  The chunk that holds the last input publishes the output row count.
__global__ void __launch_bounds__(kThreads)
conv_rank_kernel(const int4 *__restrict__ indices, const int *n_dev, int64_t n_cap, Geom g, int64_t out_cap,
                 Slot *table, const int *__restrict__ cand_slot, unsigned long long *scan_state, int *ticket,
                 int4 *out_indices, int *n_out_dev, int *status) {
  __shared__ int smem[kThreads / 32 + 1];
  __shared__ int s_chunk, s_word;
  const int n = live_count(n_dev, n_cap);
  const int chunks = live_chunks(n);
  while (true) {
    if (threadIdx.x == 0) s_chunk = atomicAdd(ticket, 1);
    __syncthreads();
    const int c = s_chunk;
    __syncthreads();
    if (chunks == 0) {
      if (c == 0 && threadIdx.x == 0) *n_out_dev = 0;
      break;
    }
    if (c >= chunks) break;
    const int base = c * kChunk;
    uint32_t wmask[kItemsPerThread];
    int ex[kItemsPerThread], tot[kItemsPerThread];
#pragma unroll
    for (int p = 0; p < kItemsPerThread; ++p) {
      const int j = base + p * kThreads + threadIdx.x;
      uint32_t m = 0;
      if (j < n) {
        for (int e = 0; e < g.emax; ++e) {
          const int s = cand_slot[(size_t)j * g.emax + e];
          if (s >= 0 && *reinterpret_cast<volatile int *>(&table[s].val) == j * g.emax + e) m |= 1u << e;
        }
      }
      wmask[p] = m;
      ex[p] = block_exclusive_scan(__popc(m), smem, tot[p]);
    }
    int chunk_total = 0;
#pragma unroll
    for (int p = 0; p < kItemsPerThread; ++p) chunk_total += tot[p];
    int running = lookback_exclusive(scan_state, c, chunk_total, &s_word);
    if (c == chunks - 1 && threadIdx.x == 0) {
      int total = running + chunk_total;
      if (total > out_cap) {
        if (status) atomicOr(status, FV2P_STATUS_OUT_OVERFLOW);
        total = (int)out_cap;
      }
      *n_out_dev = total;
    }
#pragma unroll
    for (int p = 0; p < kItemsPerThread; ++p) {
      const int j = base + p * kThreads + threadIdx.x;
      int row = running + ex[p];
      running += tot[p];
      uint32_t rest = wmask[p];
      if (rest) {
        const int4 q = __ldg(&indices[j]);
        const int pos[3] = {q.y, q.z, q.w};
        const Candidates cs = candidate_range(g, pos);
        while (rest) {
          const int e = __ffs(rest) - 1;
          rest &= rest - 1;
          int o[3], k;
          candidate_at(g, cs, pos, e, o, k);
          const int s = cand_slot[(size_t)j * g.emax + e];
          if (row < out_cap) {
            out_indices[row] = make_int4(q.x, o[0], o[1], o[2]);
            table[s].val = ~row;
          } else {
            table[s].val = kValEmpty;
          }
          ++row;
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------- strided: stage 3
// nbr[k][out_row] = j for every candidate of every input (the map was filled with -1 beforehand).
__global__ void __launch_bounds__(kThreads)
conv_nbr_kernel(const int4 *__restrict__ indices, const int *n_dev, int64_t n_cap, Geom g,
                const Slot *__restrict__ table, const int *__restrict__ cand_slot, int *nbr, int64_t nbr_stride) {
  const int n = live_count(n_dev, n_cap);
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
    const int4 q = __ldg(&indices[j]);
    const int pos[3] = {q.y, q.z, q.w};
    const Candidates cs = candidate_range(g, pos);
    for (int e = 0; e < cs.total && e < g.emax; ++e) {
      const int s = cand_slot[(size_t)j * g.emax + e];
      if (s < 0) continue;
      const int v = table[s].val;
      if (v >= 0) continue;  // over capacity
      int o[3], k;
      candidate_at(g, cs, pos, e, o, k);
      nbr[(size_t)k * nbr_stride + (~v)] = j;
    }
  }
}

// --------------------------------------------------------------------------------- compaction (pair lists)
// counts[k][chunk] = entries >= 0 of matrix row k inside the chunk
__global__ void __launch_bounds__(kThreads)
mat_counts_kernel(const int *__restrict__ mat, int64_t mat_stride, const int *n_dev, int64_t n_cap, int kvol,
                  int *counts, int n_chunks) {
  const int n = live_count(n_dev, n_cap);
  const int work = live_chunks(n) * kvol;
  for (int w = blockIdx.x; w < work; w += gridDim.x) {
    const int c = w / kvol, k = w - c * kvol;
    int total = 0;
    for (int p = 0; p < kItemsPerThread; ++p) {
      const int i = c * kChunk + p * kThreads + threadIdx.x;
      total += __syncthreads_count(i < n && mat[(size_t)k * mat_stride + i] >= 0);
    }
    if (threadIdx.x == 0) counts[(size_t)k * n_chunks + c] = total;
  }
}

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

bool mirror_symmetric(const Geom &g) {
  bool mirror = true;
  for (int a = 0; a < 3; ++a) mirror = mirror && (g.ksize[a] % 2 == 1) && g.dil[a] == 1;
  return mirror;
}

// strided rulebook scratch: candidate slots, scan state, ticket
struct ConvWs {
  int *cand_slot;
  unsigned long long *scan_state;
  int *ticket;  // [0] ticket of conv_rank
  size_t zero_off, zero_bytes;  // region that must be zero when the call starts (scan state + ticket)
  size_t bytes;
};

ConvWs carve_conv(void *ws, int64_t n_in_cap, int emax_bound) {
  ConvWs w;
  Carver c(ws);
  const int n_chunks = (int)((n_in_cap + kChunk - 1) / kChunk) + 1;
  w.cand_slot = c.take<int>((size_t)(n_in_cap > 0 ? n_in_cap : 1) * emax_bound);
  w.scan_state = c.take<unsigned long long>(n_chunks);
  w.zero_off = (size_t)(reinterpret_cast<char *>(w.scan_state) - static_cast<char *>(ws));
  w.ticket = c.take<int>(64);
  w.zero_bytes = c.used - w.zero_off;
  w.bytes = c.used + 256;
  return w;
}

// pair-list scratch: input-major matrix (strided / non-mirror), per-chunk counts, totals
struct PairWs {
  int *mat, *counts, *row_totals, *scalars;
  int n_chunks;
  size_t bytes;
};

PairWs carve_pairs(void *ws, int64_t n_in_cap, int kvol) {
  PairWs w;
  Carver c(ws);
  w.n_chunks = (int)((n_in_cap + kChunk - 1) / kChunk);
  if (w.n_chunks < 1) w.n_chunks = 1;
  w.scalars = c.take<int>(64);
  w.mat = c.take<int>((size_t)(n_in_cap > 0 ? n_in_cap : 1) * kvol);
  w.counts = c.take<int>((size_t)kvol * w.n_chunks);
  w.row_totals = c.take<int>(FV2P_MAX_KVOL + 1);
  w.bytes = c.used + 256;
  return w;
}

// mat (row-major [kvol][stride]) -> reference-layout pair lists
void launch_compaction(const int *mat, int64_t mat_stride, const int *n_dev, int64_t n_cap, int kvol, int mirror,
                       const PairWs &w, int *pairs, int64_t pair_stride, int *pair_num, cudaStream_t stream) {
  const int grid = persistent_grid(kGeoCtasPerSm);
  const int *n_live = n_dev;
  if (!n_live) {  // host-known count: publish it once so the chunk scans can bound themselves the same way
    launch_set_scalar(w.scalars, (int)n_cap, stream);
    n_live = w.scalars;
  }
  mat_counts_kernel<<<grid, kThreads, 0, stream>>>(mat, mat_stride, n_dev, n_cap, kvol, w.counts, w.n_chunks);
  launch_scan_chunk_counts(w.counts, kvol, w.n_chunks, n_live, (int64_t)w.n_chunks * kChunk, w.row_totals, stream);
  compact_pairs_kernel<<<grid, kThreads, 0, stream>>>(mat, mat_stride, n_dev, n_cap, kvol, mirror, w.counts,
                                                      w.row_totals, w.n_chunks, pairs, pair_stride, pair_num);
}

// first-generation workspace = table + conv scratch + pair scratch, carved in this order
struct LegacyWs {
  Slot *table;
  void *conv_ws, *pair_ws;
  int64_t table_rows;
  size_t conv_bytes, pair_bytes, bytes;
};

LegacyWs carve_legacy(void *ws, int64_t n_in_cap, int64_t n_out_cap, int kvol) {
  LegacyWs w;
  Carver c(ws);
  int64_t cand = n_in_cap * kvol;
  int64_t tab = cand < n_out_cap ? cand : n_out_cap;
  if (tab < n_in_cap) tab = n_in_cap;
  w.table_rows = tab;
  w.table = c.take<Slot>(table_slots_cap(tab));
  w.conv_bytes = carve_conv(nullptr, n_in_cap, 32 < kvol ? 32 : kvol).bytes;
  w.conv_ws = c.take<char>(w.conv_bytes);
  w.pair_bytes = carve_pairs(nullptr, n_in_cap, kvol).bytes;
  w.pair_ws = c.take<char>(w.pair_bytes);
  w.bytes = c.used + 256;
  return w;
}

}  // namespace
}  // namespace fv2p

using namespace fv2p;

// ============================================================================================ v2 interface
extern "C" size_t fv2p_table_bytes(int64_t row_cap) {
  if (row_cap < 0) return 0;
  return (size_t)table_slots_cap(row_cap) * sizeof(Slot);
}

extern "C" size_t fv2p_conv_neighbours_workspace_bytes(int64_t n_in_cap, int kvol) {
  if (n_in_cap < 0 || kvol < 1 || kvol > FV2P_MAX_KVOL) return 0;
  return carve_conv(nullptr, n_in_cap, 32 < kvol ? 32 : kvol).bytes;
}

extern "C" size_t fv2p_pairs_workspace_bytes(int64_t n_in_cap, int kvol) {
  if (n_in_cap < 0 || kvol < 1 || kvol > FV2P_MAX_KVOL) return 0;
  return carve_pairs(nullptr, n_in_cap, kvol).bytes;
}

extern "C" int fv2p_geometry_prefill(const fv2p_prefill_item *items, int count, fv2p_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FV2P_REQUIRE(count >= 0 && (count == 0 || items), "geometry_prefill: null items");
  FillJob job;
  job.count = 0;
  for (int i = 0; i < count; ++i) {
    const fv2p_prefill_item &it = items[i];
    if (!it.ptr) continue;
    FV2P_REQUIRE((reinterpret_cast<uintptr_t>(it.ptr) & 15) == 0, "geometry_prefill: item %d is not 16-byte aligned", i);
    if (job.count >= kMaxFillRanges - 1) {  // flush a full job and start the next one
      int st = launch_fill(job, stream);
      if (st) return st;
      job.count = 0;
    }
    switch (it.kind) {
      case FV2P_PREFILL_TABLE:
        add_fill_table(job, it.ptr, it.a);
        break;
      case FV2P_PREFILL_NBR_ALL:  // a = kvol, b = stride (multiple of 4)
        FV2P_REQUIRE((it.b & 3) == 0, "geometry_prefill: nbr stride must be a multiple of 4");
        add_fill(job, it.ptr, (size_t)it.a * it.b * 4, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
        break;
      case FV2P_PREFILL_NBR_MIRROR: {  // rows K/2+1 .. K-1 of a submanifold map
        FV2P_REQUIRE((it.b & 3) == 0, "geometry_prefill: nbr stride must be a multiple of 4");
        const int64_t first = it.a / 2 + 1;
        add_fill(job, static_cast<char *>(it.ptr) + (size_t)first * it.b * 4, (size_t)(it.a - first) * it.b * 4,
                 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
        break;
      }
      case FV2P_PREFILL_CONV_WS: {  // a = n_in_cap, b = kvol
        ConvWs w = carve_conv(it.ptr, it.a, 32 < it.b ? 32 : (int)it.b);
        add_fill(job, static_cast<char *>(it.ptr) + w.zero_off, w.zero_bytes, 0u, 0u, 0u, 0u);
        break;
      }
      case FV2P_PREFILL_GROUP_WS: {  // a = n_cap
        size_t off = 0, bytes = 0;
        group_rows_zero_region(it.a, &off, &bytes);
        add_fill(job, static_cast<char *>(it.ptr) + off, bytes, 0u, 0u, 0u, 0u);
        break;
      }
      default:
        FV2P_REQUIRE(false, "geometry_prefill: unknown item kind %d", it.kind);
    }
  }
  return launch_fill(job, stream);
}

extern "C" int fv2p_table_build(const int32_t *indices, int64_t n_cap, const int32_t *n_dev, const int32_t *shape3,
                                void *table, int64_t table_row_cap, int32_t *status_dev, int flags,
                                fv2p_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FV2P_REQUIRE(shape3 && table, "table_build: null argument");
  FV2P_REQUIRE(n_cap >= 0 && n_cap < (1ll << 26) && table_row_cap >= n_cap, "table_build: bad capacities");
  if (!(flags & FV2P_FLAG_PREFILLED)) {
    FillJob job;
    job.count = 0;
    add_fill_table(job, table, table_row_cap);
    int st = launch_fill(job, stream);
    if (st) return st;
  }
  if (n_cap == 0) return FV2P_OK;
  FV2P_REQUIRE(indices, "table_build: null indices");
  table_insert_kernel<<<persistent_grid(kGeoCtasPerSm), kThreads, 0, stream>>>(
      reinterpret_cast<const int4 *>(indices), n_dev, n_cap, shape3[0], shape3[1], shape3[2],
      static_cast<Slot *>(table), table_slots_cap(table_row_cap) - 1, status_dev);
  FV2P_LAUNCH_CHECK("table_build");
  return FV2P_OK;
}

extern "C" int fv2p_subm_neighbours(const int32_t *indices, int64_t n_cap, const int32_t *n_dev, int batch,
                                    const int32_t *shape3, const int32_t *ksize3, const int32_t *dilation3,
                                    const void *table, int64_t table_row_cap, int32_t *nbr, int64_t nbr_stride,
                                    int flags, fv2p_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FV2P_REQUIRE(shape3 && ksize3 && table && nbr, "subm_neighbours: null argument");
  FV2P_REQUIRE(batch >= 1 && n_cap >= 0 && n_cap < (1ll << 26), "subm_neighbours: bad batch or row count");
  FV2P_REQUIRE(nbr_stride >= n_cap, "subm_neighbours: nbr_stride < row capacity");
  Geom g;
  // spconv_ops.h:76-80: submanifold forces stride 1 and padding ksize/2
  int st = fill_geom(g, shape3, ksize3, nullptr, nullptr, dilation3, "subm_neighbours");
  if (st) return st;
  if (n_cap == 0) return FV2P_OK;
  FV2P_REQUIRE(indices, "subm_neighbours: null indices");
  const int grid = persistent_grid(kGeoCtasPerSm);
  const int4 *ind4 = reinterpret_cast<const int4 *>(indices);
  const Slot *tab = static_cast<const Slot *>(table);
  const uint32_t tmask = table_slots_cap(table_row_cap) - 1;
  if (mirror_symmetric(g)) {
    if (!(flags & FV2P_FLAG_PREFILLED) && g.kvol > 1) {
      const int first = g.kvol / 2 + 1;
      fill_i32_kernel<<<grid, kThreads, 0, stream>>>(nbr + (size_t)first * nbr_stride,
                                                     (int64_t)(g.kvol - first) * nbr_stride, -1);
    }
    subm_probe_sym_kernel<<<grid, kThreads, 0, stream>>>(ind4, n_dev, n_cap, g, tab, tmask, nbr, nbr_stride);
  } else {
    subm_probe_full_kernel<<<grid, kThreads, 0, stream>>>(ind4, n_dev, n_cap, g, tab, tmask, nbr, nbr_stride);
  }
  FV2P_LAUNCH_CHECK("subm_neighbours");
  return FV2P_OK;
}

extern "C" int fv2p_conv_neighbours(const int32_t *indices, int64_t n_cap, const int32_t *n_dev, int batch,
                                    const int32_t *out_shape3, const int32_t *ksize3, const int32_t *stride3,
                                    const int32_t *pad3, const int32_t *dilation3, int32_t *out_indices,
                                    int64_t out_cap, int32_t *n_out_dev, void *table, int64_t table_row_cap,
                                    int32_t *nbr, int64_t nbr_stride, int32_t *status_dev, void *workspace,
                                    size_t workspace_bytes, int flags, fv2p_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FV2P_REQUIRE(out_shape3 && ksize3 && stride3 && pad3, "conv_neighbours: null geometry");
  FV2P_REQUIRE(batch >= 1 && n_cap >= 0 && n_cap < (1ll << 26), "conv_neighbours: bad batch or row count");
  FV2P_REQUIRE(n_out_dev && out_indices && table, "conv_neighbours: null output pointer");
  Geom g;
  int st = fill_geom(g, out_shape3, ksize3, stride3, pad3, dilation3, "conv_neighbours");
  if (st) return st;
  FV2P_REQUIRE(n_cap * (int64_t)g.emax < (1ll << 31), "conv_neighbours: too many rows for 32-bit bids");
  FV2P_REQUIRE(!nbr || nbr_stride >= out_cap, "conv_neighbours: nbr_stride < output capacity");
  FV2P_REQUIRE(out_cap >= 1 && table_row_cap >= 1, "conv_neighbours: zero output capacity");
  ConvWs w = carve_conv(workspace, n_cap, 32 < g.kvol ? 32 : g.kvol);
  if (!workspace || workspace_bytes < w.bytes) {
    set_error("conv_neighbours: workspace too small (%zu < %zu)", workspace_bytes, w.bytes);
    return FV2P_ERR_WORKSPACE;
  }
  const int grid = persistent_grid(kGeoCtasPerSm);
  if (!(flags & FV2P_FLAG_PREFILLED)) {
    FillJob job;
    job.count = 0;
    add_fill_table(job, table, table_row_cap);
    add_fill(job, static_cast<char *>(workspace) + w.zero_off, w.zero_bytes, 0u, 0u, 0u, 0u);
    st = launch_fill(job, stream);
    if (st) return st;
    if (nbr) fill_i32_kernel<<<grid, kThreads, 0, stream>>>(nbr, (int64_t)g.kvol * nbr_stride, -1);
  }
  const int4 *ind4 = reinterpret_cast<const int4 *>(indices);
  Slot *tab = static_cast<Slot *>(table);
  const uint32_t tmask = table_slots_cap(table_row_cap) - 1;
  if (n_cap > 0) {
    FV2P_REQUIRE(indices, "conv_neighbours: null indices");
    conv_insert_kernel<<<grid, kThreads, 0, stream>>>(ind4, n_dev, n_cap, g, tab, tmask, w.cand_slot, status_dev);
  }
  conv_rank_kernel<<<grid, kThreads, 0, stream>>>(ind4, n_dev, n_cap, g, out_cap, tab, w.cand_slot, w.scan_state,
                                                  w.ticket, reinterpret_cast<int4 *>(out_indices), n_out_dev,
                                                  status_dev);
  if (nbr && n_cap > 0)
    conv_nbr_kernel<<<grid, kThreads, 0, stream>>>(ind4, n_dev, n_cap, g, tab, w.cand_slot, nbr, nbr_stride);
  FV2P_LAUNCH_CHECK("conv_neighbours");
  return FV2P_OK;
}

extern "C" int fv2p_subm_pairs(const int32_t *indices, int64_t n_cap, const int32_t *n_dev, int batch,
                               const int32_t *shape3, const int32_t *ksize3, const int32_t *dilation3,
                               const void *table, int64_t table_row_cap, const int32_t *nbr, int64_t nbr_stride,
                               int32_t *pairs, int64_t pair_stride, int32_t *pair_num, void *workspace,
                               size_t workspace_bytes, fv2p_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FV2P_REQUIRE(shape3 && ksize3, "subm_pairs: null geometry");
  FV2P_REQUIRE(batch >= 1 && n_cap >= 0 && n_cap < (1ll << 26), "subm_pairs: bad batch or row count");
  Geom g;
  int st = fill_geom(g, shape3, ksize3, nullptr, nullptr, dilation3, "subm_pairs");
  if (st) return st;
  FV2P_REQUIRE(!pairs || pair_stride >= n_cap, "subm_pairs: pair_stride < row capacity");
  if (n_cap == 0) {
    if (pair_num) cudaMemsetAsync(pair_num, 0, sizeof(int) * g.kvol, stream);
    return FV2P_OK;
  }
  PairWs w = carve_pairs(workspace, n_cap, g.kvol);
  if (!workspace || workspace_bytes < w.bytes) {
    set_error("subm_pairs: workspace too small (%zu < %zu)", workspace_bytes, w.bytes);
    return FV2P_ERR_WORKSPACE;
  }
  if (mirror_symmetric(g)) {
    // the pair list of offset k (ascending input row, geometry.h:281-295) is row K-1-k of the output-major map
    FV2P_REQUIRE(nbr && nbr_stride >= n_cap, "subm_pairs: the neighbour map is needed for mirror-symmetric kernels");
    launch_compaction(nbr, nbr_stride, n_dev, n_cap, g.kvol, 1, w, pairs, pair_stride, pair_num, stream);
  } else {
    FV2P_REQUIRE(indices && table, "subm_pairs: indices and table are needed for this geometry");
    input_side_kernel<<<persistent_grid(kGeoCtasPerSm), kThreads, 0, stream>>>(
        reinterpret_cast<const int4 *>(indices), n_dev, n_cap, g, static_cast<const Slot *>(table),
        table_slots_cap(table_row_cap) - 1, w.mat, n_cap);
    launch_compaction(w.mat, n_cap, n_dev, n_cap, g.kvol, 0, w, pairs, pair_stride, pair_num, stream);
  }
  FV2P_LAUNCH_CHECK("subm_pairs");
  return FV2P_OK;
}

extern "C" int fv2p_conv_pairs(const int32_t *indices, int64_t n_cap, const int32_t *n_dev, int batch,
                               const int32_t *out_shape3, const int32_t *ksize3, const int32_t *stride3,
                               const int32_t *pad3, const int32_t *dilation3, const void *table,
                               int64_t table_row_cap, int32_t *pairs, int64_t pair_stride, int32_t *pair_num,
                               void *workspace, size_t workspace_bytes, fv2p_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FV2P_REQUIRE(out_shape3 && ksize3 && stride3 && pad3 && table, "conv_pairs: null argument");
  FV2P_REQUIRE(batch >= 1 && n_cap >= 0 && n_cap < (1ll << 26), "conv_pairs: bad batch or row count");
  Geom g;
  int st = fill_geom(g, out_shape3, ksize3, stride3, pad3, dilation3, "conv_pairs");
  if (st) return st;
  FV2P_REQUIRE(!pairs || pair_stride >= n_cap, "conv_pairs: pair_stride < row capacity");
  if (n_cap == 0) {
    if (pair_num) cudaMemsetAsync(pair_num, 0, sizeof(int) * g.kvol, stream);
    return FV2P_OK;
  }
  FV2P_REQUIRE(indices, "conv_pairs: null indices");
  PairWs w = carve_pairs(workspace, n_cap, g.kvol);
  if (!workspace || workspace_bytes < w.bytes) {
    set_error("conv_pairs: workspace too small (%zu < %zu)", workspace_bytes, w.bytes);
    return FV2P_ERR_WORKSPACE;
  }
  input_side_kernel<<<persistent_grid(kGeoCtasPerSm), kThreads, 0, stream>>>(
      reinterpret_cast<const int4 *>(indices), n_dev, n_cap, g, static_cast<const Slot *>(table),
      table_slots_cap(table_row_cap) - 1, w.mat, n_cap);
  launch_compaction(w.mat, n_cap, n_dev, n_cap, g.kvol, 0, w, pairs, pair_stride, pair_num, stream);
  FV2P_LAUNCH_CHECK("conv_pairs");
  return FV2P_OK;
}

// ============================================================================== first-generation interface
extern "C" size_t fv2p_rulebook_workspace_bytes(int64_t n_in_cap, int64_t n_out_cap, int kvol) {
  if (n_in_cap < 0 || n_out_cap < 0 || kvol < 1 || kvol > FV2P_MAX_KVOL) return 0;
  return carve_legacy(nullptr, n_in_cap, n_out_cap, kvol).bytes;
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
  int st = fill_geom(g, shape3, ksize3, nullptr, nullptr, dilation3, "rulebook_subm");
  if (st) return st;
  FV2P_REQUIRE(!pairs || pair_stride >= n_cap, "rulebook_subm: pair_stride < row capacity");
  FV2P_REQUIRE(!nbr || nbr_stride >= n_cap, "rulebook_subm: nbr_stride < row capacity");
  if (n_cap == 0) {
    if (pair_num) cudaMemsetAsync(pair_num, 0, sizeof(int) * g.kvol, stream);
    return FV2P_OK;
  }
  FV2P_REQUIRE(indices, "rulebook_subm: null indices");
  LegacyWs w = carve_legacy(workspace, n_cap, n_cap, g.kvol);
  if (!workspace || workspace_bytes < w.bytes) {
    set_error("rulebook_subm: workspace too small (%zu < %zu)", workspace_bytes, w.bytes);
    return FV2P_ERR_WORKSPACE;
  }
  const bool want_pairs = pairs || pair_num;
  const bool mirror = mirror_symmetric(g);
  st = fv2p_table_build(indices, n_cap, n_dev, shape3, w.table, w.table_rows, nullptr, 0, stream_);
  if (st) return st;
  // without a caller-provided map the mirror compaction still needs one: it lives in the pair scratch
  int *map = nbr;
  int64_t map_stride = nbr_stride;
  PairWs pw = carve_pairs(w.pair_ws, n_cap, g.kvol);
  if (!map && mirror && want_pairs) {
    map = pw.mat;
    map_stride = n_cap;
  }
  if (map) {
    st = fv2p_subm_neighbours(indices, n_cap, n_dev, batch, shape3, ksize3, dilation3, w.table, w.table_rows, map,
                              map_stride, 0, stream_);
    if (st) return st;
  }
  if (want_pairs) {
    if (mirror) {
      cudaStream_t ps = fork_stream(stream, pairs_stream_);
      launch_compaction(map, map_stride, n_dev, n_cap, g.kvol, 1, pw, pairs, pair_stride, pair_num, ps);
    } else {
      st = fv2p_subm_pairs(indices, n_cap, n_dev, batch, shape3, ksize3, dilation3, w.table, w.table_rows, nullptr, 0,
                           pairs, pair_stride, pair_num, w.pair_ws, w.pair_bytes, stream_);
      if (st) return st;
    }
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
  FV2P_REQUIRE(!pairs || pair_stride >= n_cap, "rulebook_conv: pair_stride < row capacity");
  if (n_cap == 0) {
    cudaMemsetAsync(n_out_dev, 0, sizeof(int), stream);
    if (pair_num) cudaMemsetAsync(pair_num, 0, sizeof(int) * g.kvol, stream);
    return FV2P_OK;
  }
  FV2P_REQUIRE(indices && out_cap >= 1, "rulebook_conv: null indices or zero output capacity");
  LegacyWs w = carve_legacy(workspace, n_cap, out_cap, g.kvol);
  if (!workspace || workspace_bytes < w.bytes) {
    set_error("rulebook_conv: workspace too small (%zu < %zu)", workspace_bytes, w.bytes);
    return FV2P_ERR_WORKSPACE;
  }
  st = fv2p_conv_neighbours(indices, n_cap, n_dev, batch, out_shape3, ksize3, stride3, pad3, dilation3, out_indices,
                            out_cap, n_out_dev, w.table, w.table_rows, nbr, nbr_stride, status_dev, w.conv_ws,
                            w.conv_bytes, 0, stream_);
  if (st) return st;
  if (pairs || pair_num) {
    cudaStream_t ps = fork_stream(stream, pairs_stream_);
    st = fv2p_conv_pairs(indices, n_cap, n_dev, batch, out_shape3, ksize3, stride3, pad3, dilation3, w.table,
                         w.table_rows, pairs, pair_stride, pair_num, w.pair_ws, w.pair_bytes, ps);
    if (st) return st;
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
  // the first 256 bytes carry the output count and the status word; the rest is the rulebook workspace
  int *n_out_dev = static_cast<int *>(workspace);
  int st = cuda_status(cudaMemsetAsync(n_out_dev, 0, 256, stream), "get_indice_pairs_3d");
  if (st) return st;
  st = fv2p_rulebook_conv(indices, n, nullptr, batch, out_shape3, ksize3, stride3, pad3, dilation3, out_indices,
                          out_cap, n_out_dev, pairs, n, pair_num, nbr, nbr_stride, n_out_dev + 1,
                          static_cast<char *>(workspace) + 256, workspace_bytes - 256, stream_, nullptr);
  if (st) return st;
  int host[2] = {0, 0};
  if (n > 0) {
    st = cuda_status(cudaMemcpyAsync(host, n_out_dev, 2 * sizeof(int), cudaMemcpyDeviceToHost, stream),
                     "get_indice_pairs_3d");
    if (st) return st;
    st = cuda_status(cudaStreamSynchronize(stream), "get_indice_pairs_3d");
    if (st) return st;
  }
  // more active outputs than out_cap: the reference would have allocated them; here the caller stated the capacity
  FV2P_REQUIRE(host[1] == 0, "get_indice_pairs_3d: more active outputs than out_cap (%lld)", (long long)out_cap);
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
  fill_i32_kernel<<<persistent_grid(kGeoCtasPerSm), kThreads, 0, stream>>>(nbr, (int64_t)kvol * nbr_stride, -1);
  if (pair_stride > 0) {
    dim3 grid(sm_count(), kvol);
    pairs_to_nbr_kernel<<<grid, kThreads, 0, stream>>>(pairs, pair_num, kvol, pair_stride, inverse, n_out, nbr,
                                                       nbr_stride);
  }
  FV2P_LAUNCH_CHECK("pairs_to_nbr");
  return FV2P_OK;
}
