// Voxelization + mean VFE on sm_100a.
//
// Replaces the serial numba loop of pcdet/datasets/processor/voxel_generator.py:136-207 (plus its
// 360 MB dense coor->voxel map, :114) and MeanVFE (pcdet/models/backbones_3d/vfe/mean_vfe.py:26-28)
// for a whole batch of frames per call.  The sequential semantics are reproduced order-independently:
//
//   insert   every in-range point hashes its (frame,z,y,x) into an open-addressing table; lanes of a
//            warp that hit the same voxel are aggregated with __match_any_sync so one lane per
//            distinct voxel touches the table; first[slot] = atomicMin(point index)
//   rank     a point is its voxel's FIRST arrival iff first[slot] == its index; an ordered scan of
//            those flags over the frame gives the voxel id the numba loop would have assigned.
//            The point whose rank equals max_voxels is where that loop `break`s (:198-199):
//            cut[frame] = its index, every point at or after it is dropped.  ONE kernel: 512-point chunks
//            handed out by ticket in frame-major order, a decoupled look-back scan per frame, frame totals
//            (capped at max_voxels) published for the frames behind - the collate layout's row offsets -
//            then coordinates, slot -> row, and, when the caller hands one in, the level-0 coordinate table
//            of the sparse convolutions (the first thing they would otherwise build from these rows).
//   select   each surviving point is pushed through a chain of atomicMin on sel[voxel][0..T): the chain
//            conserves the multiset, so slot r ends up holding the (r+1)-th smallest point index -- the
//            T lowest-index points of the voxel in index order, exactly what `num < max_points` keeps.
//   reduce   one thread per voxel sums the kept points in index order and divides by the count
//            (IEEE fp32 division, like torch's `/`), writes coords, mean, num_points (+ legacy voxels).
//
// Coordinates use floorf((p - lo) / size) with IEEE-RN subtract and divide: reciprocal multiplies or
// FMA contraction mis-bin points that sit on voxel faces ("hard parts").
#include <limits.h>
#include <string.h>

#include "common.cuh"

namespace fv2p {
namespace {

struct VoxGeom {
  float lo[3];    // x,y,z lower bounds
  float size[3];  // voxel size x,y,z
  int grid[3];    // gx,gy,gz
};

constexpr int kMaxFeatures = 8;

__global__ void __launch_bounds__(kThreads)
vox_clear_kernel(unsigned long long *keys, int *first, uint32_t slots, int *cut, int batch, unsigned long long *state,
                 int chunks, int *frame_total, int *ticket, uint4 *table0, uint32_t table0_slots, uint4 table0_empty) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
  for (uint32_t s = t; s < slots; s += stride) {
    keys[s] = kEmptyKey;
    first[s] = INT_MAX;
  }
  for (uint32_t s = t; s < (uint32_t)chunks; s += stride) state[s] = 0ull;
  for (uint32_t s = t; s < table0_slots; s += stride) table0[s] = table0_empty;
  if (blockIdx.x == 0) {
    for (int b = threadIdx.x; b < batch; b += blockDim.x) {
      cut[b] = INT_MAX;
      frame_total[b] = 0;
    }
    if (threadIdx.x == 0) *ticket = 0;
  }
}

__global__ void __launch_bounds__(kThreads)
vox_insert_kernel(const float *__restrict__ points, const int *__restrict__ frame_offsets, int batch,
                  int max_chunks, int nf, VoxGeom g, unsigned long long *keys, int *first,
                  uint32_t mask, int *pslot) {
  const int lane = threadIdx.x & 31;
  const int work = batch * max_chunks;
  for (int w = blockIdx.x; w < work; w += gridDim.x) {
    const int b = w / max_chunks, c = w - b * max_chunks;
    const int begin = frame_offsets[b], end = frame_offsets[b + 1];
    const int base = begin + c * kChunk;
    if (base >= end) continue;  // uniform per block
#pragma unroll 1
    for (int p = 0; p < kItemsPerThread; ++p) {
      const int i = base + p * kThreads + threadIdx.x;
      const bool live = i < end;
      bool ok = live;
      unsigned long long key = kEmptyKey - 1 - lane;  // distinct per lane: never aggregates
      if (live) {
        const float *pt = points + (size_t)i * nf;
        int cc[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          float q = floorf(__fdiv_rn(__fsub_rn(__ldg(pt + a), g.lo[a]), g.size[a]));
          ok = ok && (q >= 0.0f) && (q < (float)g.grid[a]);
          cc[a] = (int)q;
        }
        if (ok) key = voxel_key(b, cc[2], cc[1], cc[0], g.grid[2], g.grid[1], g.grid[0]);
      }
      // warp-aggregated insert: lanes are in point-index order, so the lowest lane of a group also
      // carries the group's smallest index.
      const unsigned group = __match_any_sync(0xFFFFFFFFu, key);
      const int leader = __ffs(group) - 1;
      uint32_t slot = 0;
      if (ok && lane == leader) {
        slot = table_insert(keys, mask, key);
        atomicMin(&first[slot], i);
      }
      slot = __shfl_sync(0xFFFFFFFFu, slot, leader);
      if (live) pslot[i] = ok ? (int)slot : -1;
    }
  }
}

// rank: see the file header.  state [batch * max_chunks] (look-back words, one chain per frame), frame_total [batch]
// (rows of the frame + 1; 0 = not known yet) and ticket are zero when the kernel starts.
__global__ void __launch_bounds__(kThreads)
vox_rank_kernel(const int *__restrict__ frame_offsets, int batch, int max_chunks, const int *__restrict__ pslot,
                const int *__restrict__ first, const unsigned long long *__restrict__ keys,
                unsigned long long *state, int *frame_total, int *ticket, int max_voxels, int64_t cap,
                int max_points, VoxGeom g, int *voxel_offsets, int *status, int *slot_vid, int *cut, int *coords,
                int *sel, Slot *table0, uint32_t tmask0, int D0, int H0, int W0) {
  __shared__ int smem[kThreads / 32 + 1];
  __shared__ int s_word;
  __shared__ long long s_base;
  const int work = batch * max_chunks;
  for (;;) {
    if (threadIdx.x == 0) s_word = atomicAdd(ticket, 1);
    __syncthreads();
    const int w = s_word;
    __syncthreads();
    if (w >= work) break;
    const int b = w / max_chunks, c = w - b * max_chunks;
    const int begin = frame_offsets[b], end = frame_offsets[b + 1];
    const int base = begin + c * kChunk;
    const bool empty_frame = end <= begin;
    if (base >= end && !(empty_frame && c == 0)) continue;  // an empty frame still publishes its (zero) total
    int flag[kItemsPerThread], slot[kItemsPerThread];
    int total = 0;
#pragma unroll
    for (int p = 0; p < kItemsPerThread; ++p) {
      const int i = base + p * kThreads + threadIdx.x;
      flag[p] = 0, slot[p] = -1;
      if (i < end) {
        slot[p] = pslot[i];
        flag[p] = (slot[p] >= 0) && (first[slot[p]] == i);
      }
      total += __syncthreads_count(flag[p]);
    }
    const int excl = empty_frame ? 0 : lookback_exclusive(state + (size_t)b * max_chunks, c, total, &s_word);
    const bool last_chunk = empty_frame || base + kChunk >= end;
    const int frame_rows = min(excl + total, max_voxels);
    if (last_chunk && threadIdx.x == 0) *reinterpret_cast<volatile int *>(&frame_total[b]) = frame_rows + 1;
    // rows of the frames before this one (their last chunks hold earlier tickets: they are running or done)
    if (threadIdx.x < 32) {
      long long acc = 0;
      for (int bb = threadIdx.x; bb < b; bb += 32) {
        int v;
        do {
          v = *reinterpret_cast<volatile int *>(&frame_total[bb]);
        } while (v == 0);
        acc += v - 1;
      }
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, d);
      if (threadIdx.x == 0) s_base = acc < (long long)cap ? acc : (long long)cap;
    }
    __syncthreads();
    const long long vbase = s_base;
    if (last_chunk && threadIdx.x == 0) {
      long long e = vbase + frame_rows;
      if (e > (long long)cap) {
        if (status) atomicOr(status, FV2P_STATUS_VOXEL_OVERFLOW);
        e = cap;
      }
      voxel_offsets[b + 1] = (int)e;
      if (b == 0) voxel_offsets[0] = 0;
    }
    int running = excl;
#pragma unroll
    for (int p = 0; p < kItemsPerThread; ++p) {
      const int i = base + p * kThreads + threadIdx.x;
      int tot;
      const int rank = running + block_exclusive_scan(flag[p], smem, tot);
      running += tot;
      if (flag[p]) {
        const int s = slot[p];
        const long long vid = vbase + rank;
        if (rank < max_voxels && vid < (long long)cap) {
          unsigned long long key = keys[s];
          int x = (int)(key % (unsigned)g.grid[0]);
          key /= (unsigned)g.grid[0];
          int y = (int)(key % (unsigned)g.grid[1]);
          key /= (unsigned)g.grid[1];
          int z = (int)(key % (unsigned)g.grid[2]);
          reinterpret_cast<int4 *>(coords)[vid] = make_int4(b, z, y, x);
          slot_vid[s] = (int)vid;
          for (int t = 0; t < max_points; ++t) sel[(size_t)vid * max_points + t] = INT_MAX;
          if (table0) {  // what fv2p_table_build does with these rows (rulebook.cu:table_insert_kernel)
            // every voxel is inserted exactly once here, so the value is a plain store (table_insert_kernel must
            // expect duplicate coordinates from arbitrary callers and keeps the largest row with an atomicMin)
            const uint32_t ts = slot_insert(table0, tmask0, voxel_key(b, z, y, x, D0, H0, W0));
            if (ts != 0xFFFFFFFFu) table0[ts].val = ~(int)vid;
            else if (status) atomicOr(status, FV2P_STATUS_OUT_OVERFLOW);
          }
        } else {
          slot_vid[s] = -1;
          if (rank == max_voxels) cut[b] = i;  // the numba loop breaks here (voxel_generator.py:198)
        }
      }
    }
    __syncthreads();  // s_base / s_word are rewritten by the next item
  }
}

// Ordered top-T selection by an atomicMin chain (multiset-conserving insertion network).
__global__ void __launch_bounds__(kThreads)
vox_select_kernel(const int *__restrict__ frame_offsets, int batch, int max_chunks,
                  const int *__restrict__ pslot, const int *__restrict__ slot_vid,
                  const int *__restrict__ cut, int max_points, int *sel) {
  const int work = batch * max_chunks;
  for (int w = blockIdx.x; w < work; w += gridDim.x) {
    const int b = w / max_chunks, c = w - b * max_chunks;
    const int begin = frame_offsets[b], end = frame_offsets[b + 1];
    const int base = begin + c * kChunk;
    if (base >= end) continue;
    const int stop = min(end, cut[b]);
    for (int p = 0; p < kItemsPerThread; ++p) {
      const int i = base + p * kThreads + threadIdx.x;
      if (i >= stop) continue;
      const int s = pslot[i];
      if (s < 0) continue;
      const int vid = slot_vid[s];
      if (vid < 0) continue;
      int carry = i;
      int *list = sel + (size_t)vid * max_points;
      for (int r = 0; r < max_points && carry != INT_MAX; ++r) {
        int old = atomicMin(&list[r], carry);
        carry = max(old, carry);
      }
    }
  }
}

__global__ void __launch_bounds__(kThreads)
vox_reduce_kernel(const float *__restrict__ points, int nf, const int *__restrict__ voxel_offsets,
                  int batch, int max_points, const int *__restrict__ sel, float *features,
                  int *num_points, float *voxels) {
  const int m = voxel_offsets[batch];
  for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < m; v += gridDim.x * blockDim.x) {
    float acc[kMaxFeatures];
#pragma unroll
    for (int f = 0; f < kMaxFeatures; ++f) acc[f] = 0.0f;
    int n = 0;
    for (int t = 0; t < max_points; ++t) {
      const int idx = sel[(size_t)v * max_points + t];
      const bool have = idx != INT_MAX;
      const float *pt = points + (size_t)(have ? idx : 0) * nf;
#pragma unroll
      for (int f = 0; f < kMaxFeatures; ++f) {
        if (f < nf) {
          float val = have ? __ldg(pt + f) : 0.0f;
          acc[f] = __fadd_rn(acc[f], val);
          if (voxels) voxels[((size_t)v * max_points + t) * nf + f] = val;
        }
      }
      n += have;
    }
    const float denom = (float)(n < 1 ? 1 : n);
#pragma unroll
    for (int f = 0; f < kMaxFeatures; ++f)
      if (f < nf) features[(size_t)v * nf + f] = __fdiv_rn(acc[f], denom);
    if (num_points) num_points[v] = n;
  }
}

__global__ void __launch_bounds__(kThreads)
mean_vfe_kernel(const float *__restrict__ voxels, const int *__restrict__ num_points, int64_t m,
                int max_points, int nf, float *out) {
  const int64_t total = m * nf;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total;
       e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t v = e / nf;
    const int f = (int)(e - v * nf);
    float s = 0.0f;
    for (int t = 0; t < max_points; ++t) s = __fadd_rn(s, voxels[(v * max_points + t) * nf + f]);
    const int n = num_points[v];
    out[e] = __fdiv_rn(s, (float)(n < 1 ? 1 : n));
  }
}

struct VoxWorkspace {
  unsigned long long *keys;
  int *first, *slot_vid, *pslot, *frame_total, *cut, *sel, *scratch;
  unsigned long long *state;
  uint32_t slots;
  int max_chunks;
  size_t bytes;
};

VoxWorkspace carve(void *ws, int64_t total_points, int batch, int64_t max_frame_points, int max_points,
                   int64_t cap) {
  VoxWorkspace w;
  Carver c(ws);
  w.slots = table_slots_for(total_points);
  w.max_chunks = (int)((max_frame_points + kChunk - 1) / kChunk);
  if (w.max_chunks < 1) w.max_chunks = 1;
  w.keys = c.take<unsigned long long>(w.slots);
  w.first = c.take<int>(w.slots);
  w.slot_vid = c.take<int>(w.slots);
  w.pslot = c.take<int>(total_points > 0 ? total_points : 1);
  w.state = c.take<unsigned long long>((size_t)batch * w.max_chunks);
  w.frame_total = c.take<int>(batch);
  w.cut = c.take<int>(batch);
  w.sel = c.take<int>((size_t)(cap > 0 ? cap : 1) * max_points);
  w.scratch = c.take<int>(16);
  w.bytes = c.used + 256;
  return w;
}

}  // namespace
}  // namespace fv2p

using namespace fv2p;

extern "C" size_t fv2p_voxelize_workspace_bytes(int64_t total_points, int batch, int64_t max_frame_points,
                                                int max_points, int64_t cap) {
  if (total_points < 0 || batch < 1 || max_points < 1) return 0;
  return carve(nullptr, total_points, batch, max_frame_points, max_points, cap).bytes;
}

extern "C" int fv2p_voxelize_mean(const float *points, const int32_t *frame_offsets, int64_t total_points,
                                  int batch, int64_t max_frame_points, int num_features,
                                  const float *range6, const float *vsize3, int max_points,
                                  int max_voxels, int32_t *coords, float *voxel_features,
                                  int32_t *num_points, float *voxels, int32_t *voxel_offsets, int64_t cap,
                                  int32_t *status_dev, void *workspace, size_t workspace_bytes,
                                  fv2p_stream_t stream_, fv2p_stream_t features_stream_) {
  return fv2p_voxelize_mean_table(points, frame_offsets, total_points, batch, max_frame_points, num_features, range6,
                                  vsize3, max_points, max_voxels, coords, voxel_features, num_points, voxels,
                                  voxel_offsets, cap, status_dev, workspace, workspace_bytes, stream_,
                                  features_stream_, nullptr, 0, nullptr);
}

extern "C" int fv2p_voxelize_mean_table(const float *points, const int32_t *frame_offsets, int64_t total_points,
                                        int batch, int64_t max_frame_points, int num_features,
                                        const float *range6, const float *vsize3, int max_points,
                                        int max_voxels, int32_t *coords, float *voxel_features,
                                        int32_t *num_points, float *voxels, int32_t *voxel_offsets, int64_t cap,
                                        int32_t *status_dev, void *workspace, size_t workspace_bytes,
                                        fv2p_stream_t stream_, fv2p_stream_t features_stream_, void *level0_table,
                                        int64_t table_row_cap, const int32_t *shape3) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FV2P_REQUIRE(!level0_table || (shape3 && table_row_cap >= cap && table_row_cap < (1ll << 26)),
               "voxelize: the level-0 table needs its shape and a row capacity >= cap");
  FV2P_REQUIRE(batch >= 1 && batch <= 4096, "voxelize: batch must be in [1,4096], got %d", batch);
  FV2P_REQUIRE(num_features >= 3 && num_features <= kMaxFeatures,
               "voxelize: num_features must be in [3,%d], got %d", kMaxFeatures, num_features);
  FV2P_REQUIRE(max_points >= 1 && max_points <= 64, "voxelize: max_points must be in [1,64]");
  FV2P_REQUIRE(max_voxels >= 1, "voxelize: max_voxels must be positive");
  FV2P_REQUIRE(total_points >= 0 && total_points < (1ll << 30), "voxelize: total_points out of range");
  FV2P_REQUIRE(max_frame_points >= 0 && max_frame_points <= total_points + 0,
               "voxelize: max_frame_points must not exceed total_points");
  FV2P_REQUIRE(frame_offsets && coords && voxel_features && voxel_offsets && range6 && vsize3,
               "voxelize: null pointer argument");
  FV2P_REQUIRE(total_points == 0 || points, "voxelize: null points");
  VoxGeom g;
  for (int a = 0; a < 3; ++a) {
    g.lo[a] = range6[a];
    g.size[a] = vsize3[a];
    FV2P_REQUIRE(vsize3[a] > 0.0f, "voxelize: voxel size must be positive");
    // grid_size = round((hi - lo) / size) in fp32, round-half-even (voxel_generator.py:25-27)
    g.grid[a] = (int)rintf((range6[3 + a] - range6[a]) / vsize3[a]);
    FV2P_REQUIRE(g.grid[a] >= 1, "voxelize: empty grid along axis %d", a);
  }
  VoxWorkspace w = carve(workspace, total_points, batch, max_frame_points, max_points, cap);
  if (workspace_bytes < w.bytes || !workspace) {
    set_error("voxelize: workspace too small (%zu < %zu)", workspace_bytes, w.bytes);
    return FV2P_ERR_WORKSPACE;
  }
  const int grid = persistent_grid(kGeoCtasPerSm);
  Slot *table0 = static_cast<Slot *>(level0_table);
  const uint32_t t0_slots = table0 ? table_slots_cap(table_row_cap) : 0u;
  Slot empty_slot;
  empty_slot.key = kEmptyKey, empty_slot.val = kValEmpty, empty_slot.aux = 0;
  uint4 empty4;
  static_assert(sizeof(Slot) == sizeof(uint4), "Slot is one 16-byte word");
  memcpy(&empty4, &empty_slot, sizeof(uint4));
  vox_clear_kernel<<<grid, kThreads, 0, stream>>>(w.keys, w.first, w.slots, w.cut, batch, w.state,
                                                  batch * w.max_chunks, w.frame_total, w.scratch,
                                                  reinterpret_cast<uint4 *>(table0), t0_slots, empty4);
  vox_insert_kernel<<<grid, kThreads, 0, stream>>>(points, frame_offsets, batch, w.max_chunks, num_features, g,
                                                   w.keys, w.first, w.slots - 1, w.pslot);
  // 40 registers: six CTAs per SM hold more of the chunks' dependent loads (point -> slot -> first arrival -> key)
  // in flight than the default four
  const int rank_work = batch * w.max_chunks;
  const int rank_cap = persistent_grid(6);
  vox_rank_kernel<<<rank_work < rank_cap ? rank_work : rank_cap, kThreads, 0, stream>>>(frame_offsets, batch, w.max_chunks, w.pslot, w.first, w.keys, w.state,
                                                 w.frame_total, w.scratch, max_voxels, cap, max_points, g,
                                                 voxel_offsets, status_dev, w.slot_vid, w.cut, coords, w.sel, table0,
                                                 t0_slots ? t0_slots - 1 : 0u, shape3 ? shape3[0] : 0,
                                                 shape3 ? shape3[1] : 0, shape3 ? shape3[2] : 0);
  // coords and voxel_offsets are final here; the point selection and the means can leave the caller's chain
  stream = fork_stream(stream, features_stream_);
  vox_select_kernel<<<grid, kThreads, 0, stream>>>(frame_offsets, batch, w.max_chunks, w.pslot, w.slot_vid, w.cut,
                                                   max_points, w.sel);
  vox_reduce_kernel<<<grid, kThreads, 0, stream>>>(points, num_features, voxel_offsets, batch, max_points, w.sel,
                                                   voxel_features, num_points, voxels);
  FV2P_LAUNCH_CHECK("voxelize");
  return FV2P_OK;
}

extern "C" int fv2p_voxel_generate(const float *points, int64_t num_points_in, int num_features,
                                   const float *range6, const float *vsize3, int max_points, int max_voxels,
                                   int32_t *coords, float *voxel_features, int32_t *num_points, float *voxels,
                                   int64_t cap, int32_t *num_voxels_host, void *workspace,
                                   size_t workspace_bytes, fv2p_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FV2P_REQUIRE(num_voxels_host, "voxel_generate: null num_voxels_host");
  FV2P_REQUIRE(workspace && workspace_bytes >= 1024, "voxel_generate: workspace too small");
  // the first 1 KB of the workspace holds the two offset arrays
  int *offs = static_cast<int *>(workspace);
  launch_set_scalar(offs + 0, 0, stream);
  launch_set_scalar(offs + 1, (int)num_points_in, stream);
  int st = fv2p_voxelize_mean(points, offs, num_points_in, 1, num_points_in, num_features, range6, vsize3,
                              max_points, max_voxels, coords, voxel_features, num_points, voxels, offs + 4, cap,
                              nullptr, static_cast<char *>(workspace) + 1024, workspace_bytes - 1024, stream_,
                              nullptr);
  if (st) return st;
  int voff[2] = {0, 0};
  st = cuda_status(cudaMemcpyAsync(voff, offs + 4, sizeof(voff), cudaMemcpyDeviceToHost, stream), "voxel_generate");
  if (st) return st;
  st = cuda_status(cudaStreamSynchronize(stream), "voxel_generate");
  if (st) return st;
  *num_voxels_host = voff[1];
  return FV2P_OK;
}

extern "C" int fv2p_mean_vfe(const float *voxels, const int32_t *num_points, int64_t num_voxels, int max_points,
                             int num_features, float *out, fv2p_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FV2P_REQUIRE(num_voxels >= 0 && max_points >= 1 && num_features >= 1, "mean_vfe: bad sizes");
  if (num_voxels == 0) return FV2P_OK;
  FV2P_REQUIRE(voxels && num_points && out, "mean_vfe: null pointer argument");
  mean_vfe_kernel<<<persistent_grid(kGeoCtasPerSm), kThreads, 0, stream>>>(voxels, num_points, num_voxels, max_points,
                                                             num_features, out);
  FV2P_LAUNCH_CHECK("mean_vfe");
  return FV2P_OK;
}
