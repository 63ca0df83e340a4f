// Consumers of the backbone's sparse outputs that the reference runs on dense grids or by brute force
// (SURVEY.md section 8f ranks 3 and 4), rebuilt on the level's coordinate table:
//
//   voxel -> point 3-NN + interpolation   the front end of ResidualVoxelToPointDecoder.forward
//       (pcdet/models/backbones_3d/pfe/residual_v2p_decoder.py:86-116): voxel centres
//       (pcdet/utils/common_utils.py:76-92), three_nn (pointnet2_batch/src/interpolate_gpu.cu:16-58: every query
//       scans EVERY voxel of its frame, O(P*N)), inverse-distance weights and three_interpolate (:78-100), one
//       python-level loop iteration and four launches per frame.  Here: one pass over the points; a query walks
//       the shells of cells around its own cell in the hash table and stops as soon as no unseen cell can beat
//       (or tie) its third-best distance, so it looks at tens of cells instead of tens of thousands of voxels.
//       Distances use the reference's expression, candidates are ranked by (distance, row) like its strict-<
//       ascending scan, so indices and distances are bit-identical; the few queries far from every voxel fall
//       back to a warp-parallel exact scan of their frame.
//   voxel_query   (pointnet2_stack/src/voxel_query_gpu.cu:10-88) with generate_voxel2pinds
//       (pcdet/utils/spconv_utils.py:13-21): the reference first scatters row ids into a dense [B,Z,Y,X] int32
//       grid (370 MB per frame at stride 1) and then reads (2r+1)^3 cells of it per query; here the same cells are
//       looked up in the O(N) table, same visiting order, same first-nsample rule.
#include <float.h>

#include "common.cuh"

namespace fv2p {
namespace {

__device__ __forceinline__ int live_count(const int *n_dev, int64_t n_cap) {
  int n = n_dev ? *n_dev : (int)n_cap;
  return n < 0 ? 0 : (n > n_cap ? (int)n_cap : n);
}

// first row of every frame: rows are batch-contiguous (collate_batch), so a binary search per frame does it
__global__ void frame_starts_kernel(const int4 *__restrict__ indices, const int *n_dev, int64_t n_cap, int batch,
                                    int *starts) {
  const int n = live_count(n_dev, n_cap);
  for (int b = blockIdx.x * blockDim.x + threadIdx.x; b <= batch; b += gridDim.x * blockDim.x) {
    int lo = 0, hi = n;  // first row with batch index >= b
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (__ldg(&indices[mid]).x < b) lo = mid + 1;
      else hi = mid;
    }
    starts[b] = lo;
  }
}

struct Top3 {
  float d[3];
  int i[3];
};

// keeps the three smallest (distance, row) pairs in lexicographic order: exactly what the reference's ascending
// scan with strict < retains (interpolate_gpu.cu:40-56), whatever the visiting order
__device__ __forceinline__ void top3_insert(Top3 &t, float d, int i) {
  if (d < t.d[2] || (d == t.d[2] && i < t.i[2])) {
    if (d < t.d[1] || (d == t.d[1] && i < t.i[1])) {
      t.d[2] = t.d[1], t.i[2] = t.i[1];
      if (d < t.d[0] || (d == t.d[0] && i < t.i[0])) {
        t.d[1] = t.d[0], t.i[1] = t.i[0];
        t.d[0] = d, t.i[0] = i;
      } else {
        t.d[1] = d, t.i[1] = i;
      }
    } else {
      t.d[2] = d, t.i[2] = i;
    }
  }
}

struct NnGeom {
  float vs[3], lo[3];  // x, y, z: voxel size (already times the level's stride) and range minimum
  int D, H, W;
};

// get_voxel_centers (common_utils.py:87-91): (idx + 0.5) * voxel_size + range_min, three separately rounded ops
__device__ __forceinline__ float centre(int idx, float vs, float lo) {
  return __fadd_rn(__fmul_rn(__fadd_rn((float)idx, 0.5f), vs), lo);
}

// interpolate_gpu.cu:44 / voxel_query_gpu.cu:63 as nvcc compiles them (SASS of the reference's kernels built for
// sm_100: FMUL dy*dy, FFMA dx*dx + ., FFMA dz*dz + .).  Spelled with intrinsics: left to the compiler, a loop-invariant
// square gets hoisted and added unfused, which changes the last bit of the distance.
__device__ __forceinline__ float dist2_ref(float ux, float uy, float uz, float x, float y, float z) {
  const float dx = __fsub_rn(ux, x), dy = __fsub_rn(uy, y), dz = __fsub_rn(uz, z);
  return __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
}

constexpr int kInitIdx = 0x7FFFFFFF;

__global__ void __launch_bounds__(kThreads)
three_nn_hash_kernel(const float4 *__restrict__ points, int64_t n_points, const Slot *__restrict__ table,
                     uint32_t tmask, NnGeom g, int batch, int rmax, float *d2_out, int *idx_out, int *leftover,
                     int *n_leftover) {
  const float vs_min = fminf(g.vs[0], fminf(g.vs[1], g.vs[2]));
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n_points;
       p += (int64_t)gridDim.x * blockDim.x) {
    const float4 q = __ldg(&points[p]);  // (batch, x, y, z)
    const int b = (int)q.x;
    Top3 t;
    t.d[0] = t.d[1] = t.d[2] = FLT_MAX;
    t.i[0] = t.i[1] = t.i[2] = kInitIdx;
    bool resolved = false;
    if (b >= 0 && b < batch) {
      int cx = (int)floorf((q.y - g.lo[0]) / g.vs[0]), cy = (int)floorf((q.z - g.lo[1]) / g.vs[1]),
          cz = (int)floorf((q.w - g.lo[2]) / g.vs[2]);
      cx = min(max(cx, 0), g.W - 1), cy = min(max(cy, 0), g.H - 1), cz = min(max(cz, 0), g.D - 1);
      for (int r = 0; r <= rmax && !resolved; ++r) {
        const int z0 = max(cz - r, 0), z1 = min(cz + r, g.D - 1);
        const int y0 = max(cy - r, 0), y1 = min(cy + r, g.H - 1);
        for (int z = z0; z <= z1; ++z) {
          const bool zface = (z == cz - r) || (z == cz + r);
          for (int y = y0; y <= y1; ++y) {
            const bool face = zface || (y == cy - r) || (y == cy + r);
            // on a face of the shell every x of the row belongs to it, otherwise only the two ends
            const int step = (face || r == 0) ? 1 : 2 * r;
            for (int x = cx - r; x <= cx + r; x += step) {
              if (x < 0 || x >= g.W) continue;
              const int v = slot_lookup(table, tmask, voxel_key(b, z, y, x, g.D, g.H, g.W));
              if (v >= 0) continue;
              const float d = dist2_ref(q.y, q.z, q.w, centre(x, g.vs[0], g.lo[0]), centre(y, g.vs[1], g.lo[1]),
                                        centre(z, g.vs[2], g.lo[2]));
              top3_insert(t, d, ~v);
            }
          }
        }
        // every cell outside the cube of radius r is at least (r + 0.5) cells away along some axis
        const float reach = ((float)r + 0.5f) * vs_min;
        resolved = t.d[2] < reach * reach * 0.9999f;
      }
    } else {
      resolved = true;  // a point of no frame sees no voxel, like the reference's empty scan
    }
    if (resolved) {
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        d2_out[p * 3 + j] = t.d[j];
        idx_out[p * 3 + j] = t.i[j];
      }
    } else {
      leftover[atomicAdd(n_leftover, 1)] = (int)p;
    }
  }
}

// Exact scan of the query's frame, one warp per leftover query (lanes stride over the rows, then three rounds of
// "smallest head over the lanes").
__global__ void __launch_bounds__(kThreads)
three_nn_scan_kernel(const float4 *__restrict__ points, const int4 *__restrict__ indices,
                     const int *__restrict__ starts, NnGeom g, const int *__restrict__ leftover,
                     const int *__restrict__ n_leftover, float *d2_out, int *idx_out) {
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const int total = *n_leftover;
  for (int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < total; w += warps) {
    const int p = leftover[w];
    const float4 q = __ldg(&points[p]);
    const int b = (int)q.x;
    Top3 t;
    t.d[0] = t.d[1] = t.d[2] = FLT_MAX;
    t.i[0] = t.i[1] = t.i[2] = kInitIdx;
    for (int row = starts[b] + lane; row < starts[b + 1]; row += 32) {
      const int4 c = __ldg(&indices[row]);  // (b, z, y, x)
      const float d = dist2_ref(q.y, q.z, q.w, centre(c.w, g.vs[0], g.lo[0]), centre(c.z, g.vs[1], g.lo[1]),
                                centre(c.y, g.vs[2], g.lo[2]));
      top3_insert(t, d, row);
    }
    int head = 0;
    for (int j = 0; j < 3; ++j) {
      float d = head < 3 ? t.d[head] : FLT_MAX;
      int i = head < 3 ? t.i[head] : kInitIdx;
      float bd = d;
      int bi = i;
#pragma unroll
      for (int s = 16; s > 0; s >>= 1) {
        const float od = __shfl_xor_sync(0xFFFFFFFFu, bd, s);
        const int oi = __shfl_xor_sync(0xFFFFFFFFu, bi, s);
        if (od < bd || (od == bd && oi < bi)) bd = od, bi = oi;
      }
      if (d == bd && i == bi && bi != kInitIdx) ++head;  // this lane's head was taken
      if (lane == 0) {
        d2_out[(size_t)p * 3 + j] = bd;
        idx_out[(size_t)p * 3 + j] = bi;
      }
    }
  }
}

// ThreeNN.forward's sqrt + the frame-local index of the reference (it searches one frame at a time), then
// top3_interpolate (pointnet2_utils.py:315-320): w = (1 / (dist + 1e-8)) / sum, out = sum_j w_j * feats[idx_j].
// Fewer than three voxels in the frame: the reference leaves best = 1e40 -> (float) inf and index 0.
__global__ void __launch_bounds__(kThreads)
three_nn_finish_kernel(const float4 *__restrict__ points, int64_t n_points, const int *__restrict__ starts, int batch,
                       const float *__restrict__ d2, const int *__restrict__ idx_global, float *dist_out,
                       int *idx_out, const float *__restrict__ feats, int channels, float *out) {
  const int lane = threadIdx.x & 31;
  const int64_t n_round = (n_points + 31) & ~int64_t(31);  // whole warps iterate together (shuffles below)
  const int last_row = max(starts[batch] - 1, 0);
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n_round;
       p += (int64_t)gridDim.x * blockDim.x) {
    const bool live = p < n_points;
    float w[3] = {0.f, 0.f, 0.f};
    int gi[3] = {0, 0, 0};
    if (live) {
      const int b = (int)__ldg(&points[p]).x;
      const int first = (b >= 0 && b < batch) ? starts[b] : 0;
      float dist[3];
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const int i = idx_global[p * 3 + j];
        const bool have = i != kInitIdx;
        dist[j] = have ? __fsqrt_rn(d2[p * 3 + j]) : __int_as_float(0x7F800000);
        gi[j] = have ? i : first;
        if (dist_out) dist_out[p * 3 + j] = dist[j];
        if (idx_out) idx_out[p * 3 + j] = gi[j] - first;
        w[j] = __fdiv_rn(1.0f, __fadd_rn(dist[j], 1e-8f));
        gi[j] = min(gi[j], last_row);  // a frame without voxels: keep the read inside the feature matrix
      }
      const float norm = __fadd_rn(__fadd_rn(w[0], w[1]), w[2]);
#pragma unroll
      for (int j = 0; j < 3; ++j) w[j] = __fdiv_rn(w[j], norm);
    }
    if (!out) continue;
    // the warp interpolates its 32 points one after the other, lanes across the channels (coalesced rows)
    const int64_t p0 = p - lane;
    for (int s = 0; s < 32; ++s) {
      const float w0 = __shfl_sync(0xFFFFFFFFu, w[0], s), w1 = __shfl_sync(0xFFFFFFFFu, w[1], s),
                  w2 = __shfl_sync(0xFFFFFFFFu, w[2], s);
      const int g0 = __shfl_sync(0xFFFFFFFFu, gi[0], s), g1 = __shfl_sync(0xFFFFFFFFu, gi[1], s),
                g2 = __shfl_sync(0xFFFFFFFFu, gi[2], s);
      if (p0 + s >= n_points) break;
      const float *f0 = feats + (size_t)g0 * channels, *f1 = feats + (size_t)g1 * channels,
                  *f2 = feats + (size_t)g2 * channels;
      float *o = out + (size_t)(p0 + s) * channels;
      for (int c = lane; c < channels; c += 32)  // interpolate_gpu.cu:98-99
        o[c] = w0 * __ldg(f0 + c) + w1 * __ldg(f1 + c) + w2 * __ldg(f2 + c);
    }
  }
}

// voxel_query_gpu.cu:10-88 with the dense point_indices grid replaced by the table
__global__ void __launch_bounds__(kThreads)
voxel_query_kernel(int64_t M, int R1, int R2, int R3, int nsample, float radius, int z_range, int y_range,
                   int x_range, const float *__restrict__ new_xyz, const float *__restrict__ xyz,
                   const int4 *__restrict__ new_coords, const Slot *__restrict__ table, uint32_t tmask, int *idx) {
  for (int64_t pt = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; pt < M; pt += (int64_t)gridDim.x * blockDim.x) {
    int *out = idx + pt * nsample;
    const float radius2 = radius * radius;
    const float new_x = new_xyz[pt * 3 + 0], new_y = new_xyz[pt * 3 + 1], new_z = new_xyz[pt * 3 + 2];
    const int4 nc = __ldg(&new_coords[pt]);  // (batch, z, y, x)
    int cnt = 0;
    for (int l = 0; l < nsample; ++l) out[l] = 0;  // the reference's zero-initialised output
    for (int dz = -z_range; dz <= z_range; ++dz) {
      const int z = nc.y + dz;
      if (z < 0 || z >= R1) continue;
      for (int dy = -y_range; dy <= y_range; ++dy) {
        const int y = nc.z + dy;
        if (y < 0 || y >= R2) continue;
        for (int dx = -x_range; dx <= x_range; ++dx) {
          const int x = nc.w + dx;
          if (x < 0 || x >= R3) continue;
          const int v = slot_lookup(table, tmask, voxel_key(nc.x, z, y, x, R1, R2, R3));
          if (v >= 0) continue;
          const int neighbor_idx = ~v;
          const float x_per = xyz[(size_t)neighbor_idx * 3 + 0], y_per = xyz[(size_t)neighbor_idx * 3 + 1],
                      z_per = xyz[(size_t)neighbor_idx * 3 + 2];
          const float dist2 = dist2_ref(x_per, y_per, z_per, new_x, new_y, new_z);
          if (dist2 > radius2) continue;
          if (cnt < nsample) {
            if (cnt == 0)
              for (int l = 0; l < nsample; ++l) out[l] = neighbor_idx;
            out[cnt] = neighbor_idx;
            ++cnt;
          }
        }
      }
    }
    if (cnt == 0) out[0] = -1;
  }
}

struct NnWs {
  int *starts, *leftover, *n_leftover, *idx_global;
  float *d2;
  size_t bytes;
};

NnWs carve_nn(void *ws, int batch, int64_t n_points) {
  NnWs w;
  Carver c(ws);
  const size_t p = (size_t)(n_points > 0 ? n_points : 1);
  w.starts = c.take<int>(batch + 2);
  w.n_leftover = c.take<int>(4);
  w.leftover = c.take<int>(p);
  w.idx_global = c.take<int>(p * 3);
  w.d2 = c.take<float>(p * 3);
  w.bytes = c.used + 256;
  return w;
}

}  // namespace
}  // namespace fv2p

using namespace fv2p;

extern "C" size_t fv2p_voxel_three_nn_workspace_bytes(int batch, int64_t n_points) {
  if (batch < 1 || n_points < 0) return 0;
  return carve_nn(nullptr, batch, n_points).bytes;
}

extern "C" int fv2p_voxel_three_nn(const float *point_coords, int64_t n_points, const int32_t *voxel_indices,
                                   int64_t n_cap, const int32_t *n_dev, int batch, const int32_t *shape3,
                                   const void *table, int64_t table_row_cap, const float *voxel_size3,
                                   const float *range_min3, float *dist, int32_t *idx, const float *features,
                                   int channels, float *interpolated, void *workspace, size_t workspace_bytes,
                                   fv2p_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FV2P_REQUIRE(shape3 && voxel_size3 && range_min3 && table, "voxel_three_nn: null argument");
  FV2P_REQUIRE(batch >= 1 && n_points >= 0 && n_points < (1ll << 31) / 3 && n_cap >= 0 && n_cap < (1ll << 26),
               "voxel_three_nn: bad sizes");
  FV2P_REQUIRE(!interpolated || (features && channels >= 1), "voxel_three_nn: features needed for interpolation");
  if (n_points == 0) return FV2P_OK;
  FV2P_REQUIRE(point_coords && voxel_indices, "voxel_three_nn: null pointer argument");
  NnWs w = carve_nn(workspace, batch, n_points);
  if (!workspace || workspace_bytes < w.bytes) {
    set_error("voxel_three_nn: workspace too small (%zu < %zu)", workspace_bytes, w.bytes);
    return FV2P_ERR_WORKSPACE;
  }
  NnGeom g;
  for (int a = 0; a < 3; ++a) {
    g.vs[a] = voxel_size3[a];
    g.lo[a] = range_min3[a];
    FV2P_REQUIRE(voxel_size3[a] > 0.0f, "voxel_three_nn: voxel size must be positive");
  }
  g.D = shape3[0], g.H = shape3[1], g.W = shape3[2];
  const int grid = persistent_grid();
  const int4 *ind4 = reinterpret_cast<const int4 *>(voxel_indices);
  const float4 *pts = reinterpret_cast<const float4 *>(point_coords);
  int st = cuda_status(cudaMemsetAsync(w.n_leftover, 0, 16, stream), "voxel_three_nn");
  if (st) return st;
  frame_starts_kernel<<<1, 128, 0, stream>>>(ind4, n_dev, n_cap, batch, w.starts);
  // shells up to radius 6 (13^3 cells) are searched in the table; beyond that the exact frame scan is cheaper
  three_nn_hash_kernel<<<grid, kThreads, 0, stream>>>(pts, n_points, static_cast<const Slot *>(table),
                                                      table_slots_cap(table_row_cap) - 1, g, batch, 6, w.d2,
                                                      w.idx_global, w.leftover, w.n_leftover);
  three_nn_scan_kernel<<<grid, kThreads, 0, stream>>>(pts, ind4, w.starts, g, w.leftover, w.n_leftover, w.d2,
                                                      w.idx_global);
  three_nn_finish_kernel<<<grid, kThreads, 0, stream>>>(pts, n_points, w.starts, batch, w.d2, w.idx_global, dist, idx,
                                                        features, channels, interpolated);
  FV2P_LAUNCH_CHECK("voxel_three_nn");
  return FV2P_OK;
}

extern "C" int fv2p_voxel_query(int64_t m, const int32_t *shape3, int nsample, float radius, const int32_t *range3,
                                const float *new_xyz, const float *xyz, const int32_t *new_coords, const void *table,
                                int64_t table_row_cap, int32_t *idx, fv2p_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FV2P_REQUIRE(shape3 && range3 && table, "voxel_query: null argument");
  FV2P_REQUIRE(m >= 0 && nsample >= 1, "voxel_query: bad sizes");
  if (m == 0) return FV2P_OK;
  FV2P_REQUIRE(new_xyz && xyz && new_coords && idx, "voxel_query: null pointer argument");
  voxel_query_kernel<<<persistent_grid(), kThreads, 0, stream>>>(
      m, shape3[0], shape3[1], shape3[2], nsample, radius, range3[0], range3[1], range3[2], new_xyz, xyz,
      reinterpret_cast<const int4 *>(new_coords), static_cast<const Slot *>(table),
      table_slots_cap(table_row_cap) - 1, idx);
  FV2P_LAUNCH_CHECK("voxel_query");
  return FV2P_OK;
}
