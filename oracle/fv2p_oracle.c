/*
 * fv2p_oracle.c -- CPU restatement of the reference hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the parity oracle for the B200 library: a plain-C restatement of the
 * algorithms of jialeli1/From-Voxel-to-Point's voxelizer, MeanVFE, spconv rulebook
 * builders and gather-GEMM-scatter convolution.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference leg may call it.  The product path
 * (from-voxel-to-point_b200/) never links, imports or executes it.
 *
 * Parity status: PINNED.  The reference ships no golden vectors for this path
 * (SURVEY.md section 4), so the oracle is pinned against the reference itself:
 * tests/golden/make_golden.py imports the reference's numba voxelizer and its compiled
 * sparse_conv_ext (oracle/build_ref.py -> oracle/_ref/) in the build container and
 * commits their outputs as tests/golden/.npz; tests/test_oracle.py checks every function
 * below against those fixtures bit-for-bit (coordinates / rulebooks) or to 1e-6 (features).
 *
 * Each function cites the reference file:line (relative to /root/reference) it follows.
 * Written from scratch; data layouts follow the reference's tensors:
 *   indices      [N,4] int32  (batch, z, y, x)
 *   indice pairs [K,2,N] int32, -1 padded ; indice num [K] int32
 *   features     [N,C] fp32 row-major ; filters [K,Cin,Cout] fp32
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------
 * Voxelizer: pcdet/datasets/processor/voxel_generator.py:75-133 (points_to_voxel) and
 * :136-207 (_points_to_voxel_reverse_kernel).  fp32 arithmetic throughout, IEEE division.
 *   grid_size = round((range[3:]-range[:3]) / voxel_size)              (:178-181)
 *   c = floor((p - range_lo) / voxel_size); reject if c<0 || c>=grid   (:188-191)
 *   coords stored reversed (z,y,x)                                      (:192)
 *   new voxel id = running count; BREAK when count >= max_voxels        (:196-199)
 *   keep a point iff the voxel holds < max_points                        (:203-206)
 * Outputs are caller-allocated at [max_voxels,...] and pre-zeroed here like :110-115.
 * Returns voxel_num.
 * ---------------------------------------------------------------------------------------- */
ORC_API int orc_voxelize(const float *points, int num_points, int num_features,
                         const float *range6, const float *vsize3, int max_points,
                         int max_voxels, float *voxels, int32_t *coors, int32_t *num_per_voxel) {
  int32_t grid[3];
  for (int a = 0; a < 3; ++a) {
    float g = (range6[3 + a] - range6[a]) / vsize3[a];
    grid[a] = (int32_t)rintf(g); /* np.round = round-half-even = rintf in default mode */
  }
  /* dense (z,y,x) lookup, -1 = empty: voxel_generator.py:114 */
  size_t cells = (size_t)grid[0] * grid[1] * grid[2];
  int32_t *lookup = (int32_t *)malloc(cells * sizeof(int32_t));
  if (!lookup) return -1;
  memset(lookup, 0xFF, cells * sizeof(int32_t));
  memset(num_per_voxel, 0, (size_t)max_voxels * sizeof(int32_t));
  memset(coors, 0, (size_t)max_voxels * 3 * sizeof(int32_t));
  memset(voxels, 0, (size_t)max_voxels * max_points * num_features * sizeof(float));

  int voxel_num = 0;
  for (int i = 0; i < num_points; ++i) {
    const float *p = points + (size_t)i * num_features;
    int32_t zyx[3];
    int ok = 1;
    for (int a = 0; a < 3; ++a) {
      volatile float q = (p[a] - range6[a]) / vsize3[a]; /* volatile: no contraction */
      float c = floorf(q);
      if (c < 0.0f || c >= (float)grid[a]) { ok = 0; break; }
      zyx[2 - a] = (int32_t)c;
    }
    if (!ok) continue;
    size_t cell = ((size_t)zyx[0] * grid[1] + zyx[1]) * grid[0] + zyx[2];
    int32_t vid = lookup[cell];
    if (vid == -1) {
      if (voxel_num >= max_voxels) break; /* drops every later point, :198-199 */
      vid = voxel_num++;
      lookup[cell] = vid;
      coors[vid * 3 + 0] = zyx[0];
      coors[vid * 3 + 1] = zyx[1];
      coors[vid * 3 + 2] = zyx[2];
    }
    int32_t cnt = num_per_voxel[vid];
    if (cnt < max_points) {
      memcpy(voxels + ((size_t)vid * max_points + cnt) * num_features, p,
             (size_t)num_features * sizeof(float));
      num_per_voxel[vid] = cnt + 1;
    }
  }
  free(lookup);
  return voxel_num;
}

/* MeanVFE: pcdet/models/backbones_3d/vfe/mean_vfe.py:26-28.
 * mean = sum_t voxels[v,t,:] / max(num_points[v], 1); the zero padding takes part in the sum. */
ORC_API void orc_mean_vfe(const float *voxels, const int32_t *num_per_voxel, int num_voxels,
                          int max_points, int num_features, float *out) {
  for (int v = 0; v < num_voxels; ++v) {
    float denom = (float)(num_per_voxel[v] < 1 ? 1 : num_per_voxel[v]);
    for (int f = 0; f < num_features; ++f) {
      float s = 0.0f;
      for (int t = 0; t < max_points; ++t)
        s += voxels[((size_t)v * max_points + t) * num_features + f];
      out[(size_t)v * num_features + f] = s / denom;
    }
  }
}

/* ------------------------------------------------------------------------------------------
 * Candidate enumeration: pcdet/ops/spconv/include/spconv/geometry.h:25-85 (getValidOutPos).
 * For one input coordinate, lists every output coordinate it contributes to together with the
 * kernel-offset index, in the reference's order: per axis from `upper` downwards, LAST axis
 * fastest.  lower/upper use C integer division (truncation), geometry.h:41-45.
 * `out` receives up to kvol records of 4 ints (z,y,x,offset).  Returns the number of VALID
 * records; `raw_index` (optional, kvol ints) receives each valid record's raw enumeration
 * index (position in the unfiltered enumeration) -- used by the CUDA tie-break, not by the
 * reference.
 * The reference walks a mixed-radix counter (geometry.h:76-83); decoding the loop index into
 * digits with the last axis fastest is the same sequence.
 * ---------------------------------------------------------------------------------------- */
ORC_API int orc_valid_out_pos(const int32_t *in_pos, const int32_t *ksize, const int32_t *stride,
                              const int32_t *pad, const int32_t *dil, const int32_t *out_shape,
                              int32_t *out, int32_t *raw_index) {
  int32_t lo[3], hi[3], cnt[3];
  int32_t total = 1;
  for (int a = 0; a < 3; ++a) {
    lo[a] = (in_pos[a] - (ksize[a] - 1) * dil[a] - 1 + stride[a] + pad[a]) / stride[a];
    hi[a] = (in_pos[a] + pad[a]) / stride[a];
    cnt[a] = (hi[a] - lo[a]) / dil[a] + 1;
    total *= cnt[a];
  }
  int n_valid = 0;
  for (int32_t e = 0; e < total; ++e) {
    int32_t digit[3];
    int32_t rest = e;
    for (int a = 2; a >= 0; --a) { digit[a] = rest % cnt[a]; rest /= cnt[a]; }
    int ok = 1;
    int32_t offset = 0, mult = 1;
    int32_t rec[3];
    for (int a = 2; a >= 0; --a) {
      int32_t v = hi[a] - digit[a] * dil[a];
      rec[a] = v;
      if (v < 0 || v > out_shape[a] - 1) ok = 0;
      /* geometry.h:69: offset += m * (in - v*stride + pad) / dilation  (left-to-right) */
      offset += mult * (in_pos[a] - v * stride[a] + pad[a]) / dil[a];
      mult *= ksize[a];
    }
    if (ok) {
      out[n_valid * 4 + 0] = rec[0];
      out[n_valid * 4 + 1] = rec[1];
      out[n_valid * 4 + 2] = rec[2];
      out[n_valid * 4 + 3] = offset;
      if (raw_index) raw_index[n_valid] = e;
      ++n_valid;
    }
  }
  return n_valid;
}

static size_t lin_index(int32_t b, const int32_t *zyx, const int32_t *shape) {
  /* tensorview.h:453-464 rowArrayIdx + spatialVolume*batch (geometry.h:179-180); 64-bit here so
   * the oracle itself is not limited to 23 frames. */
  size_t vol = (size_t)shape[0] * shape[1] * shape[2];
  return (size_t)b * vol + ((size_t)zyx[0] * shape[1] + zyx[1]) * shape[2] + zyx[2];
}

/* Submanifold rulebook: geometry.h:248-297 (getIndicePairsSubM) under spconv_ops.h:76-80
 * (stride forced to 1, padding forced to ksize/2) with the allocations of spconv_ops.h:55-62.
 * pairs [K,2,N] is filled with -1 first, num [K] with 0.  Returns N (geometry.h:296). */
ORC_API int orc_rulebook_subm(const int32_t *indices, int n, int batch, const int32_t *shape,
                              const int32_t *ksize, const int32_t *dil, int32_t *pairs,
                              int32_t *num) {
  int kvol = ksize[0] * ksize[1] * ksize[2];
  int32_t stride[3] = {1, 1, 1};
  int32_t pad[3] = {ksize[0] / 2, ksize[1] / 2, ksize[2] / 2};
  size_t cells = (size_t)batch * shape[0] * shape[1] * shape[2];
  int32_t *grid = (int32_t *)malloc(cells * sizeof(int32_t));
  if (!grid) return -1;
  memset(grid, 0xFF, cells * sizeof(int32_t));
  for (size_t i = 0; i < (size_t)kvol * 2 * n; ++i) pairs[i] = -1;
  memset(num, 0, kvol * sizeof(int32_t));
  for (int j = 0; j < n; ++j) /* geometry.h:276-280: later duplicates overwrite */
    grid[lin_index(indices[j * 4], indices + j * 4 + 1, shape)] = j;
  int32_t *cand = (int32_t *)malloc((size_t)kvol * 4 * sizeof(int32_t));
  for (int j = 0; j < n; ++j) {
    int nv = orc_valid_out_pos(indices + j * 4 + 1, ksize, stride, pad, dil, shape, cand, NULL);
    for (int e = 0; e < nv; ++e) {
      int32_t k = cand[e * 4 + 3];
      int32_t hit = grid[lin_index(indices[j * 4], cand + e * 4, shape)];
      if (hit > -1) {
        int32_t t = num[k]++;
        pairs[((size_t)k * 2 + 0) * n + t] = j;
        pairs[((size_t)k * 2 + 1) * n + t] = hit;
      }
    }
  }
  free(cand);
  free(grid);
  return n;
}

/* Strided (regular sparse) rulebook: geometry.h:145-194 (getIndicePairsConv) with the
 * allocations of spconv_ops.h:55-62,106-108.  Output rows are created in first-touch order.
 * out_indices must hold n*kvol rows of 4 ints (zero-filled here like spconv_ops.h:106-108).
 * Returns the number of active outputs. */
ORC_API int orc_rulebook_conv(const int32_t *indices, int n, int batch, const int32_t *out_shape,
                              const int32_t *ksize, const int32_t *stride, const int32_t *pad,
                              const int32_t *dil, int32_t *out_indices, int32_t *pairs,
                              int32_t *num) {
  int kvol = ksize[0] * ksize[1] * ksize[2];
  size_t cells = (size_t)batch * out_shape[0] * out_shape[1] * out_shape[2];
  int32_t *grid = (int32_t *)malloc(cells * sizeof(int32_t));
  if (!grid) return -1;
  memset(grid, 0xFF, cells * sizeof(int32_t));
  for (size_t i = 0; i < (size_t)kvol * 2 * n; ++i) pairs[i] = -1;
  memset(num, 0, kvol * sizeof(int32_t));
  memset(out_indices, 0, (size_t)n * kvol * 4 * sizeof(int32_t));
  int32_t *cand = (int32_t *)malloc((size_t)kvol * 4 * sizeof(int32_t));
  int n_out = 0;
  for (int j = 0; j < n; ++j) {
    int32_t b = indices[j * 4];
    int nv = orc_valid_out_pos(indices + j * 4 + 1, ksize, stride, pad, dil, out_shape, cand, NULL);
    for (int e = 0; e < nv; ++e) {
      int32_t k = cand[e * 4 + 3];
      size_t cell = lin_index(b, cand + e * 4, out_shape);
      if (grid[cell] == -1) { /* first touch creates the output row, geometry.h:181-187 */
        out_indices[n_out * 4 + 0] = b;
        out_indices[n_out * 4 + 1] = cand[e * 4 + 0];
        out_indices[n_out * 4 + 2] = cand[e * 4 + 1];
        out_indices[n_out * 4 + 3] = cand[e * 4 + 2];
        grid[cell] = n_out++;
      }
      int32_t t = num[k]++;
      pairs[((size_t)k * 2 + 0) * n + t] = j;
      pairs[((size_t)k * 2 + 1) * n + t] = grid[cell];
    }
  }
  free(cand);
  free(grid);
  return n_out;
}

/* ------------------------------------------------------------------------------------------
 * Convolution: spconv_ops.h:261-362 (indiceConv<float>) with the CPU gather / scatter-add of
 * src/reordering.cc:21-50.  out = 0; subM: out = X*W[centre] with centre = first argmax of
 * num (spconv_ops.h:272-277,300-304); then for k ascending (skip empty, skip centre if subM):
 * gather rows, dense [nHot,Cin]x[Cin,Cout] product, scatter-add.  `pair_stride` is the third
 * dimension of the pair tensor (N of the rulebook's input set).  The dense product is
 * torch::mm_out (MKL sgemm) in the reference; here a plain fp32 loop nest (c ascending).
 * ---------------------------------------------------------------------------------------- */
static void dense_mm(const float *a, const float *w, float *c, int rows, int cin, int cout) {
  for (int r = 0; r < rows; ++r) {
    float *cr = c + (size_t)r * cout;
    for (int o = 0; o < cout; ++o) cr[o] = 0.0f;
    const float *ar = a + (size_t)r * cin;
    for (int i = 0; i < cin; ++i) {
      float av = ar[i];
      const float *wr = w + (size_t)i * cout;
      for (int o = 0; o < cout; ++o) cr[o] += av * wr[o];
    }
  }
}

ORC_API int orc_indice_conv(const float *features, const float *filters, const int32_t *pairs,
                            const int32_t *num, int pair_stride, int n_in, int n_out, int kvol,
                            int cin, int cout, int inverse, int subm, float *out) {
  (void)n_in;
  int centre = 0, max_hot = 0;
  for (int k = 0; k < kvol; ++k)
    if (num[k] > max_hot) { max_hot = num[k]; centre = k; }
  memset(out, 0, (size_t)n_out * cout * sizeof(float));
  if (subm) dense_mm(features, filters + (size_t)centre * cin * cout, out, n_out, cin, cout);
  float *inbuf = (float *)malloc((size_t)(max_hot > 0 ? max_hot : 1) * cin * sizeof(float));
  float *outbuf = (float *)malloc((size_t)(max_hot > 0 ? max_hot : 1) * cout * sizeof(float));
  if (!inbuf || !outbuf) return -1;
  for (int k = 0; k < kvol; ++k) {
    int hot = num[k];
    if (hot <= 0 || (subm && k == centre)) continue;
    const int32_t *src = pairs + ((size_t)k * 2 + (inverse ? 1 : 0)) * pair_stride;
    const int32_t *dst = pairs + ((size_t)k * 2 + (inverse ? 0 : 1)) * pair_stride;
    for (int t = 0; t < hot; ++t)
      memcpy(inbuf + (size_t)t * cin, features + (size_t)src[t] * cin, cin * sizeof(float));
    dense_mm(inbuf, filters + (size_t)k * cin * cout, outbuf, hot, cin, cout);
    for (int t = 0; t < hot; ++t) {
      float *o = out + (size_t)dst[t] * cout;
      const float *b = outbuf + (size_t)t * cout;
      for (int c = 0; c < cout; ++c) o[c] += b[c];
    }
  }
  free(inbuf);
  free(outbuf);
  return 0;
}

/* conv.py:223-224 bias add, nn.BatchNorm1d eval (spconv_backbone.py:75: eps=1e-3), optional
 * residual add (spconv_backbone.py:65) and ReLU (:66 / :27), in that order, in place.
 * Any pointer may be NULL to skip that step.  BN follows torch's eval formula
 * y = (x-mean)*rsqrt(var+eps)*gamma + beta evaluated in fp32. */
ORC_API void orc_bias_bn_res_relu(float *x, int rows, int ch, const float *bias,
                                  const float *gamma, const float *beta, const float *mean,
                                  const float *var, float eps, const float *residual, int relu) {
  for (int r = 0; r < rows; ++r)
    for (int c = 0; c < ch; ++c) {
      float v = x[(size_t)r * ch + c];
      if (bias) v += bias[c];
      if (gamma) {
        float inv = 1.0f / sqrtf(var[c] + eps);
        v = (v - mean[c]) * inv * gamma[c] + beta[c];
      }
      if (residual) v += residual[(size_t)r * ch + c];
      if (relu && v < 0.0f) v = 0.0f;
      x[(size_t)r * ch + c] = v;
    }
}
