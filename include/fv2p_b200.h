/*
 * fv2p_b200.h -- C ABI of libfv2p_b200.so: B200 (sm_100a) voxelization + sparse 3D convolution.
 *
 * This is the drop-in boundary for the hot path of jialeli1/From-Voxel-to-Point.  It replaces the
 * reference's pybind11 module `sparse_conv_ext` (pcdet/ops/spconv/src/all.cc:22-71) for the entries
 * the path uses, and the numba voxelizer + MeanVFE, with plain `extern "C"` symbols: raw device
 * pointers, sizes and a cudaStream_t.  No torch types, no pybind.  Citations are relative to
 * /root/reference.
 *
 * Conventions
 *  - Every pointer named *_dev or documented "device" is a CUDA device pointer; small geometry arrays
 *    (shape3, ksize3, ...) are HOST pointers read before the call returns.
 *  - The library never allocates device memory.  The caller passes a workspace whose size comes from
 *    the matching *_workspace_bytes() function, and output buffers sized for the stated capacity.
 *  - Calls are stream-ordered on `stream` and never synchronise, except the two reference-shaped
 *    entries that must return a host count (fv2p_get_indice_pairs_3d, fv2p_voxel_generate).
 *  - Row counts that depend on the data live in device scalars (int32).  A kernel that would exceed
 *    a caller-stated capacity sets a bit in `status_dev` (if given) instead of writing out of bounds.
 *  - Return value: 0 = ok, <0 = FV2P_ERR_*, >0 = cudaError_t.  fv2p_last_error() gives the message
 *    (thread-local).  There is no CPU fallback: the library only does anything on an sm_100 device.
 *  - Data layouts are the reference's: indices [N,4] int32 (batch,z,y,x); indice pairs [K,2,N] int32
 *    padded with -1; pair counts [K] int32; features [N,C] row-major; filters [kD,kH,kW,Cin,Cout].
 *    K = kD*kH*kW <= FV2P_MAX_KVOL.
 *  - The neighbour map `nbr` [K, nbr_stride] int32 is this library's own conv operand: nbr[k][i] is
 *    the INPUT row feeding OUTPUT row i through kernel offset k, or -1.  It carries the same
 *    information as the pair tensor, output-major.
 *  - Coordinates must be UNIQUE (one row per voxel), which is what the voxelizer and every rulebook of this library
 *    produce.  The reference tolerates duplicate input coordinates in a particular way (its grid keeps the last
 *    duplicate, geometry.h:276-280, while its pair lists still carry one pair per duplicate, which the gather-scatter
 *    loop then accumulates); the output-major map holds ONE input row per (offset, output) cell, so with duplicates
 *    the convolution result would differ from the reference's.  Duplicates are the caller's error here.
 */
#ifndef FV2P_B200_H_
#define FV2P_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FV2P_ABI_VERSION 2
#define FV2P_MAX_KVOL 32

#define FV2P_OK 0
#define FV2P_ERR_INVALID (-1)     /* bad argument (reference: TV_ASSERT_INVALID_ARG -> ValueError) */
#define FV2P_ERR_WORKSPACE (-2)   /* workspace too small */
#define FV2P_ERR_UNSUPPORTED (-3) /* valid in the reference, not built here (e.g. transpose conv) */
#define FV2P_ERR_DEVICE (-4)      /* no sm_100 device / wrong architecture */

/* bits OR-ed into *status_dev by kernels */
#define FV2P_STATUS_OUT_OVERFLOW 1   /* more active outputs than out_cap */
#define FV2P_STATUS_VOXEL_OVERFLOW 2 /* more voxels than the output capacity */

/* conv arithmetic modes.  The reference's second dtype, fp16 (sparse_conv_ext.indice_conv_half, all.cc:43,
 * pcdet/ops/spconv/ops.py:115-126), is deliberately NOT built: the reduced-precision path of this library is bf16
 * (BASELINE.json north_star: "fp32 ... and a stated bf16 tolerance"), which keeps fp32's exponent range through 21
 * unnormalised-by-design layers; half tensors are rejected with NotImplementedError / FV2P_ERR_UNSUPPORTED rather than
 * silently converted. */
#define FV2P_MODE_F32 0      /* fp32 in/out, fp32 FMA on CUDA cores (any channel count)            */
#define FV2P_MODE_BF16_TC 1  /* bf16 in/out, tcgen05 kind::f16, fp32 accumulation in TMEM          */
#define FV2P_MODE_FP32_TC 2   /* fp32 in/out on tcgen05: operands split into bf16 hi + lo, three products per term
                               * (hi*hi + hi*lo + lo*hi), fp32 accumulation; <= 1.3e-5 relative at 27*128 terms    */
#define FV2P_MODE_TF32X3_TC FV2P_MODE_FP32_TC /* its name while the split was tf32 + bf16 (same value, same contract) */
#define FV2P_MODE_BF16_SIMT 3 /* bf16 in/out on CUDA cores (first layer, odd channel counts)       */
#define FV2P_MODE_F32_IN_BF16_OUT 4 /* fp32 in, bf16 out on CUDA cores (entry layer of bf16 path)  */

typedef void *fv2p_stream_t; /* cudaStream_t */

#if defined(__GNUC__)
#define FV2P_API __attribute__((visibility("default")))
#else
#define FV2P_API
#endif

FV2P_API int fv2p_abi_version(void);
FV2P_API const char *fv2p_last_error(void);
/* Fails with FV2P_ERR_DEVICE unless the current device is compute capability 10.x. */
FV2P_API int fv2p_device_check(int *sm_count, int *cc_major, int *cc_minor);

/* ---------------------------------------------------------------------------------------------
 * Voxelization + mean VFE in one pass.
 * Replaces VoxelGenerator.generate (pcdet/datasets/processor/voxel_generator.py:35-39,75-207, numba,
 * CPU, one frame per call) followed by MeanVFE.forward (pcdet/models/backbones_3d/vfe/mean_vfe.py:
 * 14-31), for a whole batch of frames per call, and writes the collate_batch layout
 * (pcdet/datasets/dataset.py:164-169): coords [M,4] = (batch,z,y,x), frames contiguous.
 *   points        device [total_points, num_features] fp32, frames concatenated
 *   frame_offsets device [batch+1] int32, frame b = rows [off[b], off[b+1])
 *   max_frame_points  host upper bound of the largest frame (used to size the work decomposition)
 *   range6/vsize3 host fp32 (the reference casts both to fp32, voxel_generator.py:22-24)
 *   coords        device [cap,4] int32 out
 *   voxel_features device [cap,num_features] fp32 out (mean of the kept points)
 *   num_points    device [cap] int32 out, or NULL
 *   voxels        device [cap,max_points,num_features] fp32 out (the legacy padded tensor), or NULL
 *   voxel_offsets device [batch+1] int32 out: frame b owns voxel rows [voff[b], voff[b+1]); voff[batch]=M
 *   cap           rows available in the outputs (>= min(total_points, batch*max_voxels) is always enough)
 * Semantics are the numba loop's: voxel ids in first-arrival order of the points, at most max_points
 * lowest-index points kept per voxel, and the `break` when voxel number max_voxels would be opened
 * (every later point of that frame is dropped).
 *   features_stream  NULL, or a second stream: coords / voxel_offsets are complete on `stream` when the call
 *                 returns control to it, while voxel_features / num_points / voxels are finished on
 *                 features_stream (forked from `stream` inside the call; the caller joins it), so that the
 *                 rulebooks, which only read coordinates, need not wait for the means.
 * ------------------------------------------------------------------------------------------- */
FV2P_API size_t fv2p_voxelize_workspace_bytes(int64_t total_points, int batch, int64_t max_frame_points,
                                     int max_points, int64_t cap);
FV2P_API int fv2p_voxelize_mean(const float *points, const int32_t *frame_offsets, int64_t total_points,
                       int batch, int64_t max_frame_points, int num_features, const float *range6,
                       const float *vsize3, int max_points, int max_voxels, int32_t *coords,
                       float *voxel_features, int32_t *num_points, float *voxels,
                       int32_t *voxel_offsets, int64_t cap, int32_t *status_dev, void *workspace,
                       size_t workspace_bytes, fv2p_stream_t stream, fv2p_stream_t features_stream);

/* The same call, which ALSO builds the level-0 coordinate table of the sparse convolutions (what fv2p_table_build would
 * make of `coords` right afterwards) while it assigns the voxel rows: one launch and one pass over the coordinates
 * less at the head of the step.  level0_table: fv2p_table_bytes(table_row_cap) bytes, cleared by this call;
 * table_row_cap >= cap; shape3 = the spatial shape [D,H,W] the convolutions use for these coordinates
 * (spconv_backbone.py:80: grid_size[::-1] + [1,0,0]).  level0_table == NULL: exactly fv2p_voxelize_mean. */
FV2P_API int fv2p_voxelize_mean_table(const float *points, const int32_t *frame_offsets, int64_t total_points,
                                      int batch, int64_t max_frame_points, int num_features,
                                      const float *range6, const float *vsize3, int max_points,
                                      int max_voxels, int32_t *coords, float *voxel_features,
                                      int32_t *num_points, float *voxels, int32_t *voxel_offsets, int64_t cap,
                                      int32_t *status_dev, void *workspace, size_t workspace_bytes,
                                      fv2p_stream_t stream, fv2p_stream_t features_stream, void *level0_table,
                                      int64_t table_row_cap, const int32_t *shape3);

/* Reference-shaped single-frame entry (VoxelGenerator.generate): synchronises and returns the voxel
 * count through *num_voxels_host.  Outputs as above with batch = 1 (coords still [M,4]). */
FV2P_API int fv2p_voxel_generate(const float *points, int64_t num_points_in, int num_features,
                        const float *range6, const float *vsize3, int max_points, int max_voxels,
                        int32_t *coords, float *voxel_features, int32_t *num_points, float *voxels,
                        int64_t cap, int32_t *num_voxels_host, void *workspace, size_t workspace_bytes,
                        fv2p_stream_t stream);

/* MeanVFE.forward alone (mean_vfe.py:26-28) for callers that already hold the padded tensor. */
FV2P_API int fv2p_mean_vfe(const float *voxels, const int32_t *num_points, int64_t num_voxels, int max_points,
                  int num_features, float *out, fv2p_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Rulebooks, second generation (round 2): what the convolution reads is the output-major neighbour map
 * nbr [K, nbr_stride]; the reference-layout pair tensors are built on demand (fv2p_subm_pairs / fv2p_conv_pairs).
 * Replace getIndicePair<3> (include/spconv/spconv_ops.h:28-141) and its functors (src/indice.cc,
 * src/indice_cuda.cu, include/spconv/indice.cu.h, geometry.h:25-297).  Results are bit-identical to the reference's
 * CPU path (the deterministic one).
 *
 * A "table" is a caller-owned open-addressing map (b,z,y,x) -> row of fv2p_table_bytes(row_cap) bytes.  The table a
 * strided rulebook fills for its OUTPUT coordinates is the input table of the next level's submanifold rulebook, so
 * a backbone builds one table per level.  Calls with FV2P_FLAG_PREFILLED expect their buffers to have been prepared
 * by ONE fv2p_geometry_prefill launch at the start of the pass (tables cleared, scan states zeroed, -1 fills);
 * without the flag each call prepares its own buffers first.
 *   subm_neighbours: odd kernels with dilation 1 probe only offsets k <= K/2 (a hit (i,k)->j is also the pair
 *                    (j,K-1-k)->i); rows K/2+1..K-1 of `nbr` must hold -1 on entry (FV2P_PREFILL_NBR_MIRROR).
 *   conv_neighbours: candidates in getValidOutPos order, atomicMin bids, ONE decoupled look-back scan ranks the
 *                    winners = first-touch output rows (geometry.h:181-187); `nbr` must hold -1 on entry.
 * ------------------------------------------------------------------------------------------- */
#define FV2P_FLAG_PREFILLED 1

#define FV2P_PREFILL_TABLE 1      /* ptr = table, a = row_cap                                   */
#define FV2P_PREFILL_NBR_ALL 2    /* ptr = nbr, a = kvol, b = nbr_stride (multiple of 4): all -1 */
#define FV2P_PREFILL_NBR_MIRROR 3 /* ptr = nbr, a = kvol, b = nbr_stride: rows K/2+1..K-1 = -1   */
#define FV2P_PREFILL_CONV_WS 4    /* ptr = conv_neighbours workspace, a = n_in_cap, b = kvol     */
#define FV2P_PREFILL_GROUP_WS 5   /* ptr = group_rows workspace, a = n_cap                       */
typedef struct {
  int32_t kind;
  int32_t reserved;
  void *ptr; /* device, 16-byte aligned */
  int64_t a, b;
} fv2p_prefill_item;

FV2P_API int fv2p_geometry_prefill(const fv2p_prefill_item *items, int count, fv2p_stream_t stream);

FV2P_API size_t fv2p_table_bytes(int64_t row_cap);
/* Inserts rows [0, n) of `indices` (level-0 coordinates) into a cleared table. */
FV2P_API int fv2p_table_build(const int32_t *indices, int64_t n_cap, const int32_t *n_dev, const int32_t *shape3,
                              void *table, int64_t table_row_cap, int32_t *status_dev, int flags,
                              fv2p_stream_t stream);
FV2P_API int fv2p_subm_neighbours(const int32_t *indices, int64_t n_cap, const int32_t *n_dev, int batch,
                                  const int32_t *shape3, const int32_t *ksize3, const int32_t *dilation3,
                                  const void *table, int64_t table_row_cap, int32_t *nbr, int64_t nbr_stride,
                                  int flags, fv2p_stream_t stream);
FV2P_API size_t fv2p_conv_neighbours_workspace_bytes(int64_t n_in_cap, int kvol);
FV2P_API int fv2p_conv_neighbours(const int32_t *indices, int64_t n_cap, const int32_t *n_dev, int batch,
                                  const int32_t *out_shape3, const int32_t *ksize3, const int32_t *stride3,
                                  const int32_t *pad3, const int32_t *dilation3, int32_t *out_indices,
                                  int64_t out_cap, int32_t *n_out_dev, void *table, int64_t table_row_cap,
                                  int32_t *nbr, int64_t nbr_stride, int32_t *status_dev, void *workspace,
                                  size_t workspace_bytes, int flags, fv2p_stream_t stream);
/* Reference-layout pair lists on demand: pairs [K,2,pair_stride] (-1 tail up to the live row count), pair_num [K].
 * subm_pairs reads the neighbour map (mirror-symmetric kernels) or indices + table (other geometries); conv_pairs
 * reads the input coordinates and the OUTPUT table of the strided rulebook. */
FV2P_API size_t fv2p_pairs_workspace_bytes(int64_t n_in_cap, int kvol);
FV2P_API int fv2p_subm_pairs(const int32_t *indices, int64_t n_cap, const int32_t *n_dev, int batch,
                             const int32_t *shape3, const int32_t *ksize3, const int32_t *dilation3,
                             const void *table, int64_t table_row_cap, const int32_t *nbr, int64_t nbr_stride,
                             int32_t *pairs, int64_t pair_stride, int32_t *pair_num, void *workspace,
                             size_t workspace_bytes, fv2p_stream_t stream);
FV2P_API int fv2p_conv_pairs(const int32_t *indices, int64_t n_cap, const int32_t *n_dev, int batch,
                             const int32_t *out_shape3, const int32_t *ksize3, const int32_t *stride3,
                             const int32_t *pad3, const int32_t *dilation3, const void *table,
                             int64_t table_row_cap, int32_t *pairs, int64_t pair_stride, int32_t *pair_num,
                             void *workspace, size_t workspace_bytes, fv2p_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Rulebooks, first-generation entry points (kept; same kernels underneath, each call builds its own table inside
 * `workspace` and, when asked for them, the pair lists).
 *
 *   indices   device [n_cap,4] int32;  the live row count is *n_dev if n_dev != NULL, else n_cap
 *   pairs     device [K,2,pair_stride] int32 out or NULL; rows >= the live count are not touched
 *             beyond index n-1 (the -1 tail is written up to n)
 *   pair_num  device [K] int32 out or NULL
 *   nbr       device [K,nbr_stride] int32 out or NULL (see header comment)
 *   pairs_stream  NULL, or a second stream for the compaction of pairs / pair_num: it is forked from `stream`
 *             once the neighbour map is complete, so that consumers of `nbr` need not wait for the pair lists.
 *             The caller joins it and must leave `workspace` alone until it has drained.
 * ------------------------------------------------------------------------------------------- */
FV2P_API size_t fv2p_rulebook_workspace_bytes(int64_t n_in_cap, int64_t n_out_cap, int kvol);

FV2P_API int fv2p_rulebook_subm(const int32_t *indices, int64_t n_cap, const int32_t *n_dev, int batch,
                       const int32_t *shape3, const int32_t *ksize3, const int32_t *dilation3,
                       int32_t *pairs, int64_t pair_stride, int32_t *pair_num, int32_t *nbr,
                       int64_t nbr_stride, void *workspace, size_t workspace_bytes,
                       fv2p_stream_t stream, fv2p_stream_t pairs_stream);

/*   out_indices device [out_cap,4] int32 out;  n_out_dev device int32 out (live output rows)      */
FV2P_API int fv2p_rulebook_conv(const int32_t *indices, int64_t n_cap, const int32_t *n_dev, int batch,
                       const int32_t *out_shape3, const int32_t *ksize3, const int32_t *stride3,
                       const int32_t *pad3, const int32_t *dilation3, int32_t *out_indices,
                       int64_t out_cap, int32_t *n_out_dev, int32_t *pairs, int64_t pair_stride,
                       int32_t *pair_num, int32_t *nbr, int64_t nbr_stride, int32_t *status_dev,
                       void *workspace, size_t workspace_bytes, fv2p_stream_t stream,
                       fv2p_stream_t pairs_stream);

/* Reference-shaped entry: sparse_conv_ext.get_indice_pairs_3d (all.cc:26, spconv_ops.h:28-33), same
 * argument order.  subm != 0 forces stride 1 / padding ksize/2 like spconv_ops.h:76-80 and copies
 * `indices` to out_indices.  Synchronises; *num_act_out_host receives the output row count.
 * pairs is [K,2,n] (pair_stride = n).  transpose != 0 -> FV2P_ERR_UNSUPPORTED. */
FV2P_API int fv2p_get_indice_pairs_3d(const int32_t *indices, int64_t n, int batch, const int32_t *out_shape3,
                             const int32_t *spatial_shape3, const int32_t *ksize3,
                             const int32_t *stride3, const int32_t *pad3, const int32_t *dilation3,
                             const int32_t *out_pad3, int subm, int transpose, int32_t *out_indices,
                             int64_t out_cap, int32_t *pairs, int32_t *pair_num, int32_t *nbr,
                             int64_t nbr_stride, int32_t *num_act_out_host, void *workspace,
                             size_t workspace_bytes, fv2p_stream_t stream);

/* Output-major neighbour map from a reference-layout pair tensor (any producer). */
FV2P_API int fv2p_pairs_to_nbr(const int32_t *pairs, const int32_t *pair_num, int kvol, int64_t pair_stride,
                      int inverse, int64_t n_out, int32_t *nbr, int64_t nbr_stride,
                      fv2p_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Sparse convolution forward with fused epilogue.  Replaces indiceConv<T> (spconv_ops.h:261-362:
 * per-offset gather -> torch::mm_out -> scatter-add), the bias add of conv.py:223-224, and the
 * BatchNorm1d(eval) + residual + ReLU that follow in the backbones (spconv_backbone.py:25-27,57-66):
 *     out[i,:] = act( (sum_k X[nbr[k][i],:] * W[k] + bias) * scale + shift + residual[i,:] )
 * Any of bias/scale/shift/residual may be NULL.  scale/shift are the folded eval BatchNorm:
 * scale = gamma / sqrt(var + eps), shift = beta - mean * scale.
 *   features  device [n_in_cap,cin]   (fp32 or bf16 by mode); n_in_cap = rows the buffer holds (the tensor-core
 *             modes describe it to the TMA unit; every neighbour index must be < n_in_cap)
 *   weight    device: FV2P_MODE_F32 / *_SIMT: [K,cin,cout] fp32 (the reference layout, flattened);
 *             tensor-core modes: the packed image written by fv2p_pack_weight
 *   row_perm  NULL, or the row order from fv2p_sort_rows_by_mask (tensor-core modes only): `nbr` is then the map
 *             permuted the same way (nbr_sorted) and sorted position t writes output row row_perm[t]; results are
 *             identical either way
 *   tile_order NULL, or fv2p_sort_rows_by_mask's tile list (needs row_perm): (tile, offset mask) pairs in the order
 *             in which the 128-row tiles are handed out (most active offsets first)
 *   sched     NULL (tiles dealt round-robin to the CTAs), or two device int32 words, zero on entry, that the
 *             tensor-core kernel uses as its tile counter and leaves zeroed again; one pair per launch that may be
 *             in flight at the same time
 *   n_out_cap rows of `out`/`nbr` columns; live count *n_out_dev if given
 * ------------------------------------------------------------------------------------------- */
FV2P_API int fv2p_conv_fwd(const void *features, int64_t n_in_cap, const void *weight, const int32_t *nbr, int64_t nbr_stride,
                  const int32_t *row_perm, const int32_t *tile_order, int32_t *sched, int kvol, int64_t n_out_cap,
                  const int32_t *n_out_dev, int cin, int cout, const float *bias, const float *scale,
                  const float *shift, const void *residual, int relu, int mode, void *out, fv2p_stream_t stream);

/* Row order for the tensor-core conv: output rows grouped by a 12-bit digest of their neighbour mask (bit k = offset
 * k has a neighbour; digest = which x-offsets occur | which (z,y) lines of the kernel have a neighbour; `kx` = extent
 * of the fastest kernel axis).  perm[t] = output row at position t; nbr_sorted[k][t] = nbr[k][perm[t]] (optional).
 * Tiles cut from this order need less than half the pipeline stages (rows of a tile share their active offsets); the
 * conv result does not change, and the order inside a group is whatever the atomic cursors hand out (any permutation
 * gives bit-identical conv results).  tile_order (optional, 2 * (ceil(n_cap / 128) + 1) int32): (tile, OR of the
 * tile's masks) pairs by descending population count, ties by tile.  No reference counterpart (spconv_ops.h:308-357
 * works on per-offset pair lists).  fv2p_sort_rows_by_mask is the first-generation name (kx inferred from kvol). */
FV2P_API size_t fv2p_group_rows_workspace_bytes(int64_t n_cap);
FV2P_API int fv2p_group_rows(const int32_t *nbr, int64_t nbr_stride, int kvol, int kx, int64_t n_cap,
                             const int32_t *n_dev, int32_t *perm, int32_t *nbr_sorted, int64_t sorted_stride,
                             int32_t *tile_order, void *workspace, size_t workspace_bytes, int flags,
                             fv2p_stream_t stream);
FV2P_API size_t fv2p_sort_rows_workspace_bytes(int64_t n_cap);
FV2P_API int fv2p_sort_rows_by_mask(const int32_t *nbr, int64_t nbr_stride, int kvol, int64_t n_cap,
                                    const int32_t *n_dev, int32_t *perm, int32_t *nbr_sorted,
                                    int64_t sorted_stride, int32_t *tile_order, void *workspace,
                                    size_t workspace_bytes, fv2p_stream_t stream);

/* Producer of the gathered A tile in the tensor-core kernels: -1 = auto (default: TMA gather for stages of one
 * offset, cp.async for packed stages), 0 = LSU (swizzled cp.async) everywhere, 1 = TMA (cp.async.bulk.tensor
 * tile::gather4) where possible.  Same results either way; a tuning knob kept for measurement
 * (profiles/r2_notes.md).  Process-wide, not stream-ordered. */
FV2P_API int fv2p_tc_gather_mode(int mode);

/* Packed weight image for the tensor-core modes (done once per layer, device to device). */
FV2P_API size_t fv2p_pack_weight_bytes(int kvol, int cin, int cout, int mode);
FV2P_API int fv2p_pack_weight(const float *weight_f32, int kvol, int cin, int cout, int mode, void *packed,
                     fv2p_stream_t stream);

/* Reference-shaped entry: sparse_conv_ext.indice_conv_fp32 (all.cc:38, spconv_ops.h:261-263) on a
 * reference-layout pair tensor; workspace holds the temporary neighbour map
 * (fv2p_indice_conv_workspace_bytes).  Result rows not touched by any pair are zero like the
 * reference's torch::zeros output (spconv_ops.h:294). */
FV2P_API size_t fv2p_indice_conv_workspace_bytes(int kvol, int64_t num_act_out);
FV2P_API int fv2p_indice_conv_fp32(const float *features, const float *filters, const int32_t *pairs,
                          const int32_t *pair_num, int64_t pair_stride, int64_t num_act_out,
                          int inverse, int subm, int kvol, int cin, int cout, float *out,
                          void *workspace, size_t workspace_bytes, fv2p_stream_t stream);

/* Filter gradient of the sparse convolution (training path).  Replaces the per-offset
 * gather + gather + torch::mm_out and the host loop over the D2H-copied pair counts of indiceConvBackward
 * (include/spconv/spconv_ops.h:365-457, :378, :399-436) with ONE launch and no host synchronisation:
 *     grad_filters[k][ci][co] = sum_i features[nbr[k][i]][ci] * grad_out[i][co]
 * on the forward pass's output-major neighbour map.  fp32; grad_filters [K,cin,cout] is zeroed by the call; partial
 * sums are combined with atomics, so results match the reference to rounding.  The input gradient is
 * fv2p_conv_fwd on the transposed map with W^T (no kernel of its own). */
FV2P_API int fv2p_conv_grad_filters(const float *features, const float *grad_out, const int32_t *nbr,
                                    int64_t nbr_stride, int kvol, int64_t n_out_cap, const int32_t *n_out_dev,
                                    int cin, int cout, float *grad_filters, fv2p_stream_t stream);

/* Training-mode BatchNorm1d over the N active rows of a feature matrix.  Replaces the
 * `nn.BatchNorm1d(eps=1e-3, momentum=0.01)` the reference's backbones apply to `.features` after every conv
 * (pcdet/models/backbones_3d/spconv_backbone.py:75, :193) when the module is in training mode: batch statistics
 * (biased variance to normalise, unbiased into running_var), y = (x - mean) * invstd * weight + bias, optional ReLU.
 * x, y [n, channels] fp32 row-major (y may alias x); weight / bias / running_* may be NULL (affine=False /
 * track_running_stats=False); `momentum` is the factor actually applied (the caller resolves momentum=None).
 * save_mean / save_invstd [channels] are what the backward needs.  n < 2 is an error, as in torch.
 * `workspace`: fv2p_batchnorm_workspace_bytes(channels) bytes, 8-byte aligned, ZERO on first use; the calls hand it
 * back zeroed.  Two launches, no host synchronisation; fp64 column sums, so results match torch to rounding. */
FV2P_API size_t fv2p_batchnorm_workspace_bytes(int channels);
FV2P_API int fv2p_batchnorm_train_fwd(const float *x, int64_t n, int channels, const float *weight, const float *bias,
                                      float eps, float momentum, float *running_mean, float *running_var, int relu,
                                      float *y, float *save_mean, float *save_invstd, void *workspace,
                                      size_t workspace_bytes, fv2p_stream_t stream);
/* Backward of the above (without the ReLU): grad_bias = sum grad_out, grad_weight = sum grad_out * xhat,
 * grad_input = weight * invstd * (grad_out - grad_bias / n - xhat * grad_weight / n).  grad_weight / grad_bias
 * [channels] are always written (also when weight is NULL). */
FV2P_API int fv2p_batchnorm_train_bwd(const float *x, const float *grad_out, int64_t n, int channels,
                                      const float *weight, const float *save_mean, const float *save_invstd,
                                      float *grad_input, float *grad_weight, float *grad_bias, void *workspace,
                                      size_t workspace_bytes, fv2p_stream_t stream);

/* SparseConvTensor.dense() (pcdet/ops/spconv/structure.py:57-66) in channels-first layout
 * [batch, C, D, H, W]; `dense` must be zero-filled by the caller. features fp32. */
FV2P_API int fv2p_dense_ncdhw(const float *features, const int32_t *indices, int64_t n_cap,
                     const int32_t *n_dev, int channels, const int32_t *shape3, float *dense,
                     fv2p_stream_t stream);

/* HeightCompression.forward (pcdet/models/backbones_2d/map_to_bev/height_compression.py:20-25): the dense() of the
 * stride-8 output viewed as [batch, channels*D, H, W] - the same bytes as [batch, channels, D, H, W].  Every
 * element of `spatial_features` is written (zero where no row exists); the live rows are *n_dev if given, else
 * n_cap; rows are (batch, z, y, x).  elem_bytes 4 = fp32, 2 = bf16.  With a workspace of
 * fv2p_height_compression_workspace_bytes (a cell -> row map) the map is written in one coalesced pass; without
 * one (NULL) it is zero fill + scatter.  Stream-ordered, no host sync: can follow the backbone in the same CUDA
 * graph. */
FV2P_API size_t fv2p_height_compression_workspace_bytes(int batch, const int32_t *shape3);
FV2P_API int fv2p_height_compression(const void *features, const int32_t *indices, int64_t n_cap,
                            const int32_t *n_dev, int batch, int channels, const int32_t *shape3,
                            int elem_bytes, void *spatial_features, void *workspace, size_t workspace_bytes,
                            fv2p_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Consumers of the sparse outputs, on the level's coordinate table instead of dense grids.
 *
 * fv2p_voxel_three_nn: front end of ResidualVoxelToPointDecoder.forward
 * (pcdet/models/backbones_3d/pfe/residual_v2p_decoder.py:86-116): for every query point the three nearest voxel
 * CENTRES of its own frame (get_voxel_centers, pcdet/utils/common_utils.py:76-92; three_nn,
 * pcdet/ops/pointnet2/pointnet2_batch/src/interpolate_gpu.cu:16-58, which scans every voxel per query), and
 * optionally the inverse-distance interpolation of the voxel features (top3_interpolate,
 * pointnet2_batch/pointnet2_utils.py:292-326; three_interpolate, interpolate_gpu.cu:78-100).
 *   point_coords  device [P,4] fp32 (batch, x, y, z)        (bottom_point_coords)
 *   voxel_indices device [n_cap,4] int32 (batch, z, y, x), frames contiguous; live rows *n_dev or n_cap
 *   table         the level's table (rows of voxel_indices), shape3 = the level's (D,H,W)
 *   voxel_size3   host fp32 (x,y,z), ALREADY multiplied by the level's downsample factor; range_min3 host fp32 (x,y,z)
 *   dist [P,3] fp32 (sqrt of the squared distance, like ThreeNN.forward), idx [P,3] int32 = row inside the point's
 *   frame (the reference searches one frame at a time); either may be NULL
 *   interpolated [P,channels] fp32 out or NULL; features [n_cap,channels] fp32
 * Indices and distances are bit-identical to the reference kernel (ties by lowest row).
 *
 * fv2p_voxel_query: voxel_query_kernel_stack (pcdet/ops/pointnet2/pointnet2_stack/src/voxel_query_gpu.cu:10-88) with
 * the dense point_indices grid of generate_voxel2pinds (pcdet/utils/spconv_utils.py:13-21) replaced by the table:
 * same cells in the same order, first nsample hits within `radius`, idx[0] = -1 for an empty ball.
 *   shape3 = (Z,Y,X) of the grid; range3 = (z_range, y_range, x_range); new_xyz [M,3], xyz [N,3] fp32;
 *   new_coords [M,4] int32 (batch,z,y,x); idx [M,nsample] int32 out (every entry is written)
 * ------------------------------------------------------------------------------------------- */
FV2P_API size_t fv2p_voxel_three_nn_workspace_bytes(int batch, int64_t n_points);
FV2P_API int fv2p_voxel_three_nn(const float *point_coords, int64_t n_points, const int32_t *voxel_indices,
                                 int64_t n_cap, const int32_t *n_dev, int batch, const int32_t *shape3,
                                 const void *table, int64_t table_row_cap, const float *voxel_size3,
                                 const float *range_min3, float *dist, int32_t *idx, const float *features,
                                 int channels, float *interpolated, void *workspace, size_t workspace_bytes,
                                 fv2p_stream_t stream);
FV2P_API int fv2p_voxel_query(int64_t m, const int32_t *shape3, int nsample, float radius, const int32_t *range3,
                              const float *new_xyz, const float *xyz, const int32_t *new_coords, const void *table,
                              int64_t table_row_cap, int32_t *idx, fv2p_stream_t stream);

/* Copies the live rows (count on the device) of a capacity-sized row buffer; row_bytes must be a multiple of 16.
 * Used to snapshot a step's result so that its D2H copy overlaps the next step. */
FV2P_API int fv2p_copy_rows(const void *src, void *dst, int64_t row_bytes, int64_t n_cap, const int32_t *n_dev,
                            fv2p_stream_t stream);

/* dtype helpers used by the bf16 path */
FV2P_API int fv2p_cast_f32_to_bf16(const float *src, void *dst, int64_t count, fv2p_stream_t stream);
FV2P_API int fv2p_cast_bf16_to_f32(const void *src, float *dst, int64_t count, fv2p_stream_t stream);

/* The library also exports a few fv2p_debug_* symbols (timing ablations, %globaltimer stamps, role timers in
 * profiling builds; see profiles/run_layer.py and profiles/timeline.py).  They are profiling hooks, not part of
 * this interface, and may change without an ABI version bump. */

#ifdef __cplusplus
}
#endif
#endif /* FV2P_B200_H_ */
