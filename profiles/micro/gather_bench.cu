// Microbenchmark: how many scattered rows per cycle can one SM pull from L2 into shared memory?
// The sparse conv gathers 128 rows x (64..128 B) per pipeline stage; this measures the gather alone (no MMA) for
// the candidate producers, on all SMs at once, rows drawn at random from a table that fits L2.
//   mode 0  cp.async 16 B (LDGSTS), 8 (or 4) lanes per row, G commit groups in flight per warp
//   mode 1  ld.global.nc.v4 -> registers (U loads in flight per lane) -> st.shared
//   mode 2  TMA tile::gather4 (4 rows per instruction), S mbarrier slots in flight per warp
//   mode 3  half the warps mode 0, half mode 2 (do the two paths add up?)
//   mode 4  ld.global.nc 32 B per lane (v8.f32) -> st.shared
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o profiles/micro/_build/gather_bench profiles/micro/gather_bench.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void tma_gather4(uint32_t dst, const CUtensorMap *map, int col, int r0, int r1, int r2, int r3,
                                            uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes"
               " [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(dst), "l"(map), "r"(col), "r"(r0), "r"(r1), "r"(r2),
               "r"(r3), "r"(bar) : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}

constexpr int kWarpBuf = 8192;  // bytes of smem each warp cycles through

// Each warp gathers `rows_per_warp` rows (indices idx[warp_global * rows_per_warp ...]) of `row_bytes` each.
__global__ void __launch_bounds__(1024, 1)
gather_kernel(const __grid_constant__ CUtensorMap map, const uint8_t *__restrict__ table, const int *__restrict__ idx,
              int rows_per_warp, int row_bytes, int mode, int depth, unsigned long long *cycles, float *sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bars[32 * 8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  uint8_t *buf = smem + (size_t)warp * kWarpBuf;
  const uint32_t buf_u32 = smem_u32(buf);
  const int gw = blockIdx.x * nwarps + warp;
  const int *my = idx + (size_t)gw * rows_per_warp;
  if (lane == 0)
    for (int s = 0; s < 8; ++s) mbar_init(smem_u32(&bars[warp * 8 + s]), 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  const long long t0 = clock64();
  int m = mode;
  if (mode == 3) m = (warp & 1) ? 2 : 0;
  float acc = 0.f;
  if (m == 0) {
    const int lanes_per_row = row_bytes / 16;
    const int rows_per_instr = 32 / lanes_per_row;
    const int chunk = lane % lanes_per_row, r_in = lane / lanes_per_row;
    const int instr_per_group = 8;
    int group = 0;
    for (int r0 = 0; r0 < rows_per_warp; r0 += rows_per_instr * instr_per_group, ++group) {
#pragma unroll
      for (int u = 0; u < instr_per_group; ++u) {
        const int r = r0 + u * rows_per_instr + r_in;
        const int src = my[r];
        cp_async16(buf_u32 + (uint32_t)(((r * row_bytes) + chunk * 16) & (kWarpBuf - 1)),
                   table + (size_t)src * row_bytes + chunk * 16);
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      // keep `depth` groups in flight
      if (depth == 1) asm volatile("cp.async.wait_group 1;" ::: "memory");
      else if (depth == 2) asm volatile("cp.async.wait_group 2;" ::: "memory");
      else if (depth == 4) asm volatile("cp.async.wait_group 4;" ::: "memory");
      else asm volatile("cp.async.wait_group 7;" ::: "memory");
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  } else if (m == 1) {
    const int lanes_per_row = row_bytes / 16;
    const int rows_per_instr = 32 / lanes_per_row;
    const int chunk = lane % lanes_per_row, r_in = lane / lanes_per_row;
    constexpr int U = 8;
    for (int r0 = 0; r0 < rows_per_warp; r0 += rows_per_instr * U) {
      float4 v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int r = r0 + u * rows_per_instr + r_in;
        const int src = my[r];
        v[u] = __ldg(reinterpret_cast<const float4 *>(table + (size_t)src * row_bytes + chunk * 16));
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int r = r0 + u * rows_per_instr + r_in;
        *reinterpret_cast<float4 *>(buf + (((r * row_bytes) + chunk * 16) & (kWarpBuf - 1))) = v[u];
      }
    }
  } else if (m == 4) {
    const int lanes_per_row = row_bytes / 32;
    const int rows_per_instr = 32 / lanes_per_row;
    const int chunk = lane % lanes_per_row, r_in = lane / lanes_per_row;
    constexpr int U = 4;
    for (int r0 = 0; r0 < rows_per_warp; r0 += rows_per_instr * U) {
      float v[U][8];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int r = r0 + u * rows_per_instr + r_in;
        const int src = my[r];
        const uint8_t *p = table + (size_t)src * row_bytes + chunk * 32;
        asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=f"(v[u][0]), "=f"(v[u][1]), "=f"(v[u][2]), "=f"(v[u][3]), "=f"(v[u][4]), "=f"(v[u][5]),
                       "=f"(v[u][6]), "=f"(v[u][7]) : "l"(p));
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int r = r0 + u * rows_per_instr + r_in;
        float4 *d = reinterpret_cast<float4 *>(buf + (((r * row_bytes) + chunk * 32) & (kWarpBuf - 1)));
        d[0] = make_float4(v[u][0], v[u][1], v[u][2], v[u][3]);
        d[1] = make_float4(v[u][4], v[u][5], v[u][6], v[u][7]);
      }
    }
  } else {
    // TMA gather4: lane l of the warp issues rows 4l..4l+3 of a 128-row batch; `depth` batches in flight per warp
    const int batch_rows = 128 * row_bytes <= kWarpBuf ? 128 : kWarpBuf / row_bytes;  // rows per batch (fits the buffer)
    const int lanes_used = batch_rows / 4;
    const int elem = 4;
    int batch = 0;
    for (int r0 = 0; r0 < rows_per_warp; r0 += batch_rows, ++batch) {
      const int s = batch % depth;
      const uint32_t bar = smem_u32(&bars[warp * 8 + s]);
      if (batch >= depth) mbar_wait(bar, ((batch / depth) - 1) & 1);
      if (lane == 0) mbar_expect(bar, (uint32_t)batch_rows * row_bytes);
      __syncwarp();
      if (lane < lanes_used) {
        const int4 rows = *reinterpret_cast<const int4 *>(&my[r0 + 4 * lane]);
        // all slots of a warp write the same buffer region (contents are irrelevant here)
        tma_gather4(buf_u32 + (uint32_t)(4 * lane) * row_bytes, &map, 0, rows.x, rows.y, rows.z, rows.w, bar);
      }
      (void)elem;
    }
    const int total = (rows_per_warp + batch_rows - 1) / batch_rows;
    for (int b = (total > depth ? total - depth : 0); b < total; ++b)
      mbar_wait(smem_u32(&bars[warp * 8 + b % depth]), (b / depth) & 1);
  }
  __syncthreads();
  const long long t1 = clock64();
  acc += buf[lane * 16];
  if (threadIdx.x == 0) cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
  if (acc == 12345.678f) sink[0] = acc;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char **argv) {
  const int table_rows = argc > 1 ? atoi(argv[1]) : 120000;
  int dev_sms = 0;
  CK(cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, 0));
  void *fnp = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q));
  EncodeTiledFn encode = (EncodeTiledFn)fnp;
  CK(cudaFuncSetAttribute(gather_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * kWarpBuf > 200 * 1024 ? 200 * 1024 : 32 * kWarpBuf));
  unsigned long long *cyc;
  float *sink;
  CK(cudaMalloc(&cyc, sizeof(unsigned long long) * dev_sms));
  CK(cudaMalloc(&sink, 4));
  printf("# SMs %d, table rows %d; rows/cyc/SM and B/cyc/SM from the slowest CTA's clock64\n", dev_sms, table_rows);
  printf("%-28s %6s %6s %6s %10s %10s %10s\n", "mode", "rowB", "warps", "depth", "cyc/row", "B/cyc/SM", "GB/s chip");
  const char *names[] = {"cp.async16", "ldg.v4->sts", "tma gather4", "cp.async + tma (half/half)", "ldg.v8->sts"};
  for (int row_bytes : {128, 64}) {
    uint8_t *table;
    CK(cudaMalloc(&table, (size_t)table_rows * row_bytes));
    CK(cudaMemset(table, 1, (size_t)table_rows * row_bytes));
    CUtensorMap map;
    cuuint64_t gdim[2] = {(cuuint64_t)row_bytes / 4, (cuuint64_t)table_rows};
    cuuint64_t gstride[1] = {(cuuint64_t)row_bytes};
    cuuint32_t box[2] = {(cuuint32_t)row_bytes / 4, 1};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, table, gdim, gstride, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
    for (int mode : {0, 1, 4, 2, 3}) {
      for (int warps : {4, 8, 16}) {
        for (int depth : {2, 4, 7}) {
          if ((mode == 1 || mode == 4) && depth != 2) continue;
          const int rows_per_warp = 16384;
          const size_t n_idx = (size_t)dev_sms * warps * rows_per_warp;
          std::vector<int> h(n_idx);
          uint32_t x = 12345u;
          for (size_t i = 0; i < n_idx; ++i) {
            x = x * 1664525u + 1013904223u;
            h[i] = (int)((x >> 8) % (uint32_t)table_rows);
          }
          int *idx;
          CK(cudaMalloc(&idx, n_idx * 4));
          CK(cudaMemcpy(idx, h.data(), n_idx * 4, cudaMemcpyHostToDevice));
          for (int rep = 0; rep < 2; ++rep)
            gather_kernel<<<dev_sms, warps * 32, warps * kWarpBuf>>>(map, table, idx, rows_per_warp, row_bytes, mode,
                                                                      depth, cyc, sink);
          CK(cudaDeviceSynchronize());
          std::vector<unsigned long long> hc(dev_sms);
          CK(cudaMemcpy(hc.data(), cyc, sizeof(unsigned long long) * dev_sms, cudaMemcpyDeviceToHost));
          unsigned long long mx = 0;
          for (auto c : hc) mx = c > mx ? c : mx;
          const double rows_sm = (double)warps * rows_per_warp;
          const double cyc_row = (double)mx / rows_sm;
          printf("%-28s %6d %6d %6d %10.2f %10.1f %10.0f\n", names[mode], row_bytes, warps, depth, cyc_row,
                 row_bytes / cyc_row, row_bytes / cyc_row * dev_sms * 1.9);
          CK(cudaFree(idx));
        }
      }
    }
    CK(cudaFree(table));
  }
  return 0;
}
