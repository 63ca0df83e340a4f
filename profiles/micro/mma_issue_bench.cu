// Microbenchmark: what does the single-thread tcgen05.mma issue loop of the sparse conv cost per pipeline stage?
// One CTA per SM, operands already in shared memory (contents irrelevant), no producers: per stage the issuing
// thread does what conv_tc.cu's MMA warp does - [wait on an already-complete mbarrier] [fence] J x tcgen05.mma,
// tcgen05.commit to a ring of barriers - and waits for the commit of 8 stages ago.  Reports cycles per stage.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o profiles/micro/_build/mma_issue_bench profiles/micro/mma_issue_bench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
template <bool kTf32>
__device__ __forceinline__ void tc_mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  if (kTf32)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
// A operand from tensor memory (kind::f16)
__device__ __forceinline__ void tc_mma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
               ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ uint32_t elect_one_sync() {
  uint32_t pred = 0, laneid = 0;
  asm volatile("{\n\t.reg .b32 %%rx;\n\t.reg .pred %%px;\n\telect.sync %%rx|%%px, %2;\n\t@%%px mov.s32 %1, 1;\n\tmov.s32 %0, %%rx;\n\t}"
               : "+r"(laneid), "+r"(pred) : "r"(0xFFFFFFFFu));
  return pred;
}
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t sbo, uint32_t layout) {
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46) |
         ((uint64_t)layout << 61);
}
__host__ __device__ constexpr uint32_t instr_desc(int n, bool tf32) {
  return (1u << 4) | ((tf32 ? 2u : 1u) << 7) | ((tf32 ? 2u : 1u) << 10) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
}

// variant bits: 1 = wait on a ready barrier before each stage, 2 = tcgen05.fence::after per stage,
// 4 = read a flag word from smem per stage, 8 = __syncwarp per stage with the whole warp looping (lane 0 issues),
// 16 = whole warp loops and the issue is guarded by elect.sync instead of lane == 0 (CUTLASS's form),
// 32 = commit only every 4th stage, 64 = never wait for a slot inside the loop (pure issue cost),
// 128 = wait for the stage's own commit right away (issue -> execute -> arrive latency),
// 256 = one elect.sync outside the stage loop: the elected thread alone runs it (no per-stage elect / syncwarp),
//       no slot waits; with 32 also no commits except every 4th stage
// 512 = with 256: the A operand comes from TMEM (columns 128.. of the allocation) instead of shared memory (kind::f16 only);
// 1024 = with 256: alternate SS (tf32 / f16 as the template says) and TS f16 MMAs, the 3xTF32 path's mix
template <bool kTf32>
__global__ void __launch_bounds__(128, 1)
issue_kernel(int n, int mmas_per_stage, int stages, int variant, unsigned long long *cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bars[16];
  __shared__ uint32_t tmem_slot;
  __shared__ volatile int flags[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < 16; ++s) mbar_init(smem_u32(&bars[s]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    for (int s = 0; s < 8; ++s) flags[s] = 0;
  }
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  if (warp == 1) {
    // bars[8..15] are "full" barriers kept permanently complete for phase 0 by one arrival each
    if (lane == 0)
      for (int s = 0; s < 8; ++s) mbar_arrive(smem_u32(&bars[8 + s]));
    __syncwarp();
    const uint32_t idesc = instr_desc(n, kTf32);
    const uint64_t a0 = smem_desc(smem_u32(smem), 1024, 2), b0 = smem_desc(smem_u32(smem + 16384), 1024, 2);
    const bool whole_warp = (variant & (8 | 16)) != 0;
    const bool use_elect = (variant & 16) != 0;
    const bool sparse_commit = (variant & 32) != 0;
    const long long t0 = clock64();
    if (variant & 256) {
      if (elect_one_sync()) {
        for (int st = 0; st < stages; ++st) {
          const int s = st & 7;
          if (variant & 512) {
            const uint32_t idesc16 = instr_desc(n, false);
            for (int j = 0; j < mmas_per_stage; ++j)
              tc_mma_ts(tmem, tmem + 128u + 8u * (j & 3), b0 + 2 * (j & 3), idesc16, (st | j) ? 1u : 0u);
          } else if (variant & 1024) {
            const uint32_t idesc16 = instr_desc(n, false);
            for (int j = 0; j < mmas_per_stage; ++j) {
              if (j & 1) tc_mma_ts(tmem, tmem + 128u + 8u * ((j >> 1) & 3), b0 + 2 * (j & 3), idesc16, 1u);
              else tc_mma<kTf32>(tmem, a0 + 2 * ((j >> 1) & 3), b0 + 2 * ((j >> 1) & 3), idesc, (st | j) ? 1u : 0u);
            }
          } else {
            for (int j = 0; j < mmas_per_stage; ++j)
              tc_mma<kTf32>(tmem, a0 + 2 * (j & 3), b0 + 2 * (j & 3), idesc, (st | j) ? 1u : 0u);
          }
          if (!sparse_commit || (s & 3) == 3) tc_commit(smem_u32(&bars[s]));
        }
      }
      __syncwarp();
      if (lane == 0) mbar_wait(smem_u32(&bars[(stages - 1) & 7]), ((stages - 1) >> 3) & 1);
    } else if (lane == 0 || whole_warp) {
      for (int st = 0; st < stages; ++st) {
        const int s = st & 7;
        if (st >= 8 && !(variant & (64 | 128)) && (!sparse_commit || (s & 3) == 3))
          mbar_wait(smem_u32(&bars[s]), ((st >> 3) - 1) & 1);  // the smem slot is free again
        if (variant & 1) mbar_wait(smem_u32(&bars[8 + s]), 0);
        int f = 0;
        if (variant & 4) f = flags[s];
        if (variant & 2) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (use_elect) {
          if (elect_one_sync()) {
            for (int j = 0; j < mmas_per_stage; ++j)
              tc_mma<kTf32>(tmem, a0 + 2 * (j & 3), b0 + 2 * (j & 3), idesc, (st | j | f) ? 1u : 0u);
            if (!sparse_commit || (s & 3) == 3) tc_commit(smem_u32(&bars[s]));
          }
        } else if (lane == 0) {
          for (int j = 0; j < mmas_per_stage; ++j)
            tc_mma<kTf32>(tmem, a0 + 2 * (j & 3), b0 + 2 * (j & 3), idesc, (st | j | f) ? 1u : 0u);
          if (!sparse_commit || (s & 3) == 3) tc_commit(smem_u32(&bars[s]));
        }
        if (variant & 128) mbar_wait(smem_u32(&bars[s]), (st >> 3) & 1);
        if (whole_warp) __syncwarp();
      }
      for (int st = (stages > 8 ? stages - 8 : 0); st < stages; ++st)
        if (!(variant & 128) && (!sparse_commit || (st & 3) == 3)) mbar_wait(smem_u32(&bars[st & 7]), (st >> 3) & 1);
    }
    const long long t1 = clock64();
    if (lane == 0) cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
  }
}

int main() {
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  CK(cudaFuncSetAttribute(issue_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  CK(cudaFuncSetAttribute(issue_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  unsigned long long *cyc, h[256];
  CK(cudaMalloc(&cyc, sizeof(unsigned long long) * sms));
  const int stages = 4096;
  printf("%-6s %4s %5s %8s %12s %10s\n", "kind", "N", "mmas", "variant", "cyc/stage", "cyc/mma");
  for (int tf32 = 0; tf32 < 2; ++tf32)
    for (int n : {16, 32, 64, 128})
      for (int mmas : {4, 6, 8, 12})
        for (int variant : {256, 256 + 512, 256 + 1024}) {
          if (tf32 && (variant & 512)) continue;    // the TS form is f16 only here
          if (!tf32 && (variant & 1024)) continue;  // the mix is the fp32 path's
          for (int rep = 0; rep < 2; ++rep) {
            if (tf32) issue_kernel<true><<<sms, 128, 64 * 1024>>>(n, mmas, stages, variant, cyc);
            else issue_kernel<false><<<sms, 128, 64 * 1024>>>(n, mmas, stages, variant, cyc);
          }
          CK(cudaDeviceSynchronize());
          CK(cudaMemcpy(h, cyc, sizeof(unsigned long long) * sms, cudaMemcpyDeviceToHost));
          unsigned long long mx = 0;
          for (int i = 0; i < sms; ++i) mx = h[i] > mx ? h[i] : mx;
          printf("%-6s %4d %5d %8d %12.1f %10.1f\n", tf32 ? "tf32" : "f16", n, mmas, variant, (double)mx / stages,
                 (double)mx / stages / mmas);
        }
  return 0;
}
