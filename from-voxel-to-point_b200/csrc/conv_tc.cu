// Sparse convolution forward on the 5th-generation tensor cores (tcgen05), sm_100a only.
//
// Replaces, per layer, the reference's 26x(gather kernel + cuBLAS GEMM + scatter-add kernel) + bias +
// BatchNorm1d + ReLU (pcdet/ops/spconv/include/spconv/spconv_ops.h:294-357, conv.py:223-224,
// spconv_backbone.py:25-27,57-66) with ONE persistent, warp-specialised, output-stationary implicit GEMM:
//
//   CTA tile      128 output rows x Cout, accumulator in TMEM (fp32, Cout columns, double buffered so the
//                 epilogue of tile t overlaps the contraction of tile t+1)
//   K loop        the kernel offsets that have at least one neighbour in the tile  x  Cin in 128-byte slices
//   A operand     the gathered input rows: warps 4-11 issue 16-byte cp.async straight into the UMMA canonical
//                 K-major (no-swizzle) layout and hand the stage over with cp.async.mbarrier.arrive (the
//                 arrival fires when the copies land, the threads never wait); a missing neighbour is a
//                 zero-filled cp.async, so there is no predication in the MMA and no scatter afterwards
//   B operand     W[k] slices, pre-packed once per layer into the exact shared-memory image and pulled with
//                 one TMA bulk copy (cp.async.bulk, mbarrier complete_tx) per stage
//   MMA           one elected lane of warp 12 issues tcgen05.mma (kind::f16 for bf16, kind::tf32 for fp32),
//                 tcgen05.commit releases the smem stage / publishes the accumulator through mbarriers
//   epilogue      warps 0-3 read TMEM with tcgen05.ld (one accumulator row per thread), apply
//                 bias + folded BatchNorm + residual + ReLU and store 16-byte vectors
//
// fp32 path = 3xTF32: A is split in shared memory into hi = rn_tf32(A) and lo = rn_tf32(A - hi) by four
// transform warps (13-16) between the gather and the MMA, W is packed as hi/lo images, and each K step issues A_lo*W_hi + A_hi*W_lo + A_hi*W_hi.  The
// dropped lo*lo term and the rounding of the lo parts are O(2^-22) relative and unbiased, far inside 1e-4.
#include "common.cuh"

namespace fv2p {
namespace {

constexpr int kTileM = 128;
constexpr int kEpiThreads = 128;     // warps 0-3
constexpr int kGatherThreads = 256;  // warps 4-11: two per scheduler so one warp's smem/L2 latency hides behind the other
constexpr int kGatherWarps = kGatherThreads / 32;
constexpr int kMmaWarp = (kEpiThreads + kGatherThreads) / 32;
constexpr int kXformThreads = 128;  // warps 13-16, fp32 (3xTF32) kernels only
constexpr int kTcThreadsBase = kEpiThreads + kGatherThreads + 32;
constexpr int kMaxStages = 8;
constexpr int kSmemBudget = 212 * 1024;

// ------------------------------------------------------------------------------------------ PTX
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
// The mbarrier arrival is triggered when every cp.async this thread issued so far has landed (.noinc: it is one
// of the arrivals the barrier was initialised with).  Non-blocking: the thread moves on to the next stage.
__device__ __forceinline__ void cp_async_arrive(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
template <bool kTf32>
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                       uint32_t accumulate) {
  if constexpr (kTf32) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
// Round-to-nearest fp32 -> tf32 (kept in an fp32 container).  Truncation instead would bias every product the
// same way and the bias grows linearly with the 27*Cin-term reduction; rounding keeps the split unbiased.
__device__ __forceinline__ float tf32_rn(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
// 16 consecutive fp32 accumulator columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float *v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// UMMA shared-memory descriptor, K-major, no swizzle (cute/arch/mma_sm100_desc.hpp SmemDescriptor):
// [0,14) start>>4, [16,30) leading byte offset>>4 (between the two 16-byte K chunks of one MMA),
// [32,46) stride byte offset>>4 (between 8-row groups), [46,48) version=1, [61,64) layout type 0.
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) |
         (1ull << 46);
}
// Instruction descriptor (InstrDescriptor): c=F32 [4,6), a/b format [7,10)/[10,13), K-major both,
// n>>3 at [17,23), m>>4 at [24,29).
__host__ __device__ constexpr uint32_t instr_desc(int n, bool tf32) {
  return (1u << 4) | ((tf32 ? 2u : 1u) << 7) | ((tf32 ? 2u : 1u) << 10) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(kTileM >> 4) << 24);
}

template <bool kTf32, int N>
struct Cfg {
  static constexpr int kABytes = kTileM * 128;                   // one 128-byte slice per row
  static constexpr int kWBytes = N * 128;
  static constexpr int kStageBytes = (kTf32 ? 2 : 1) * (kABytes + kWBytes);
  static constexpr int kNbrBytes = FV2P_MAX_KVOL * kTileM * 4;
  static constexpr int kStagesRaw = (kSmemBudget - kNbrBytes - 1024) / kStageBytes;
  static constexpr int kStages = kStagesRaw > kMaxStages ? kMaxStages : kStagesRaw;
  static constexpr int kSmemBytes = kStages * kStageBytes + kNbrBytes + 1024 + 1024;  // + barriers + align slack
  static constexpr int kTmemCols = 2 * N < 32 ? 32 : 2 * N;  // N in {16,32,64,128} -> power of two
  static constexpr int kThreads = kTcThreadsBase + (kTf32 ? kXformThreads : 0);
  static_assert(kStages >= 3, "pipeline too shallow");
};

struct Epilogue {
  const float *bias, *scale, *shift;
  const void *residual;
  void *out;
  int relu;
};

template <bool kTf32, int N>
__global__ void __launch_bounds__((Cfg<kTf32, N>::kThreads), 1)
conv_tc_kernel(const void *__restrict__ features, const uint8_t *__restrict__ wpacked,
               const int *__restrict__ nbr, int64_t nbr_stride, int kvol, int64_t n_out_cap,
               const int *__restrict__ n_out_dev, int cin, Epilogue ep) {
  using C = Cfg<kTf32, N>;
  constexpr int kElem = kTf32 ? 4 : 2;
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t *stage_base = smem;
  int *nbr_s = reinterpret_cast<int *>(smem + C::kStages * C::kStageBytes);
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + C::kStages * C::kStageBytes + C::kNbrBytes);
  // barrier layout: full[kStages], empty[kStages], landed[kStages] (fp32 only), tmem_full[2], tmem_empty[2]
  const uint32_t bar_full = smem_u32(bars), bar_empty = bar_full + 8 * C::kStages;
  const uint32_t bar_landed = bar_empty + 8 * C::kStages;
  const uint32_t bar_tfull = bar_landed + 8 * C::kStages, bar_tempty = bar_tfull + 16;
  volatile int *stage_flags = reinterpret_cast<volatile int *>(bars + 3 * C::kStages + 4);  // [kStages]
  volatile uint32_t *tile_mask = reinterpret_cast<volatile uint32_t *>(stage_flags + C::kStages);
  uint32_t *tmem_slot = const_cast<uint32_t *>(tile_mask) + 1;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int n_out = n_out_dev ? *n_out_dev : (int)n_out_cap;
  if (n_out > n_out_cap) n_out = (int)n_out_cap;
  const int n_tiles = (n_out + kTileM - 1) / kTileM;
  const int row_bytes = min(cin * kElem, 128);  // bytes of one row consumed per stage
  const int chunks = row_bytes >> 4;            // 16-byte chunks per row per stage
  const int slices = (cin * kElem) / row_bytes; // stages per kernel offset
  const uint32_t w_stage_bytes = (uint32_t)N * row_bytes;

  if (threadIdx.x == 0) {
    for (int s = 0; s < C::kStages; ++s) {
      // bf16: 256 cp.async arrivals + the expect_tx arrive of the weight copy.  fp32: the gathers land on
      // `landed` (+1 plain arrive that publishes the stage flags), the transform warps arrive on `full`.
      mbar_init(bar_full + 8 * s, (kTf32 ? kXformThreads : kGatherThreads) + 1);
      mbar_init(bar_empty + 8 * s, 1);
      mbar_init(bar_landed + 8 * s, kGatherThreads + 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_tfull + 8 * a, 1);
      mbar_init(bar_tempty + 8 * a, kEpiThreads);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)C::kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= 4 && warp < kMmaWarp) {
    // =============================== gather producers ===============================
    const int tid = threadIdx.x - kEpiThreads;
    const uint8_t *feat = static_cast<const uint8_t *>(features);
    const size_t feat_row_bytes = (size_t)cin * kElem;
    uint32_t issued = 0;
    // Item (it, tid) of a stage is the 16-byte unit  it*256 + tid  of the canonical layout
    // ((row/8)*chunks + chunk)*8 + row%8.  With 256 threads the chunk of a thread is constant and only the
    // row advances with `it`, so eight consecutive lanes write 128 contiguous bytes of shared memory (no bank
    // conflicts) while lanes 8 apart read the next 16 bytes of the same global rows (full 32-byte sectors).
    const int cshift = __ffs(chunks) - 1;            // chunks is 2, 4 or 8
    const int my_chunk = (tid >> 3) & (chunks - 1);
    const int my_row0 = ((tid >> (3 + cshift)) << 3) + (tid & 7);
    const int rows_per_it = kGatherThreads >> cshift;  // 128, 64 or 32 rows per pass
    const int iters = chunks >> 1;                     // kTileM * chunks / kGatherThreads
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int row0 = tile * kTileM;
      // ---- neighbour rows of this tile -> smem, and which offsets feed anything
      asm volatile("bar.sync 1, 256;" ::: "memory");  // everyone finished reading nbr_s of the previous tile
      if (tid == 0) *tile_mask = 0u;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      {
        const int r = tid & (kTileM - 1);
        const int row = row0 + r;
        uint32_t mine = 0;
        for (int k = tid >> 7; k < kvol; k += 2) {  // warps 4-7 take the even offsets, 8-11 the odd ones
          int src = -1;
          if (row < n_out) src = __ldg(&nbr[(size_t)k * nbr_stride + row]);
          nbr_s[k * kTileM + r] = src;
          if (__ballot_sync(0xFFFFFFFFu, src >= 0)) mine |= 1u << k;
        }
        if (lane == 0 && mine) atomicOr(const_cast<uint32_t *>(tile_mask), mine);
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      uint32_t mask = *tile_mask;
      if (mask == 0u) mask = 1u;  // a tile nothing feeds still has to produce (zero) accumulators
      const uint32_t first_k = __ffs(mask) - 1;
      const uint32_t last_k = 31 - __clz(mask);
      while (mask) {
        const int k = __ffs(mask) - 1;
        mask &= mask - 1;
        int src_row[4];
#pragma unroll
        for (int it = 0; it < 4; ++it)
          src_row[it] = it < iters ? nbr_s[k * kTileM + it * rows_per_it + my_row0] : -1;
        for (int sl = 0; sl < slices; ++sl) {
          const uint32_t s = issued % C::kStages;
          mbar_wait(bar_empty + 8 * s, ((issued / C::kStages) & 1) ^ 1);
          uint8_t *a_dst = stage_base + (size_t)s * C::kStageBytes;
          if (tid == 0) {
            stage_flags[s] = ((k == (int)first_k && sl == 0) ? 1 : 0) | ((k == (int)last_k && sl == slices - 1) ? 2 : 0);
            const uint8_t *wsrc = wpacked + ((size_t)k * slices + sl) * w_stage_bytes * (kTf32 ? 2 : 1);
            uint8_t *w_dst = a_dst + (kTf32 ? 2 : 1) * C::kABytes;
            mbar_arrive_expect_tx(bar_full + 8 * s, w_stage_bytes * (kTf32 ? 2 : 1));
            bulk_g2s(smem_u32(w_dst), wsrc, w_stage_bytes * (kTf32 ? 2 : 1), bar_full + 8 * s);
            if constexpr (kTf32) mbar_arrive(bar_landed + 8 * s);
          }
          const uint32_t a_u32 = smem_u32(a_dst) + tid * 16;
          const size_t col_off = (size_t)sl * row_bytes + my_chunk * 16;
#pragma unroll
          for (int it = 0; it < 4; ++it) {
            if (it < iters) {
              const int src = src_row[it];
              const uint8_t *p = feat + (src >= 0 ? (size_t)src * feat_row_bytes + col_off : 0);
              cp_async16(a_u32 + it * (kGatherThreads * 16), p, src >= 0 ? 16u : 0u);
            }
          }
          cp_async_arrive(kTf32 ? bar_landed + 8 * s : bar_full + 8 * s);
          ++issued;
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // =============================== MMA issuer ===============================
    constexpr uint32_t idesc = instr_desc(N, kTf32);
    const uint32_t sbo = (uint32_t)chunks * 128u;
    const int ksteps = row_bytes >> 5;  // 32 bytes of K per MMA (16 bf16 / 8 tf32)
    uint32_t consumed = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      mbar_wait(bar_tempty + 8 * acc, acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * N);
      bool last = false;
      while (!last) {
        const uint32_t s = consumed % C::kStages;
        mbar_wait(bar_full + 8 * s, (consumed / C::kStages) & 1);
        tc_fence_after();
        const int flags = stage_flags[s];
        last = (flags & 2) != 0;
        if (lane == 0) {
          const uint32_t a_addr = smem_u32(stage_base + (size_t)s * C::kStageBytes);
          const uint32_t w_addr = a_addr + (kTf32 ? 2 : 1) * C::kABytes;
          uint32_t accumulate = (flags & 1) ? 0u : 1u;
          for (int j = 0; j < ksteps; ++j) {
            const uint64_t a_hi = smem_desc(a_addr + j * 256, 128, sbo);
            const uint64_t b_hi = smem_desc(w_addr + j * 256, 128, sbo);
            if constexpr (kTf32) {
              const uint64_t a_lo = smem_desc(a_addr + C::kABytes + j * 256, 128, sbo);
              const uint64_t b_lo = smem_desc(w_addr + w_stage_bytes + j * 256, 128, sbo);
              tc_mma<true>(d_tmem, a_lo, b_hi, idesc, accumulate);
              tc_mma<true>(d_tmem, a_hi, b_lo, idesc, 1u);
              tc_mma<true>(d_tmem, a_hi, b_hi, idesc, 1u);
            } else {
              tc_mma<false>(d_tmem, a_hi, b_hi, idesc, accumulate);
            }
            accumulate = 1u;
          }
          tc_commit(bar_empty + 8 * s);            // smem stage reusable once these MMAs retire
          if (last) tc_commit(bar_tfull + 8 * acc);  // accumulator complete
        }
        __syncwarp();
        ++consumed;
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  } else if (warp > kMmaWarp) {
    // =============================== fp32 split (warps 13-16, 3xTF32 kernels only) ===============================
    // A_raw <- hi = rn_tf32(x), A_lo <- rn_tf32(x - hi).  Done by warps that have no cp.async in flight, so
    // the proxy fence that makes the generic-proxy stores visible to the tensor core does not stall on gathers.
    if constexpr (kTf32) {
      const int t = threadIdx.x - (kTcThreadsBase);
      uint32_t done = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        bool last = false;
        while (!last) {
          const uint32_t s = done % C::kStages;
          mbar_wait(bar_landed + 8 * s, (done / C::kStages) & 1);
          last = (stage_flags[s] & 2) != 0;
          uint8_t *a_hi = stage_base + (size_t)s * C::kStageBytes;
          uint8_t *a_lo = a_hi + C::kABytes;
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            if (it < chunks) {
              const int off = (it * kXformThreads + t) * 16;
              float4 x = *reinterpret_cast<float4 *>(a_hi + off);
              float4 h, l;
              h.x = tf32_rn(x.x), h.y = tf32_rn(x.y), h.z = tf32_rn(x.z), h.w = tf32_rn(x.w);
              l.x = tf32_rn(x.x - h.x), l.y = tf32_rn(x.y - h.y), l.z = tf32_rn(x.z - h.z), l.w = tf32_rn(x.w - h.w);
              *reinterpret_cast<float4 *>(a_hi + off) = h;
              *reinterpret_cast<float4 *>(a_lo + off) = l;
            }
          }
          fence_proxy_async();
          mbar_arrive(bar_full + 8 * s);
          ++done;
        }
      }
    }
  } else {
    // =============================== epilogue (warps 0-3) ===============================
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      mbar_wait(bar_tfull + 8 * acc, acc_phase);
      tc_fence_after();
      const int row = tile * kTileM + warp * 32 + lane;
      const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(acc * N);
#pragma unroll 1
      for (int c0 = 0; c0 < N; c0 += 16) {
        float v[16];
        tmem_ld16(taddr + c0, v);  // warp-collective: executed by all lanes even for rows past the end
        if (row < n_out) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float x = v[i];
            if (ep.bias) x += __ldg(&ep.bias[c0 + i]);
            if (ep.scale) x = fmaf(x, __ldg(&ep.scale[c0 + i]), __ldg(&ep.shift[c0 + i]));
            v[i] = x;
          }
          if constexpr (kTf32) {
            float *o = static_cast<float *>(ep.out) + (size_t)row * N + c0;
            if (ep.residual) {
              const float4 *rs = reinterpret_cast<const float4 *>(static_cast<const float *>(ep.residual) +
                                                                 (size_t)row * N + c0);
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                float4 t = __ldg(rs + q);
                v[4 * q] += t.x, v[4 * q + 1] += t.y, v[4 * q + 2] += t.z, v[4 * q + 3] += t.w;
              }
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              float4 t;
              t.x = v[4 * q], t.y = v[4 * q + 1], t.z = v[4 * q + 2], t.w = v[4 * q + 3];
              if (ep.relu) t.x = fmaxf(t.x, 0.f), t.y = fmaxf(t.y, 0.f), t.z = fmaxf(t.z, 0.f), t.w = fmaxf(t.w, 0.f);
              reinterpret_cast<float4 *>(o)[q] = t;
            }
          } else {
            __nv_bfloat16 *o = static_cast<__nv_bfloat16 *>(ep.out) + (size_t)row * N + c0;
            if (ep.residual) {
              const uint4 *rs = reinterpret_cast<const uint4 *>(static_cast<const __nv_bfloat16 *>(ep.residual) +
                                                               (size_t)row * N + c0);
#pragma unroll
              for (int q = 0; q < 2; ++q) {
                uint4 t = __ldg(rs + q);
                const __nv_bfloat162 *h = reinterpret_cast<const __nv_bfloat162 *>(&t);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  float2 f = __bfloat1622float2(h[e]);
                  v[8 * q + 2 * e] += f.x;
                  v[8 * q + 2 * e + 1] += f.y;
                }
              }
            }
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              uint4 t;
              __nv_bfloat162 *h = reinterpret_cast<__nv_bfloat162 *>(&t);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                float a = v[8 * q + 2 * e], b = v[8 * q + 2 * e + 1];
                if (ep.relu) a = fmaxf(a, 0.f), b = fmaxf(b, 0.f);
                h[e] = __floats2bfloat162_rn(a, b);
              }
              reinterpret_cast<uint4 *>(o)[q] = t;
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(bar_tempty + 8 * acc);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)C::kTmemCols)
                 : "memory");
  }
}

// Packs W [K,cin,cout] fp32 into per-(offset, 128-byte slice) shared-memory images of the B operand
// (N x Kslice, K-major canonical layout: ((n/8)*chunks + chunk)*8 + n%8 sixteen-byte units).
template <bool kTf32>
__global__ void __launch_bounds__(kThreads)
pack_weight_kernel(const float *__restrict__ w, int kvol, int cin, int cout, uint8_t *packed) {
  constexpr int kElem = kTf32 ? 4 : 2;
  constexpr int kPerChunk = 16 / kElem;
  const int row_bytes = min(cin * kElem, 128);
  const int chunks = row_bytes >> 4;
  const int slices = (cin * kElem) / row_bytes;
  const int per_slice = row_bytes / kElem;  // input channels per slice
  const int64_t total = (int64_t)kvol * cin * cout;
  const size_t image = (size_t)cout * row_bytes;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int n = (int)(e % cout);
    const int ci = (int)((e / cout) % cin);
    const int k = (int)(e / ((int64_t)cout * cin));
    const int sl = ci / per_slice, within = ci % per_slice;
    const int c = within / kPerChunk, t = within % kPerChunk;
    const size_t unit = ((size_t)(n >> 3) * chunks + c) * 8 + (n & 7);
    const float val = w[e];
    uint8_t *base = packed + ((size_t)k * slices + sl) * image * (kTf32 ? 2 : 1);
    if constexpr (kTf32) {
      const float hi = tf32_rn(val);
      reinterpret_cast<float *>(base + unit * 16)[t] = hi;
      reinterpret_cast<float *>(base + image + unit * 16)[t] = tf32_rn(val - hi);
    } else {
      reinterpret_cast<__nv_bfloat16 *>(base + unit * 16)[t] = __float2bfloat16_rn(val);
    }
  }
}

bool tc_shape_ok(int cin, int cout) {
  const bool n_ok = cout == 16 || cout == 32 || cout == 64 || cout == 128;
  const bool k_ok = cin == 16 || cin == 32 || cin == 64 || cin == 128 || cin == 256;
  return n_ok && k_ok;
}

template <bool kTf32, int N>
int launch_one(const void *features, const void *weight, const int *nbr, int64_t nbr_stride, int kvol,
               int64_t n_out_cap, const int *n_out_dev, int cin, const Epilogue &ep, cudaStream_t stream) {
  using C = Cfg<kTf32, N>;
  static bool configured = false;
  if (!configured) {
    int st = cuda_status(cudaFuncSetAttribute(conv_tc_kernel<kTf32, N>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              C::kSmemBytes),
                         "conv_fwd(tc) smem attribute");
    if (st) return st;
    configured = true;
  }
  int64_t tiles = (n_out_cap + kTileM - 1) / kTileM;
  int grid = (int)(tiles < sm_count() ? tiles : sm_count());
  if (grid < 1) grid = 1;
  conv_tc_kernel<kTf32, N><<<grid, C::kThreads, C::kSmemBytes, stream>>>(
      features, static_cast<const uint8_t *>(weight), nbr, nbr_stride, kvol, n_out_cap, n_out_dev, cin, ep);
  return cuda_status(cudaGetLastError(), "conv_fwd(tc)");
}

}  // namespace

int launch_conv_tc(const void *features, const void *weight, const int *nbr, int64_t nbr_stride, int kvol,
                   int64_t n_out_cap, const int *n_out_dev, int cin, int cout, const float *bias,
                   const float *scale, const float *shift, const void *residual, int relu, int mode, void *out,
                   cudaStream_t stream) {
  if (!tc_shape_ok(cin, cout)) {
    set_error("conv_fwd: tensor-core modes need cin in {16,32,64,128,256} and cout in {16,32,64,128} (got %d->%d)",
              cin, cout);
    return FV2P_ERR_INVALID;
  }
  Epilogue ep{bias, scale, shift, residual, out, relu};
  const bool tf32 = mode == FV2P_MODE_TF32X3_TC;
#define FV2P_TC(NN)                                                                                           \
  return tf32 ? launch_one<true, NN>(features, weight, nbr, nbr_stride, kvol, n_out_cap, n_out_dev, cin, ep,  \
                                     stream)                                                                 \
              : launch_one<false, NN>(features, weight, nbr, nbr_stride, kvol, n_out_cap, n_out_dev, cin, ep, \
                                      stream)
  switch (cout) {
    case 16: FV2P_TC(16);
    case 32: FV2P_TC(32);
    case 64: FV2P_TC(64);
    default: FV2P_TC(128);
  }
#undef FV2P_TC
}

}  // namespace fv2p

using namespace fv2p;

extern "C" size_t fv2p_pack_weight_bytes(int kvol, int cin, int cout, int mode) {
  if (kvol < 1 || kvol > FV2P_MAX_KVOL || !tc_shape_ok(cin, cout)) return 0;
  if (mode == FV2P_MODE_BF16_TC) return (size_t)kvol * cin * cout * 2;
  if (mode == FV2P_MODE_TF32X3_TC) return (size_t)kvol * cin * cout * 4 * 2;
  return 0;
}

extern "C" int fv2p_pack_weight(const float *weight_f32, int kvol, int cin, int cout, int mode, void *packed,
                                fv2p_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FV2P_REQUIRE(weight_f32 && packed, "pack_weight: null pointer argument");
  FV2P_REQUIRE(fv2p_pack_weight_bytes(kvol, cin, cout, mode) > 0, "pack_weight: unsupported shape %d x %d->%d mode %d",
               kvol, cin, cout, mode);
  if (mode == FV2P_MODE_TF32X3_TC)
    pack_weight_kernel<true><<<persistent_grid(), kThreads, 0, stream>>>(weight_f32, kvol, cin, cout,
                                                                         static_cast<uint8_t *>(packed));
  else
    pack_weight_kernel<false><<<persistent_grid(), kThreads, 0, stream>>>(weight_f32, kvol, cin, cout,
                                                                          static_cast<uint8_t *>(packed));
  FV2P_LAUNCH_CHECK("pack_weight");
  return FV2P_OK;
}
