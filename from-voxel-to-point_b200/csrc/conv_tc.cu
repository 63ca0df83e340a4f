// Sparse convolution forward on the 5th-generation tensor cores (tcgen05), sm_100a only.
//
// Replaces, per layer, the reference's 26x(gather kernel + cuBLAS GEMM + scatter-add kernel) + bias +
// BatchNorm1d + ReLU (pcdet/ops/spconv/include/spconv/spconv_ops.h:294-357, conv.py:223-224,
// spconv_backbone.py:25-27,57-66) with ONE persistent, warp-specialised, output-stationary implicit GEMM:
//
//   CTA tile      128 output rows x Cout, accumulator in TMEM (fp32, Cout columns, double buffered so the
//                 epilogue of tile t overlaps the contraction of tile t+1)
//   K loop        the kernel offsets that have at least one neighbour in the tile  x  Cin in 128-byte slices
//   A operand     the gathered input rows in the 128B/64B/32B-swizzled K-major tile tcgen05.mma reads.  Two
//                 interchangeable producers (profiles/r1_notes.md has the measurements):
//                   * LSU (default): 16-byte cp.async (LDGSTS), eight lanes per row so that every warp instruction
//                     reads whole 128-byte lines and, thanks to the swizzle, writes conflict-free; a missing
//                     neighbour is a zero-fill cp.async; the stage is handed over with
//                     cp.async.mbarrier.arrive, so the issuing warp never waits for its copies;
//                   * TMA: cp.async.bulk.tensor ... tile::gather4 pulls four arbitrary feature rows per
//                     instruction; a missing neighbour is an out-of-range row index that the TMA zero fills.
//                 Warps 4-11; each ring slot is owned by one warp (two, half the rows each, when the ring is
//                 shorter than eight), so the waits and arrivals of one stage overlap the copy issue of the next
//                 ones.  Either way the MMA needs no predication and nothing is scattered afterwards.
//   B operand     W[k] slices, pre-packed once per layer into the exact (swizzled) shared-memory image and
//                 pulled with one TMA bulk copy per stage by warp 14 into a ring of their own (three slots; the
//                 gathered tiles have a deeper ring of 16 KB slots, since their round trip is the long one)
//   MMA           one elected thread of warp 12 runs the whole tcgen05.mma issue loop (kind::f16 for bf16,
//                 A from TMEM for fp32); tcgen05.commit releases the smem stage / publishes the accumulator
//   epilogue      warps 0-3 read TMEM with tcgen05.ld (one accumulator row per thread), apply
//                 bias + folded BatchNorm + residual + ReLU and store 16-byte vectors
//   tiles         warp 13 is the scheduler: it claims tiles in the order fv2p_sort_rows_by_mask computed (most
//                 active offsets first = longest-processing-time-first list scheduling; with mask-sorted rows a
//                 tile takes 1..27 offsets, and dealing tiles round-robin left the busiest CTA with 1.3-2.2x the
//                 mean work: 180 stages against a mean of 81 on KITTI's 128->128 layers), and streams the rows of
//                 the neighbour map the tile needs - only the offsets in its mask - into a shared-memory ring with
//                 bulk copies, a few tiles ahead of the producers.  Tile ids reach the epilogue through a small
//                 ring; a sentinel tile / stage ends the stream for every role.
//
// fp32 path = split bf16 ("3xBF16"): x = hi + lo (+ <= 2^-18 |x|) with hi = bf16_rn(x), lo = bf16_rn(x - hi), the same
// for W.  Per 16-channel K step THREE kind::f16 MMAs whose A operand sits in TMEM:
//   A_hi*W_hi + A_hi*W_lo + A_lo*W_hi        (the dropped A_lo*W_lo is <= 2^-18 of the product)
// - the transform warps (15-18, one per TMEM lane quadrant, thread = tile row) read the landed fp32 row from shared
// memory, split it and write [hi(16) | lo(16)] as bf16 pairs to the stage's TMEM columns with tcgen05.st (32 columns
// per A slot); B = the weight row [W_hi | W_lo] (bf16) of the packed image.  Measured against a float64 contraction:
// 3.5e-6 ... 1.3e-5 relative at 27*16 ... 27*128 terms, 1.0-1.3e-5 on the stride-8 output of the 21-layer backbone
// at benchmark size (bar 1e-4).
// History: (1) hi/lo tiles in shared memory and three tf32 passes: bound by the shared-memory pipe (16 KB landed +
// 48 KB split traffic + 12 x 8 KB operand reads per stage ~ 1500 cycles at 128 B/cycle; cvt.rna pairs another ~500);
// (2) round 1 and most of round 2, "3xTF32": A_hi*W_hi as kind::tf32 straight from the landed tile (a tf32 MMA ignores
// the low 13 mantissa bits) + ONE bf16 MMA of K = 16 per 8 channels for both corrections, A = [lo | hi] in TMEM - two
// MMAs per EIGHT channels, the tf32 one at half rate and reading 4 KB of A from shared memory; (3) now three bf16
// MMAs per SIXTEEN channels, none of which reads A from shared memory: 25 % less tensor time on the 128-wide layers,
// ~25 % less issue / operand-feed time on the narrow ones (micro/mma_issue_bench.cu, N = 32: 6 TS MMAs 363 cycles
// against 459 for the 4 + 4 mix), half the weight bytes - and a SMALLER error (fewer accumulating MMAs, each of which
// truncates toward zero): waymo_b4 fp32 2.98 -> 2.72 ms per step.
#include <cuda.h>

#include "common.cuh"

namespace fv2p {
namespace {

constexpr int kTileM = 128;
constexpr int kEpiThreads = 128;   // warps 0-3
constexpr int kProdThreads = 256;  // warps 4-11: neighbour prefetch + gather issue (each ring slot has one owner warp)
constexpr int kProdWarps = kProdThreads / 32;
constexpr int kMmaWarp = (kEpiThreads + kProdThreads) / 32;  // warp 12
constexpr int kSchedWarp = kMmaWarp + 1;                     // warp 13: tile scheduler + neighbour-map loader
constexpr int kWLoadWarp = kSchedWarp + 1;                   // warp 14: weight-slice loader (W ring)
constexpr int kXformThreads = 128;                           // warps 15-18, fp32 (split bf16) kernels only
constexpr int kTcThreadsBase = kEpiThreads + kProdThreads + 96;
constexpr int kMaxStages = 8;
constexpr int kSmemLimit = 232448;                 // 227 KB of dynamic shared memory per CTA
constexpr int kSmemGuest = 2048;                   // left free so that a geometry CTA (rulebook, grouping) can co-reside
constexpr int kSmemMisc = 2048;                    // barriers and small rings + 1024-byte alignment slack
// Epilogue staging: each of the four epilogue warps transposes 16-column chunks of its 32 accumulator rows through
// shared memory, so that its global accesses are whole contiguous row segments instead of one row per lane.
constexpr int kEpiRowPitch = 80;                   // 64 bytes of fp32 + 16: 16-byte accesses of 8 rows hit 8 bank groups
constexpr int kEpiStageBytes = 4 * 32 * kEpiRowPitch;
constexpr int kNbrBufInts = FV2P_MAX_KVOL * 128;   // one tile of the neighbour map: [offset][128 rows]
constexpr int kTileRing = 16;  // > kMaxStages + 2: how far the producers can run ahead of the epilogue, in tiles
constexpr int kFlagFirst = 1, kFlagLast = 2, kFlagStop = 4;

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
// Four rows (r0..r3) x one box of columns starting at `col` -> 4 consecutive swizzled smem rows.
__device__ __forceinline__ void tma_gather4(uint32_t dst, const CUtensorMap *map, int col, int r0, int r1, int r2,
                                            int r3, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(dst),
      "l"(map), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
// The mbarrier arrival fires when every cp.async this thread issued so far has landed (.noinc: it is one of the
// arrivals the barrier was initialised with).  Non-blocking.
__device__ __forceinline__ void cp_async_arrive(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
// One lane of the (converged) warp; ptxas keeps what the elected thread computes in uniform registers, which is what
// the tcgen05 / bulk-copy instructions take (a `lane == 0` branch makes it wrap each of them in an ELECT loop).
__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred = 0, laneid = 0;
  asm volatile(
      "{\n\t.reg .b32 %%rx;\n\t.reg .pred %%px;\n\t"
      "elect.sync %%rx|%%px, %2;\n\t"
      "@%%px mov.s32 %1, 1;\n\t"
      "mov.s32 %0, %%rx;\n\t}"
      : "+r"(laneid), "+r"(pred)
      : "r"(0xFFFFFFFFu));
  return pred != 0;
}
// Programmatic dependent launch: let the next kernel on the stream start its set-up / wait for the previous one's
// results (no-ops when the kernel was not launched with the attribute).
__device__ __forceinline__ void griddep_launch_dependents() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
template <bool kUnused>
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                       uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// A operand from tensor memory (lane = tile row, two 16-bit elements per 32-bit column, K-consecutive), B from
// shared memory
__device__ __forceinline__ void tc_mma_ts_f16(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo_elem, float hi_elem) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo_elem, hi_elem);  // .x (low half) = first K element
  return *reinterpret_cast<uint32_t *>(&v);
}
// 16 consecutive 32-bit columns of this thread's TMEM lane (the warp covers its 32-lane quadrant)
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t *r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
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

// Explicit shared-state-space accesses: pointers derived from the aligned dynamic shared memory base are generic to
// the compiler, which then emits generic LD / ST (address-space check per access) in the gather and epilogue loops.
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ float4 lds128f(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ int4 lds128i(uint32_t addr) {
  int4 v;
  asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ int lds32i(uint32_t addr) {
  int v;
  asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}

// Small int arrays in shared memory that the roles use as mailboxes (ordered by the mbarriers around them): volatile
// accesses through the shared window (a `volatile int *` compiles to generic LD/ST.STRONG.SYS).
struct SmemInts {
  uint32_t base;
  __device__ __forceinline__ int get(int i) const {
    int v;
    asm volatile("ld.volatile.shared.s32 %0, [%1];" : "=r"(v) : "r"(base + 4u * (uint32_t)i) : "memory");
    return v;
  }
  __device__ __forceinline__ void set(int i, int v) const {
    asm volatile("st.volatile.shared.s32 [%0], %1;" ::"r"(base + 4u * (uint32_t)i), "r"(v) : "memory");
  }
};

// Split form for software pipelining: the load is issued into `r` without waiting; tmem_ld_wait makes every register
// of `r` depend on the tcgen05.wait::ld (in/out operands), so no consumer can be scheduled above it.
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t *r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait(uint32_t *r) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :
               : "memory");
}

// UMMA shared-memory descriptor, K-major, swizzled (cute/arch/mma_sm100_desc.hpp SmemDescriptor):
// [0,14) start>>4, [16,30) leading byte offset>>4 (1 for swizzled K-major), [32,46) stride byte offset>>4
// (distance between 8-row groups = 8 * row_bytes), [46,48) version=1, [61,64) layout: 2/4/6 = 128B/64B/32B swizzle.
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t sbo, uint32_t layout) {
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46) |
         ((uint64_t)layout << 61);
}
// Instruction descriptor (InstrDescriptor): c=F32 [4,6), a/b format [7,10)/[10,13), K-major both,
// n>>3 at [17,23), m>>4 at [24,29).
__host__ __device__ constexpr uint32_t instr_desc(int n, bool tf32) {
  return (1u << 4) | ((tf32 ? 2u : 1u) << 7) | ((tf32 ? 2u : 1u) << 10) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(kTileM >> 4) << 24);
}

// Timing experiments only (profiles/run_layer.py --debug): 3 = no weight copy, 4 = no MMA, 6 = no epilogue body,
// 7 = CTA 0 stamps %globaltimer at entry/exit into g_tc_stamps (profiles/timeline.py).  0 in production.
__device__ int g_tc_debug = 0;
constexpr int kStampSlots = 256;
__device__ unsigned long long g_tc_stamps[kStampSlots][4];  // start, end, cin * 1000 + N, output pointer
__device__ unsigned int g_tc_stamp_n = 0;
#ifdef FV2P_TC_TIMERS
// Role timers (profiling builds only: FV2P_EXTRA_NVCC_FLAGS=-DFV2P_TC_TIMERS python build.py --force).
// Per CTA: 0 total, 1 stages, 2 tiles, 3 producer-warp-0 wait(empty), 4 producer-warp-0 issue, 5 producer tile
// prologue, 6 MMA wait(full), 7 MMA wait(tmem empty), 8 MMA issue, 9 epilogue wait, 10 epilogue work,
// 11 transform wait, 12 transform work.
__device__ unsigned long long g_tc_timers[160][16];
#define TC_T0() const long long tc_t0_ = clock64()
#define TC_ACC(var) var += clock64() - tc_t0_
#define TC_TIMER_DECL(var) long long var = 0
#else
#define TC_T0()
#define TC_ACC(var)
#define TC_TIMER_DECL(var)
#endif
__device__ __forceinline__ unsigned long long global_timer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// Host-side choice of the A-tile producer: -1 = auto (measured best per shape), 0 = LSU (cp.async), 1 = TMA gather4.
int g_tc_gather_mode = -1;
int g_tc_pdl = 1;  // programmatic dependent launch of the tensor-core conv kernels (debug switch: fv2p_debug_pdl)

// kPacked: rows narrower than 128 bytes (bf16 cin 16/32, fp32 cin 16).  The first generation spent one pipeline stage
// per kernel offset whatever the row width, and the narrow layers ran at the issue loop's fixed cost per stage
// (~320 cycles of MMA-thread time + barrier round trips for 8-32 cycles of tensor work; role timers in
// profiles/r2_notes.md).  Packed stages put G = 128 / row_bytes offsets side by side in ONE 128-byte-swizzled A tile
// - row r = [x(k0) | x(k1) | ...] - and issue one MMA per 32 bytes of it against that offset's weight image, which for
// these widths is small enough (<= 108 KB for all 27 offsets) to stay RESIDENT in shared memory: no weight ring, no
// weight traffic per stage, and 2-4x fewer stages, barriers and commits per tile.
template <bool kFp32, int N, bool kPacked = false>
struct Cfg {
  static constexpr int kABytes = kTileM * 128;  // one (up to) 128-byte slice per row
  static constexpr int kWBytes = N * 128;
  // Two rings.  A: the gathered tile (bf16 rows, or the raw fp32 rows), kStages slots of 16 KB - deep, because a
  // slot's round trip (gather issue -> L2 -> [split] -> MMA -> commit, ~5000 cycles) sets the pace of the kernel.
  // W: the weight slice of the stage (N rows of 128 B: bf16 channels, or for fp32 [W_hi | W_lo] bf16 halves),
  // kWStages slots - one sequential bulk copy each, so three are enough.  (One ring of A+W slots was 4 deep at
  // N = 128.)
  static constexpr int kWSlotBytes = kWBytes;  // fp32: [W_hi | W_lo] bf16 halves of one 128-byte row per output channel
  static constexpr int kWStages = kPacked ? 0 : 3;
  // packed: all (<= 27) offsets' images, sized for the widest packed row (64 bytes: bf16 cin 32 / fp32 cin 16)
  static constexpr int kMaxPackedKvol = 27;
  static constexpr int kWResBytes = kPacked ? kMaxPackedKvol * N * 64 : 0;
  // The scheduler warp streams neighbour-map tiles into a ring of kNbrBufs buffers ahead of the producers.
  static constexpr int kNbrBufs = kPacked ? (kWResBytes > 100000 ? 2 : 3) : (N == 128 ? 2 : (N == 64 ? 3 : 4));
  static constexpr int kNbrBytes = kNbrBufs * kNbrBufInts * 4;
  // fp32: the tensor core never reads the landed fp32 tile; the transform warps split it into the stage's
  // kAColsPerStage TMEM columns (per 16-channel K step: 8 columns of hi pairs, 8 of lo pairs).
  static constexpr int kAColsPerStage = 32;
  // The geometry of later levels (and of the next batch) runs on other streams under the feature pass; it only
  // gets onto an SM if the conv CTA leaves it some shared memory.  The fp32 128-wide layers come last, when little
  // geometry is left, and take it all.
  static constexpr int kSmemAvail = kSmemLimit - ((kFp32 && N == 128) ? 0 : kSmemGuest);
  static constexpr int kStagesSmem =
      (kSmemAvail - kSmemMisc - kEpiStageBytes - kNbrBytes - kWStages * kWSlotBytes - kWResBytes) / kABytes;
  static constexpr int kStagesTmem = kFp32 ? (512 - 2 * N) / kAColsPerStage : kMaxStages;
  static constexpr int kStagesRaw = kStagesSmem < kStagesTmem ? kStagesSmem : kStagesTmem;
  static constexpr int kStages = kStagesRaw > kMaxStages ? kMaxStages : kStagesRaw;
  static constexpr int kWRegion = kPacked ? kWResBytes : kWStages * kWSlotBytes;
  static constexpr int kSmemBytes = kStages * kABytes + kWRegion + kNbrBytes + kEpiStageBytes + kSmemMisc;
  static constexpr int kTmemCols = kFp32 ? 512 : (2 * N < 32 ? 32 : 2 * N);  // power of two
  static constexpr int kThreads = kTcThreadsBase + (kFp32 ? kXformThreads : 0);
  // Register cap: leaves >= 10 K of the 64 K registers so that one 256-thread geometry CTA (<= 40 registers per
  // thread) fits next to the conv CTA.
  static constexpr int kMaxRegs = kFp32 ? 88 : 112;
  static_assert(kStages >= 3, "pipeline too shallow");
};

// Producer warps that share A-ring slot `s` (warp w fills slot w % stages): 1 or 2, half the rows each.
template <int kStages>
__host__ __device__ constexpr int slot_parts(int s) {
  return (s + kStages < kProdWarps) ? 2 : 1;
}

struct Epilogue {
  const float *bias, *scale, *shift;
  const void *residual;
  void *out;
  int relu;
};

template <bool kFp32, int N, bool kPacked>
__global__ void __launch_bounds__((Cfg<kFp32, N, kPacked>::kThreads)) __maxnreg__((Cfg<kFp32, N, kPacked>::kMaxRegs))
conv_tc_kernel(const __grid_constant__ CUtensorMap feat_map, const void *__restrict__ feat_ptr,
               const uint8_t *__restrict__ wpacked,
               const int *__restrict__ nbr, int64_t nbr_stride, const int *__restrict__ row_perm,
               const int *__restrict__ tile_order, int *sched, int kvol, int64_t n_out_cap,
               const int *__restrict__ n_out_dev, int cin, int oob_row, int use_tma_arg, Epilogue ep) {
  using C = Cfg<kFp32, N, kPacked>;
  constexpr int kElem = kFp32 ? 4 : 2;
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t *stage_base = smem;                               // A ring
  uint8_t *w_base = smem + C::kStages * C::kABytes;         // W ring, or the resident images (packed)
  int *nbr_s = reinterpret_cast<int *>(w_base + C::kWRegion);
  uint8_t *epi_stage = reinterpret_cast<uint8_t *>(nbr_s) + C::kNbrBytes;
  uint64_t *bars = reinterpret_cast<uint64_t *>(epi_stage + kEpiStageBytes);
  // barrier layout: full[kStages], empty[kStages], landed[kStages] (fp32 only), tmem_full[2], tmem_empty[2],
  // nbr_full[kNbrBufs], nbr_empty[kNbrBufs], w_full[max(kWStages, 1)] (packed: [0] = the resident images have landed)
  constexpr int kWBars = C::kWStages > 0 ? C::kWStages : 1;
  const uint32_t bar_full = smem_u32(bars), bar_empty = bar_full + 8 * C::kStages;
  const uint32_t bar_landed = bar_empty + 8 * C::kStages;
  const uint32_t bar_tfull = bar_landed + 8 * C::kStages, bar_tempty = bar_tfull + 16;
  const uint32_t bar_nfull = bar_tempty + 16, bar_nempty = bar_nfull + 8 * C::kNbrBufs;
  const uint32_t bar_wfull = bar_nempty + 8 * C::kNbrBufs;
  int *misc = reinterpret_cast<int *>(bars + 3 * C::kStages + 4 + 2 * C::kNbrBufs + kWBars);
  const SmemInts stage_flags{smem_u32(misc)};                   // [kStages]
  const SmemInts stage_ks{stage_flags.base + 4u * C::kStages};  // [kStages] packed: the stage's offsets, one byte each
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(misc + 2 * C::kStages);
  const SmemInts tile_ring{smem_u32(tmem_slot + 1)};            // [kTileRing] tile id per sequence number
  const SmemInts tile_info{tile_ring.base + 4u * kTileRing};    // [kNbrBufs][2]: tile id, offset mask
  // [1] number of tile_ring entries the scheduler has published (release / acquire): lets the epilogue look one tile
  // ahead of the accumulator it is waiting for
  const uint32_t tiles_posted = tile_info.base + 4u * 2u * C::kNbrBufs;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int dbg = g_tc_debug;
#ifdef FV2P_TC_TIMERS
  const long long tc_cta_t0 = clock64();
#endif
  unsigned int stamp_slot = 0;
  if (dbg == 7 && blockIdx.x == 0 && threadIdx.x == 0) {
    stamp_slot = atomicAdd(&g_tc_stamp_n, 1u) % kStampSlots;
    g_tc_stamps[stamp_slot][0] = global_timer();
    g_tc_stamps[stamp_slot][2] = (unsigned long long)(cin * 1000 + N);
    g_tc_stamps[stamp_slot][3] = (unsigned long long)(uintptr_t)ep.out;
  }
  const bool use_tma = use_tma_arg == 1;
  griddep_launch_dependents();
  int n_out = n_out_dev ? *n_out_dev : (int)n_out_cap;
  if (n_out > n_out_cap) n_out = (int)n_out_cap;
  const int n_tiles = (n_out + kTileM - 1) / kTileM;
  // w_row_bytes: bytes of one input row consumed per MMA group = row pitch of a weight image (= TMA box width);
  // row_bytes: row pitch of the A tile - the same, except packed stages, whose 128-byte rows hold kGroup offsets.
  const int w_row_bytes = min(cin * kElem, 128);
  const int row_bytes = kPacked ? 128 : w_row_bytes;
  const int slices = kPacked ? 1 : (cin * kElem) / w_row_bytes;  // stages per kernel offset
  const int group = kPacked ? 128 / w_row_bytes : 1;             // offsets per stage
  const uint32_t a_stage_bytes = (uint32_t)kTileM * row_bytes;
  const uint32_t w_stage_bytes = (uint32_t)N * w_row_bytes;

  if (threadIdx.x == 0) {
    for (int s = 0; s < C::kStages; ++s) {
      // A side of a stage.  The owning warps' cp.async arrivals (32 each) + the primary owner's arrive, which
      // publishes the stage flags (TMA gather: that arrive carries the expected bytes instead), land on `full`
      // for bf16 and on `landed` for fp32, where the 128 transform threads then complete `full`.
      const uint32_t a_arrivals = use_tma ? 1u : 32u * (uint32_t)slot_parts<C::kStages>(s) + 1u;
      mbar_init(bar_full + 8 * s, kFp32 ? (uint32_t)kXformThreads : a_arrivals);
      mbar_init(bar_empty + 8 * s, 1);
      mbar_init(bar_landed + 8 * s, a_arrivals);
    }
    // W slots are released by the same tcgen05.commit as the A slot of the stage that used them (bar_empty): one
    // commit per stage instead of two
    for (int w = 0; w < kWBars; ++w) mbar_init(bar_wfull + 8 * w, 1);  // the expect_tx arrive of the bulk copy's issuer
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_tfull + 8 * a, 1);
      mbar_init(bar_tempty + 8 * a, kEpiThreads);
    }
    for (int b = 0; b < C::kNbrBufs; ++b) {
      mbar_init(bar_nfull + 8 * b, 1);            // the scheduler's (expect_tx) arrive; bulk copies complete the bytes
      // every producer warp and (with a weight ring) the W loader are done with the buffer
      mbar_init(bar_nempty + 8 * b, kProdWarps + (kPacked ? 0 : 1));
    }
    asm volatile("st.shared.s32 [%0], %1;" ::"r"(tiles_posted), "r"(0) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&feat_map) : "memory");
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
    // =============================== producers: gather issue ===============================
    // No block-level synchronisation: a warp takes the tile's neighbour rows and offset mask from the scheduler's
    // ring, issues the stages whose ring slot it owns, and hands the buffer back.
    const int pwarp = (threadIdx.x - kEpiThreads) >> 5;
    // Warp w fills ring slot w % kStages, always the same one, so that each empty barrier is waited on in program
    // order by its owners (the parity wait cannot alias, whatever the drift between warps).  Slots with a second
    // owner (w + kStages is still a producer warp) are filled half and half; a third owner would idle.
    const int my_slot = pwarp % C::kStages, my_part = pwarp / C::kStages;
    const uint8_t *feat = static_cast<const uint8_t *>(feat_ptr);
    const size_t feat_row_bytes = (size_t)cin * kElem;
    uint32_t issued = 0;
    // LSU gather geometry: `lanes_per_row` = 16-byte chunks per row slice (8, 4 or 2); one warp instruction covers
    // 32/lanes_per_row rows.  Swizzle<B,4,3>: chunk ^= (row >> (3-B)) & (chunks-1) with B = log2(chunks).
    const int chunks = row_bytes >> 4;
    const int cshift = __ffs(chunks) - 1;
    const int my_chunk = lane & (chunks - 1);
    const int my_row0 = lane >> cshift;
    const int rows_per_instr = 32 >> cshift;
    TC_TIMER_DECL(tm_pwait);
    TC_TIMER_DECL(tm_pissue);
    TC_TIMER_DECL(tm_ppro);
    TC_TIMER_DECL(tm_tiles);
    griddep_wait();  // the features gathered below are the previous layer's output
    for (uint32_t seq = 0;; ++seq) {
      const uint32_t nb = seq % C::kNbrBufs;
      {
        TC_T0();
        mbar_wait(bar_nfull + 8 * nb, (seq / C::kNbrBufs) & 1);
        TC_ACC(tm_ppro);
      }
      const int tile = tile_info.get(2 * nb);
      uint32_t mask = (uint32_t)tile_info.get(2 * nb + 1);
      const int *nbr_b = nbr_s + nb * kNbrBufInts;
      const uint32_t nbr_b32 = smem_u32(nbr_b);
      if (tile < 0) {
        // Sentinel stage: same arrivals as a real stage, no copies; tells the other roles to stop.
        const uint32_t s = issued % C::kStages;
        if ((int)s == my_slot && my_part < 2) {
          mbar_wait(bar_empty + 8 * s, ((issued / C::kStages) & 1) ^ 1);
          const uint32_t a_bar = kFp32 ? bar_landed + 8 * s : bar_full + 8 * s;
          if (lane == 0 && my_part == 0) {
            stage_flags.set(s, kFlagStop);
            mbar_arrive(a_bar);
          }
          __syncwarp();
          if (!use_tma) cp_async_arrive(a_bar);
        }
        break;
      }
#ifdef FV2P_TC_TIMERS
      tm_tiles += 1;
#endif
      if (mask == 0u) mask = 1u;  // a tile nothing feeds still has to produce (zero) accumulators
      if constexpr (kPacked) {
        // `group` offsets per stage, side by side in the 128-byte rows of the A tile: 16-byte chunk c of row r holds
        // chunk c % cpo of the input row that feeds r through the (c / cpo)-th offset of the stage.
        const int cpo = w_row_bytes >> 4;  // chunks per offset: 2 or 4
        const int my_q = my_chunk / cpo;
        const uint8_t *src_base = feat + (size_t)(my_chunk % cpo) * 16;
        const int n_st = (__popc(mask) + group - 1) / group;
        for (int g = 0; g < n_st; ++g, ++issued) {
          uint32_t ks = 0xFFFFFFFFu;
          int my_k = -1;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            if (q < group && mask) {
              const int k = __ffs(mask) - 1;
              mask &= mask - 1;
              ks = (ks & ~(0xFFu << (8 * q))) | ((uint32_t)k << (8 * q));
              if (q == my_q) my_k = k;
            }
          }
          const uint32_t s = issued % C::kStages;
          if ((int)s != my_slot || my_part >= 2) continue;
          const int rows_here = (s + C::kStages < kProdWarps) ? kTileM / 2 : kTileM;  // rows this warp gathers
          const int row_base = my_part * rows_here;
          {
            TC_T0();
            mbar_wait(bar_empty + 8 * s, ((issued / C::kStages) & 1) ^ 1);
            TC_ACC(tm_pwait);
          }
          TC_T0();
          const uint32_t a_u32 = smem_u32(stage_base + (size_t)s * C::kABytes);
          const uint32_t a_bar = kFp32 ? bar_landed + 8 * s : bar_full + 8 * s;
          if (lane == 0 && my_part == 0) {
            stage_flags.set(s, (g == 0 ? kFlagFirst : 0) | (g == n_st - 1 ? kFlagLast : 0));
            stage_ks.set(s, (int)ks);
            mbar_arrive(a_bar);
          }
          __syncwarp();
          if (my_k >= 0) {  // lanes of an unused slot of the last group copy nothing (no MMA reads those columns)
            const int rpg = rows_here >> 2;  // rows per 8-lane group
            const int r_first = row_base + (lane >> 3) * rpg;
            const uint32_t idx4 = nbr_b32 + (uint32_t)(my_k * kTileM + r_first) * 4u;
            const uint32_t dst0 = a_u32 + (uint32_t)r_first * 128u;
            for (int q4 = 0; q4 < (rpg >> 2); ++q4) {
              const int4 v = lds128i(idx4 + 16u * (uint32_t)q4);
              const int srcs[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const int r = r_first + 4 * q4 + u;
                const uint32_t swz = (uint32_t)(my_chunk ^ (r & 7));
                const int src = srcs[u];
                cp_async16(dst0 + (uint32_t)(4 * q4 + u) * 128u + (swz << 4),
                           src_base + (src >= 0 ? (size_t)src * feat_row_bytes : 0), src >= 0 ? 16u : 0u);
              }
            }
          }
          cp_async_arrive(a_bar);
          TC_ACC(tm_pissue);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_nempty + 8 * nb);  // this warp no longer reads the tile's neighbour rows
        continue;
      }
      const uint32_t first_k = __ffs(mask) - 1;
      const uint32_t last_k = 31 - __clz(mask);
      while (mask) {
        const int k = __ffs(mask) - 1;
        mask &= mask - 1;
        for (int sl = 0; sl < slices; ++sl, ++issued) {
          const uint32_t s = issued % C::kStages;
          if ((int)s != my_slot || my_part >= 2) continue;
          const int rows_here = (s + C::kStages < kProdWarps) ? kTileM / 2 : kTileM;  // rows this warp gathers
          const int row_base = my_part * rows_here;
          {
            TC_T0();
            mbar_wait(bar_empty + 8 * s, ((issued / C::kStages) & 1) ^ 1);
            TC_ACC(tm_pwait);
          }
          TC_T0();
          const uint32_t a_u32 = smem_u32(stage_base + (size_t)s * C::kABytes);
          const uint32_t a_bar = kFp32 ? bar_landed + 8 * s : bar_full + 8 * s;
          if (lane == 0 && my_part == 0) {
            stage_flags.set(s, ((k == (int)first_k && sl == 0) ? kFlagFirst : 0) |
                             ((k == (int)last_k && sl == slices - 1) ? kFlagLast : 0));
            if (use_tma) mbar_arrive_expect_tx(a_bar, a_stage_bytes);
            else mbar_arrive(a_bar);
          }
          __syncwarp();
          if (use_tma) {
            // lane l gathers four of this warp's rows; a missing neighbour becomes an out-of-range row (zero fill)
            if (4 * lane < rows_here) {
              const int r4 = row_base + 4 * lane;
              int4 rows = *reinterpret_cast<const int4 *>(&nbr_b[k * kTileM + r4]);
              rows.x = rows.x < 0 ? oob_row : rows.x, rows.y = rows.y < 0 ? oob_row : rows.y;
              rows.z = rows.z < 0 ? oob_row : rows.z, rows.w = rows.w < 0 ? oob_row : rows.w;
              tma_gather4(a_u32 + (uint32_t)r4 * row_bytes, &feat_map, sl * (row_bytes / kElem), rows.x, rows.y,
                          rows.z, rows.w, a_bar);
            }
          } else {
            // Lane group g = lane / chunks owns the contiguous rows [g*rpg, (g+1)*rpg): its row indices come in
            // with 16-byte shared loads, and every warp instruction still reads whole rows (full 128-byte lines for
            // 128-byte slices) and writes conflict-free thanks to the swizzle.
            const uint8_t *src_base = feat + (size_t)sl * row_bytes + my_chunk * 16;
            if (chunks < 8) {  // narrow rows: interleaved rows per instruction measured faster (0.046 vs 0.053 ms, 16->16)
              const uint32_t rows_k = nbr_b32 + (uint32_t)(k * kTileM) * 4u;
#pragma unroll 4
              for (int r = row_base + my_row0; r < row_base + rows_here; r += rows_per_instr) {
                const int src = lds32i(rows_k + 4u * (uint32_t)r);
                const uint32_t swz = (uint32_t)(my_chunk ^ ((r >> (3 - cshift)) & (chunks - 1)));
                cp_async16(a_u32 + (uint32_t)r * row_bytes + (swz << 4),
                           src_base + (src >= 0 ? (size_t)src * feat_row_bytes : 0), src >= 0 ? 16u : 0u);
              }
              cp_async_arrive(a_bar);
              TC_ACC(tm_pissue);
              continue;
            }
            const int rpg = rows_here >> (5 - cshift);  // rows per lane group: 32, 16 or 8 (half with two owners)
            const int r_first = row_base + (lane >> cshift) * rpg;
            const uint32_t idx4 = nbr_b32 + (uint32_t)(k * kTileM + r_first) * 4u;
            const uint32_t dst0 = a_u32 + (uint32_t)r_first * row_bytes;
            for (int q = 0; q < (rpg >> 2); ++q) {
              const int4 v = lds128i(idx4 + 16u * (uint32_t)q);
              const int srcs[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const int r = r_first + 4 * q + u;
                const uint32_t swz = (uint32_t)(my_chunk ^ ((r >> (3 - cshift)) & (chunks - 1)));
                const int src = srcs[u];
                cp_async16(dst0 + (uint32_t)(4 * q + u) * row_bytes + (swz << 4),
                           src_base + (src >= 0 ? (size_t)src * feat_row_bytes : 0), src >= 0 ? 16u : 0u);
              }
            }
            cp_async_arrive(a_bar);
          }
          TC_ACC(tm_pissue);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_nempty + 8 * nb);  // this warp no longer reads the tile's neighbour rows
    }
#ifdef FV2P_TC_TIMERS
    if (threadIdx.x == kEpiThreads) {
      g_tc_timers[blockIdx.x][1] = issued;
      g_tc_timers[blockIdx.x][2] = tm_tiles;
      g_tc_timers[blockIdx.x][3] = tm_pwait;
      g_tc_timers[blockIdx.x][4] = tm_pissue;
      g_tc_timers[blockIdx.x][5] = tm_ppro;
    }
#endif
  } else if (warp == kSchedWarp) {
    // =============================== tile scheduler + neighbour-map loader ===============================
    // Runs up to kNbrBufs tiles ahead of the producers: claims the next tile, and brings the rows of the
    // neighbour map the tile needs (only the offsets in its mask) into the ring with bulk copies that complete
    // on the buffer's barrier.  Tile positions index `tile_order` (heaviest first).  Claiming from a shared
    // counter only would let the first CTAs take several of the heaviest tiles each when there are 1-2 tiles
    // per CTA, so the first two rounds are dealt statically in snake order - position b, then 2G-1-b: the CTA
    // with the lightest first tile gets the heaviest second one - and the rest comes from the counter
    // (round-robin without scheduler scratch).
    const int2 *order2 = reinterpret_cast<const int2 *>(tile_order);
    const bool bulk_ok = (nbr_stride & 3) == 0 && (reinterpret_cast<uintptr_t>(nbr) & 15) == 0;
    int fetches = 0;
    int2 pending = make_int2(-1, 0);  // lane 0: the claim in flight
    auto claim = [&]() {
      if (lane == 0) {
        const int G = (int)gridDim.x, b = (int)blockIdx.x;
        int i;
        if (fetches == 0) i = b;
        else if (fetches == 1) i = 2 * G - 1 - b;
        else if (sched) {
          // the counter words may still be in use by the previous launch of the same layer (programmatic dependent
          // launch lets this kernel start before it has finished)
          if (fetches == 2) griddep_wait();
          i = 2 * G + atomicAdd(&sched[0], 1);
        }
        else i = fetches * G + b;
        ++fetches;
        pending = i >= n_tiles ? make_int2(-1, 0) : (order2 ? __ldg(&order2[i]) : make_int2(i, 0));
      }
    };
    claim();
    for (uint32_t seq = 0;; ++seq) {
      const int tile = __shfl_sync(0xFFFFFFFFu, pending.x, 0);
      uint32_t mask = (uint32_t)__shfl_sync(0xFFFFFFFFu, pending.y, 0);
      if (tile >= 0) claim();  // the next claim (atomic + list entry) is in flight while this tile is set up
      const uint32_t nb = seq % C::kNbrBufs;
      mbar_wait(bar_nempty + 8 * nb, ((seq / C::kNbrBufs) & 1) ^ 1);
      int *dst = nbr_s + nb * kNbrBufInts;
      if (lane == 0) {
        tile_ring.set(seq % kTileRing, tile);
        asm volatile("st.release.cta.shared.s32 [%0], %1;" ::"r"(tiles_posted), "r"((int)seq + 1) : "memory");
      }
      if (tile < 0) {
        if (lane == 0) {
          tile_info.set(2 * nb, -1);
          tile_info.set(2 * nb + 1, 0);
          mbar_arrive(bar_nfull + 8 * nb);
        }
        break;
      }
      const int row0 = tile * kTileM;
      const int valid = min(kTileM, n_out - row0);
      if (bulk_ok && order2 && mask != 0u && valid == kTileM) {
        if (elect_one_sync()) {
          tile_info.set(2 * nb, tile);
          tile_info.set(2 * nb + 1, (int)mask);
          mbar_arrive_expect_tx(bar_nfull + 8 * nb, (uint32_t)__popc(mask) * (kTileM * 4u));
          uint32_t m = mask;
          while (m) {
            const int k = __ffs(m) - 1;
            m &= m - 1;
            bulk_g2s(smem_u32(dst + k * kTileM), nbr + (size_t)k * nbr_stride + row0, kTileM * 4u, bar_nfull + 8 * nb);
          }
        }
        __syncwarp();
      } else {
        // Partial last tile, callers without a tile list, unaligned maps: plain loads by the whole warp (lane l
        // serves rows 4l..4l+3), rows past the end read as "no neighbour", mask derived from the data.
        uint32_t found = 0;
        for (int k0 = 0; k0 < kvol; k0 += 8) {  // eight offsets' loads in flight at a time
          int v[8][4];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int r = 4 * lane + u;
              v[q][u] = (k0 + q < kvol && r < valid) ? __ldg(&nbr[(size_t)(k0 + q) * nbr_stride + row0 + r]) : -1;
            }
          }
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            if (k0 + q < kvol) {
              *reinterpret_cast<int4 *>(dst + (k0 + q) * kTileM + 4 * lane) = make_int4(v[q][0], v[q][1], v[q][2], v[q][3]);
              if (__ballot_sync(0xFFFFFFFFu, (v[q][0] & v[q][1] & v[q][2] & v[q][3]) >= 0)) found |= 1u << (k0 + q);
            }
          }
        }
        __syncwarp();
        if (lane == 0) {
          tile_info.set(2 * nb, tile);
          tile_info.set(2 * nb + 1, (int)found);
          mbar_arrive(bar_nfull + 8 * nb);
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // =============================== MMA issuer ===============================
    constexpr uint32_t idesc = instr_desc(N, false);
    constexpr uint32_t idesc_corr = instr_desc(N, false);  // fp32 path: bf16 correction MMAs
    (void)idesc_corr;
    const uint32_t sbo = 8u * (uint32_t)row_bytes;
    const uint32_t layout = row_bytes == 128 ? 2u : (row_bytes == 64 ? 4u : 6u);
    const uint32_t w_sbo = 8u * (uint32_t)w_row_bytes;
    const uint32_t w_layout = w_row_bytes == 128 ? 2u : (w_row_bytes == 64 ? 4u : 6u);
    const int ksteps = row_bytes >> 5;  // 32 bytes of the A row per bf16 MMA; fp32 rows: 64 bytes (16 channels) per K step
    // This loop is a single thread's instruction stream, so it is kept short: the descriptor of a slot is the
    // descriptor of slot 0 plus a constant in the 16-byte start-address field (no carry: smem is < 256 KB).
    const uint64_t a_desc0 = smem_desc(smem_u32(stage_base), sbo, layout);
    const uint64_t w_desc0 = smem_desc(smem_u32(w_base), w_sbo, w_layout);
    constexpr uint64_t kStageStep = (uint64_t)(C::kABytes >> 4);
    constexpr uint64_t kWSlotStep = (uint64_t)(C::kWSlotBytes >> 4);
    const uint64_t w_lo_off = (uint64_t)(w_row_bytes >> 5);  // fp32: W_lo is the second half of every weight row
    const uint64_t w_img_step = (uint64_t)(w_stage_bytes >> 4);  // packed: one offset's image
    const int kpo_shift = (w_row_bytes >> 5) - 1;  // packed: MMA K steps per offset = 1 << kpo_shift (1 or 2)
    (void)w_img_step, (void)kpo_shift, (void)kWSlotStep;
    const uint32_t a_tmem0 = tmem_base + 2u * (uint32_t)N;  // fp32: TMEM ring of split A tiles behind the accumulators
    // One elected thread runs the whole loop on its own (waits included).  Re-electing per stage with the warp
    // waiting and re-synchronising around the issue costs ~300 cycles per stage and ~50 per tcgen05.mma
    // (profiles/micro/mma_issue_bench.cu: 4 MMAs + commit per stage take 760 cycles that way, 280 this way).
    if (elect_one_sync()) {
      uint32_t s = 0, phase = 0;      // A ring slot
      uint32_t ws = 0, wphase = 0;    // W ring slot
      uint64_t a_desc = a_desc0, b_desc = w_desc0;
      uint32_t acc = 0, acc_phase = 0;
      TC_TIMER_DECL(tm_mfull);
      TC_TIMER_DECL(tm_mtmem);
      TC_TIMER_DECL(tm_missue);
      if constexpr (kPacked) mbar_wait(bar_wfull, 0);  // the resident weight images
      for (;;) {
        {
          TC_T0();
          mbar_wait(bar_full + 8 * s, phase);
          TC_ACC(tm_mfull);
        }
        const int flags = stage_flags.get(s);
        {
          TC_T0();
          if (flags & (kFlagFirst | kFlagStop)) mbar_wait(bar_tempty + 8 * acc, acc_phase ^ 1);  // accumulator drained
          TC_ACC(tm_mtmem);
        }
        TC_T0();
        tc_fence_after();
        if (flags & kFlagStop) {
          mbar_arrive(bar_tfull + 8 * acc);  // wakes the epilogue, which finds the sentinel tile id
#ifdef FV2P_TC_TIMERS
          g_tc_timers[blockIdx.x][6] = tm_mfull;
          g_tc_timers[blockIdx.x][7] = tm_mtmem;
          g_tc_timers[blockIdx.x][8] = tm_missue;
#endif
          break;
        }
        const uint32_t d_tmem = tmem_base + acc * (uint32_t)N;
        uint32_t accumulate = (flags & kFlagFirst) ? 0u : 1u;
        if constexpr (kPacked) {
          // one MMA per 32 bytes of the A row: K step t belongs to the (t >> kpo_shift)-th offset of the stage and
          // reads that offset's resident weight image
          const uint32_t ks = (uint32_t)stage_ks.get(s);
          if (dbg != 4) {
            if constexpr (kFp32) {
              // fp32 rows of 64 bytes: two offsets per stage, 16 channels = ONE K step of the split operands each
#pragma unroll
              for (int q = 0; q < 2; ++q) {
                const uint32_t k = (ks >> (8 * q)) & 0xFFu;
                if (k != 0xFFu) {
                  const uint64_t b = w_desc0 + (uint64_t)k * w_img_step;
                  const uint32_t a_hi = a_tmem0 + s * (uint32_t)C::kAColsPerStage + 16u * (uint32_t)q;
                  tc_mma_ts_f16(d_tmem, a_hi, b, idesc_corr, accumulate);
                  tc_mma_ts_f16(d_tmem, a_hi, b + w_lo_off, idesc_corr, 1u);
                  tc_mma_ts_f16(d_tmem, a_hi + 8u, b, idesc_corr, 1u);
                  accumulate = 1u;
                }
              }
            } else {
#pragma unroll
              for (int t = 0; t < 4; ++t) {
                const uint32_t k = (ks >> (8 * (t >> kpo_shift))) & 0xFFu;
                if (k != 0xFFu) {
                  const uint64_t b = w_desc0 + (uint64_t)k * w_img_step + (uint64_t)(2 * (t & ((1 << kpo_shift) - 1)));
                  tc_mma<false>(d_tmem, a_desc + (uint64_t)(2 * t), b, idesc, accumulate);
                  accumulate = 1u;
                }
              }
            }
          }
        } else {
          {
            TC_T0();
            mbar_wait(bar_wfull + 8 * ws, wphase);  // the stage's weight slice
            TC_ACC(tm_mfull);
          }
          tc_fence_after();
          if (dbg != 4) {
            if constexpr (kFp32) {
              // 16 input channels per step, three bf16 MMAs with the A operand in TMEM: hi*hi + hi*lo + lo*hi
#pragma unroll
              for (int j = 0; j < 2; ++j) {
                if (2 * j < ksteps) {
                  const uint64_t b = b_desc + (uint64_t)(2 * j);  // 16 bf16 = 32 bytes of K in the W_hi half
                  const uint32_t a_hi = a_tmem0 + s * (uint32_t)C::kAColsPerStage + 16u * (uint32_t)j;
                  tc_mma_ts_f16(d_tmem, a_hi, b, idesc_corr, accumulate);
                  tc_mma_ts_f16(d_tmem, a_hi, b + w_lo_off, idesc_corr, 1u);
                  tc_mma_ts_f16(d_tmem, a_hi + 8u, b, idesc_corr, 1u);
                  accumulate = 1u;
                }
              }
            } else {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                if (j < ksteps) {
                  const uint64_t adv = (uint64_t)(2 * j);  // 32 bytes of K
                  tc_mma<false>(d_tmem, a_desc + adv, b_desc + adv, idesc, accumulate);
                  accumulate = 1u;
                }
              }
            }
          }
        }
        tc_commit(bar_empty + 8 * s);  // the A slot AND the W slot of this stage are reusable once these MMAs retire
        if (flags & kFlagLast) {
          tc_commit(bar_tfull + 8 * acc);  // accumulator complete
          acc ^= 1u;
          if (acc == 0u) acc_phase ^= 1u;
        }
        a_desc += kStageStep;
        if (++s == (uint32_t)C::kStages) {
          s = 0;
          phase ^= 1;
          a_desc = a_desc0;
        }
        if constexpr (!kPacked) {
          b_desc += kWSlotStep;
          if (++ws == (uint32_t)C::kWStages) {
            ws = 0;
            wphase ^= 1;
            b_desc = w_desc0;
          }
        }
        TC_ACC(tm_missue);
      }
    }
    __syncwarp();
  } else if (warp == kWLoadWarp) {
    // =============================== weight-slice loader ===============================
    // One thread walks the same stage sequence as the producers (tile masks from the scheduler's ring) and keeps
    // the W ring full: stage i's slice goes to slot i % kWStages with one bulk copy.  A single in-order waiter per
    // barrier, so the W ring may be shallower than the A ring without a parity wait ever being two phases ahead.
    if (elect_one_sync()) {
      const uint32_t wb = w_stage_bytes;
      if constexpr (kPacked) {
        // every offset's image once, resident for the whole kernel (weights do not depend on the previous layer, so
        // this does not wait for it either)
        mbar_arrive_expect_tx(bar_wfull, dbg == 3 ? 0u : wb * (uint32_t)kvol);
        if (dbg != 3)
          for (int k = 0; k < kvol; ++k)
            bulk_g2s(smem_u32(w_base + (size_t)k * wb), wpacked + (size_t)k * wb, wb, bar_wfull);
      } else {
        uint32_t n = 0;
        for (uint32_t seq = 0;; ++seq) {
          const uint32_t nb = seq % C::kNbrBufs;
          mbar_wait(bar_nfull + 8 * nb, (seq / C::kNbrBufs) & 1);
          const int tile = tile_info.get(2 * nb);
          uint32_t mask = (uint32_t)tile_info.get(2 * nb + 1);
          if (tile < 0) break;
          if (mask == 0u) mask = 1u;
          while (mask) {
            const int k = __ffs(mask) - 1;
            mask &= mask - 1;
            for (int sl = 0; sl < slices; ++sl, ++n) {
              const uint32_t w = n % C::kWStages;
              if (n >= (uint32_t)C::kWStages) {
                // slot w was last read by stage n - kWStages; that stage's commit (bar_empty of ITS A slot) frees it.
                // kWStages <= kStages, so the barrier cannot complete another phase before stage n has its weights.
                const uint32_t m = n - (uint32_t)C::kWStages;
                mbar_wait(bar_empty + 8 * (m % C::kStages), (m / C::kStages) & 1);
              }
              mbar_arrive_expect_tx(bar_wfull + 8 * w, dbg == 3 ? 0u : wb);  // dbg 3: no weight copy
              if (dbg != 3)
                bulk_g2s(smem_u32(w_base + (size_t)w * C::kWSlotBytes), wpacked + ((size_t)k * slices + sl) * wb, wb,
                         bar_wfull + 8 * w);
            }
          }
          mbar_arrive(bar_nempty + 8 * nb);
        }
      }
    }
    __syncwarp();
  } else if (warp > kWLoadWarp) {
    // =============================== fp32 split (warps 15-18, fp32 kernels only) ===============================
    // Thread = tile row (the warp's TMEM lane quadrant is warp % 4): reads its landed fp32 row slice from the
    // swizzled tile, splits it and stores hi / lo to the stage's TMEM columns.
    if constexpr (kFp32) {
      const int t = threadIdx.x - kTcThreadsBase;
      const int quad = warp & 3;
      const int r = quad * 32 + lane;
      const int chunks = row_bytes >> 4;  // 16-byte chunks per row slice: 8, 4 (cin = 16) -- 4 floats each
      const int cshift = __ffs(chunks) - 1;
      const uint32_t swz_row = (uint32_t)((r >> (3 - cshift)) & (chunks - 1));
      const uint32_t lane_addr = tmem_base + ((uint32_t)(quad * 32) << 16) + 2u * (uint32_t)N;
      TC_TIMER_DECL(tm_xwait);
      TC_TIMER_DECL(tm_xwork);
      for (uint32_t done = 0;; ++done) {
        const uint32_t s = done % C::kStages;
        {
          TC_T0();
          mbar_wait(bar_landed + 8 * s, (done / C::kStages) & 1);
          TC_ACC(tm_xwait);
        }
        TC_T0();
        const bool stop = (stage_flags.get(s) & kFlagStop) != 0;
        if (!stop) {
          const uint32_t row = smem_u32(stage_base) + s * (uint32_t)C::kABytes + (uint32_t)r * (uint32_t)row_bytes;
          const uint32_t a_cols = lane_addr + s * (uint32_t)C::kAColsPerStage;
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            if (half * 4 < chunks) {
              float4 x[4];
#pragma unroll
              for (int c = 0; c < 4; ++c)
                x[c] = lds128f(row + (((uint32_t)(half * 4 + c) ^ swz_row) << 4));
              // 16 channels = one K step: columns 0-7 = hi (bf16 round-to-nearest of x, two per column, K-consecutive),
              // columns 8-15 = lo = bf16(x - hi); x - hi - lo <= 2^-18 |x|.  Three products per step (hi*W_hi, hi*W_lo,
              // lo*W_hi; the dropped lo*W_lo is 2^-18 of the product) instead of the first version's tf32 product + bf16
              // correction (two MMAs per EIGHT channels, one of them reading its A operand from shared memory).
              uint32_t split[16];
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                const float v[4] = {x[c].x, x[c].y, x[c].z, x[c].w};
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                  const uint32_t hi = pack_bf16x2(v[2 * e], v[2 * e + 1]);
                  split[2 * c + e] = hi;
                  split[8 + 2 * c + e] = pack_bf16x2(v[2 * e] - __uint_as_float(hi << 16),
                                                     v[2 * e + 1] - __uint_as_float(hi & 0xFFFF0000u));
                }
              }
              tmem_st16(a_cols + 16u * half, split);
            }
          }
          asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
          tc_fence_before();
        }
        mbar_arrive(bar_full + 8 * s);
        TC_ACC(tm_xwork);
        if (stop) break;
      }
#ifdef FV2P_TC_TIMERS
      if (t == 0) {
        g_tc_timers[blockIdx.x][11] = tm_xwait;
        g_tc_timers[blockIdx.x][12] = tm_xwork;
      }
#else
      (void)t;
#endif
    }
  } else {
    // =============================== epilogue (warps 0-3) ===============================
    int acc = 0;
    uint32_t acc_phase = 0;
    TC_TIMER_DECL(tm_ewait);
    TC_TIMER_DECL(tm_ework);
    griddep_wait();  // residual rows come from an earlier layer; the output buffer may still be read by the previous one
    // The epilogue's own latency chain per tile used to be: accumulator ready -> row order (a global load) -> residual
    // (a dependent global load) -> tcgen05.ld -> store, i.e. two L2 round trips before the first useful instruction;
    // with 128 x 16 / 32 outputs per tile that chain, not the contraction, paced the narrow layers (2-4 k cycles per
    // tile).  Now the tile id of the NEXT accumulator is read from the scheduler's ring while the current one is
    // processed (`tiles_posted` says whether it is there yet), its output rows are loaded one tile ahead, and the
    // first residual chunk is issued before waiting for the accumulator.
    constexpr int kLanesPerRow = kFp32 ? 4 : 2;          // lanes that share a row segment
    constexpr int kRowsPerInstr = 32 / kLanesPerRow;     // 8 / 16
    constexpr int kPasses = 32 / kRowsPerInstr;          // 4 / 2
    constexpr int kColsPerLane = 16 / kLanesPerRow;      // 4 / 8 columns of a 16-column chunk per lane
    const int sub = lane / kLanesPerRow, seg = lane % kLanesPerRow;
    const bool body = dbg != 6;
    // sorted position -> output row (identity without a row order from fv2p_sort_rows_by_mask)
    auto load_row = [&](int tile) {
      const int srow = tile * kTileM + warp * 32 + lane;
      return srow < n_out ? (row_perm ? __ldg(&row_perm[srow]) : srow) : n_out;
    };
    auto load_residual = [&](const int *grow, int c0, uint4 *res) {
#pragma unroll
      for (int q = 0; q < kPasses; ++q)
        if (grow[q] < n_out)
          res[q] = __ldg(reinterpret_cast<const uint4 *>(static_cast<const uint8_t *>(ep.residual) +
                                                         ((size_t)grow[q] * N + c0 + seg * kColsPerLane) * kElem));
    };
    int tile_next = -1, row_next = 0;
    bool have_next = false;
    for (uint32_t seq = 0;; ++seq) {
      int tile = -1, row = 0;
      int grow[kPasses];  // the output row this lane serves in pass q (the row of lane q * kRowsPerInstr + sub)
      uint4 res0[kPasses];
      const bool early = have_next;
      if (early) {
        tile = tile_next, row = row_next;
#pragma unroll
        for (int q = 0; q < kPasses; ++q) grow[q] = __shfl_sync(0xFFFFFFFFu, row, q * kRowsPerInstr + sub);
        if (body && ep.residual) load_residual(grow, 0, res0);
      }
      {
        TC_T0();
        mbar_wait(bar_tfull + 8 * acc, acc_phase);
        TC_ACC(tm_ewait);
      }
      TC_T0();
      tc_fence_after();
      if (!early) {
        tile = tile_ring.get(seq % kTileRing);
        if (tile < 0) {
#ifdef FV2P_TC_TIMERS
          if (threadIdx.x == 0) {
            g_tc_timers[blockIdx.x][9] = tm_ewait;
            g_tc_timers[blockIdx.x][10] = tm_ework;
          }
#endif
          break;
        }
        row = load_row(tile);
#pragma unroll
        for (int q = 0; q < kPasses; ++q) grow[q] = __shfl_sync(0xFFFFFFFFu, row, q * kRowsPerInstr + sub);
        if (body && ep.residual) load_residual(grow, 0, res0);
      }
      {
        // one tile ahead: only when the scheduler has already published it (it usually has - it runs ahead of the
        // producers); the sentinel is left to the slow path above, which runs after the MMA thread's final arrive
        int posted;
        asm volatile("ld.acquire.cta.shared.s32 %0, [%1];" : "=r"(posted) : "r"(tiles_posted) : "memory");
        have_next = false;
        if (posted > (int)seq + 1) {
          tile_next = tile_ring.get((seq + 1) % kTileRing);
          if (tile_next >= 0) {
            row_next = load_row(tile_next);
            have_next = true;
          }
        }
      }
      const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(acc * N);
      // Per 16-column chunk: tcgen05.ld hands every lane ITS row (32 rows x 16 fp32 per warp); the warp writes them to
      // its staging buffer and reads them back transposed - lanes side by side along a row - so that every global
      // access below covers whole contiguous row segments (fp32: 4 lanes x 16 B = the 64 bytes of the chunk, 8 rows per
      // instruction; bf16: 2 lanes x 16 B, 16 rows per instruction) and a lane needs the folded BatchNorm parameters of
      // 4 / 8 columns only (2-4 vector loads per chunk instead of 32 scalar ones).  Round 1 had one row per lane all
      // the way: 32 half-used sectors per load / store instruction and 40 memory instructions per lane and chunk -
      // the epilogue (18-40 k cycles per 128-wide tile) was busy longer than the contraction it is meant to hide
      // behind and competed with the gather for the load/store unit.  The tcgen05.ld of chunk c+1 and the residual
      // of chunk c are in flight while chunk c is transposed.
      constexpr int kChunks = N / 16;
      const uint32_t stage = smem_u32(epi_stage) + (uint32_t)warp * (32 * kEpiRowPitch);
      uint32_t acc_regs[2][16];
      tmem_ld16_issue(taddr, acc_regs[0]);  // warp-collective: executed by all lanes even for rows past the end
#pragma unroll
      for (int c = 0; c < kChunks; ++c) {
        const int c0 = c * 16;
        uint32_t *cur = acc_regs[c & 1];
        // residual segments of this chunk: issued before the transposition so that their latency hides behind it
        uint4 res[kPasses];
        if (c == 0) {
#pragma unroll
          for (int q = 0; q < kPasses; ++q) res[q] = res0[q];
        } else if (body && ep.residual) {
          load_residual(grow, c0, res);
        }
        tmem_ld_wait(cur);
        if (c + 1 < kChunks) {
          tmem_ld16_issue(taddr + c0 + 16, acc_regs[(c + 1) & 1]);
        } else {
          // the whole accumulator is in registers: the MMA thread may start the tile after next on it
          tc_fence_before();
          mbar_arrive(bar_tempty + 8 * acc);
        }
        if (!body) continue;
        __syncwarp();  // the previous chunk has been read back by every lane
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4)
          sts128(stage + (uint32_t)(lane * kEpiRowPitch + q4 * 16), cur[4 * q4], cur[4 * q4 + 1], cur[4 * q4 + 2],
                 cur[4 * q4 + 3]);
        __syncwarp();
        // folded BatchNorm (and bias) of this lane's columns
        float sc[kColsPerLane], sh[kColsPerLane], bi[kColsPerLane];
#pragma unroll
        for (int v4 = 0; v4 < kColsPerLane / 4; ++v4) {
          const int col = c0 + seg * kColsPerLane + 4 * v4;
          const float4 a = ep.scale ? __ldg(reinterpret_cast<const float4 *>(ep.scale + col)) : make_float4(1, 1, 1, 1);
          const float4 b = ep.scale ? __ldg(reinterpret_cast<const float4 *>(ep.shift + col)) : make_float4(0, 0, 0, 0);
          const float4 d = ep.bias ? __ldg(reinterpret_cast<const float4 *>(ep.bias + col)) : make_float4(0, 0, 0, 0);
          sc[4 * v4] = a.x, sc[4 * v4 + 1] = a.y, sc[4 * v4 + 2] = a.z, sc[4 * v4 + 3] = a.w;
          sh[4 * v4] = b.x, sh[4 * v4 + 1] = b.y, sh[4 * v4 + 2] = b.z, sh[4 * v4 + 3] = b.w;
          bi[4 * v4] = d.x, bi[4 * v4 + 1] = d.y, bi[4 * v4 + 2] = d.z, bi[4 * v4 + 3] = d.w;
        }
#pragma unroll
        for (int q = 0; q < kPasses; ++q) {
          const int r = q * kRowsPerInstr + sub;  // row of the warp's 32
          float v[kColsPerLane];
#pragma unroll
          for (int v4 = 0; v4 < kColsPerLane / 4; ++v4) {
            const float4 x = lds128f(stage + (uint32_t)(r * kEpiRowPitch + (seg * kColsPerLane + 4 * v4) * 4));
            v[4 * v4] = x.x, v[4 * v4 + 1] = x.y, v[4 * v4 + 2] = x.z, v[4 * v4 + 3] = x.w;
          }
          if (grow[q] >= n_out) continue;
#pragma unroll
          for (int i = 0; i < kColsPerLane; ++i) {
            float x = v[i];
            if (ep.bias) x += bi[i];
            if (ep.scale) x = fmaf(x, sc[i], sh[i]);
            v[i] = x;
          }
          uint8_t *o = static_cast<uint8_t *>(ep.out) + ((size_t)grow[q] * N + c0 + seg * kColsPerLane) * kElem;
          if constexpr (kFp32) {
            if (ep.residual) {
              v[0] += __uint_as_float(res[q].x), v[1] += __uint_as_float(res[q].y);
              v[2] += __uint_as_float(res[q].z), v[3] += __uint_as_float(res[q].w);
            }
            float4 t = make_float4(v[0], v[1], v[2], v[3]);
            if (ep.relu) t.x = fmaxf(t.x, 0.f), t.y = fmaxf(t.y, 0.f), t.z = fmaxf(t.z, 0.f), t.w = fmaxf(t.w, 0.f);
            *reinterpret_cast<float4 *>(o) = t;
          } else {
            if (ep.residual) {
              const __nv_bfloat162 *h = reinterpret_cast<const __nv_bfloat162 *>(&res[q]);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 f = __bfloat1622float2(h[e]);
                v[2 * e] += f.x;
                v[2 * e + 1] += f.y;
              }
            }
            uint4 t;
            __nv_bfloat162 *h = reinterpret_cast<__nv_bfloat162 *>(&t);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float a = v[2 * e], b = v[2 * e + 1];
              if (ep.relu) a = fmaxf(a, 0.f), b = fmaxf(b, 0.f);
              h[e] = __floats2bfloat162_rn(a, b);
            }
            *reinterpret_cast<uint4 *>(o) = t;
          }
        }
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
      TC_ACC(tm_ework);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (dbg == 7 && blockIdx.x == 0 && threadIdx.x == 0) g_tc_stamps[stamp_slot][1] = global_timer();
#ifdef FV2P_TC_TIMERS
  if (threadIdx.x == 0) g_tc_timers[blockIdx.x][0] = clock64() - tc_cta_t0;
#endif
  if (sched && threadIdx.x == 0) {
    // the last CTA to leave re-arms the scheduler words for the next launch that uses them
    __threadfence();
    if (atomicAdd(&sched[1], 1) == (int)gridDim.x - 1) {
      sched[0] = 0;
      sched[1] = 0;
    }
  }
  if (warp == kMmaWarp) {
    __syncwarp();
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)C::kTmemCols)
                 : "memory");
  }
}

// Byte offset of 16-byte chunk c of row n inside a K-major tile with `row_bytes` per row and the matching
// 128B/64B/32B swizzle (Swizzle<B,4,3>: address bits [4,4+B) ^= bits [7,7+B)).
__host__ __device__ inline size_t swizzled_offset(int n, int c, int row_bytes) {
  const size_t lin = (size_t)n * row_bytes + (size_t)c * 16;
  const int bits = row_bytes == 128 ? 3 : (row_bytes == 64 ? 2 : 1);
  const size_t x = (lin >> 7) & ((1u << bits) - 1);
  return lin ^ (x << 4);
}

// Packs W [K,cin,cout] fp32 into per-(offset, 128-byte slice) shared-memory images of the B operand
// (N rows x Kslice, K-major, swizzled like the A tile so one bulk copy drops it in place).
template <bool kFp32>
__global__ void __launch_bounds__(kThreads)
pack_weight_kernel(const float *__restrict__ w, int kvol, int cin, int cout, uint8_t *packed) {
  constexpr int kElem = kFp32 ? 4 : 2;
  constexpr int kPerChunk = 16 / kElem;
  const int row_bytes = min(cin * kElem, 128);
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
    const size_t off = swizzled_offset(n, c, row_bytes);
    const float val = w[e];
    uint8_t *base = packed + ((size_t)k * slices + sl) * image;
    if constexpr (kFp32) {
      // fp32: the row of output channel n holds the slice's input channels twice, as bf16: first half W_hi =
      // bf16(w), second half W_lo = bf16(w - W_hi) (the B operands of hi*W_hi, lo*W_hi and hi*W_lo)
      (void)off, (void)t;
      const __nv_bfloat16 hi = __float2bfloat16_rn(val);
      const int half_chunks = row_bytes >> 5;  // 16-byte chunks per half row
      reinterpret_cast<__nv_bfloat16 *>(base + swizzled_offset(n, within / 8, row_bytes))[within % 8] = hi;
      reinterpret_cast<__nv_bfloat16 *>(base + swizzled_offset(n, half_chunks + within / 8, row_bytes))[within % 8] =
          __float2bfloat16_rn(val - __bfloat162float(hi));
    } else {
      reinterpret_cast<__nv_bfloat16 *>(base + off)[t] = __float2bfloat16_rn(val);
    }
  }
}

bool tc_shape_ok(int cin, int cout) {
  const bool n_ok = cout == 16 || cout == 32 || cout == 64 || cout == 128;
  const bool k_ok = cin == 16 || cin == 32 || cin == 64 || cin == 128 || cin == 256;
  return n_ok && k_ok;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (the library links no libcuda).
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2D map over the feature matrix [rows, cin]; box = one row slice of row_bytes; gather4 fetches four boxes.
int make_feature_map(CUtensorMap *map, const void *features, int64_t rows, int cin, bool fp32) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) {
    set_error("conv_fwd(tc): cuTensorMapEncodeTiled is not available from this driver");
    return FV2P_ERR_DEVICE;
  }
  const int elem = fp32 ? 4 : 2;
  const int row_bytes = cin * elem < 128 ? cin * elem : 128;
  cuuint64_t gdim[2] = {(cuuint64_t)cin, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)cin * elem};
  cuuint32_t box[2] = {(cuuint32_t)(row_bytes / elem), 1};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMapSwizzle sw = row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                           : (row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  CUresult r = fn(map, fp32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                  const_cast<void *>(features), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("conv_fwd(tc): cuTensorMapEncodeTiled failed with %d (rows=%lld cin=%d)", (int)r, (long long)rows, cin);
    return FV2P_ERR_INVALID;
  }
  return 0;
}

int g_tc_packed = 1;  // packed stages for narrow rows (debug switch: fv2p_debug_packed)

template <bool kFp32, int N, bool kPacked>
int launch_one(const void *features, int64_t feat_rows, const void *weight, const int *nbr, int64_t nbr_stride,
               const int *row_perm, const int *tile_order, int *sched, int kvol, int64_t n_out_cap,
               const int *n_out_dev, int cin, const Epilogue &ep, cudaStream_t stream) {
  using C = Cfg<kFp32, N, kPacked>;
  static bool configured[64] = {false};  // the attribute is per device
  const int dev = current_device();
  if (!configured[dev]) {
    int st = cuda_status(cudaFuncSetAttribute(conv_tc_kernel<kFp32, N, kPacked>,
                                              cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes),
                         "conv_fwd(tc) smem attribute");
    if (st) return st;
    configured[dev] = true;
  }
  CUtensorMap map;
  int st = make_feature_map(&map, features, feat_rows, cin, kFp32);
  if (st) return st;
  // A-tile producer.  Default (auto): the TMA gather (cp.async.bulk.tensor tile::gather4: one instruction per lane
  // fetches four rows, a warp covers the stage with one) for every kernel whose stage holds ONE offset; packed stages
  // always use the swizzled cp.async gather (their rows interleave several offsets).  With the cp.async gather a
  // producer warp spends ~2600 cycles issuing the 32 copies per lane of a stage (role timers, profiles/r2_notes.md),
  // which became the longest leg of a slot's round trip once the fp32 MMA count dropped: waymo_b4 2.73 -> 2.67 ms
  // fp32, 1.79 -> 1.77 bf16, kitti_b8 1.072 -> 1.050 / 0.779 -> 0.775.  (In round 1 it was measured faster only for
  // fp32 rows of 128 bytes and more on KITTI-sized layers.)  fv2p_tc_gather_mode(0) forces cp.async.
  const int use_tma = (!kPacked && g_tc_gather_mode != 0) ? 1 : 0;
  int64_t tiles = (n_out_cap + kTileM - 1) / kTileM;
  int grid = (int)(tiles < sm_count() ? tiles : sm_count());
  if (grid < 1) grid = 1;
  // Programmatic dependent launch: the layers of a backbone follow each other on one stream, and this kernel's
  // set-up (barriers, TMEM, the scheduler's first tiles, which only read geometry) does not depend on the previous
  // layer.  Its CTAs may therefore start while the previous kernel drains; the roles that touch feature memory
  // execute griddepcontrol.wait first.
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3((unsigned)C::kThreads);
  cfg.dynamicSmemBytes = C::kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = g_tc_pdl ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  const uint8_t *wp = static_cast<const uint8_t *>(weight);
  const int oob = (int)feat_rows;
  return cuda_status(cudaLaunchKernelEx(&cfg, conv_tc_kernel<kFp32, N, kPacked>, map, features, wp, nbr, nbr_stride,
                                        row_perm, tile_order, sched, kvol, n_out_cap, n_out_dev, cin, oob, use_tma, ep),
                     "conv_fwd(tc)");
}

}  // namespace

int launch_conv_tc(const void *features, int64_t feat_rows, const void *weight, const int *nbr, int64_t nbr_stride,
                   const int *row_perm, const int *tile_order, int *sched, int kvol, int64_t n_out_cap,
                   const int *n_out_dev, int cin, int cout, const float *bias, const float *scale, const float *shift,
                   const void *residual, int relu, int mode, void *out, cudaStream_t stream) {
  if (!tc_shape_ok(cin, cout)) {
    set_error("conv_fwd: tensor-core modes need cin in {16,32,64,128,256} and cout in {16,32,64,128} (got %d->%d)",
              cin, cout);
    return FV2P_ERR_INVALID;
  }
  if (feat_rows < 1 || feat_rows >= (1ll << 31) - 1) {
    set_error("conv_fwd: tensor-core modes need the input row capacity (got %lld)", (long long)feat_rows);
    return FV2P_ERR_INVALID;
  }
  Epilogue ep{bias, scale, shift, residual, out, relu};
  const bool fp32 = mode == FV2P_MODE_FP32_TC;
  // packed stages: rows narrower than 128 bytes whose 27 weight images fit the resident area
  const int row_bytes_in = cin * (fp32 ? 4 : 2);
  const bool packed = g_tc_packed && row_bytes_in < 128 && kvol <= 27 && (fp32 ? cout <= 32 : cout <= 64);
#define FV2P_TC_ARGS features, feat_rows, weight, nbr, nbr_stride, row_perm, tile_order, sched, kvol, n_out_cap, n_out_dev, cin, ep, stream
#define FV2P_TC(NN) return fp32 ? launch_one<true, NN, false>(FV2P_TC_ARGS) : launch_one<false, NN, false>(FV2P_TC_ARGS)
#define FV2P_TCP(NN) return fp32 ? launch_one<true, NN, true>(FV2P_TC_ARGS) : launch_one<false, NN, true>(FV2P_TC_ARGS)
  if (packed) {
    switch (cout) {
      case 16: FV2P_TCP(16);
      case 32: FV2P_TCP(32);
      default:  // 64, bf16 only
        return launch_one<false, 64, true>(FV2P_TC_ARGS);
    }
  }
  switch (cout) {
    case 16: FV2P_TC(16);
    case 32: FV2P_TC(32);
    case 64: FV2P_TC(64);
    default: FV2P_TC(128);
  }
#undef FV2P_TCP
#undef FV2P_TC_ARGS
#undef FV2P_TC
}

}  // namespace fv2p

using namespace fv2p;

extern "C" int fv2p_tc_gather_mode(int mode) {
  g_tc_gather_mode = mode < 0 ? -1 : (mode > 1 ? 1 : mode);
  return FV2P_OK;
}

extern "C" __attribute__((visibility("default"))) int fv2p_debug_packed(int v) {
  g_tc_packed = v ? 1 : 0;
  return 0;
}

extern "C" __attribute__((visibility("default"))) int fv2p_debug_pdl(int v) {
  g_tc_pdl = v ? 1 : 0;
  return 0;
}

// copies the stamp table to `out` ([256][4] u64), returns the number of stamps taken and resets the counter
extern "C" __attribute__((visibility("default"))) int fv2p_debug_stamps(unsigned long long *out) {
  unsigned int n = 0, zero = 0;
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(&n, g_tc_stamp_n, sizeof(n));
  cudaMemcpyFromSymbol(out, g_tc_stamps, sizeof(unsigned long long) * kStampSlots * 4);
  cudaMemcpyToSymbol(g_tc_stamp_n, &zero, sizeof(zero));
  return (int)n;
}

#ifdef FV2P_TC_TIMERS
extern "C" __attribute__((visibility("default"))) int fv2p_debug_timers(unsigned long long *out) {
  cudaDeviceSynchronize();
  return (int)cudaMemcpyFromSymbol(out, g_tc_timers, sizeof(unsigned long long) * 160 * 16);
}
#endif

extern "C" __attribute__((visibility("default"))) int fv2p_debug_set(int v) {
  return (int)cudaMemcpyToSymbol(g_tc_debug, &v, sizeof(int));
}

extern "C" size_t fv2p_pack_weight_bytes(int kvol, int cin, int cout, int mode) {
  if (kvol < 1 || kvol > FV2P_MAX_KVOL || !tc_shape_ok(cin, cout)) return 0;
  if (mode == FV2P_MODE_BF16_TC) return (size_t)kvol * cin * cout * 2;
  if (mode == FV2P_MODE_FP32_TC) return (size_t)kvol * cin * cout * 4;  // W_hi and W_lo as bf16
  return 0;
}

extern "C" int fv2p_pack_weight(const float *weight_f32, int kvol, int cin, int cout, int mode, void *packed,
                                fv2p_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FV2P_REQUIRE(weight_f32 && packed, "pack_weight: null pointer argument");
  FV2P_REQUIRE(fv2p_pack_weight_bytes(kvol, cin, cout, mode) > 0, "pack_weight: unsupported shape %d x %d->%d mode %d",
               kvol, cin, cout, mode);
  if (mode == FV2P_MODE_FP32_TC)
    pack_weight_kernel<true><<<persistent_grid(), kThreads, 0, stream>>>(weight_f32, kvol, cin, cout,
                                                                         static_cast<uint8_t *>(packed));
  else
    pack_weight_kernel<false><<<persistent_grid(), kThreads, 0, stream>>>(weight_f32, kvol, cin, cout,
                                                                          static_cast<uint8_t *>(packed));
  FV2P_LAUNCH_CHECK("pack_weight");
  return FV2P_OK;
}
