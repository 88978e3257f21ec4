// Fused neighbour-gather + GEMM for the network's hexagonal / orthogonal convolutions on sm_100a:
// tcgen05.mma with TMEM accumulators, the im2col matrix never exists in memory.
//
//   out[m, n] = act( sum_{tap, c} x[row(m, tap), c] * wt[n, tap * Cin + c]  (+ residual[m, n]) )
//   row(m, tap) = (m / RC) * RC + nbr[(m % RC) * taps + tap]   (nbr < 0 -> the tap reads zeros)
//
// One CTA owns 256 output rows x N (N <= 256): two 128-row accumulators in tensor memory (2 x 256 fp32 columns = all
// 512) share every B stage.  Eight producer warps gather A rows (one 128-byte segment per row and tap, zero-fill
// cp.async for off-board taps) into SWIZZLE_128B K-major tiles, B arrives as one tiled TMA box per K chunk, one thread
// of the ninth warp issues the MMAs, and the producers then become the epilogue (tcgen05.ld -> + residual -> ReLU ->
// bf16 -> shared memory -> coalesced rows).
//
// PAIR = true runs two such CTAs as a cluster on one TPC with tcgen05.mma.cta_group::2: one instruction multiplies
// both CTAs' 128-row A tiles with a B tile that is split between the two shared memories (each CTA loads only its
// half of the weight rows), which halves the B bytes per CTA from L2 and the operand bytes each SM's tensor core reads
// from shared memory — the single-CTA form needs 160 bytes/clk of shared-memory bandwidth at the MMA rate, more than
// the 128 an SM has.  The leader CTA issues every MMA; completion is multicast to both CTAs' barriers.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace nzg {

constexpr int BLOCK_M = 256, BLOCK_K = 64, PRODUCERS = 256, THREADS = PRODUCERS + 32;
constexpr int A_HALF_BYTES = 128 * BLOCK_K * 2;  // 16 KB: one 128-row accumulator's A tile
constexpr int MAX_TAPS = 9;
constexpr int SMEM_TILES = 196608;               // 3 stages x 64 KB (single CTA) or 4 stages x 48 KB (pair)
// stages + barriers + the per-tap source-row table.  Kept under 195 KB so that the 196 KB shared-memory carve-out is
// enough and ~60 KB of the SM's 256 KB stay L1 (the gathered x slab lives there, see TAPS_INNER).
constexpr int SMEM_BYTES = SMEM_TILES + 256 + MAX_TAPS * BLOCK_M;

// HALVES = 2: a CTA owns 256 rows (two 128-row accumulators).  HALVES = 1: 128 rows, one accumulator — for small batches:
// the stages are smaller, so more K chunks are in flight (6 instead of 4 stages in the pair form) and a lone tile's K loop
// is no longer bounded by "two chunks per L2 round trip".
template <bool PAIR, int HALVES>
struct Cfg {
  static constexpr int B_BYTES = (PAIR ? 128 : 256) * BLOCK_K * 2;     // this CTA's rows of the weight chunk (at most)
  static constexpr int STAGE_BYTES = HALVES * A_HALF_BYTES + B_BYTES;  // 48 / 64 KB (HALVES = 2), 32 / 48 KB (HALVES = 1)
  static constexpr int STAGES = SMEM_TILES / STAGE_BYTES;              // 4 / 3, 6 / 4
  static_assert(STAGES * STAGE_BYTES <= SMEM_TILES && STAGES <= 6, "stage memory");
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
// this CTA's barrier, or (PAIR) the barrier at the same offset in the leader CTA of the pair: shared window addresses
// of the two CTAs differ in bit 24 (cute::Sm100MmaPeerBitMask), the leader is the even one
template <bool PAIR>
__device__ __forceinline__ uint32_t leader_addr(const void* p) {
  return PAIR ? (smem_u32(p) & 0xFEFFFFFFu) : smem_u32(p);
}
template <bool PAIR>
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
  // default semantics (release at CTA scope), as cutlass::arch::ClusterBarrier::arrive(cta_id): a cluster-scope release
  // costs a MEMBAR that waits for every cp.async still in flight and invalidates L1 (ncu: 2.1 membar stall cycles per
  // issue, the pipeline collapses).  The data never leaves this CTA's shared memory — only the count crosses.
  if (PAIR) asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(leader_addr<true>(b)) : "memory");
  else asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
template <bool CLUSTER_SCOPE>
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  const uint32_t addr = smem_u32(b);
  uint32_t ok;
  do {
    if (CLUSTER_SCOPE)
      asm volatile(
          "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
    else
      asm volatile(
          "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
  } while (!ok);
}
// K-major operand tile, SWIZZLE_128B, rows of 128 bytes, 8-row groups 1024 bytes apart (cute::UMMA::SmemDescriptor)
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3ffffu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
template <bool PAIR>
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  if (PAIR)
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
  else
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
// completion of all MMAs issued so far -> one arrival on `bar` (PAIR: on the barrier at that offset in both CTAs)
template <bool PAIR>
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  if (PAIR)
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
  else
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
template <bool PAIR>
__device__ __forceinline__ void block_or_cluster_sync() {
  if (PAIR) {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  } else {
    __syncthreads();
  }
}

struct Params {
  const __nv_bfloat16* x;         // [rows, cin]   cells-major, channels last
  const int32_t* nbr;             // [RC, taps]
  const __nv_bfloat16* wt;        // [n_pad, taps * cin]  (= W^T, K contiguous)
  const __nv_bfloat16* residual;  // [rows, ldo] or null
  __nv_bfloat16* out;             // [rows, ldo]
  int rows, RC, taps, cin, n_pad, ldo, relu_in, relu_out;
  long long* trace;               // optional: per-chunk clock64 stamps of CTA 0 (profiling aid)
};

// TAPS_INNER = false (default): K is tap-major (the weight matrix's own order), the gathers bypass L1 (cp.async.cg).
// TAPS_INNER = true (experiment, kept for comparison): K runs channel-chunk-major with the taps innermost, so the `taps`
// consecutive chunks of one 64-channel slice gather the SAME 256 x 128-byte slab of x and only the first touch goes to L2
// (cp.async.ca, the 32 KB slab fits in the L1 left beside the stages).  Measured: L2 sectors -45 %, L1 hit rate 59 %, but
// an L1-allocating LDGSTS issues three times slower (1100 vs 330 cycles per chunk), a net loss.
// GROUPS: the eight producer warps work as GROUPS independent teams that take the K chunks in turn (chunk kc belongs to team
// kc % GROUPS).  A producer thread's iteration — wait for its copies, proxy fence, barrier arrive, wait for a free stage,
// issue — costs ~450 cycles of latency besides the copies themselves (measured: a lone tile's K loop ran at ~740 cycles per
// chunk even with every copy a zero-byte one); with teams that latency is paid once per GROUPS chunks and the loop runs at
// the rate of the LSU (~9 cycles per 512-byte LDGSTS) or of the tensor pipe.  AHEAD counts a team's own chunks.
// TMA_A: the A rows are gathered by the TMA unit (cp.async.bulk.tensor.2d.tile::gather4: four rows of 64 channels per
// instruction, swizzled on arrival, rows beyond the tensor read as zeros) instead of 16-byte cp.async copies: one producer
// warp issues 32 (64) instructions per chunk and the bytes are counted on the stage barrier — no wait_group / proxy fence /
// arrive sequence in the producers.  `tm_x` describes x as [rows, cin] with a 64 x 1 box.
template <bool TAPS_INNER, bool PAIR, int AHEAD, int HALVES, int GROUPS = 1, bool TMA_A = false>
__global__ void __launch_bounds__(THREADS, 1) hexconv_kernel(const __grid_constant__ CUtensorMap tm_w,
                                                             const __grid_constant__ CUtensorMap tm_x, const Params p) {
  using C = Cfg<PAIR, HALVES>;
  constexpr int STAGES = C::STAGES, STAGE_BYTES = C::STAGE_BYTES;
  constexpr int BM = 128 * HALVES;   // rows of this CTA's tile
  constexpr int TPG = PRODUCERS / GROUPS;     // threads of a producer team
  constexpr int RSTEP = TPG / 8;              // tile rows between two copies of one thread (8 threads per 128-byte row)
  constexpr int NJ = 4 * HALVES * GROUPS;     // 16-byte copies per producer thread and K chunk of its team
  static_assert(AHEAD >= 1 && AHEAD * GROUPS <= STAGES && (GROUPS == 1 ? AHEAD < STAGES : true), "chunks in flight");
  static_assert(!TAPS_INNER || GROUPS == 1, "the taps-innermost experiment keeps one team");
  static_assert(!TMA_A || (!TAPS_INNER && GROUPS == 1), "the TMA gather has one producer warp");
  static_assert(RSTEP % 8 == 0 && TPG % 32 == 0, "team shape");
  extern __shared__ __align__(1024) unsigned char smem[];  // SWIZZLE_128B atoms are 1024-byte aligned
  uint64_t* bars = (uint64_t*)(smem + SMEM_TILES);
  uint64_t* full = bars;               // [STAGES]  producers (of both CTAs) + TMA bytes -> MMA (PAIR: the leader's copy is used)
  uint64_t* empty = bars + STAGES;     // [STAGES]  MMA (tcgen05.commit) -> producers
  uint64_t* accum = bars + 2 * STAGES; // MMA -> epilogue
  uint32_t* tmem_slot = (uint32_t*)(bars + 2 * STAGES + 1);
  signed char* srcdelta = (signed char*)(smem + SMEM_TILES + 256);  // [taps][BLOCK_M]: source row - own row, -128 = zeros

  // Programmatic dependent launch: the next kernel in the stream (the next layer) may be scheduled right away — it sets
  // itself up (barriers, tensor memory, source-row table) on SMs this layer leaves free and then waits for this grid.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int K = p.taps * p.cin, n_chunks = K / BLOCK_K, chunks_per_tap = p.cin / BLOCK_K;
  const size_t m0 = (size_t)blockIdx.x * BM;
  uint32_t cta_rank = 0;
  if (PAIR) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(cta_rank));
  // blockIdx.y selects a slice of p.n_pad output channels (small batches are split over the channels so that more SMs work
  // on them and one CTA's K loop runs at the MMA rate of a narrower tile; p.n_pad is the width of ONE slice)
  const int col0 = (int)blockIdx.y * p.n_pad;
  const bool traced = p.trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0;

  {
    for (int i = tid; i < p.taps * BM; i += THREADS) {
      const int tap = i / BM, r = i - tap * BM;
      const size_t m = m0 + r;
      int d = -128;
      if (m < (size_t)p.rows) {
        const int cell = (int)(m % p.RC);
        const int nb = __ldg(p.nbr + cell * p.taps + tap);
        if (nb >= 0) d = nb - cell;
      }
      srcdelta[i] = (signed char)d;
    }
  }
  if (tid == 0) {
    // full: one arrival per warp of the chunk's producer team (of both CTAs) + the leader's expect_tx arrival for the B bytes
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], TMA_A ? 1 : (PAIR ? 2 : 1) * (TPG / 32) + 1); mbar_init(&empty[s], 1); }
    mbar_init(accum, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_w) : "memory");
    if (TMA_A) asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_x) : "memory");
  }
  if (warp == PRODUCERS / 32) {  // the MMA warp owns the tensor-memory allocation: 2 x 256 f32 columns
    if (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  block_or_cluster_sync<PAIR>();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp < PRODUCERS / 32) {
    // ===================== producers: gather A, stream B ==========================================
    // everything above touched only static data (neighbour table, tensor map); x / residual / out belong to the layers
    // before: wait for the previous kernel (no-op when this launch is not a programmatic dependent)
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (TMA_A) {
      {
        // TMA operands are warp-uniform: a warp issues one gather per active lane in turn, so the 32 row groups of a tile are
        // spread over the eight producer warps (four lanes each).  Group g = tile rows 4g .. 4g + 3 (+ 128 for the second
        // accumulator): destination = 512 contiguous bytes of the swizzled tile.
        const int g = warp * 4 + (lane & 3);
        const bool active = lane < 4;
        const int b_rows = PAIR ? p.n_pad >> 1 : p.n_pad;
        const uint32_t smem_base = smem_u32(smem);
        const uint32_t chunk_bytes = (uint32_t)((PAIR ? 2 : 1) * BM * BLOCK_K * 2 + p.n_pad * BLOCK_K * 2);
        int tap = 0, in_tap = 0;
        int rows4[HALVES][4];
        auto tap_rows = [&](int tp) {
          if (tp >= p.taps) return;
#pragma unroll
          for (int h = 0; h < HALVES; ++h) {
            const int r = h * 128 + 4 * g;
            const uint32_t d4 = *(const uint32_t*)(srcdelta + tp * BM + r);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int d = (int)(signed char)((d4 >> (8 * i)) & 0xffu);
              rows4[h][i] = d == -128 ? p.rows : (int)m0 + r + i + d;  // a row beyond the tensor reads as zeros
            }
          }
        };
        tap_rows(0);
        for (int kc = 0; kc < n_chunks; ++kc) {
          const int s = kc % STAGES;
          if (kc >= STAGES) mbar_wait<false>(&empty[s], ((kc / STAGES) - 1) & 1);
          if (traced && tid == 0) p.trace[kc * 4 + 0] = clock64();
          const uint32_t st = smem_base + s * STAGE_BYTES;
          const uint32_t bar = leader_addr<PAIR>(&full[s]);
          if (tid == 0) {
            if (!PAIR || cta_rank == 0)
              asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(chunk_bytes) : "memory");
            const int b_col = kc * BLOCK_K;
            if (PAIR)
              asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                           ::"r"(st + HALVES * A_HALF_BYTES), "l"(&tm_w), "r"(b_col), "r"(col0 + (int)cta_rank * b_rows), "r"(bar) : "memory");
            else
              asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                           ::"r"(st + HALVES * A_HALF_BYTES), "l"(&tm_w), "r"(b_col), "r"(col0), "r"(bar) : "memory");
          }
          const int col = in_tap * BLOCK_K;
          if (active) {
#pragma unroll
            for (int h = 0; h < HALVES; ++h) {
              const uint32_t dst = st + h * A_HALF_BYTES + g * 512;
              if (PAIR)
                asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                             ::"r"(dst), "l"(&tm_x), "r"(col), "r"(rows4[h][0]), "r"(rows4[h][1]), "r"(rows4[h][2]), "r"(rows4[h][3]), "r"(bar) : "memory");
              else
                asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                             ::"r"(dst), "l"(&tm_x), "r"(col), "r"(rows4[h][0]), "r"(rows4[h][1]), "r"(rows4[h][2]), "r"(rows4[h][3]), "r"(bar) : "memory");
            }
          }
          if (traced && tid == 0) p.trace[kc * 4 + 3] = clock64();
          if (++in_tap == chunks_per_tap) { in_tap = 0; tap_rows(++tap); }
        }
      }
    } else {
    const int grp = tid / TPG, tg = tid - grp * TPG;  // producer team, thread in the team
    const int c16 = tg & 7;         // which 16-byte piece of a 128-byte row
    const int r0 = tg >> 3;         // rows r0 + RSTEP j
    const int sw = (c16 ^ (r0 & 7)) * 16;  // swizzled piece offset: (r0 + RSTEP j) & 7 == r0 & 7
    // cp.async (16 bytes) straight into the swizzled tiles; an off-board tap or a row beyond the end is a zero-fill
    // copy (src-size 0).  AHEAD chunks (of this team) are kept in flight per thread: a chunk is published (proxy fence +
    // barrier arrive) once cp.async.wait_group says its copies have landed, while the following ones load.
    uint32_t dst_a[NJ];
    const char* src_a[NJ];   // source row of the current tap, first channel slice (TAPS_INNER: unused)
    uint32_t bytes_a[NJ];
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int r = r0 + RSTEP * j, half = r >> 7, rr = r & 127;
      dst_a[j] = half * A_HALF_BYTES + (rr >> 3) * 1024 + (rr & 7) * 128 + sw;
      src_a[j] = (const char*)p.x;
      bytes_a[j] = 0u;
    }
    // source rows of a tap from the shared-memory table (a global look-up on the issue path stalls the pipeline for an L2
    // round trip under load: +1500 cycles per tap in the trace); computed right after the last chunk of the tap before
    auto tap_sources = [&](int tp) {
      if (tp >= p.taps) return;
      const signed char* dl = srcdelta + tp * BM + r0;
      const char* base = (const char*)(p.x + (m0 + r0) * (size_t)p.cin + c16 * 8);
      const long long row_bytes = (long long)p.cin * 2;
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const int d = dl[RSTEP * j];
        const bool ok = d != -128;
        src_a[j] = ok ? base + (long long)(RSTEP * j + d) * row_bytes : (const char*)p.x;
        bytes_a[j] = ok ? 16u : 0u;
      }
    };
    // this team's next chunk: tap and 64-channel slice inside the tap (tap-major K order)
    int tap = grp / chunks_per_tap, in_tap = grp - tap * chunks_per_tap;
    if (!TAPS_INNER) tap_sources(tap);
    const int b_rows = PAIR ? p.n_pad >> 1 : p.n_pad;   // weight rows this CTA loads per chunk
    const uint32_t smem_base = smem_u32(smem);
    for (int kc = grp; kc < n_chunks + AHEAD * GROUPS; kc += GROUPS) {
      // publish the chunk issued AHEAD iterations ago first: the MMA must not wait for it behind this iteration's stage
      // wait (the stage about to be refilled is freed by the MMA of chunk kc - STAGES)
      if (kc >= AHEAD * GROUPS) {
        asm volatile("cp.async.wait_group %0;" ::"n"(AHEAD - 1) : "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> async proxy (UMMA)
        __syncwarp();
        if (lane == 0) mbar_arrive<PAIR>(&full[(kc - AHEAD * GROUPS) % STAGES]);
        if (traced && tg == 0) p.trace[(kc - AHEAD * GROUPS) * 4 + 1] = clock64();  // chunk published by the team's first warp
      }
      if (kc < n_chunks) {
        const int s = kc % STAGES;
        if (kc >= STAGES) mbar_wait<false>(&empty[s], ((kc / STAGES) - 1) & 1);
        if (traced && tg == 0) p.trace[kc * 4 + 0] = clock64();  // stage free, issue starts
        int b_col;  // column of this chunk in the weight matrix [n_pad, taps * cin]
        const uint32_t st = smem_base + s * STAGE_BYTES;
        if (TAPS_INNER) {
          const int cc = kc / p.taps, tp = kc - cc * p.taps;  // 64-channel slice, tap
          b_col = tp * p.cin + cc * BLOCK_K;
          const signed char* dl = srcdelta + tp * BM + r0;
          const char* base = (const char*)(p.x + (m0 + r0) * (size_t)p.cin + cc * BLOCK_K + c16 * 8);
          const long long row_bytes = (long long)p.cin * 2;
#pragma unroll
          for (int j = 0; j < NJ; ++j) {
            const int d = dl[RSTEP * j];
            const bool ok = d != -128;
            const char* src = ok ? base + (long long)(RSTEP * j + d) * row_bytes : (const char*)p.x;
            asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(st + dst_a[j]), "l"(src), "r"(ok ? 16u : 0u) : "memory");
          }
        } else {
          b_col = kc * BLOCK_K;
          const int ch_off = in_tap * (BLOCK_K * 2);  // byte offset of this chunk's channel slice inside a row
#pragma unroll
          for (int j = 0; j < NJ; ++j)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(st + dst_a[j]), "l"(src_a[j] + ch_off), "r"(bytes_a[j]) : "memory");
        }
        if (tg == 0) {
          // B: one tiled TMA box per chunk (64 x b_rows), counted in bytes on the (leader's) full barrier.  The leader
          // announces the bytes of both halves; its own arrival keeps the phase open until it has done so.
          const uint32_t bar = leader_addr<PAIR>(&full[s]);
          if (!PAIR || cta_rank == 0)
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)p.n_pad * BLOCK_K * 2) : "memory");
          if (PAIR)
            asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                         ::"r"(st + HALVES * A_HALF_BYTES), "l"(&tm_w), "r"(b_col), "r"(col0 + (int)cta_rank * b_rows), "r"(bar) : "memory");
          else
            asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                         ::"r"(st + HALVES * A_HALF_BYTES), "l"(&tm_w), "r"(b_col), "r"(col0), "r"(bar) : "memory");
        }
        if (!TAPS_INNER) {  // the team's next chunk; a new tap's source rows are looked up off the stage-free -> issue path
          in_tap += GROUPS;
          if (in_tap >= chunks_per_tap) {
            do { in_tap -= chunks_per_tap; ++tap; } while (in_tap >= chunks_per_tap);
            tap_sources(tap);
          }
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      if (traced && tg == 0 && kc < n_chunks) p.trace[kc * 4 + 3] = clock64();  // copies of chunk kc issued
    }
    }
    // ===================== epilogue: TMEM -> registers -> global ===================================
    // Each warp stages its 32 rows x 512 bytes in stage memory.  Global traffic is row-wise (one coalesced 512-byte
    // request per row), TMEM traffic is thread = row; the 16-byte pieces of a row are XOR-swizzled with the row number so
    // that both access patterns are free of bank conflicts.  The residual rows are fetched while the last chunks are
    // still in the tensor pipe: the warps reuse the stages of chunks n - STAGES, n - STAGES + 1, ... (WPS warps of
    // 16 KB each per stage) as soon as the MMAs have read them.
    // HALVES = 1: one accumulator; warps q and q + 4 share the 32 rows of lane quarter q and take one half of the columns each
    constexpr int WPS = STAGE_BYTES / (32 * 512);        // 16 KB row groups that fit one stage: 4 / 3 (HALVES = 2), 3 / 2
    const int half = HALVES == 2 ? warp >> 2 : 0;        // accumulator
    const int q = warp & 3;                              // TMEM lane quarter of this warp
    const int row0 = half * 128 + q * 32;                // first tile row of this warp
    const size_t mrow0 = m0 + row0;
    const int n_lim = max(0, min(p.n_pad, p.ldo - col0)); // output columns of this CTA's slice that exist in the row
    const int n_mid = HALVES == 2 ? n_lim : (((n_lim >> 4) + 1) >> 1) << 4;  // column split between warps q and q + 4
    const int n_beg = (HALVES == 1 && warp >= 4) ? n_mid : 0;
    const int n_end = (HALVES == 1 && warp < 4) ? n_mid : n_lim;
    const int c_beg = n_beg >> 3, c_end = n_end >> 3;    // 16-byte pieces of an output row that this warp moves
    const __nv_bfloat16* res0 = p.residual ? p.residual + col0 : nullptr;
    __nv_bfloat16* out0 = p.out + col0;
    const bool early = p.residual != nullptr && n_chunks >= STAGES;
    const int group = HALVES == 2 ? warp : q;            // 16 KB staging group (32 rows x 512 bytes)
    const int kc_reuse = early ? n_chunks - STAGES + group / WPS : group / WPS;
    unsigned char* stg = smem + (kc_reuse % STAGES) * STAGE_BYTES + (group % WPS) * (32 * 512);
    if (p.residual) {
      if (early) mbar_wait<false>(&empty[kc_reuse % STAGES], (kc_reuse / STAGES) & 1);
      else mbar_wait<false>(accum, 0);
      const uint32_t stg_u32 = smem_u32(stg);
#pragma unroll 8
      for (int r = 0; r < 32; ++r) {
        const int c = (lane ^ r) & 31;
        if (mrow0 + r < (size_t)p.rows && c >= c_beg && c < c_end)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(stg_u32 + r * 512 + lane * 16),
                       "l"(res0 + (mrow0 + r) * p.ldo + c * 8) : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
    mbar_wait<false>(accum, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (traced && tid == 0) p.trace[n_chunks * 4 + 0] = clock64();  // epilogue starts
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
    const uint32_t taddr0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * 256);
    auto finish16 = [&](const uint32_t (&r)[16], int n0) {  // 16 columns of this thread's row: + residual, ReLU, bf16, stage
      float f[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(r[i]);
      const int c0 = n0 >> 3;
      uint4* s0 = (uint4*)(stg + lane * 512 + ((c0 ^ lane) & 31) * 16);
      uint4* s1 = (uint4*)(stg + lane * 512 + (((c0 + 1) ^ lane) & 31) * 16);
      if (p.residual) {
        const uint4 a = *s0, b = *s1;
        const __nv_bfloat162* ha = (const __nv_bfloat162*)&a;
        const __nv_bfloat162* hb = (const __nv_bfloat162*)&b;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 xa = __bfloat1622float2(ha[i]), xb = __bfloat1622float2(hb[i]);
          f[2 * i] += xa.x; f[2 * i + 1] += xa.y; f[8 + 2 * i] += xb.x; f[8 + 2 * i + 1] += xb.y;
        }
      }
      if (p.relu_out) {
#pragma unroll
        for (int i = 0; i < 16; ++i) f[i] = fmaxf(f[i], 0.f);
      }
      uint4 o0, o1;
      __nv_bfloat162* h0 = (__nv_bfloat162*)&o0;
      __nv_bfloat162* h1 = (__nv_bfloat162*)&o1;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        h0[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
        h1[i] = __floats2bfloat162_rn(f[8 + 2 * i], f[8 + 2 * i + 1]);
      }
      *s0 = o0;
      *s1 = o1;
    };
#define NZ_TMEM_LD16(r, addr)                                                                                              \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"    \
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), \
                 "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])                  \
               : "r"(addr) : "memory")
    for (int n0 = n_beg; n0 < n_end; n0 += 32) {  // two TMEM loads in flight per wait
      uint32_t ra[16], rb[16];
      const bool second = n0 + 16 < n_end;  // warp-uniform
      NZ_TMEM_LD16(ra, taddr0 + (uint32_t)n0);
      if (second) NZ_TMEM_LD16(rb, taddr0 + (uint32_t)n0 + 16u);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      finish16(ra, n0);
      if (second) finish16(rb, n0 + 16);
    }
#undef NZ_TMEM_LD16
    __syncwarp();
#pragma unroll 8
    for (int r = 0; r < 32; ++r) {
      const int c = (lane ^ r) & 31;
      if (mrow0 + r < (size_t)p.rows && c >= c_beg && c < c_end)
        *(uint4*)(out0 + (mrow0 + r) * p.ldo + c * 8) = *(const uint4*)(stg + r * 512 + lane * 16);
    }
    if (traced && tid == 0) p.trace[n_chunks * 4 + 1] = clock64();  // epilogue done
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  } else {
    // ===================== MMA issuer: one thread (PAIR: of the leader CTA) ==========================
    if (lane == 0 && cta_rank == 0) {
      // kind::f16, A = B = bf16, D = f32, both K-major, N = n_pad, M = 128 (one CTA) or 256 (128 rows from each CTA)
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.n_pad >> 3) << 17) | ((PAIR ? 256u >> 4 : 128u >> 4) << 24);
      for (int kc = 0; kc < n_chunks; ++kc) {
        const int s = kc % STAGES;
        mbar_wait<false>(&full[s], (kc / STAGES) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (traced) p.trace[kc * 4 + 2] = clock64();  // MMA sees the chunk
        const uint32_t a0 = smem_u32(smem + s * STAGE_BYTES), a1 = a0 + A_HALF_BYTES, b0 = a0 + HALVES * A_HALF_BYTES;
#pragma unroll
        for (int k = 0; k < BLOCK_K / 16; ++k) {  // UMMA_K = 16 bf16 = 32 bytes inside the 128-byte swizzle atom
          const uint64_t db = umma_desc(b0 + k * 32);
          umma_bf16<PAIR>(tmem_base, umma_desc(a0 + k * 32), db, idesc, (kc | k) != 0);
          if (HALVES == 2) umma_bf16<PAIR>(tmem_base + 256, umma_desc(a1 + k * 32), db, idesc, (kc | k) != 0);
        }
        umma_commit<PAIR>(&empty[s]);  // frees the stage (in both CTAs) once these MMAs have read it
      }
      umma_commit<PAIR>(accum);
    }
    __syncwarp();
  }
  block_or_cluster_sync<PAIR>();
  if (warp == PRODUCERS / 32) {
    if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

}  // namespace nzg
