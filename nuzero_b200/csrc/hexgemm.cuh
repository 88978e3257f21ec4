// Fused neighbour-gather + GEMM for the network's hexagonal / orthogonal convolutions on sm_100a:
// tcgen05.mma with TMEM accumulators, the im2col matrix never exists in memory.
//
//   out[m, n] = act( sum_{tap, c} x[row(m, tap), c] * wt[n, tap * Cin + c]  (+ residual[m, n]) )
//   row(m, tap) = (m / RC) * RC + nbr[(m % RC) * taps + tap]   (nbr < 0 -> the tap reads zeros)
//
// One CTA computes a 256 x N tile (N <= 256): two 128-row accumulators share every B stage, so a K
// chunk of 64 moves 32 KB of A + 32 KB of B for 2 x (128 x 256 x 64) MMAs — the same bytes per FLOP as a
// 256 x 256 2-CTA cuBLAS tile.  Eight producer warps gather A rows (128 bytes each) and B rows with
// 16-byte loads into SWIZZLE_128B K-major tiles, one thread of the ninth warp issues the MMAs, the
// producers turn into the epilogue (tcgen05.ld -> + residual -> ReLU -> bf16 -> global).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace nzg {

constexpr int BLOCK_M = 256, BLOCK_K = 64, STAGES = 3, PRODUCERS = 256, THREADS = PRODUCERS + 32;
constexpr int A_HALF_BYTES = 128 * BLOCK_K * 2;             // 16 KB: one 128-row accumulator's A tile
constexpr int B_BYTES = 256 * BLOCK_K * 2;                  // 32 KB
constexpr int STAGE_BYTES = 2 * A_HALF_BYTES + B_BYTES;     // 64 KB
constexpr int MAX_TAPS = 9;
// stages + barriers + the per-tap source-row table.  Kept under 195 KB so that the 196 KB shared-memory carve-out is
// enough and ~60 KB of the SM's 256 KB stay L1 (the gathered x slab lives there, see TAPS_INNER).
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 256 + MAX_TAPS * BLOCK_M;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  const uint32_t addr = smem_u32(b);
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!ok);
}
// K-major operand tile, SWIZZLE_128B, rows of 128 bytes, 8-row groups 1024 bytes apart (cute::UMMA::SmemDescriptor)
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3ffffu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
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

// TAPS_INNER: K runs channel-chunk-major with the taps innermost, so the `taps` consecutive chunks of one 64-channel
// slice gather the SAME 256 x 128-byte slab of x (each row once per tap that reaches it) and only the first touch goes
// to L2 — the copies allocate in L1 (cp.async.ca) and the slab (32 KB) fits beside the 3 x 64 KB of stages.  With
// TAPS_INNER = false K is tap-major (the weight matrix's own order) and the copies bypass L1 (cp.async.cg).
template <bool TAPS_INNER>
__global__ void __launch_bounds__(THREADS, 1) hexconv_kernel(const __grid_constant__ CUtensorMap tm_w, const Params p) {
  extern __shared__ __align__(1024) unsigned char smem[];  // SWIZZLE_128B atoms are 1024-byte aligned
  uint64_t* bars = (uint64_t*)(smem + STAGES * STAGE_BYTES);
  uint64_t* full = bars;               // [STAGES]  producers -> MMA
  uint64_t* empty = bars + STAGES;     // [STAGES]  MMA (tcgen05.commit) -> producers
  uint64_t* accum = bars + 2 * STAGES; // MMA -> epilogue
  uint32_t* tmem_slot = (uint32_t*)(bars + 2 * STAGES + 1);
  signed char* srcdelta = (signed char*)(smem + STAGES * STAGE_BYTES + 256);  // [taps][BLOCK_M]: source row - own row, -128 = zeros

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int K = p.taps * p.cin, n_chunks = K / BLOCK_K, chunks_per_tap = p.cin / BLOCK_K;
  const size_t m0 = (size_t)blockIdx.x * BLOCK_M;

  if (TAPS_INNER) {
    for (int i = tid; i < p.taps * BLOCK_M; i += THREADS) {
      const int tap = i / BLOCK_M, r = i - tap * BLOCK_M;
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
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], PRODUCERS / 32 + 1); mbar_init(&empty[s], 1); }
    mbar_init(accum, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_w) : "memory");
  }
  if (warp == PRODUCERS / 32) {  // the MMA warp owns the tensor-memory allocation: 2 x 256 f32 columns
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp < PRODUCERS / 32) {
    // ===================== producers: gather A, stream B ==========================================
    const int c16 = tid & 7;        // which 16-byte piece of a 128-byte row
    const int r0 = tid >> 3;        // rows r0 + 32 j
    const int sw = (c16 ^ (r0 & 7)) * 16;  // swizzled piece offset: (r0 + 32 j) & 7 == r0 & 7
    int cell[8];
    long long rowbase[8];           // (m / RC) * RC, or -1 beyond the last row
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const size_t m = m0 + r0 + 32 * j;
      if (m < (size_t)p.rows) { cell[j] = (int)(m % p.RC); rowbase[j] = (long long)(m - cell[j]); }
      else { cell[j] = 0; rowbase[j] = -1; }
    }
    // cp.async (16 bytes, L2 only) straight into the swizzled tiles; an off-board tap or a row beyond the end is
    // a zero-fill copy (src-size 0).  Two chunks are kept in flight per thread: chunk kc is published (proxy
    // fence + barrier arrive) once cp.async.wait_group says its copies have landed, while kc+1 and kc+2 load.
    // Address generation is kept off the critical path: per row a running source pointer (+128 bytes per chunk,
    // re-derived from the neighbour table only when the tap changes) and a constant shared-memory offset.
    constexpr int AHEAD = 2;
    uint32_t dst_a[8];
    const char* src_a[8];
    uint32_t bytes_a[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int r = r0 + 32 * j, half = r >> 7, rr = r & 127;
      dst_a[j] = half * A_HALF_BYTES + (rr >> 3) * 1024 + (rr & 7) * 128 + sw;
      src_a[j] = (const char*)p.x;
      bytes_a[j] = 0u;
    }
    const uint32_t b_bytes = (uint32_t)p.n_pad * BLOCK_K * 2;
    const uint32_t smem_base = smem_u32(smem);
    int in_tap = 0, tap = 0;
    for (int kc = 0; kc < n_chunks + AHEAD; ++kc) {
      // publish chunk kc - AHEAD first: its copies were issued AHEAD iterations ago, and the MMA must not wait for it
      // behind this iteration's stage wait (the stage about to be refilled is freed by the MMA of chunk kc - STAGES)
      if (kc >= AHEAD) {
        asm volatile("cp.async.wait_group %0;" ::"n"(AHEAD - 1) : "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> async proxy (UMMA)
        __syncwarp();
        if (lane == 0) mbar_arrive(&full[(kc - AHEAD) % STAGES]);
        if (p.trace && blockIdx.x == 0 && tid == 0) p.trace[(kc - AHEAD) * 4 + 1] = clock64();  // chunk published by warp 0
      }
      if (kc < n_chunks) {
        const int s = kc % STAGES;
        if (kc >= STAGES) mbar_wait(&empty[s], ((kc / STAGES) - 1) & 1);
        if (p.trace && blockIdx.x == 0 && tid == 0) p.trace[kc * 4 + 0] = clock64();  // stage free, issue starts
        int b_col;  // column of this chunk in the weight matrix [n_pad, taps * cin]
        const uint32_t st = smem_base + s * STAGE_BYTES;
        if (TAPS_INNER) {
          const int cc = kc / p.taps, tp = kc - cc * p.taps;  // 64-channel slice, tap
          b_col = tp * p.cin + cc * BLOCK_K;
          const signed char* dl = srcdelta + tp * BLOCK_M + r0;
          const char* base = (const char*)(p.x + (m0 + r0) * (size_t)p.cin + cc * BLOCK_K + c16 * 8);
          const long long row_bytes = (long long)p.cin * 2;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int d = dl[32 * j];
            const bool ok = d != -128;
            const char* src = ok ? base + (long long)(32 * j + d) * row_bytes : (const char*)p.x;
            asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(st + dst_a[j]), "l"(src), "r"(ok ? 16u : 0u) : "memory");
          }
        } else {
          b_col = kc * BLOCK_K;
          if (in_tap == 0) {  // new tap: look the source rows up again
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              bytes_a[j] = 0u;
              src_a[j] = (const char*)p.x;
              if (rowbase[j] >= 0) {
                const int src = __ldg(p.nbr + cell[j] * p.taps + tap);
                if (src >= 0) { src_a[j] = (const char*)(p.x + (size_t)(rowbase[j] + src) * p.cin + c16 * 8); bytes_a[j] = 16u; }
              }
            }
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(st + dst_a[j]), "l"(src_a[j]), "r"(bytes_a[j]) : "memory");
            src_a[j] += bytes_a[j] * 8;  // next 64 channels of the same row (zero-fill rows stay put)
          }
        }
        if (tid == 0) {  // B: one tiled TMA box per chunk (64 x n_pad), counted in bytes on the same barrier
          const uint32_t bar = smem_u32(&full[s]);
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(b_bytes) : "memory");
          asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                       ::"r"(st + 2 * A_HALF_BYTES), "l"(&tm_w), "r"(b_col), "r"(0), "r"(bar) : "memory");
        }
        if (++in_tap == chunks_per_tap) { in_tap = 0; ++tap; }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
    // ===================== epilogue: TMEM -> registers -> global ===================================
    // Each warp stages its 32 rows x 512 bytes in stage memory.  Global traffic is row-wise (one coalesced 512-byte
    // request per row), TMEM traffic is thread = row; the 16-byte pieces of a row are XOR-swizzled with the row number so
    // that both access patterns are free of bank conflicts.  The residual rows are fetched while the last two chunks
    // are still in the tensor pipe: warps 0-3 reuse the stage of chunk n-3 as soon as its MMAs have read it, warps 4-7
    // the stage of chunk n-2.
    const int half = warp >> 2, q = warp & 3;            // accumulator, TMEM lane quarter of this warp
    const int row0 = half * 128 + q * 32;                // first tile row of this warp
    const size_t mrow0 = m0 + row0;
    const int nch = min(p.ldo, p.n_pad) >> 3;            // 16-byte pieces per output row that this kernel produces
    const bool early = p.residual != nullptr && n_chunks >= STAGES;
    const int kc_reuse = early ? n_chunks - STAGES + half : 0;
    unsigned char* stg = smem + (early ? kc_reuse % STAGES : half) * STAGE_BYTES + q * (32 * 512);
    if (p.residual) {
      if (early) mbar_wait(&empty[kc_reuse % STAGES], (kc_reuse / STAGES) & 1);
      else mbar_wait(accum, 0);
      const uint32_t stg_u32 = smem_u32(stg);
#pragma unroll 8
      for (int r = 0; r < 32; ++r) {
        const int c = (lane ^ r) & 31;
        if (mrow0 + r < (size_t)p.rows && c < nch)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(stg_u32 + r * 512 + lane * 16),
                       "l"(p.residual + (mrow0 + r) * p.ldo + c * 8) : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
    mbar_wait(accum, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (p.trace && blockIdx.x == 0 && tid == 0) p.trace[n_chunks * 4 + 0] = clock64();  // epilogue starts
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
    const uint32_t taddr0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * 256);
    for (int n0 = 0; n0 < p.n_pad; n0 += 16) {
      uint32_t r[16];
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
            "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
          : "r"(taddr0 + (uint32_t)n0)
          : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (n0 < p.ldo) {
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
      }
    }
    __syncwarp();
#pragma unroll 8
    for (int r = 0; r < 32; ++r) {
      const int c = (lane ^ r) & 31;
      if (mrow0 + r < (size_t)p.rows && c < nch)
        *(uint4*)(p.out + (mrow0 + r) * p.ldo + c * 8) = *(const uint4*)(stg + r * 512 + lane * 16);
    }
    if (p.trace && blockIdx.x == 0 && tid == 0) p.trace[n_chunks * 4 + 1] = clock64();  // epilogue done
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  } else {
    // ===================== MMA issuer: one thread ===================================================
    if (lane == 0) {
      // kind::f16, A = B = bf16, D = f32, both K-major, M = 128, N = n_pad  (cute::UMMA::InstrDescriptor)
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.n_pad >> 3) << 17) | ((128u >> 4) << 24);
      for (int kc = 0; kc < n_chunks; ++kc) {
        const int s = kc % STAGES;
        mbar_wait(&full[s], (kc / STAGES) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (p.trace && blockIdx.x == 0) p.trace[kc * 4 + 2] = clock64();  // MMA sees the chunk
        const uint32_t a0 = smem_u32(smem + s * STAGE_BYTES), a1 = a0 + A_HALF_BYTES, b0 = a0 + 2 * A_HALF_BYTES;
#pragma unroll
        for (int k = 0; k < BLOCK_K / 16; ++k) {  // UMMA_K = 16 bf16 = 32 bytes inside the 128-byte swizzle atom
          const uint64_t db = umma_desc(b0 + k * 32);
          umma_bf16(tmem_base, umma_desc(a0 + k * 32), db, idesc, (kc | k) != 0);
          umma_bf16(tmem_base + 256, umma_desc(a1 + k * 32), db, idesc, (kc | k) != 0);
        }
        umma_commit(&empty[s]);  // frees the stage once these MMAs have read it
      }
      umma_commit(accum);
    }
    __syncwarp();
  }
  __syncthreads();
  if (warp == PRODUCERS / 32) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}


// ---------------------------------------------------------------------------------------------------------
// TMA-fed variant: the producer is ONE warp.  A rows come in with cp.async.bulk.tensor ... tile::gather4 (four
// arbitrary rows of x per instruction, 64 channels wide, written with the 128-byte swizzle; a row index beyond
// the tensor reads zeros, which is how off-board taps and the ragged last tile are handled), B with one tiled
// TMA box per chunk.  Completion is counted in bytes on the stage's mbarrier, so there are no per-thread
// copies, no proxy fences and no LSU work in the main loop.
// ---------------------------------------------------------------------------------------------------------
constexpr int TMA_THREADS = 64 + 256;  // warp 0 producer, warp 1 MMA, warps 2..9 epilogue

__global__ void __launch_bounds__(TMA_THREADS, 1)
hexconv_tma_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w, const Params p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t* bars = (uint64_t*)(smem + STAGES * STAGE_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* accum = bars + 2 * STAGES;
  uint32_t* tmem_slot = (uint32_t*)(bars + 2 * STAGES + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int K = p.taps * p.cin, n_chunks = K / BLOCK_K, chunks_per_tap = p.cin / BLOCK_K;
  const size_t m0 = (size_t)blockIdx.x * BLOCK_M;

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(accum, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_x) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_w) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer ==============================================================
    // lane l owns the 4-row groups l and l + 32 of the 256-row tile
    int cell[8];
    long long rowbase[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const size_t m = m0 + (size_t)((lane + 32 * (j >> 2)) * 4 + (j & 3));
      if (m < (size_t)p.rows) { cell[j] = (int)(m % p.RC); rowbase[j] = (long long)(m - cell[j]); }
      else { cell[j] = 0; rowbase[j] = -1; }
    }
    uint32_t dst[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int r = (lane + 32 * h) * 4, half = r >> 7, rr = r & 127;
      dst[h] = half * A_HALF_BYTES + (rr >> 3) * 1024 + (rr & 7) * 128;
    }
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t stage_bytes = 2 * A_HALF_BYTES + (uint32_t)p.n_pad * BLOCK_K * 2;
    int idx[8];
    int in_tap = 0, tap = 0;
    for (int kc = 0; kc < n_chunks; ++kc) {
      const int s = kc % STAGES;
      if (kc >= STAGES) mbar_wait(&empty[s], ((kc / STAGES) - 1) & 1);
      if (in_tap == 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          idx[j] = p.rows;  // beyond the tensor: TMA fills zeros
          if (rowbase[j] >= 0) {
            const int src = __ldg(p.nbr + cell[j] * p.taps + tap);
            if (src >= 0) idx[j] = (int)(rowbase[j] + src);
          }
        }
      }
      const uint32_t st = smem_base + s * STAGE_BYTES, bar = smem_u32(&full[s]);
      if (lane == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(stage_bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     ::"r"(st + 2 * A_HALF_BYTES), "l"(&tm_w), "r"(kc * BLOCK_K), "r"(0), "r"(bar) : "memory");
      }
      __syncwarp();
      const int col = in_tap * BLOCK_K;
#pragma unroll
      for (int h = 0; h < 2; ++h)
        asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                     ::"r"(st + dst[h]), "l"(&tm_x), "r"(col), "r"(idx[4 * h]), "r"(idx[4 * h + 1]), "r"(idx[4 * h + 2]),
                     "r"(idx[4 * h + 3]), "r"(bar) : "memory");
      if (++in_tap == chunks_per_tap) { in_tap = 0; ++tap; }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: one thread =====================================================
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.n_pad >> 3) << 17) | ((128u >> 4) << 24);
      for (int kc = 0; kc < n_chunks; ++kc) {
        const int s = kc % STAGES;
        mbar_wait(&full[s], (kc / STAGES) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a0 = smem_u32(smem + s * STAGE_BYTES), a1 = a0 + A_HALF_BYTES, b0 = a0 + 2 * A_HALF_BYTES;
#pragma unroll
        for (int k = 0; k < BLOCK_K / 16; ++k) {
          const uint64_t db = umma_desc(b0 + k * 32);
          umma_bf16(tmem_base, umma_desc(a0 + k * 32), db, idesc, (kc | k) != 0);
          umma_bf16(tmem_base + 256, umma_desc(a1 + k * 32), db, idesc, (kc | k) != 0);
        }
        umma_commit(&empty[s]);
      }
      umma_commit(accum);
    }
    __syncwarp();
  } else {
    // ===================== epilogue: TMEM -> registers -> global ======================================
    mbar_wait(accum, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int q = warp & 3, half = (warp - 2) >> 2;  // TMEM lane quarter of this warp, accumulator
    const size_t m = m0 + (size_t)half * 128 + q * 32 + lane;
    const uint32_t taddr0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * 256);
    for (int n0 = 0; n0 < p.n_pad; n0 += 16) {
      uint32_t r[16];
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
            "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
          : "r"(taddr0 + (uint32_t)n0)
          : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (m < (size_t)p.rows && n0 < p.ldo) {
        float f[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(r[i]);
        if (p.residual) {
          const uint4* rp = (const uint4*)(p.residual + m * p.ldo + n0);
          const uint4 a = rp[0], b = rp[1];
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
        uint4* op = (uint4*)(p.out + m * p.ldo + n0);
        op[0] = o0;
        op[1] = o1;
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

}  // namespace nzg
