// Shared device helpers for the sm_100a self-play search kernels.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/nz_engine.h"

#define NZ_CTA_THREADS 128

namespace nz {

// Everything the kernels need, passed by value (fits the 4 KB kernel-parameter space easily).
struct View {
  // search config
  int G, P, max_depth, max_children, sims, training, policy_is_prob, auto_advance, games_per_slot;
  int max_sims_per_launch, record_detail, n_softmax_moves, compact, max_levels;
  int V;            // leaves one game may have waiting at the network (1 = the reference's sequencing)
  int ctable_len, tape_moves, tape_width, arena_words;
  int A;            // number of actions of the bound game
  int leaf_elems;   // C*R*Cc
  int state_words;  // compact game state, 32-bit words
  int scratch_extra; // bytes of per-game shared-memory tables behind Game::Scratch (SCS stack table)
  int slab_bytes, slab_words_bytes;  // per-tile shared-memory slab of the search kernels and its word part (host-computed)
  double pb_c_base, pb_c_init, value_factor, noise_frac, noise_alpha, noise_beta, eps_softmax, eps_random;
  unsigned long long seed;
  // node pool: one 32-byte record per node (= one DRAM sector), index = g*P + node.
  //   first 16 B : prior (f64; f32-chain games store the exact f32 value) | W value sum (f64)
  //   second 16 B: N visits (i32) | flags (bit 0: prior is a noised f64; bits 16-31: the action that leads here) |
  //                first child (u32) | n_children (u32)
  // A tree level of select is ONE contiguous run of records (one 256-bit load per child), the backup is two
  // fire-and-forget reductions (RED.ADD) per path node.
  // Slot layout: node 0 is ALWAYS the root of the current move, nodes [1, 1 + K_root) are ALWAYS its children (re-rooting
  // moves the chosen child's record to 0 and its child run to 1..), nodes [g0, P) are the general pool (two halves when
  // compaction is on).  The first tree level therefore has an address that is known before the control block arrives.
  uint4* node;
  int g0;          // first node of the general pool: round_up_even(1 + max(max_children, tile width))
  int half_nodes;  // compaction: nodes per half of the general pool (even)
  // per-slot
  uint32_t* ctl;      // [G][NZ_CTL_WORDS]
  uint32_t* path;     // [G][V][max_depth]
  uint32_t* gstate;   // [G][1 + V][state_words]: root state, leaf state(s)
  uint32_t* pend;     // [G][V][2]: (leaf node, path length) of the pending leaves (V > 1 only)
  const double2* ctable; // [ctable_len] (c(N), sqrt(N)) computed by the host libm
  const double* gamma_tape;
  const double* unif_tape;
  uint32_t* arena;
  uint32_t* arena_top;  // [0] = words used, [1] = dropped records, [2] = records written
  uint32_t* rec_index;  // [rec_index_len]: arena offset of every record, in the order arena_top[2] counted them
  int rec_index_len;
  const void* gstatic;  // game-specific static tables (SCS scenario), device
  // per-node game states (nz_config.node_state_cache): row g*P + node holds the compact state AT that node, written when the
  // node is expanded; null = the descent replays the game from the root
  uint32_t* nstate;
  int nstate_words;  // row stride in words (state_words rounded up to 16 bytes)
  // dense leaf batch (nz_engine_attach_cache): a leaf that needs the network takes the next free row of the leaf tensor
  // (one atomic per leaf) instead of row g, so the network runs on a prefix of the batch that holds nothing but work
  int dense;
  // Two LANES of dense rows: launch k uses lane k & 1 (its own leaf / policy / value tensors, counters and row map), so the
  // network can evaluate the rows of launch k while launch k + 1 searches on; a game that parked in lane L (bit 31 of its
  // LEAF control word) is skipped by the launches of the other lane and consumes its row two launches later.
  int lane;
  uint32_t* dense_count;  // [lane][4]: [0] rows handed out by the lane's current launch, [1] games parked by it (zeroed before the launch)
  uint32_t dense_target;  // a slot starts no further simulation once the launch has handed out this many rows ...
  uint32_t park_target;   // ... or once this many games wait for the network (dense_count[1]: own row or a shared one)
  int32_t* dense_rows;    // [lane][G]: dense row -> game slot
  // in-kernel inference cache (Explorer.evaluate consults the cache before every inference, Explorer.py:146-155):
  // open-addressing table keyed by the compact leaf state + scenario map; the search kernel reads ready entries and
  // claims entries for the states it sends to the network (nz_cache_insert_dense completes them)
  uint32_t* cache_keys;        // [cap][cache_kw]; null = no cache
  int32_t* cache_meta;         // [cap] 0 empty, 1 key being written, 3 pending (waits for the network in dense row cache_row), 2 ready
  int32_t* cache_row;          // [cap] dense row of a pending entry | lane << 30
  const void* cache_pol;       // [cap][A] policy dtype
  const float* cache_val;      // [cap]
  uint32_t cache_mask;
  int cache_kw;                // state_words + 1
  // published expansions: (action, prior) list of the state of entry p, written by the first game that expands it
  int32_t* cache_exp_meta;     // [cap] 0 nothing yet, K + 1 published; null = off
  uint16_t* cache_exp_act;     // [cap][cache_exp_width]
  double* cache_exp_prior;     // [cap][cache_exp_width]
  int cache_exp_width;
};

// hash of a cache key held as `kw` words (the last one is the scenario map); every lane of a TILE-wide group calls it.
// The per-word mixes are XOR-combined, so the value does not depend on the group width.
template <int TILE>
__device__ __forceinline__ uint32_t cache_hash_tile(const uint32_t* key, int kw, int lane, unsigned mask) {
  uint32_t h = 0u;
  for (int i = lane; i < kw; i += TILE) {
    uint32_t x = key[i] * 0x9E3779B1u + (uint32_t)i * 0x85EBCA77u;
    x ^= x >> 15; x *= 0x2C1B3C6Du; x ^= x >> 12;
    h ^= x;
  }
#pragma unroll
  for (int off = TILE / 2; off > 0; off >>= 1) h ^= __shfl_xor_sync(mask, h, off, TILE);
  h ^= h >> 16; h *= 0x7FEB352Du; h ^= h >> 15; h *= 0x846CA68Bu; h ^= h >> 16;
  return h;
}

// A tile of TILE consecutive lanes owns one game (TILE = 32: the classic warp per game; small
// games pack several games into one warp so that no lane idles on a 9-action board).
template <int TILE>
struct Tl {
  static_assert(TILE == 1 || TILE == 2 || TILE == 4 || TILE == 8 || TILE == 16 || TILE == 32, "tile width");
  int tl;         // lane within the tile
  int shift;      // first lane of the tile inside the warp
  unsigned mask;  // participation mask of the tile
  __device__ __forceinline__ Tl() {
    int lane;
    // read the lane id once through a volatile asm: the compiler otherwise re-materialises the slow
    // special-register read (S2R) at every use (8 % of the stall samples in the first profile)
    asm volatile("mov.u32 %0, %%laneid;" : "=r"(lane));
    tl = lane & (TILE - 1);
    shift = lane - tl;
    mask = TILE == 32 ? 0xffffffffu : (((1u << (TILE & 31)) - 1u) << shift);
  }
  // TILE == 1 (one thread per game): every collective is the identity
  template <class T>
  __device__ __forceinline__ T bcast(T v, int src) const { if (TILE == 1) return v; return __shfl_sync(mask, v, src, TILE); }
  __device__ __forceinline__ unsigned ballot(bool p) const {
    if (TILE == 1) return p ? 1u : 0u;
    unsigned b = __ballot_sync(mask, p);
    return TILE == 32 ? b : ((b >> shift) & ((1u << (TILE & 31)) - 1u));
  }
  __device__ __forceinline__ void sync() const { if (TILE != 1) __syncwarp(mask); }
  __device__ __forceinline__ unsigned rmax(unsigned v) const { if (TILE == 1) return v; return __reduce_max_sync(mask, v); }
  __device__ __forceinline__ int imax(int v) const { if (TILE == 1) return v; return __reduce_max_sync(mask, v); }
  __device__ __forceinline__ int isum(int v) const { if (TILE == 1) return v; return __reduce_add_sync(mask, v); }
  __device__ __forceinline__ double sum(double v) const {
#pragma unroll
    for (int off = TILE / 2; off > 0; off >>= 1) v += __shfl_xor_sync(mask, v, off, TILE);
    return v;
  }
  __device__ __forceinline__ float sum(float v) const {
#pragma unroll
    for (int off = TILE / 2; off > 0; off >>= 1) v += __shfl_xor_sync(mask, v, off, TILE);
    return v;
  }
  __device__ __forceinline__ float fmax(float v) const {
#pragma unroll
    for (int off = TILE / 2; off > 0; off >>= 1) v = fmaxf(v, __shfl_xor_sync(mask, v, off, TILE));
    return v;
  }
};

// order-preserving map double -> u64 (greater double <=> greater key); -0.0 must be canonicalised
// by the caller (x + 0.0) because IEEE compares it equal to +0.0
__device__ __forceinline__ unsigned long long order_key(double x) {
  long long b = __double_as_longlong(x);
  return (unsigned long long)(b ^ ((b >> 63) | (long long)0x8000000000000000ull));
}

// ---- Philox4x32-10 counter-based generator ------------------------------------------------------
struct Philox {
  __device__ static inline void round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
    uint32_t hi0 = __umulhi(M0, c[0]), lo0 = M0 * c[0];
    uint32_t hi1 = __umulhi(M1, c[2]), lo1 = M1 * c[2];
    uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
  }
  __device__ static inline void gen(unsigned long long key, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                    uint32_t (&out)[4]) {
    uint32_t k0 = (uint32_t)key, k1 = (uint32_t)(key >> 32);
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      round(out, k0, k1);
      k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
  }
};

__device__ __forceinline__ double u01(uint32_t hi, uint32_t lo) {
  // 53-bit uniform in (0,1)
  unsigned long long x = (((unsigned long long)hi << 32) | lo) >> 11;
  return ((double)x + 0.5) * (1.0 / 9007199254740992.0);
}

// Gamma(alpha, scale) by Marsaglia–Tsang (alpha<1 boosted through alpha+1), one stream per call.
__device__ __noinline__ double philox_gamma(unsigned long long key, uint32_t c0, uint32_t c1, uint32_t c2,
                                            double alpha, double scale) {
  double a = alpha < 1.0 ? alpha + 1.0 : alpha;
  double d = a - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
  uint32_t r[4];
  double out = 0.0;
  for (uint32_t it = 0; it < 64; ++it) {
    Philox::gen(key, c0, c1, c2, it * 2u, r);
    double u1 = u01(r[0], r[1]), u2 = u01(r[2], r[3]);
    double x = sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);  // Box–Muller
    double v = 1.0 + c * x;
    if (v <= 0.0) continue;
    v = v * v * v;
    Philox::gen(key, c0, c1, c2, it * 2u + 1u, r);
    double u = u01(r[0], r[1]);
    if (log(u) < 0.5 * x * x + d - d * v + d * log(v)) {
      out = d * v;
      if (alpha < 1.0) out *= pow(u01(r[2], r[3]), 1.0 / alpha);
      break;
    }
  }
  return out * scale;
}

__device__ __forceinline__ float load_policy(const void* p, int dtype, size_t i) {
  return dtype == NZ_BF16 ? __bfloat162float(((const __nv_bfloat16*)p)[i]) : ((const float*)p)[i];
}
__device__ __forceinline__ void store_leaf(void* p, int dtype, size_t i, float v) {
  if (dtype == NZ_BF16) ((__nv_bfloat16*)p)[i] = __float2bfloat16_rn(v);
  else ((float*)p)[i] = v;
}

}  // namespace nz
