// C ABI of the sm_100a self-play search engine (see include/nz_engine.h).
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "common.cuh"
#include "game_scs.cuh"
#include "game_ttt.cuh"
#include "hexgemm.cuh"
#include "mcts.cuh"

namespace nz {

static thread_local char g_err[512] = "";
static int fail(const char* fmt, const char* a = "", long b = 0) {
  snprintf(g_err, sizeof(g_err), fmt, a, b);
  return -1;
}
static int cuda_fail(cudaError_t e, const char* what) {
  snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
  return -2;
}

struct Buf { size_t off, bytes; };

}  // namespace nz

struct nz_engine {
  nz_config cfg;
  nz::View view;
  std::map<std::string, nz::Buf> bufs;
  size_t total;
  bool bound;
  nz::ScsHost scs;  // parsed scenario (empty for TTT)
  int A, C, R, CC, planes, state_words;
  size_t adv_smem, commit_smem, reset_smem, env_smem;
};

namespace nz {

static size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

static void add_buf(nz_engine* e, const char* name, size_t bytes) {
  e->bufs[name] = Buf{e->total, bytes};
  e->total = align_up(e->total + bytes);
}

// ---- environment kernels: the Game interface over n compact states -----------------------------
template <class Game>
__global__ void __launch_bounds__(NZ_CTA_THREADS)
env_kernel(const __grid_constant__ View v, int op, uint32_t* states, const int32_t* map_ids, const int32_t* actions, int32_t* iout,
           uint8_t* mask_out, void* enc_out, int dtype, int n) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int TILE = Game::TILE;
  const typename Game::T t;
  const int tile_in_cta = threadIdx.x / TILE;
  const int i = blockIdx.x * (NZ_CTA_THREADS / TILE) + tile_in_cta;
  if (i >= n) return;
  const int nwords = (v.A + 31) >> 5;
  const size_t scr_bytes = Game::scratch_bytes(v);
  const size_t slab = scr_bytes + (((size_t)nwords * 4 + 15) & ~(size_t)15);
  typename Game::Scratch reg;
  typename Game::Scratch& sc = Game::SMEM ? *(typename Game::Scratch*)(smem_raw + tile_in_cta * slab) : reg;
  uint32_t* words = (uint32_t*)(smem_raw + tile_in_cta * slab + scr_bytes);
  const int map = map_ids ? map_ids[i] : 0;
  uint32_t* st = states + (size_t)i * v.state_words;
  if (op == 0) {  // reset
    Game::reset(sc, v, map, t);
    t.sync();
    Game::save(sc, st, v, t);
    return;
  }
  Game::load(sc, st, v, map, t);
  t.sync();
  if (op == 1) {  // step, with the reference's legality check (SCS_Game.py:379-382)
    const int a = actions[i];
    bool ok = !Game::terminal(sc) && a >= 0 && a < v.A;
    if (ok) {
      Game::legal(sc, v, map, words, t);
      ok = (words[a >> 5] >> (a & 31)) & 1u;
    }
    if (ok) {
      Game::step(sc, v, map, a, t);
      t.sync();
      Game::save(sc, st, v, t);
    }
    if (t.tl == 0) iout[i] = ok ? 0 : 1;
  } else if (op == 2) {  // possible_actions
    Game::legal(sc, v, map, words, t);
    for (int a = t.tl; a < v.A; a += TILE) mask_out[(size_t)i * v.A + a] = (words[a >> 5] >> (a & 31)) & 1u;
  } else if (op == 3) {  // generate_network_input
    Game::encode(sc, v, map, enc_out, dtype, (size_t)i, t);
  } else if (op == 4) {  // is_terminal, get_terminal_value, get_current_player, get_length
    if (t.tl == 0) {
      iout[4 * i + 0] = Game::terminal(sc) ? 1 : 0;
      iout[4 * i + 1] = Game::terminal_value(sc);
      iout[4 * i + 2] = Game::to_play(sc);
      iout[4 * i + 3] = Game::length(sc);
    }
  }
}

// ---- replay decode: move records -> training tuples (Training/ReplayBuffer.py:31-36) ------------------
// One tile per record.  The reference stores, per position, the network input of the root state
// (Gamer.py:65-66), the visit-count policy over ALL actions (store_search_statistics: N_child / sum N, 0 for the
// other actions) and later the game's terminal value.  Here the record holds the compact root state and the
// (action, visits) pairs; this kernel re-encodes the planes as float32 and scatters the policy row, both straight
// into the replay buffer's rows `dst_rows[i]`.
template <class Game>
__global__ void __launch_bounds__(NZ_CTA_THREADS)
replay_decode_kernel(const __grid_constant__ View v, const uint32_t* words, const int64_t* offsets, const int64_t* dst_rows,
                     float* states_out, float* policy_out, int n) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int TILE = Game::TILE;
  const typename Game::T t;
  const int tile_in_cta = threadIdx.x / TILE;
  const int i = blockIdx.x * (NZ_CTA_THREADS / TILE) + tile_in_cta;
  if (i >= n) return;
  const int nwords = (v.A + 31) >> 5;
  const size_t scr_bytes = Game::scratch_bytes(v);
  const size_t slab = scr_bytes + (((size_t)nwords * 4 + 15) & ~(size_t)15);
  typename Game::Scratch reg;
  typename Game::Scratch& sc = Game::SMEM ? *(typename Game::Scratch*)(smem_raw + tile_in_cta * slab) : reg;
  const uint32_t* r = words + offsets[i];
  const size_t row = (size_t)dst_rows[i];
  const int K = (int)(r[2] >> 16);
  const int map = (int)(r[7] >> 20);
  Game::load(sc, r + NZ_REC_HDR, v, map, t);
  t.sync();
  Game::encode(sc, v, map, states_out, NZ_F32, row, t);
  float* pol = policy_out + row * (size_t)v.A;
  for (int a = t.tl; a < v.A; a += TILE) pol[a] = 0.f;
  const uint32_t* c = r + NZ_REC_HDR + v.state_words;
  long long part = 0;
  for (int k = t.tl; k < K; k += TILE) part += (long long)c[2 * k + 1];
  for (int off = TILE / 2; off > 0; off >>= 1) part += __shfl_xor_sync(t.mask, part, off, TILE);
  t.sync();
  const double total = (double)part;  // sum of the root children's visit counts (SCS_Game.py:1518 / tic_tac_toe.py:178)
  for (int k = t.tl; k < K; k += TILE) pol[c[2 * k]] = (float)__ddiv_rn((double)c[2 * k + 1], total);
}

template <class Game>
static int launch_replay_decode(nz_engine* e, const uint32_t* words, const int64_t* offsets, const int64_t* dst_rows,
                                float* states_out, float* policy_out, int n, cudaStream_t st) {
  if (n <= 0) return 0;
  constexpr int per = NZ_CTA_THREADS / Game::TILE;
  const int blocks = (n + per - 1) / per;
  replay_decode_kernel<Game><<<blocks, NZ_CTA_THREADS, e->env_smem, st>>>(e->view, words, offsets, dst_rows, states_out,
                                                                                 policy_out, n);
  cudaError_t err = cudaGetLastError();
  return err == cudaSuccess ? 0 : cuda_fail(err, "nz_replay_decode launch");
}

// ---- device inference cache (Utils/Caches/DictCache.py: state -> (policy, value), exact keys) ---------------------
// Open-addressing table in HBM keyed by the compact leaf state (+ scenario map).  nz_cache_lookup serves the rows whose
// leaf state was evaluated before (copies the stored network output into the engine's policy / value rows) and lists the
// others; the host runs the network on the listed rows only and nz_cache_insert stores their outputs.  Keys are compared
// word for word, so a hit returns exactly what the network returned for that state (the reference's KeylessCache can
// return another state's output on an id collision, KeylessCache.py:61-72 — this one cannot).
struct CacheView {
  uint32_t* keys;     // [cap][kw]
  int32_t* meta;      // [cap] 0 empty, 1 being written, 2 ready
  void* pol;          // [cap][A] policy_dtype
  float* val;         // [cap]
  uint32_t mask;      // cap - 1
  int kw;             // state_words + 1 (scenario map)
  int32_t* row;       // [cap] dense row of a pending entry (in-kernel form), or null
};

__device__ __forceinline__ uint32_t cache_hash(const uint32_t* key, int kw, int lane) {
  return cache_hash_tile<32>(key, kw, lane, 0xffffffffu);  // the same function the search kernel probes with
}

// one warp per leaf row; rows whose game is not waiting for the network are skipped
__global__ void __launch_bounds__(128) cache_lookup_kernel(const __grid_constant__ View v, CacheView c, void* policy, float* value,
                                                           int policy_dtype, int32_t* miss_rows, int32_t* counters,
                                                           const unsigned char* leaf, unsigned char* leaf_stage, int row_bytes) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 4 + warp;
  if (row >= v.G * v.V) return;
  const int g = row / v.V, j = row - g * v.V;
  const uint32_t* ctl = v.ctl + (size_t)g * NZ_CTL_WORDS;
  if (ctl[NZ_CTL_PHASE] != NZ_PHASE_LEAF_PENDING) return;
  if (v.V > 1 && j >= (int)ctl[NZ_CTL_N_PENDING]) return;
  extern __shared__ uint32_t key_smem[];
  uint32_t* key = key_smem + warp * c.kw;
  const uint32_t* st = v.gstate + ((size_t)g * (1 + v.V) + 1 + j) * v.state_words;
  for (int i = lane; i < v.state_words; i += 32) key[i] = st[i];
  if (lane == 0) key[v.state_words] = ctl[NZ_CTL_MAP];
  __syncwarp();
  uint32_t p = cache_hash(key, c.kw, lane) & c.mask;
  bool hit = false;
  for (int probe = 0; probe < 64; ++probe) {
    const int m = c.meta[p];
    if (m == 0) break;
    if (m == 2) {
      bool same = true;
      for (int i = lane; i < c.kw; i += 32) same &= c.keys[(size_t)p * c.kw + i] == key[i];
      if (__all_sync(0xffffffffu, same)) { hit = true; break; }
    }
    p = (p + 1) & c.mask;
  }
  if (hit) {
    const size_t A = (size_t)v.A;
    if (policy_dtype == NZ_BF16) {
      const __nv_bfloat16* src = (const __nv_bfloat16*)c.pol + (size_t)p * A;
      __nv_bfloat16* dst = (__nv_bfloat16*)policy + (size_t)row * A;
      for (size_t a = lane; a < A; a += 32) dst[a] = src[a];
    } else {
      const float* src = (const float*)c.pol + (size_t)p * A;
      float* dst = (float*)policy + (size_t)row * A;
      for (size_t a = lane; a < A; a += 32) dst[a] = src[a];
    }
    if (lane == 0) { value[row] = c.val[p]; atomicAdd(counters + 1, 1); }
  } else {
    int idx = 0;
    if (lane == 0) {
      idx = atomicAdd(counters, 1);
      miss_rows[idx] = row;
    }
    idx = __shfl_sync(0xffffffffu, idx, 0);
    if (leaf_stage) {  // the missed rows are handed to the network as one dense batch: row idx of the staging tensor
      const unsigned char* src = leaf + (size_t)row * row_bytes;
      unsigned char* dst = leaf_stage + (size_t)idx * row_bytes;
      if ((row_bytes & 3) == 0) {
        for (int i = lane; i < (row_bytes >> 2); i += 32) ((uint32_t*)dst)[i] = ((const uint32_t*)src)[i];
      } else {
        for (int i = lane; i < (row_bytes >> 1); i += 32) ((uint16_t*)dst)[i] = ((const uint16_t*)src)[i];
      }
    }
  }
}

// one warp per evaluated row: store (key, policy, value); an equal key that is already there (a duplicate in the batch, or
// a racing warp) is left alone, a full neighbourhood drops the entry
__global__ void __launch_bounds__(128) cache_insert_kernel(const __grid_constant__ View v, CacheView c, const void* policy_in,
                                                           const float* value_in, int policy_dtype, const int32_t* rows, int n,
                                                           void* policy_out, float* value_out, int dense) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i0 = blockIdx.x * 4 + warp;
  if (i0 >= n) return;
  // dense (= lane + 1): the network's rows ARE the engine's rows (row i0 belongs to game slot rows[i0]; nothing to scatter)
  const int row = dense ? rows[i0] * v.V : rows[i0];
  const int g = row / v.V, j = row - g * v.V;
  // policy_out != null: the network's outputs sit in the dense batch (row i0): scatter them to the engine's row first
  const void* policy = policy_in;
  const float* value = value_in;
  size_t prow = (size_t)row;
  if (policy_out) {
    const size_t A = (size_t)v.A;
    if (policy_dtype == NZ_BF16) {
      const __nv_bfloat16* src = (const __nv_bfloat16*)policy_in + (size_t)i0 * A;
      __nv_bfloat16* dst = (__nv_bfloat16*)policy_out + (size_t)row * A;
      for (size_t a = lane; a < A; a += 32) dst[a] = src[a];
    } else {
      const float* src = (const float*)policy_in + (size_t)i0 * A;
      float* dst = (float*)policy_out + (size_t)row * A;
      for (size_t a = lane; a < A; a += 32) dst[a] = src[a];
    }
    if (lane == 0) value_out[row] = value_in[i0];
    prow = (size_t)i0;
  }
  if (dense) prow = (size_t)i0;
  extern __shared__ uint32_t key_smem[];
  uint32_t* key = key_smem + warp * c.kw;
  const uint32_t* st = v.gstate + ((size_t)g * (1 + v.V) + 1 + j) * v.state_words;
  for (int i = lane; i < v.state_words; i += 32) key[i] = st[i];
  if (lane == 0) key[v.state_words] = v.ctl[(size_t)g * NZ_CTL_WORDS + NZ_CTL_MAP];
  __syncwarp();
  uint32_t p = cache_hash(key, c.kw, lane) & c.mask;
  auto key_equal = [&](uint32_t q) {
    bool same = true;
    for (int i = lane; i < c.kw; i += 32) same &= c.keys[(size_t)q * c.kw + i] == key[i];
    return __all_sync(0xffffffffu, same) != 0;
  };
  auto fill = [&](uint32_t q) {  // (key,) policy row and value, then READY
    for (int i = lane; i < c.kw; i += 32) c.keys[(size_t)q * c.kw + i] = key[i];
    const size_t A = (size_t)v.A;
    if (policy_dtype == NZ_BF16) {
      const __nv_bfloat16* src = (const __nv_bfloat16*)policy + prow * A;
      __nv_bfloat16* dst = (__nv_bfloat16*)c.pol + (size_t)q * A;
      for (size_t a = lane; a < A; a += 32) dst[a] = src[a];
    } else {
      const float* src = (const float*)policy + prow * A;
      float* dst = (float*)c.pol + (size_t)q * A;
      for (size_t a = lane; a < A; a += 32) dst[a] = src[a];
    }
    if (lane == 0) c.val[q] = value[prow];
    __threadfence();
    __syncwarp();
    if (lane == 0) atomicExch(c.meta + q, 2);
  };
  // Dense form: the search kernel has usually claimed the entry of this row already (PENDING, cache_row = row | lane << 30): it
  // is completed only if its KEY is this row's key too — an entry never changes its key, which the published expansion lists
  // rely on.  An equal key that is already READY (two games claimed the same state at the same moment) does not end the walk:
  // this row's own pending entry further along the probe sequence is completed as well, so no entry stays pending for ever.
  const int32_t tag = (int32_t)((uint32_t)i0 | ((uint32_t)(dense > 0 ? dense - 1 : 0) << 30));
  bool found_ready = false;
  for (int probe = 0; probe < 64; ++probe) {
    int m = 0;
    if (lane == 0) {
      m = *(volatile int32_t*)(c.meta + p);
      if (m == 0 && !found_ready) m = atomicCAS(c.meta + p, 0, 1) == 0 ? -1 : *(volatile int32_t*)(c.meta + p);
    }
    m = __shfl_sync(0xffffffffu, m, 0);
    if (m == -1) {  // an empty slot, now ours
      fill(p);
      return;
    }
    if (m == 0) return;  // end of the chain and the state is stored already
    if (m == 3 && dense && c.row && c.row[p] == tag && key_equal(p)) {
      fill(p);
      return;
    }
    if (m == 2 && key_equal(p)) {
      if (!dense) return;
      found_ready = true;
    }
    p = (p + 1) & c.mask;
  }
}

// ---- deterministic dyadic stub network (parity protocol) -------------------------
__global__ void stubnet_kernel(const void* leaf, int leaf_dtype, const int32_t* salt, const uint32_t* uid,
                               int uid_stride, int salt_uid_mul, int n, int F, int A, void* policy_out, int policy_dtype,
                               float* value_out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = blockIdx.x * (blockDim.x >> 5) + warp;
  if (i >= n) return;
  const long long P = 65521;
  long long a1 = 0, a2 = 0;
  for (int f = lane; f < F; f += 32) {
    float x = leaf_dtype == NZ_BF16 ? __bfloat162float(((const __nv_bfloat16*)leaf)[(size_t)i * F + f])
                                    : ((const float*)leaf)[(size_t)i * F + f];
    long long q = (long long)rintf(x * 64.0f);
    a1 += q * (long long)((f * 37 + 11) % 251 + 1);
    a2 += q * (long long)((f * 101 + 7) % 241 + 1);
  }
  for (int off = 16; off > 0; off >>= 1) {
    a1 += __shfl_xor_sync(0xffffffffu, a1, off);
    a2 += __shfl_xor_sync(0xffffffffu, a2, off);
  }
  long long sl = (salt ? (long long)salt[i] : 0) + (uid ? (long long)uid[(size_t)i * uid_stride] * salt_uid_mul : 0);
  long long s1 = (a1 + sl) % P, s2 = (a2 + 3 * sl) % P;
  if (s1 < 0) s1 += P;
  if (s2 < 0) s2 += P;
  for (int a = lane; a < A; a += 32) {
    long long m1 = ((long long)a * 40503 + 12345) % P, m2 = ((long long)a * 30011 + 54321) % P,
              m3 = ((long long)a * 977 + 101) % P;
    int h = (int)(((s1 * m1 + s2 * m2 + m3) % P) % 255) + 1;
    float p = (float)h * (1.0f / 256.0f);
    if (policy_dtype == NZ_BF16) ((__nv_bfloat16*)policy_out)[(size_t)i * A + a] = __float2bfloat16_rn(p);
    else ((float*)policy_out)[(size_t)i * A + a] = p;
  }
  if (lane == 0) {
    int k = (int)(((s1 * 7 + s2 * 13 + 5) % P) % 255) - 127;
    value_out[i] = (float)k * (1.0f / 128.0f);
  }
}

// ---- neighbour-table im2col for the network's hexagonal / orthogonal convolutions -----------------
// x: [B, RC, C] bf16 (cells-major, channels last), nbr: [RC, K] source cell of each tap (-1 = off the
// board), out: [B, RC, K, C].  One thread moves 16 bytes; optional ReLU on the way (fuses the
// activation that precedes the convolution).  The GEMM that follows is a plain library GEMM.
__global__ void __launch_bounds__(256) im2col_kernel(const uint4* __restrict__ x, const int32_t* __restrict__ nbr,
                                                     uint4* __restrict__ out, int B, int RC, int K, int C8, int relu) {
  const size_t total = (size_t)B * RC * K * C8;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % C8);
    size_t r = i / C8;
    const int k = (int)(r % K);
    r /= K;
    const int cell = (int)(r % RC);
    const size_t b = r / RC;
    const int src = nbr[cell * K + k];
    uint4 val = make_uint4(0u, 0u, 0u, 0u);
    if (src >= 0) {
      val = x[(b * RC + src) * C8 + c8];
      if (relu) {
        __nv_bfloat162* h = (__nv_bfloat162*)&val;
        const __nv_bfloat162 z = __float2bfloat162_rn(0.f);
#pragma unroll
        for (int j = 0; j < 4; ++j) h[j] = __hmax2(h[j], z);
      }
    }
    out[i] = val;
  }
}

// small feature counts (Tic-Tac-Toe: 18): one THREAD per row, no shuffles — 512 warps for 16384 rows.  The feature weights
// and action multipliers (pure functions of the index) are tabulated once per CTA in shared memory, and the per-action
// hash is reduced term by term so that everything stays in 32-bit arithmetic (s, m < 65521: s * m < 2^32); the values are
// those of the 64-bit expressions of the oracle (oracle/stubnet_np.py) bit for bit.
__global__ void __launch_bounds__(128) stubnet_small_kernel(const void* leaf, int leaf_dtype, const int32_t* salt,
                                                            const uint32_t* uid, int uid_stride, int salt_uid_mul, int n,
                                                            int F, int A, void* policy_out, int policy_dtype,
                                                            float* value_out) {
  __shared__ uint32_t w1[32], w2[32], m1[32], m2[32], m3[32];
  const uint32_t P = 65521u;
  if (threadIdx.x < 32) {
    const uint32_t k = threadIdx.x;
    w1[k] = (k * 37u + 11u) % 251u + 1u;
    w2[k] = (k * 101u + 7u) % 241u + 1u;
    m1[k] = (k * 40503u + 12345u) % P;
    m2[k] = (k * 30011u + 54321u) % P;
    m3[k] = (k * 977u + 101u) % P;
  }
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  long long a1 = 0, a2 = 0;  // |q| <= 64 * |x|: sums of a few thousand at most, but x is unconstrained -> keep 64 bits
  for (int f = 0; f < F; ++f) {
    const float x = leaf_dtype == NZ_BF16 ? __bfloat162float(((const __nv_bfloat16*)leaf)[(size_t)i * F + f])
                                          : ((const float*)leaf)[(size_t)i * F + f];
    const long long q = (long long)rintf(x * 64.0f);
    a1 += q * (long long)w1[f];
    a2 += q * (long long)w2[f];
  }
  const long long sl = (salt ? (long long)salt[i] : 0) + (uid ? (long long)uid[(size_t)i * uid_stride] * salt_uid_mul : 0);
  long long s1l = (a1 + sl) % (long long)P, s2l = (a2 + 3 * sl) % (long long)P;
  if (s1l < 0) s1l += P;
  if (s2l < 0) s2l += P;
  const uint32_t s1 = (uint32_t)s1l, s2 = (uint32_t)s2l;
  for (int a = 0; a < A; ++a) {
    const uint32_t h = ((s1 * m1[a]) % P + (s2 * m2[a]) % P + m3[a]) % P;  // == (s1*m1 + s2*m2 + m3) % P
    const float p = (float)((int)(h % 255u) + 1) * (1.0f / 256.0f);
    if (policy_dtype == NZ_BF16) ((__nv_bfloat16*)policy_out)[(size_t)i * A + a] = __float2bfloat16_rn(p);
    else ((float*)policy_out)[(size_t)i * A + a] = p;
  }
  value_out[i] = (float)((int)(((s1 * 7u + s2 * 13u + 5u) % P) % 255u) - 127) * (1.0f / 128.0f);
}

// draws of the device noise generator, for statistical tests (the throughput-mode root noise is not
// parity-checked against numpy's MT19937 stream, only against the Gamma distribution's moments)
__global__ void gamma_probe_kernel(double* out, int n, double alpha, double scale, unsigned long long seed) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = philox_gamma(seed ^ ((unsigned long long)(i >> 6) * 0x9E3779B97F4A7C15ull), (uint32_t)(i & 63), 1u, 0u, alpha, scale);
}

template <class Game>
static int launch_advance(nz_engine* e, void* leaf, const void* pol, const float* val, cudaStream_t st) {
  constexpr int per = NZ_CTA_THREADS / Game::TILE;
  const int blocks = (e->cfg.n_games + per - 1) / per;
  if (e->view.V > 1)
    advance_vl_kernel<Game><<<blocks, NZ_CTA_THREADS, e->adv_smem, st>>>(e->view, leaf, pol, val, e->cfg.leaf_dtype,
                                                                                 e->cfg.policy_dtype);
  else if (e->view.dense)
    advance_kernel<Game, true><<<blocks, NZ_CTA_THREADS, e->adv_smem, st>>>(e->view, leaf, pol, val, e->cfg.leaf_dtype,
                                                                                    e->cfg.policy_dtype);
  else
    advance_kernel<Game, false><<<blocks, NZ_CTA_THREADS, e->adv_smem, st>>>(e->view, leaf, pol, val, e->cfg.leaf_dtype,
                                                                                     e->cfg.policy_dtype);
  cudaError_t err = cudaGetLastError();
  return err == cudaSuccess ? 0 : cuda_fail(err, "nz_advance launch");
}

template <class Game>
static int launch_commit(nz_engine* e, const int32_t* actions, cudaStream_t st) {
  constexpr int per = NZ_CTA_THREADS / Game::TILE;
  const int blocks = (e->cfg.n_games + per - 1) / per;
  commit_kernel<Game><<<blocks, NZ_CTA_THREADS, e->commit_smem, st>>>(e->view, actions);
  cudaError_t err = cudaGetLastError();
  return err == cudaSuccess ? 0 : cuda_fail(err, "nz_commit_moves launch");
}

template <class Game>
static int launch_reset(nz_engine* e, cudaStream_t st) {
  constexpr int per = NZ_CTA_THREADS / Game::TILE;
  const int blocks = (e->cfg.n_games + per - 1) / per;
  reset_kernel<Game><<<blocks, NZ_CTA_THREADS, e->reset_smem, st>>>(e->view);
  cudaError_t err = cudaGetLastError();
  return err == cudaSuccess ? 0 : cuda_fail(err, "nz_reset launch");
}

template <class Game>
static int launch_env(nz_engine* e, int op, uint32_t* states, const int32_t* map_ids, const int32_t* actions,
                      int32_t* iout, uint8_t* mask_out, void* enc_out, int dtype, int n, cudaStream_t st) {
  if (n <= 0) return 0;
  constexpr int per = NZ_CTA_THREADS / Game::TILE;
  const int blocks = (n + per - 1) / per;
  env_kernel<Game><<<blocks, NZ_CTA_THREADS, e->env_smem, st>>>(e->view, op, states, map_ids, actions, iout,
                                                                        mask_out, enc_out, dtype, n);
  cudaError_t err = cudaGetLastError();
  return err == cudaSuccess ? 0 : cuda_fail(err, "nz_env launch");
}

template <class Game>
static int setup_smem(nz_engine* e) {
  constexpr size_t per = NZ_CTA_THREADS / Game::TILE;
  const size_t scr = Game::scratch_bytes(e->view);
  const int nwords = (e->A + 31) >> 5;
  e->view.slab_bytes = (int)tile_slab_bytes<Game>(e->view);
  e->view.slab_words_bytes = (int)tile_slab_words_bytes<Game>(e->view);
  e->adv_smem = per * tile_slab_bytes<Game>(e->view);
  e->commit_smem = per * (scr + (((size_t)e->state_words * 4 + 15) & ~(size_t)15));
  e->reset_smem = per * scr;
  e->env_smem = per * (scr + (((size_t)nwords * 4 + 15) & ~(size_t)15));
  if (e->adv_smem > 200 * 1024) return fail("per-CTA shared memory too large (%s%ld bytes)", "", (long)e->adv_smem);
  cudaError_t err = cudaSuccess;
  if (e->adv_smem > 48 * 1024)
    err = cudaFuncSetAttribute(advance_kernel<Game, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->adv_smem);
  if (err == cudaSuccess && e->adv_smem > 48 * 1024)
    err = cudaFuncSetAttribute(advance_kernel<Game, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->adv_smem);
  if (err == cudaSuccess && e->adv_smem > 48 * 1024)
    err = cudaFuncSetAttribute(advance_vl_kernel<Game>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->adv_smem);
  if (err == cudaSuccess && e->env_smem > 48 * 1024)
    err = cudaFuncSetAttribute(env_kernel<Game>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->env_smem);
  if (err == cudaSuccess && e->env_smem > 48 * 1024)
    err = cudaFuncSetAttribute(replay_decode_kernel<Game>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->env_smem);
  // no CUDA device in the build container: attribute calls fail there, which is fine for a
  // create/layout-only use; launches report their own errors.
  (void)err;
  cudaGetLastError();
  return 0;
}

}  // namespace nz

#define NZ_GAME_SWITCH(e, FN, ...) \
  ((e)->cfg.game_kind == NZ_GAME_TTT ? FN<nz::TTT>(__VA_ARGS__) : FN<nz::SCS>(__VA_ARGS__))

template <bool TAPS_INNER, bool PAIR, int AHEAD, int HALVES = 2, int GROUPS = 1, bool TMA_A = false>
static cudaError_t nz_hexconv_launch(const CUtensorMap& tm_w, const CUtensorMap& tm_x, const nzg::Params& p, int tiles, int nsplit,
                                     cudaStream_t stream) {
  auto kern = nzg::hexconv_kernel<TAPS_INNER, PAIR, AHEAD, HALVES, GROUPS, TMA_A>;
  static bool attr_done = false;  // one flag per instantiation
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, nzg::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    attr_done = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(PAIR ? (unsigned)((tiles + 1) / 2 * 2) : (unsigned)tiles, (unsigned)nsplit);
  cfg.blockDim = dim3(nzg::THREADS);
  cfg.dynamicSmemBytes = nzg::SMEM_BYTES;
  cfg.stream = stream;
  static int pdl = -1;  // NZ_HEXCONV_PDL=0 switches programmatic dependent launch off
  if (pdl < 0) {
    const char* e = getenv("NZ_HEXCONV_PDL");
    pdl = (e && e[0] == '0') ? 0 : 1;
  }
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = PAIR ? 2 : 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = pdl;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  return cudaLaunchKernelEx(&cfg, kern, tm_w, tm_x, p);
}

extern "C" {

const char* nz_last_error(void) { return nz::g_err; }
int nz_abi_version(void) { return NZ_ABI_VERSION; }
size_t nz_config_bytes(void) { return sizeof(nz_config); }

int nz_engine_create(const nz_config* cfg, nz_engine** out) {
  using namespace nz;
  if (!cfg || !out) return fail("null argument");
  if (cfg->abi_version != NZ_ABI_VERSION) return fail("ABI version mismatch");
  if (cfg->n_games <= 0 || cfg->pool_nodes < 2 || cfg->max_depth < 2 || cfg->mcts_simulations <= 0)
    return fail("bad sizes in nz_config");
  if (cfg->max_depth > 4096) return fail("max_depth too large");
  if (cfg->ctable_len <= 0) return fail("ctable_len must be > 0");
  if (cfg->max_sims_per_launch <= 0) return fail("max_sims_per_launch must be > 0");
  if (cfg->virtual_loss_width < 0 || cfg->virtual_loss_width > 8) return fail("virtual_loss_width must be 0..8");
  if (cfg->virtual_loss_width > 1 && cfg->max_levels_per_launch > 0)
    return fail("virtual_loss_width > 1 cannot be combined with max_levels_per_launch");
  nz_engine* e = new nz_engine();
  e->cfg = *cfg;
  e->cfg.scs_desc = nullptr;
  e->bound = false;
  e->total = 0;
  bool prior64;
  if (cfg->game_kind == NZ_GAME_TTT) {
    e->A = TTT::A; e->C = TTT::C; e->R = TTT::R; e->CC = TTT::CC; e->planes = TTT::PLANES;
    e->state_words = TTT::STATE_WORDS;
    prior64 = true;
  } else if (cfg->game_kind == NZ_GAME_SCS) {
    if (!cfg->scs_desc || cfg->scs_desc_len <= 0) { delete e; return fail("SCS needs scs_desc"); }
    if (scs_parse(cfg->scs_desc, cfg->scs_desc_len, e->scs, g_err, sizeof(g_err)) != 0) { delete e; return -1; }
    e->A = e->scs.A; e->C = e->scs.C; e->R = e->scs.R; e->CC = e->scs.CC; e->planes = e->scs.planes;
    e->state_words = e->scs.state_words;
    prior64 = false;
  } else {
    delete e;
    return fail("unknown game_kind");
  }
  if (e->A > 65535) { delete e; return fail("action space too large for 16-bit action ids"); }
  if (cfg->max_children <= 0 || cfg->max_children > 65535) { delete e; return fail("bad max_children"); }
  if (cfg->n_games > (1 << 20)) { delete e; return fail("at most 2^20 game slots per engine (move records keep the slot in 20 bits)"); }
  const size_t G = cfg->n_games, P = cfg->pool_nodes, V = cfg->virtual_loss_width > 1 ? cfg->virtual_loss_width : 1;
  (void)prior64;
  // node 0 = root, nodes 1 .. reserved = the root's children (the search kernel loads the first 32 of them before it
  // knows anything about the slot), general pool from g0 (even: child runs start on 64-byte boundaries)
  const int reserved = cfg->max_children > 32 ? cfg->max_children : 32;
  const int g0 = (1 + reserved + 1) & ~1;
  const bool compact = cfg->compact_on_reroot && cfg->auto_advance;
  const int half_nodes = compact ? (((int)P - g0) / 2) & ~1 : 0;
  if ((int)P < g0 + 4 || (compact && half_nodes < 2)) {
    delete e;
    return fail("pool_nodes too small: need at least %s%ld nodes per slot", "", (long)(g0 + 4));
  }
  add_buf(e, "nodes", G * P * 32);
  add_buf(e, "ctl", G * NZ_CTL_WORDS * 4);
  add_buf(e, "path", G * V * (size_t)cfg->max_depth * 4);
  add_buf(e, "gstate", G * (1 + V) * (size_t)e->state_words * 4);
  add_buf(e, "pend", G * V * 2 * 4);
  add_buf(e, "ctable", (size_t)cfg->ctable_len * 16);
  add_buf(e, "gamma_tape", cfg->tape_moves > 0 ? G * (size_t)cfg->tape_moves * cfg->tape_width * 8 : 8);
  add_buf(e, "unif_tape", cfg->tape_moves > 0 ? G * (size_t)cfg->tape_moves * 3 * 8 : 8);
  add_buf(e, "arena", (size_t)(cfg->arena_words > 0 ? cfg->arena_words : 1) * 4);
  add_buf(e, "arena_top", 16);
  add_buf(e, "rec_index", ((size_t)(cfg->arena_words > 0 ? cfg->arena_words : 1) / (NZ_REC_HDR + 3) + 1) * 4);
  add_buf(e, "scs_static", cfg->game_kind == NZ_GAME_SCS ? e->scs.static_bytes() : 8);
  const bool nstate = cfg->node_state_cache && cfg->game_kind == NZ_GAME_SCS;
  if (nstate && (P & 1)) { delete e; return fail("node_state_cache needs an even pool_nodes"); }
  const size_t nstate_words = ((size_t)e->state_words + 3) & ~(size_t)3;
  add_buf(e, "nstate", nstate ? (G * P / 2 + 1) * nstate_words * 4 : 16);  // one row per possible child run (runs start on even nodes)
  add_buf(e, "dense_count", 32);
  add_buf(e, "dense_rows", 2 * G * V * 4);

  View& v = e->view;
  memset(&v, 0, sizeof(v));
  v.G = cfg->n_games; v.P = cfg->pool_nodes; v.max_depth = cfg->max_depth; v.max_children = cfg->max_children;
  v.sims = cfg->mcts_simulations; v.training = cfg->training; v.policy_is_prob = cfg->policy_is_prob;
  v.auto_advance = cfg->auto_advance; v.games_per_slot = cfg->games_per_slot;
  v.max_sims_per_launch = cfg->max_sims_per_launch; v.record_detail = cfg->record_detail;
  v.V = (int)V;
  v.n_softmax_moves = cfg->number_of_softmax_moves;
  v.compact = compact ? 1 : 0;
  v.g0 = g0; v.half_nodes = half_nodes;
  v.max_levels = cfg->max_levels_per_launch > 0 ? cfg->max_levels_per_launch : 0x7fffffff;
  v.ctable_len = cfg->ctable_len; v.tape_moves = cfg->tape_moves; v.tape_width = cfg->tape_width;
  v.arena_words = cfg->arena_words;
  v.A = e->A; v.leaf_elems = e->C * e->R * e->CC; v.state_words = e->state_words;
  v.scratch_extra = cfg->game_kind == NZ_GAME_SCS ? (int)e->scs.occ_bytes() : 0;
  v.pb_c_base = cfg->pb_c_base; v.pb_c_init = cfg->pb_c_init; v.value_factor = cfg->value_factor;
  v.noise_frac = cfg->root_exploration_fraction; v.noise_alpha = cfg->root_dist_alpha;
  v.noise_beta = cfg->root_dist_beta; v.eps_softmax = cfg->epsilon_softmax_exploration;
  v.eps_random = cfg->epsilon_random_exploration; v.seed = cfg->seed;
  int rc = NZ_GAME_SWITCH(e, nz::setup_smem, e);
  if (rc != 0) { delete e; return rc; }
  *out = e;
  g_err[0] = 0;
  return 0;
}

void nz_engine_destroy(nz_engine* eng) { delete eng; }

size_t nz_engine_workspace_bytes(const nz_engine* eng) { return eng ? eng->total : 0; }

int nz_engine_buffer(const nz_engine* eng, const char* name, size_t* offset, size_t* bytes) {
  if (!eng || !name) return nz::fail("null argument");
  auto it = eng->bufs.find(name);
  if (it == eng->bufs.end()) return nz::fail("unknown buffer '%s'", name);
  if (offset) *offset = it->second.off;
  if (bytes) *bytes = it->second.bytes;
  return 0;
}

int nz_engine_bind(nz_engine* eng, void* ws, size_t bytes) {
  using namespace nz;
  if (!eng || !ws) return fail("null argument");
  if (bytes < eng->total) return fail("workspace too small (%s%ld bytes needed)", "", (long)eng->total);
  if (((uintptr_t)ws & 255u) != 0) return fail("workspace must be 256-byte aligned");
  unsigned char* b = (unsigned char*)ws;
  View& v = eng->view;
  auto at = [&](const char* n) { return (void*)(b + eng->bufs[n].off); };
  v.node = (uint4*)at("nodes");
  v.ctl = (uint32_t*)at("ctl");
  v.path = (uint32_t*)at("path");
  v.gstate = (uint32_t*)at("gstate");
  v.pend = (uint32_t*)at("pend");
  v.ctable = (const double2*)at("ctable");
  v.gamma_tape = (const double*)at("gamma_tape");
  v.unif_tape = (const double*)at("unif_tape");
  v.arena = (uint32_t*)at("arena");
  v.arena_top = (uint32_t*)at("arena_top");
  v.rec_index = (uint32_t*)at("rec_index");
  v.rec_index_len = (int)(eng->bufs["rec_index"].bytes / 4);
  v.gstatic = at("scs_static");
  v.nstate = (eng->cfg.node_state_cache && eng->cfg.game_kind == NZ_GAME_SCS) ? (uint32_t*)at("nstate") : nullptr;
  v.nstate_words = (eng->state_words + 3) & ~3;
  v.dense_count = (uint32_t*)at("dense_count");
  v.dense_rows = (int32_t*)at("dense_rows");
  eng->bound = true;
  return 0;
}

int nz_game_shape(const nz_engine* eng, int32_t* out6) {
  if (!eng || !out6) return nz::fail("null argument");
  out6[0] = eng->planes; out6[1] = eng->R; out6[2] = eng->CC;
  out6[3] = eng->C; out6[4] = eng->R; out6[5] = eng->CC;
  return 0;
}

int nz_env_state_words(const nz_engine* eng) { return eng ? eng->state_words : -1; }

/* The SCS scenario tables the kernels read (terrain, schedule, maps) as a byte image the caller
 * uploads into the "scs_static" buffer. */
int nz_scs_static_image(const nz_engine* eng, void* host_out, size_t bytes) {
  if (!eng || !host_out) return nz::fail("null argument");
  if (eng->cfg.game_kind != NZ_GAME_SCS) return nz::fail("not an SCS engine");
  if (bytes < eng->scs.static_bytes()) return nz::fail("buffer too small");
  eng->scs.write_image(host_out);
  return 0;
}

#define NZ_REQUIRE_BOUND(eng) \
  if (!(eng) || !(eng)->bound) return nz::fail("engine not bound to a workspace")

int nz_reset(nz_engine* eng, void* stream) {
  NZ_REQUIRE_BOUND(eng);
  return NZ_GAME_SWITCH(eng, nz::launch_reset, eng, (cudaStream_t)stream);
}

int nz_advance(nz_engine* eng, void* leaf_out, const void* policy_in, const float* value_in, void* stream) {
  NZ_REQUIRE_BOUND(eng);
  if (!leaf_out || !policy_in || !value_in) return nz::fail("null tensor pointer");
  if (eng->view.dense) {
    cudaError_t err = cudaMemsetAsync(eng->view.dense_count + 4 * eng->view.lane, 0, 16, (cudaStream_t)stream);
    if (err != cudaSuccess) return nz::cuda_fail(err, "nz_advance dense counter reset");
  }
  return NZ_GAME_SWITCH(eng, nz::launch_advance, eng, leaf_out, policy_in, value_in, (cudaStream_t)stream);
}

int nz_commit_moves(nz_engine* eng, const int32_t* actions, void* stream) {
  NZ_REQUIRE_BOUND(eng);
  return NZ_GAME_SWITCH(eng, nz::launch_commit, eng, actions, (cudaStream_t)stream);
}

#define NZ_ENV(op, states, maps, acts, iout, mask, enc, dt) \
  NZ_GAME_SWITCH(eng, nz::launch_env, eng, op, (uint32_t*)(states), maps, acts, iout, mask, enc, dt, n, (cudaStream_t)stream)

int nz_env_reset(nz_engine* eng, uint32_t* states, const int32_t* map_ids, int n, void* stream) {
  NZ_REQUIRE_BOUND(eng);
  return NZ_ENV(0, states, map_ids, nullptr, nullptr, nullptr, nullptr, 0);
}
int nz_env_step(nz_engine* eng, uint32_t* states, const int32_t* map_ids, const int32_t* actions, int32_t* err_out,
                int n, void* stream) {
  NZ_REQUIRE_BOUND(eng);
  if (!actions || !err_out) return nz::fail("null argument");
  return NZ_ENV(1, states, map_ids, actions, err_out, nullptr, nullptr, 0);
}
int nz_env_mask(nz_engine* eng, const uint32_t* states, const int32_t* map_ids, uint8_t* mask_out, int n, void* stream) {
  NZ_REQUIRE_BOUND(eng);
  return NZ_ENV(2, states, map_ids, nullptr, nullptr, mask_out, nullptr, 0);
}
int nz_env_encode(nz_engine* eng, const uint32_t* states, const int32_t* map_ids, void* out, int dtype, int n,
                  void* stream) {
  NZ_REQUIRE_BOUND(eng);
  return NZ_ENV(3, states, map_ids, nullptr, nullptr, nullptr, out, dtype);
}
int nz_env_status(nz_engine* eng, const uint32_t* states, const int32_t* map_ids, int32_t* out, int n, void* stream) {
  NZ_REQUIRE_BOUND(eng);
  return NZ_ENV(4, states, map_ids, nullptr, out, nullptr, nullptr, 0);
}

int nz_replay_decode(nz_engine* eng, const uint32_t* words, const int64_t* offsets, const int64_t* dst_rows, float* states_out,
                     float* policy_out, int n, void* stream) {
  NZ_REQUIRE_BOUND(eng);
  if (!words || !offsets || !dst_rows || !states_out || !policy_out) return nz::fail("null argument");
  return NZ_GAME_SWITCH(eng, nz::launch_replay_decode, eng, words, offsets, dst_rows, states_out, policy_out, n, (cudaStream_t)stream);
}

int nz_cache_lookup(nz_engine* eng, uint32_t* keys, int32_t* meta, void* cache_policy, float* cache_value, int capacity_log2,
                    void* policy, float* value, int32_t* miss_rows, int32_t* counters, const void* leaf, void* leaf_stage,
                    void* stream) {
  NZ_REQUIRE_BOUND(eng);
  if (!keys || !meta || !cache_policy || !cache_value || !policy || !value || !miss_rows || !counters) return nz::fail("null argument");
  if (capacity_log2 < 4 || capacity_log2 > 30) return nz::fail("capacity_log2 must be 4..30");
  nz::CacheView c{keys, meta, cache_policy, cache_value, (1u << capacity_log2) - 1u, eng->state_words + 1};
  const int rows = eng->view.G * eng->view.V;
  nz::cache_lookup_kernel<<<(rows + 3) / 4, 128, 4 * c.kw * sizeof(uint32_t), (cudaStream_t)stream>>>(
      eng->view, c, policy, value, eng->cfg.policy_dtype, miss_rows, counters, (const unsigned char*)leaf,
      (unsigned char*)(leaf ? leaf_stage : nullptr), eng->view.leaf_elems * (eng->cfg.leaf_dtype == NZ_BF16 ? 2 : 4));
  cudaError_t err = cudaGetLastError();
  return err == cudaSuccess ? 0 : nz::cuda_fail(err, "nz_cache_lookup launch");
}

int nz_cache_insert(nz_engine* eng, uint32_t* keys, int32_t* meta, void* cache_policy, float* cache_value, int capacity_log2,
                    const void* policy, const float* value, const int32_t* rows, int n, void* policy_out, float* value_out,
                    void* stream) {
  NZ_REQUIRE_BOUND(eng);
  if (!keys || !meta || !cache_policy || !cache_value || !policy || !value || !rows) return nz::fail("null argument");
  if (capacity_log2 < 4 || capacity_log2 > 30) return nz::fail("capacity_log2 must be 4..30");
  if (n <= 0) return 0;
  nz::CacheView c{keys, meta, cache_policy, cache_value, (1u << capacity_log2) - 1u, eng->state_words + 1, nullptr};
  nz::cache_insert_kernel<<<(n + 3) / 4, 128, 4 * c.kw * sizeof(uint32_t), (cudaStream_t)stream>>>(
      eng->view, c, policy, value, eng->cfg.policy_dtype, rows, n, policy_out, policy_out ? value_out : nullptr, 0);
  cudaError_t err = cudaGetLastError();
  return err == cudaSuccess ? 0 : nz::cuda_fail(err, "nz_cache_insert launch");
}

int nz_engine_attach_cache(nz_engine* eng, uint32_t* keys, int32_t* meta, int32_t* cache_row, const void* cache_policy,
                           const float* cache_value, int capacity_log2, int miss_target, int park_target) {
  NZ_REQUIRE_BOUND(eng);
  nz::View& v = eng->view;
  if (!keys) {  // detach: leaves go back to row g of the leaf tensor
    v.cache_keys = nullptr; v.cache_meta = nullptr; v.cache_row = nullptr; v.cache_pol = nullptr; v.cache_val = nullptr;
    v.dense = 0;
    v.lane = 0;
    v.cache_exp_meta = nullptr; v.cache_exp_act = nullptr; v.cache_exp_prior = nullptr; v.cache_exp_width = 0;
    return 0;
  }
  if (!meta || !cache_row || !cache_policy || !cache_value) return nz::fail("null argument");
  if (capacity_log2 < 4 || capacity_log2 > 30) return nz::fail("capacity_log2 must be 4..30");
  if (v.V > 1) return nz::fail("the in-kernel inference cache needs virtual_loss_width <= 1");
  v.cache_keys = keys; v.cache_meta = meta; v.cache_row = cache_row; v.cache_pol = cache_policy; v.cache_val = cache_value;
  v.cache_mask = (1u << capacity_log2) - 1u;
  v.cache_kw = eng->state_words + 1;
  v.dense = 1;
  v.dense_target = miss_target > 0 ? (uint32_t)miss_target : 0xffffffffu;
  v.park_target = park_target > 0 ? (uint32_t)park_target : 0xffffffffu;
  return 0;
}

int nz_engine_attach_expansions(nz_engine* eng, int32_t* exp_meta, uint16_t* exp_actions, double* exp_priors, int width) {
  NZ_REQUIRE_BOUND(eng);
  nz::View& v = eng->view;
  if (!exp_meta) {
    v.cache_exp_meta = nullptr; v.cache_exp_act = nullptr; v.cache_exp_prior = nullptr; v.cache_exp_width = 0;
    return 0;
  }
  if (!exp_actions || !exp_priors || width <= 0) return nz::fail("null argument");
  if (!v.cache_keys) return nz::fail("nz_engine_attach_expansions: attach the cache first (nz_engine_attach_cache)");
  v.cache_exp_meta = exp_meta; v.cache_exp_act = exp_actions; v.cache_exp_prior = exp_priors; v.cache_exp_width = width;
  return 0;
}

int nz_engine_set_lane(nz_engine* eng, int lane) {
  NZ_REQUIRE_BOUND(eng);
  if (lane != 0 && lane != 1) return nz::fail("lane must be 0 or 1");
  eng->view.lane = lane;
  return 0;
}

int nz_cache_insert_dense(nz_engine* eng, uint32_t* keys, int32_t* meta, void* cache_policy, float* cache_value, int capacity_log2,
                          const void* policy, const float* value, int n, int lane, void* stream) {
  NZ_REQUIRE_BOUND(eng);
  if (!keys || !meta || !cache_policy || !cache_value || !policy || !value) return nz::fail("null argument");
  if (capacity_log2 < 4 || capacity_log2 > 30) return nz::fail("capacity_log2 must be 4..30");
  if (!eng->view.dense) return nz::fail("nz_cache_insert_dense: no cache attached (nz_engine_attach_cache)");
  if (lane != 0 && lane != 1) return nz::fail("lane must be 0 or 1");
  if (n <= 0) return 0;
  nz::CacheView c{keys, meta, cache_policy, cache_value, (1u << capacity_log2) - 1u, eng->state_words + 1, eng->view.cache_row};
  nz::cache_insert_kernel<<<(n + 3) / 4, 128, 4 * c.kw * sizeof(uint32_t), (cudaStream_t)stream>>>(
      eng->view, c, policy, value, eng->cfg.policy_dtype, eng->view.dense_rows + (size_t)lane * eng->view.G, n, nullptr, nullptr, 1 + lane);
  cudaError_t err = cudaGetLastError();
  return err == cudaSuccess ? 0 : nz::cuda_fail(err, "nz_cache_insert_dense launch");
}

int nz_im2col_bf16(const void* x, const int32_t* nbr, void* out, int batch, int cells, int taps, int channels, int relu,
                   void* stream) {
  if (!x || !nbr || !out) return nz::fail("null tensor pointer");
  if (channels % 8 != 0) return nz::fail("nz_im2col_bf16: channels must be a multiple of 8");
  const size_t total = (size_t)batch * cells * taps * (channels / 8);
  if (total == 0) return 0;
  size_t blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;  // grid-stride: 16 resident CTAs of 256 threads on each of the 148 SMs
  nz::im2col_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const uint4*)x, nbr, (uint4*)out, batch, cells, taps,
                                                                   channels / 8, relu);
  cudaError_t err = cudaGetLastError();
  return err == cudaSuccess ? 0 : nz::cuda_fail(err, "nz_im2col_bf16 launch");
}

static long long* nz_hexconv_trace = nullptr;  // set through nz_hexconv_set_trace (profiling aid)

typedef CUresult (*nz_encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                       const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                       CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int nz_make_tmap(CUtensorMap* tm, const void* base, uint64_t cols, uint64_t rows, uint32_t box_cols, uint32_t box_rows) {
  static nz_encode_tiled_fn enc = nullptr;
  if (!enc) {
    cudaDriverEntryPointQueryResult q;
    void* fn = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn)
      return nz::fail("cuTensorMapEncodeTiled is not available from the driver");
    enc = (nz_encode_tiled_fn)fn;
  }
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t es[2] = {1, 1};
  const CUresult rc = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, es,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return rc == CUDA_SUCCESS ? 0 : nz::fail("cuTensorMapEncodeTiled failed (%s%ld)", "", (long)rc);
}

int nz_hexconv_bf16(const void* x, const int32_t* nbr, const void* wt, const void* residual, void* out, int rows, int cells,
                    int taps, int cin, int n_pad, int ldo, int flags, int relu_out, void* stream) {
  if (!x || !nbr || !wt || !out) return nz::fail("null tensor pointer");
  if (cin % 64 != 0 || n_pad % 16 != 0 || n_pad < 16 || n_pad > 256 || ldo % 16 != 0 || taps < 1 || taps > nzg::MAX_TAPS ||
      cells < 1 || rows < 1)
    return nz::fail("nz_hexconv_bf16: need cin % 64 == 0, 16 <= n_pad <= 256 (multiple of 16), ldo % 16 == 0, taps <= 9");
  if (flags & 1) return nz::fail("nz_hexconv_bf16: relu_in is not supported (apply relu_out in the producing layer)");
  nzg::Params p;
  p.x = (const __nv_bfloat16*)x; p.nbr = nbr; p.wt = (const __nv_bfloat16*)wt;
  p.residual = (const __nv_bfloat16*)residual; p.out = (__nv_bfloat16*)out;
  p.rows = rows; p.RC = cells; p.taps = taps; p.cin = cin; p.n_pad = n_pad; p.ldo = ldo; p.relu_in = 0; p.relu_out = relu_out;
  p.trace = nz_hexconv_trace;
  int tiles = (rows + nzg::BLOCK_M - 1) / nzg::BLOCK_M;
  // variants (for comparison and as a fallback): bit 1 = taps innermost in K with L1-allocating gathers, bit 2 = one CTA
  // per tile (tcgen05 cta_group::1) instead of the CTA pair; the pair form needs an even split of the weight rows into
  // 8-row atoms
  static bool pair_ok = true;  // cleared when the device refuses the 2-CTA cluster launch (e.g. a partitioned GPU)
  bool pair = !(flags & 4) && n_pad % 32 == 0 && pair_ok;
  const bool taps_inner = (flags & 2) != 0;
  // small batches: split the output channels over two CTAs (pairs) per row tile when that still fits one wave — more SMs
  // work on the layer and each K loop runs at the MMA rate of a 128-wide tile (a layer's latency is what bounds the forward
  // on the few hundred rows the inference cache leaves over).  bit 4 switches the split off.
  // ... and 128-row tiles with one accumulator (HALVES = 1: smaller stages, six K chunks in flight) while even those fit one
  // wave: a lone tile's K loop is bounded by L2 round trips per chunks in flight, not by the tensor pipe.
  int nsplit = 1, halves = 2;
  if (!(flags & 16) && !taps_inner && !((flags & 8) != 0)) {
    const int tiles128 = (rows + 127) / 128;
    const int units1 = pair ? (tiles128 + 1) / 2 : tiles128, limit = pair ? 74 : 148;
    if (units1 <= limit) {
      halves = 1;
      tiles = tiles128;
      if (n_pad % 64 == 0 && units1 * 2 <= limit) nsplit = 2;
    }
  }
  if (halves == 2 && !(flags & 16)) {
    const int units = pair ? (tiles + 1) / 2 : tiles;
    if (n_pad % 64 == 0 && units * 2 <= (pair ? 74 : 148)) nsplit = 2;
  }
  p.n_pad = n_pad / nsplit;
  CUtensorMap tm_w;
  if (nz_make_tmap(&tm_w, wt, (uint64_t)taps * cin, (uint64_t)n_pad, 64, (uint32_t)(pair ? p.n_pad / 2 : p.n_pad)) != 0) return -1;
  CUtensorMap tm_x;  // x as [rows, cin] with a 64-channel x 1-row box: the operand of the TMA row gather
  if (nz_make_tmap(&tm_x, x, (uint64_t)cin, (uint64_t)rows, 64, 1) != 0) return -1;
  cudaError_t err;
  const bool long_ahead = (flags & 8) != 0;  // bit 3 (pair only): publish a chunk three iterations after its issue, not two
  cudaStream_t st = (cudaStream_t)stream;
  if (halves == 1) {
    // pair form: two producer teams that take the K chunks in turn, two chunks of each in flight (measured 15.9 against
    // 17.6 us per layer on 1600 rows; one team with 2-5 chunks in flight: 17.3-17.7, four teams: 16.3)
    // (the TMA row gather of the large-batch form is no faster here — 17.2 us —: a lone tile's K loop is bound by its chain of
    // 4 x 28 dependent-issue UMMAs at ~200 cycles each, whatever fills the stages)
    err = pair ? nz_hexconv_launch<false, true, 2, 1, 2>(tm_w, tm_x, p, tiles, nsplit, st) : nz_hexconv_launch<false, false, 2, 1>(tm_w, tm_x, p, tiles, nsplit, st);
  } else if (pair) {
    if (long_ahead) err = taps_inner ? nz_hexconv_launch<true, true, 3>(tm_w, tm_x, p, tiles, nsplit, st) : nz_hexconv_launch<false, true, 3>(tm_w, tm_x, p, tiles, nsplit, st);
    // default: the A rows arrive by TMA row gather (tile::gather4), 1127 against 1107 TFLOP/s with cp.async gathers (bit 5)
    else if (!taps_inner && !(flags & 32)) err = nz_hexconv_launch<false, true, 1, 2, 1, true>(tm_w, tm_x, p, tiles, nsplit, st);
    else err = taps_inner ? nz_hexconv_launch<true, true, 2>(tm_w, tm_x, p, tiles, nsplit, st) : nz_hexconv_launch<false, true, 2>(tm_w, tm_x, p, tiles, nsplit, st);
  } else {
    err = taps_inner ? nz_hexconv_launch<true, false, 2>(tm_w, tm_x, p, tiles, nsplit, st) : nz_hexconv_launch<false, false, 2>(tm_w, tm_x, p, tiles, nsplit, st);
  }
  if (err == cudaSuccess) err = cudaGetLastError();
  if (err != cudaSuccess && pair && !(flags & 4)) {
    // still a B200 tcgen05 kernel, only without the pairing: one CTA per tile (cta_group::1)
    cudaGetLastError();
    pair_ok = false;
    return nz_hexconv_bf16(x, nbr, wt, residual, out, rows, cells, taps, cin, n_pad, ldo, flags | 4, relu_out, stream);
  }
  return err == cudaSuccess ? 0 : nz::cuda_fail(err, "nz_hexconv_bf16 launch");
}

int nz_hexconv_set_trace(void* dev_buffer) {
  nz_hexconv_trace = (long long*)dev_buffer;
  return 0;
}

int nz_noise_probe(double* out, int n, double alpha, double scale, uint64_t seed, void* stream) {
  if (!out || n <= 0) return nz::fail("bad argument");
  nz::gamma_probe_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(out, n, alpha, scale, seed);
  cudaError_t err = cudaGetLastError();
  return err == cudaSuccess ? 0 : nz::cuda_fail(err, "nz_noise_probe launch");
}

int nz_stubnet_forward(const void* leaf, int leaf_dtype, const int32_t* salt, const uint32_t* uid, int uid_stride, int salt_uid_mul,
                       int n, int n_features, int n_actions, void* policy_out, int policy_dtype, float* value_out,
                       void* stream) {
  if (!leaf || !policy_out || !value_out) return nz::fail("null tensor pointer");
  if (n <= 0) return 0;
  if (n_features <= 32 && n_actions <= 32) {
    nz::stubnet_small_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
        leaf, leaf_dtype, salt, uid, uid_stride, salt_uid_mul, n, n_features, n_actions, policy_out, policy_dtype, value_out);
  } else {
    const int wpb = 4;
    nz::stubnet_kernel<<<(n + wpb - 1) / wpb, 32 * wpb, 0, (cudaStream_t)stream>>>(
        leaf, leaf_dtype, salt, uid, uid_stride, salt_uid_mul, n, n_features, n_actions, policy_out, policy_dtype, value_out);
  }
  cudaError_t err = cudaGetLastError();
  return err == cudaSuccess ? 0 : nz::cuda_fail(err, "nz_stubnet_forward launch");
}

}  // extern "C"
