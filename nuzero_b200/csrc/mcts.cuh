// Tile-per-game PUCT search over per-slot node pools in HBM (sm_100a).
//
// A tile of Game::TILE lanes (a full warp for SCS, 8 lanes for Tic-Tac-Toe) owns one game slot for
// the whole launch: it consumes the pending network row (expand + backup), then runs simulations —
// select with a REDUX arg-max over the children, game step, terminal backup — until a leaf needs the
// network, and commits finished moves itself in auto mode.  Per-game sequencing is exactly the
// reference's (one simulation in flight per game, Search/Explorer.py:49-62); the batch comes from
// the number of games.  With a network that emits probabilities (the parity stub) every result is
// bit-identical to the reference; with logits the softmax is computed in f32 like scipy's, but not
// with scipy's summation order (priors agree to a few ulp, tests/test_gpu_logits_parity.py).
//
// Round-2 layout: node 0 of a slot is always the root and nodes 1..K its children (the root's child count and
// visit count live in the control block, so a launch needs no load to find the first tree level); backup is
// fire-and-forget RED.ADD (N += 1, W += v) instead of a read-modify-write round trip; child runs start on
// 64-byte boundaries.  Measured and rejected (profiles/r2_notes.md): loading the first level and prefetching a
// pending leaf's inputs before the control block arrives (no gain: the launch is bound by instruction issue,
// not by the length of the load chain), and ld.global.cg for node records (LDG.STRONG.GPU on sm_100: -22 %).
//
// Round-2 additions, all result-identical (tests/test_gpu_scs_parity.py, test_gpu_cache.py):
//  * per-run game states (Game::NODE_STATE, View::nstate): an expanded node keeps its compact game state, the descent selects
//    on node records alone and leaf_state() steps the game once from the leaf's parent instead of once per tree level;
//  * advance_kernel<Game, DENSE = true>: the inference cache is consulted inside the launch (cache_probe: HIT expands in
//    place and the game runs on, OWN claims the entry and takes the next dense row of the leaf tensor, SHARE waits for the
//    row another game of the launch already claimed for the same state), expansions are published next to their entries
//    (expand(): the first expander writes the (action, prior) list, later ones copy it), and a launch may end when a batch is
//    full or enough games are parked; two lanes of rows let the network call of launch k run beside launch k + 1.
#pragma once
#include "common.cuh"

namespace nz {

#define NZ_REC_HDR 12

struct Slot {  // tile-uniform registers: the ctl words the simulation loop touches
  uint32_t phase, pool_top, sims_done, root_N0, root_K, path_len, leaf, err;  // words 0-7, written back together
  uint32_t noised, map, half;
  uint32_t noised0, half0;  // as loaded: the two words are written back only when they changed
  uint32_t d_sims, d_levels, d_scanned, d_terminal;  // deltas of this launch
};
// cold ctl words (move, uid, games_done, chosen, n_pending, counters) are read and written in place by
// the few code paths that need them (once per move)

__device__ __forceinline__ void slot_load(Slot& s, const uint32_t* ctl) {
  // words 0-7 and 8-15 of the control block as two 256-bit loads
  unsigned long long a, b, c, d, e, f, g2, h;
  asm volatile("ld.global.v4.u64 {%0, %1, %2, %3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(ctl) : "memory");
  asm volatile("ld.global.v4.u64 {%0, %1, %2, %3}, [%4];" : "=l"(e), "=l"(f), "=l"(g2), "=l"(h) : "l"(ctl + 8) : "memory");
  s.phase = (uint32_t)a; s.pool_top = (uint32_t)(a >> 32); s.sims_done = (uint32_t)b; s.root_N0 = (uint32_t)(b >> 32);
  s.root_K = (uint32_t)c; s.path_len = (uint32_t)(c >> 32); s.leaf = (uint32_t)d; s.err = (uint32_t)(d >> 32);
  s.noised = (uint32_t)e;          // NZ_CTL_NOISED = 8
  s.map = (uint32_t)(e >> 32);     // NZ_CTL_MAP = 9
  s.half = (uint32_t)h;            // NZ_CTL_HALF = 14
  s.noised0 = s.noised; s.half0 = s.half;
  s.d_sims = s.d_levels = s.d_scanned = s.d_terminal = 0;
}

__device__ __forceinline__ void slot_store(const Slot& s, uint32_t* ctl, int tl) {
  if (tl == 0) {
    const unsigned long long a = (unsigned long long)s.phase | ((unsigned long long)s.pool_top << 32);
    const unsigned long long b = (unsigned long long)s.sims_done | ((unsigned long long)s.root_N0 << 32);
    const unsigned long long c = (unsigned long long)s.root_K | ((unsigned long long)s.path_len << 32);
    const unsigned long long d = (unsigned long long)s.leaf | ((unsigned long long)s.err << 32);
    asm volatile("st.global.v4.u64 [%4], {%0, %1, %2, %3};" :: "l"(a), "l"(b), "l"(c), "l"(d), "l"(ctl) : "memory");
    if (s.noised != s.noised0) ctl[NZ_CTL_NOISED] = s.noised;
    if (s.half != s.half0) ctl[NZ_CTL_HALF] = s.half;
    // statistics: fire-and-forget reductions (no load, nothing to wait for at the end of the kernel)
    if (s.d_sims) atomicAdd(ctl + NZ_CTL_N_SIMS, s.d_sims);
    if (s.d_levels) atomicAdd(ctl + NZ_CTL_N_LEVELS, s.d_levels);
    if (s.d_scanned) atomicAdd(ctl + NZ_CTL_N_SCANNED, s.d_scanned);
    if (s.d_terminal) atomicAdd(ctl + NZ_CTL_N_TERMINAL, s.d_terminal);
  }
}

// exploration bias c(N) = log((N + base + 1) / base) + init (Explorer.py:103-108) and sqrt(N)
// (:110-112) come from one 16-byte table entry computed on the host with the libm the reference
// uses; beyond the table the device functions are used and the slot is flagged.
__device__ __forceinline__ double2 bias_sqrt(const View& v, int n, uint32_t& err) {
  if (n < v.ctable_len) return v.ctable[n];
  err |= NZ_ERR_CTABLE;
  return make_double2(log(((double)n + v.pb_c_base + 1.0) / v.pb_c_base) + v.pb_c_init, __dsqrt_rn((double)n));
}

// ---- node record accessors --------------------------------------------------------------------------
// {prior f64 | W f64 | N i32 | flags u32 (bit 0 noised f64 prior, bits 16-31 action) | first child u32 | n_children u32}
struct NodeRec { double prior, W; int N; uint32_t flags, base, K; };
// whole 32-byte record in ONE 256-bit access (sm_100: LDG.E.256 / STG.E.256), L2 only (see the header)
__device__ __forceinline__ NodeRec ld_node(const View& v, size_t i) {
  unsigned long long a, b, c, d;
#if defined(NZ_NODE_LD_CG)  // experiment: ld.global.cg is LDG.STRONG.GPU on sm_100 — measured 22 % slower than the weak load
  asm volatile("ld.global.cg.v4.u64 {%0, %1, %2, %3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(v.node + 2 * i) : "memory");
#elif defined(NZ_NODE_LD_NA)  // experiment: weak load that does not allocate in L1
  asm volatile("ld.global.L1::no_allocate.v4.u64 {%0, %1, %2, %3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(v.node + 2 * i) : "memory");
#else
  asm volatile("ld.global.v4.u64 {%0, %1, %2, %3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(v.node + 2 * i) : "memory");
#endif
  NodeRec r;
  r.prior = __longlong_as_double((long long)a);
  r.W = __longlong_as_double((long long)b);
  r.N = (int)(uint32_t)c; r.flags = (uint32_t)(c >> 32);
  r.base = (uint32_t)d; r.K = (uint32_t)(d >> 32);
  return r;
}
__device__ __forceinline__ void st_node(const View& v, size_t i, double prior, double W, int N, uint32_t flags, uint32_t base, uint32_t K) {
  const unsigned long long a = (unsigned long long)__double_as_longlong(prior), b = (unsigned long long)__double_as_longlong(W);
  const unsigned long long c = (unsigned long long)(uint32_t)N | ((unsigned long long)flags << 32);
  const unsigned long long d = (unsigned long long)base | ((unsigned long long)K << 32);
  asm volatile("st.global.v4.u64 [%4], {%0, %1, %2, %3};" :: "l"(a), "l"(b), "l"(c), "l"(d), "l"(v.node + 2 * i) : "memory");
}
// Record copies (re-rooting, compaction) go through two 128-bit halves: ptxas 12.9 turns a 256-bit load whose registers
// feed a 256-bit store unchanged into a copy of the first 8 bytes only (SASS: LDG.E.ENL2.256 RZ, R, [..], 0x3 + STG.E.64).
struct NodeRaw {
  uint4 lo, hi;  // lo: prior, W   hi: N, flags, first child, n_children
  __device__ __forceinline__ int N() const { return (int)hi.x; }
  __device__ __forceinline__ uint32_t flags() const { return hi.y; }
  __device__ __forceinline__ uint32_t base() const { return hi.z; }
  __device__ __forceinline__ uint32_t K() const { return hi.w; }
};
__device__ __forceinline__ NodeRaw ld_raw(const View& v, size_t i) {
  NodeRaw r;
  r.lo = __ldcg(v.node + 2 * i);
  r.hi = __ldcg(v.node + 2 * i + 1);
  return r;
}
__device__ __forceinline__ void st_raw(const View& v, size_t i, const NodeRaw& r) {
  v.node[2 * i] = r.lo;
  v.node[2 * i + 1] = r.hi;
}
__device__ __forceinline__ void clear_node(const View& v, size_t i) { st_node(v, i, 0.0, 0.0, 0, 0u, 0u, 0u); }
__device__ __forceinline__ int* node_N_ptr(const View& v, size_t i) { return (int*)(v.node + 2 * i + 1); }
__device__ __forceinline__ double* node_W_ptr(const View& v, size_t i) { return (double*)(v.node + 2 * i) + 1; }
__device__ __forceinline__ double* node_prior_ptr(const View& v, size_t i) { return (double*)(v.node + 2 * i); }
__device__ __forceinline__ uint32_t* node_flags_ptr(const View& v, size_t i) { return (uint32_t*)(v.node + 2 * i + 1) + 1; }
__device__ __forceinline__ uint32_t* node_base_ptr(const View& v, size_t i) { return (uint32_t*)(v.node + 2 * i + 1) + 2; }
// cold-path field reads (L2, coherent with the REDs of the backup)
__device__ __forceinline__ int ld_N(const View& v, size_t i) { return __ldcg(node_N_ptr(v, i)); }
__device__ __forceinline__ double ld_W(const View& v, size_t i) { return __ldcg(node_W_ptr(v, i)); }
__device__ __forceinline__ uint32_t ld_flags(const View& v, size_t i) { return __ldcg(node_flags_ptr(v, i)); }
// the child range of a node as one aligned 8-byte store (expand)
__device__ __forceinline__ void st_range(const View& v, size_t i, uint32_t base, uint32_t K) {
  *(uint2*)node_base_ptr(v, i) = make_uint2(base, K);
}
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" :: "l"(p)); }
__device__ __forceinline__ int ld_acquire_s32(const int32_t* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// game state of an EXPANDED node (View::nstate), keyed by the node's child run: run_base = slot offset + first child, which is
// even for every run of the general pool (the root's own state lives in "gstate"), so P / 2 rows per slot suffice; a row
// moves only when the run moves (compaction), never at re-rooting.  16-byte aligned.
__device__ __forceinline__ uint32_t* nstate_row(const View& v, size_t run_base) {
  return v.nstate + (run_base >> 1) * (size_t)v.nstate_words;
}
__device__ __forceinline__ void nstate_copy(const View& v, size_t dst, size_t src) {  // one lane copies one row
  const uint4* a = (const uint4*)nstate_row(v, src);
  uint4* b = (uint4*)nstate_row(v, dst);
  for (int w = 0; w < (v.nstate_words >> 2); ++w) b[w] = __ldcg(a + w);
}

// PUCT score of one child (Explorer.score, Explorer.py:114-130), the two arithmetic chains of
// SURVEY.md §8a spelled with non-fusing intrinsics so that no FMA contraction can change a bit.
__device__ __forceinline__ double score_f64(double prior, double u, double c, double q) {
  return __dadd_rn(__dmul_rn(__dmul_rn(prior, u), c), q);
}
__device__ __forceinline__ double score_f32(float prior, double u, double c, double q) {
  float s = __fadd_rn(__fmul_rn(__fmul_rn(prior, __double2float_rn(u)), __double2float_rn(c)), __double2float_rn(q));
  return (double)s;
}

// ---- backup (Explorer.backpropagate, Explorer.py:132-135): N += 1, W += value, no sign flip -----
// Two reductions per path node, performed by the L2 (RED.ADD.S32 / RED.ADD.F64: an IEEE round-to-nearest add, the same
// `W += value` the reference performs); nothing is loaded and nothing is waited for.
// dn / add_value: the virtual-loss mode splits the update in two — the visit (dn = 1, no value) when a descent parks
// its leaf at the network, the value (dn = 0) when the network's answer arrives.
template <int TILE>
__device__ __forceinline__ void backup(const View& v, size_t nb, const uint32_t* path, int n_path, double value,
                                       const Tl<TILE>& t, int dn = 1, bool add_value = true) {
  for (int i = t.tl; i < n_path; i += TILE) {
    const size_t idx = nb + path[i];
#ifdef NZ_BACKUP_LDST  // experiment: read-modify-write instead of reductions
    if (dn) *node_N_ptr(v, idx) = __ldcg(node_N_ptr(v, idx)) + dn;
    if (add_value) *node_W_ptr(v, idx) = __dadd_rn(__ldcg(node_W_ptr(v, idx)), value);
#else
    if (dn) asm volatile("red.global.add.s32 [%0], %1;" :: "l"(node_N_ptr(v, idx)), "r"(dn) : "memory");
    if (add_value) asm volatile("red.global.add.f64 [%0], %1;" :: "l"(node_W_ptr(v, idx)), "d"(value) : "memory");
#endif
  }
  t.sync();
}

// iterate the set bits of a tile-shared bit set in ascending order; body(a, rank) runs on one lane
template <int TILE, class F>
__device__ __forceinline__ int for_each_valid(const Tl<TILE>& t, const uint32_t* words, int nwords, F body) {
  int running = 0;
  if (TILE == 1) {  // one thread per game: walk the set bits
    for (int w = 0; w < nwords; ++w) {
      uint32_t bits = words[w];
      while (bits) {
        const int b = __ffs(bits) - 1;
        bits &= bits - 1u;
        body(w * 32 + b, running++);
      }
    }
    return running;
  }
  for (int w = 0; w < nwords; ++w) {
    const uint32_t bits = words[w];
    if (bits == 0u) continue;
#pragma unroll
    for (int c = 0; c < 32; c += TILE) {
      const uint32_t chunk = TILE == 32 ? bits : ((bits >> c) & ((1u << (TILE & 31)) - 1u));
      if (chunk == 0u) continue;
      if ((chunk >> t.tl) & 1u) body(w * 32 + c + t.tl, running + __popc(chunk & ((1u << t.tl) - 1u)));
      running += __popc(chunk);
    }
  }
  return running;
}

// end of the part of the general pool the current tree may grow into (with compaction: the current half)
__device__ __forceinline__ uint32_t pool_end(const View& v, const Slot& s) {
  return v.compact ? (uint32_t)v.g0 + (s.half + 1u) * (uint32_t)v.half_nodes : (uint32_t)v.P;
}

// ---- expand (Explorer.evaluate, Explorer.py:137-181) --------------------------------------------
// Returns the network value; creates one child per legal action, ascending action order.  The children of the root go
// to nodes 1..K, everybody else's to the top of the general pool, at an even index: a child run starts on a 64-byte
// boundary, so a run of K records costs ceil(K / 2) DRAM bursts.
template <class Game>
__device__ __forceinline__ double expand(const View& v, Slot& s, uint32_t* ctl, size_t row, size_t nb, uint32_t leaf,
                                         typename Game::Scratch& scr, uint32_t* words, const void* policy_in,
                                         int policy_dtype, const float* value_in, const typename Game::T& t,
                                         uint32_t& new_base, int& new_k, int exp_slot = -1) {
  new_base = 0u;
  new_k = 0;  // stays 0 when no child is created (no legal action, or a fault)
  using PriorT = typename Game::PriorT;
  constexpr int TILE = Game::TILE;
  const int A = v.A, nwords = (A + 31) >> 5;
  const double value = (double)value_in[row];  // predicted_value.item() (Explorer.py:162); row = network row of this leaf
  // Published expansions (in-kernel inference cache): the legal actions and priors of a state are a function of the state and
  // of the network's row, so the first game that expands a cached state writes its (action, prior) list next to the table entry
  // and every later expansion of that state — thousands of games share it — copies the list instead of recomputing the legal
  // mask and the soft-max over all actions.
  int pubK = -1;
  if (exp_slot >= 0) {
    int km = 0;
    if (t.tl == 0) km = ld_acquire_s32(v.cache_exp_meta + exp_slot);
    t.sync();
    pubK = t.bcast(km, 0) - 1;
  }
  if (pubK >= 0) {
    const int K = pubK;
    if (K == 0) return value;
    uint32_t base = 1u;
    if (leaf != 0u) {
      base = (s.pool_top + 1u) & ~1u;
      if (base + (uint32_t)K > pool_end(v, s)) s.err |= NZ_ERR_POOL_FULL;
    }
    if (s.err & NZ_ERR_POOL_FULL) {
      s.phase = NZ_PHASE_ERROR;
      return value;
    }
    if (leaf != 0u) s.pool_top = base + (uint32_t)K;
    else s.root_K = (uint32_t)K;
    const size_t e0 = (size_t)exp_slot * v.cache_exp_width;
    for (int i = t.tl; i < K; i += TILE)
      st_node(v, nb + base + i, __ldcg(v.cache_exp_prior + e0 + i), 0.0, 0, (uint32_t)__ldcg(v.cache_exp_act + e0 + i) << 16, 0u, 0u);
    if (Game::NODE_STATE && v.nstate != nullptr && leaf != 0u) Game::save(scr, nstate_row(v, nb + base), v, t);
    if (t.tl == 0) {
      st_range(v, nb + leaf, base, (uint32_t)K);
      atomicAdd(ctl + NZ_CTL_N_EXPAND, 1u);
      atomicAdd(ctl + NZ_CTL_N_CREATED, (uint32_t)K);
    }
    new_base = base;
    new_k = K;
    t.sync();
    return value;
  }
  Game::legal(scr, v, (int)s.map, words, t);
  const size_t prow = row * A;

  // softmax over ALL actions when the network emits logits (Explorer.py:152/159), in f32 like scipy.special.softmax:
  // exp(x - max) / sum(exp(x - max)) with the accurate expf (scipy's summation order is not reproduced)
  float smax = 0.f, ssum = 1.f;
  if (!v.policy_is_prob) {
    float m = -INFINITY;
    for (int a = t.tl; a < A; a += TILE) m = fmaxf(m, load_policy(policy_in, policy_dtype, prow + a));
    m = t.fmax(m);
    float e = 0.f;
    for (int a = t.tl; a < A; a += TILE) e += expf(load_policy(policy_in, policy_dtype, prow + a) - m);
    smax = m;
    ssum = t.sum(e);
  }
  auto prob_of = [&](int a) -> float {
    const float x = load_policy(policy_in, policy_dtype, prow + a);
    return v.policy_is_prob ? x : __fdiv_rn(expf(x - smax), ssum);
  };

  // total over the legal actions (np.sum(probs), Explorer.py:169)
  PriorT part = (PriorT)0;
  const int K = for_each_valid(t, words, nwords, [&](int a, int) { part += (PriorT)prob_of(a); });
  PriorT total = t.sum(part);
  const bool uniform = (total == (PriorT)0);  // "network predicted zero valid actions" workaround (:171-174)
  if (uniform) total = (PriorT)K;
  const bool publish = exp_slot >= 0 && K <= v.cache_exp_width;
  if (K == 0) {  // no legal action: node stays childless and is re-evaluated each visit
    if (publish && t.tl == 0) atomicExch(v.cache_exp_meta + exp_slot, 1);
    return value;
  }
  uint32_t base = 1u;
  if (leaf != 0u) {
    base = (s.pool_top + 1u) & ~1u;
    if (base + (uint32_t)K > pool_end(v, s)) s.err |= NZ_ERR_POOL_FULL;
  }
  if (K > v.max_children) s.err |= NZ_ERR_POOL_FULL;
  if (s.err & NZ_ERR_POOL_FULL) {
    s.phase = NZ_PHASE_ERROR;
    return value;
  }
  if (leaf != 0u) s.pool_top = base + (uint32_t)K;
  else s.root_K = (uint32_t)K;
  for_each_valid(t, words, nwords, [&](int a, int rank) {
    const size_t idx = nb + base + rank;
    const PriorT p = uniform ? (PriorT)1 : (PriorT)prob_of(a);
    // prior = probs[i] / total: IEEE division in f64 or f32 like the reference's numpy scalar; an f32
    // prior is kept as the (exact) double of that float
    const double prior = (double)(PriorT)(p / total);
    st_node(v, idx, prior, 0.0, 0, (uint32_t)a << 16, 0u, 0u);
    if (publish) {
      v.cache_exp_prior[(size_t)exp_slot * v.cache_exp_width + rank] = prior;
      v.cache_exp_act[(size_t)exp_slot * v.cache_exp_width + rank] = (uint16_t)a;
    }
  });
  if (publish) {  // several games may publish the same list at once: identical values
    __threadfence();
    t.sync();
    if (t.tl == 0) atomicExch(v.cache_exp_meta + exp_slot, K + 1);
  }
  // the expanded node keeps its game state: its children's first visits step from here (leaf_state)
  if (Game::NODE_STATE && v.nstate != nullptr && leaf != 0u) Game::save(scr, nstate_row(v, nb + base), v, t);
  if (t.tl == 0) {
    st_range(v, nb + leaf, base, (uint32_t)K);
    atomicAdd(ctl + NZ_CTL_N_EXPAND, 1u);
    atomicAdd(ctl + NZ_CTL_N_CREATED, (uint32_t)K);
  }
  new_base = base;
  new_k = K;
  t.sync();
  return value;
}

// ---- root noise (Explorer.add_exploration_noise, Explorer.py:201-210) ----------------------------
template <class Game>
__device__ __noinline__ void add_root_noise(const View& v, Slot& s, uint32_t move, uint32_t uid, int g, size_t nb,
                                           const typename Game::T& t) {
  constexpr int TILE = Game::TILE;
  const int K = (int)s.root_K;
  s.noised = 0;
  if (K == 0) return;
  const double frac = v.noise_frac;
  for (int i = t.tl; i < K; i += TILE) {
    double n;
    if (v.tape_moves > 0) {
      const int m = min((int)move, v.tape_moves - 1);
      n = (i < v.tape_width) ? v.gamma_tape[((size_t)g * v.tape_moves + m) * v.tape_width + i] : 0.0;
    } else {
      n = philox_gamma(v.seed ^ ((unsigned long long)uid * 0x9E3779B97F4A7C15ull), move, 1u, (uint32_t)i,
                       v.noise_alpha, v.noise_beta);
    }
    const size_t idx = nb + 1 + i;
    const double nf = __dmul_rn(n, frac);
    double* pp = node_prior_ptr(v, idx);
    uint32_t* pf = node_flags_ptr(v, idx);
    const double p = __ldcg(pp);
    const uint32_t f = __ldcg(pf);
    if (Game::PRIOR_F64 || (f & 1u)) {
      *pp = __dadd_rn(__dmul_rn(p, 1.0 - frac), nf);
    } else {
      // np.float32 * python float -> float32 ; + np.float64 -> float64: the prior becomes a true f64
      const float scaled = __fmul_rn((float)p, __double2float_rn(1.0 - frac));
      *pp = __dadd_rn((double)scaled, nf);
      *pf = f | 1u;
    }
  }
  s.noised = 1;
  t.sync();
}

// ---- final action choice (Explorer.select_action, Explorer.py:70-97) -> child index --------------
template <class Game>
__device__ __noinline__ int choose_child(const View& v, uint32_t move, uint32_t uid, int g, size_t nb, uint32_t base,
                                         int K, int game_length, const typename Game::T& t) {
  constexpr int TILE = Game::TILE;
  // max_action (:183-185): python max with key -> first maximum -> LOWEST action on ties
  int bn = -1, bi = 0x7fffffff;
  for (int i = t.tl; i < K; i += TILE) {
    const int n = ld_N(v, nb + base + i);
    if (n > bn) { bn = n; bi = i; }
  }
  const int mx = t.imax(bn);
  const int idx = -t.imax(bn == mx ? -bi : (int)0x80000001);  // smallest index among the maxima
  if (!v.training) return idx;

  double u0, u1, u2;
  if (v.tape_moves > 0) {
    const int m = min((int)move, v.tape_moves - 1);
    const double* row = v.unif_tape + ((size_t)g * v.tape_moves + m) * 3;
    u0 = row[0]; u1 = row[1]; u2 = row[2];
  } else {
    uint32_t r[4], r2[4];
    const unsigned long long k = v.seed ^ ((unsigned long long)uid * 0x9E3779B97F4A7C15ull);
    Philox::gen(k, move, 2u, 0u, 0u, r);
    Philox::gen(k, move, 2u, 1u, 0u, r2);
    u0 = u01(r[0], r[1]); u1 = u01(r[2], r[3]); u2 = u01(r2[0], r2[1]);
  }
  int mode = 0;  // 0 max, 1 softmax over visit counts, 2 uniform over legal actions
  if (game_length < v.n_softmax_moves) mode = 1;
  else if (u0 < v.eps_softmax) mode = 1;
  else if (u1 < v.eps_random) mode = 2;
  if (mode == 0) return idx;

  // sequential on one lane, in the reference's order of operations (rare: a few % of moves)
  int pick = 0;
  if (t.tl == 0) {
    if (mode == 1) {
      // softmax_action (:187-199): scipy softmax of the raw counts, renormalise, np.random.choice
      const double mxd = (double)mx;
      double sum = 0.0;
      for (int i = 0; i < K; ++i) sum += exp((double)ld_N(v, nb + base + i) - mxd);
      double sum2 = 0.0;
      for (int i = 0; i < K; ++i) sum2 += exp((double)ld_N(v, nb + base + i) - mxd) / sum;
      double tot = 0.0;
      for (int i = 0; i < K; ++i) tot += (exp((double)ld_N(v, nb + base + i) - mxd) / sum) / sum2;
      double cdf = 0.0;
      for (int i = 0; i < K; ++i) {
        cdf += (exp((double)ld_N(v, nb + base + i) - mxd) / sum) / sum2;
        if (cdf / tot <= u2) pick = i + 1;
      }
    } else {
      const double each = 1.0 / (double)K;
      double tot = 0.0, cdf = 0.0;
      for (int i = 0; i < K; ++i) tot += each;
      for (int i = 0; i < K; ++i) {
        cdf += each;
        if (cdf / tot <= u2) pick = i + 1;
      }
    }
    pick = min(pick, K - 1);
  }
  return t.bcast(pick, 0);
}

// ---- move record ---------------------------------------------------------------------------------
template <class Game>
__device__ __noinline__ void write_record(const View& v, Slot& s, uint32_t move, uint32_t uid, int g, size_t nb,
                                          uint32_t base, int K, int child, int action, int player, const uint32_t* state_before, bool game_end, int tv,
                                          int length_after, const typename Game::T& t) {
  constexpr int TILE = Game::TILE;
  const int SW = v.state_words;
  const int len = NZ_REC_HDR + SW + 2 * K + (v.record_detail ? 4 * K : 0);
  uint32_t off = 0;
  if (t.tl == 0) off = atomicAdd(v.arena_top, (uint32_t)len);
  off = t.bcast(off, 0);
  if (off + (uint32_t)len > (uint32_t)v.arena_words) {
    if (t.tl == 0) atomicAdd(v.arena_top + 1, 1u);
    s.err |= NZ_ERR_ARENA_FULL;
    return;
  }
  if (t.tl == 0) {  // record index: lets the replay decoder (nz_replay_decode) take the records in parallel
    const uint32_t slot_i = atomicAdd(v.arena_top + 2, 1u);
    if (slot_i < (uint32_t)v.rec_index_len) v.rec_index[slot_i] = off;
  }
  uint32_t* r = v.arena + off;
  const int rootN = ld_N(v, nb);
  const double rootW = ld_W(v, nb);
  uint32_t dummy = 0;
  const double bias = bias_sqrt(v, rootN, dummy).x;
  if (t.tl == 0) {
    r[0] = (uint32_t)len;
    r[1] = uid;
    r[2] = move | ((uint32_t)K << 16);
    r[3] = (uint32_t)action | ((uint32_t)player << 16) |
           (((v.record_detail ? 1u : 0u) | (game_end ? 2u : 0u) | ((uint32_t)(tv + 1) << 2)) << 24);
    r[4] = (uint32_t)rootN;
    r[5] = (uint32_t)__double_as_longlong(rootW);
    r[6] = (uint32_t)(__double_as_longlong(rootW) >> 32);
    r[7] = (uint32_t)g | (s.map << 20);  // slot (20 bits) | scenario map (12 bits)
    r[8] = (uint32_t)__double_as_longlong(bias);
    r[9] = (uint32_t)(__double_as_longlong(bias) >> 32);
    r[10] = (uint32_t)length_after;
    r[11] = (uint32_t)child;
  }
  for (int i = t.tl; i < SW; i += TILE) r[NZ_REC_HDR + i] = state_before[i];
  uint32_t* c = r + NZ_REC_HDR + SW;
  for (int i = t.tl; i < K; i += TILE) {
    const NodeRec h = ld_node(v, nb + base + i);
    c[2 * i] = h.flags >> 16;
    c[2 * i + 1] = (uint32_t)h.N;
    if (v.record_detail) {
      const long long w = __double_as_longlong(h.W);
      const long long p = __double_as_longlong(h.prior);
      uint32_t* d = c + 2 * K + 4 * i;
      d[0] = (uint32_t)w; d[1] = (uint32_t)(w >> 32); d[2] = (uint32_t)p; d[3] = (uint32_t)(p >> 32);
    }
  }
}

// ---- keep_subtree with compaction: breadth-first copy of the current tree below the root's children (which already
// sit in nodes 1..K) into the other half of the general pool (one BFS level per outer iteration; lanes take the level's
// nodes side by side, a tile-wide exclusive scan hands out the new child ranges, each on an even index).  Dead nodes are
// simply left behind.
template <class Game>
__device__ __noinline__ void compact_subtree(const View& v, Slot& s, size_t nb, const typename Game::T& t) {
  constexpr int TILE = Game::TILE;
  const uint32_t dst0 = (uint32_t)v.g0 + (s.half ^ 1u) * (uint32_t)v.half_nodes;  // the half that does NOT hold the current tree
  const uint32_t dst_end = dst0 + (uint32_t)v.half_nodes;
  auto copy_node = [&](uint32_t dst, uint32_t src) { st_raw(v, nb + dst, ld_raw(v, nb + src)); };
  t.sync();
  uint32_t lo = 1u, hi = 1u + s.root_K, top = dst0;
  bool overflow = false;
  while (lo < hi && !overflow) {
    const uint32_t level_start = top;
    for (uint32_t i0 = lo; i0 < hi; i0 += TILE) {
      const uint32_t i = i0 + (uint32_t)t.tl;
      const bool valid = i < hi;
      NodeRec h = {};
      if (valid) h = ld_node(v, nb + i);
      const int K = (int)h.K, Ke = (K + 1) & ~1;
      int incl = Ke;
#pragma unroll
      for (int off = 1; off < TILE; off <<= 1) {
        const int n = __shfl_up_sync(t.mask, incl, off, TILE);
        if (t.tl >= off) incl += n;
      }
      const int total = t.bcast(incl, TILE - 1);
      if (top + (uint32_t)total > dst_end) { overflow = true; break; }
      const uint32_t newbase = top + (uint32_t)(incl - Ke);
      if (K > 0) {
        for (int c = 0; c < K; ++c) copy_node(newbase + (uint32_t)c, h.base + (uint32_t)c);
        if (Ke != K) clear_node(v, nb + newbase + (uint32_t)K);  // the padding slot is walked by the next level: childless
        if (Game::NODE_STATE && v.nstate != nullptr) nstate_copy(v, nb + newbase, nb + h.base);  // the run's owner state moves with it
        *node_base_ptr(v, nb + i) = newbase;
      }
      top += (uint32_t)total;
    }
    t.sync();
    lo = level_start;
    hi = top;
  }
  if (overflow) {
    s.err |= NZ_ERR_POOL_FULL;
    s.phase = NZ_PHASE_ERROR;
    return;
  }
  s.half ^= 1u;
  s.pool_top = top;
  t.sync();
}

// ---- commit a move: Training/Gamer.py:74-79 (+ restart, Gamer.play_game called again) -------------
template <class Game>
__device__ __noinline__ void commit_move(const View& v, Slot& s, uint32_t* ctl, int g, size_t nb,
                                         typename Game::Scratch& rootS, uint32_t* state_tmp, int forced_action,
                                         const typename Game::T& t) {
  constexpr int TILE = Game::TILE;
  const int K = (int)s.root_K;
  const uint32_t base = 1u;
  if (K == 0) {  // the reference would raise on max() of an empty sequence
    s.err |= NZ_ERR_ILLEGAL;
    s.phase = NZ_PHASE_ERROR;
    return;
  }
  uint32_t move = ctl[NZ_CTL_MOVE], uid = ctl[NZ_CTL_UID];
  int child;
  if (forced_action >= 0) {
    int found = -1;
    for (int i = t.tl; i < K; i += TILE)
      if ((int)(ld_flags(v, nb + base + i) >> 16) == forced_action) found = i;
    found = t.imax(found);
    if (found < 0) {
      s.err |= NZ_ERR_ILLEGAL;
      s.phase = NZ_PHASE_ERROR;
      return;
    }
    child = found;
  } else if (s.phase == NZ_PHASE_MOVE_READY) {
    child = (int)ctl[NZ_CTL_CHOSEN];
  } else {
    child = choose_child<Game>(v, move, uid, g, nb, base, K, Game::length(rootS), t);
  }
  NodeRaw rc = ld_raw(v, nb + base + child);  // the new root (every lane reads it: one broadcast request)
  const int action = (int)(rc.flags() >> 16);
  const int player = Game::to_play(rootS);
  Game::save(rootS, state_tmp, v, t);  // state before the move, for the record
  t.sync();
  Game::step(rootS, v, (int)s.map, action, t);
  const bool over = Game::terminal(rootS);
  write_record<Game>(v, s, move, uid, g, nb, base, K, child, action, player, state_tmp, over,
                     over ? Game::terminal_value(rootS) : 0, Game::length(rootS), t);
  move += 1;
  s.sims_done = 0;
  s.noised = 0;
  s.phase = NZ_PHASE_READY;
  uint32_t games_done = ctl[NZ_CTL_GAMES_DONE];
  t.sync();  // the record has read the old root and its children
  if (over) {
    games_done += 1;
    s.root_K = 0;
    s.root_N0 = 0;
    if (!v.auto_advance || (v.games_per_slot > 0 && (int)games_done >= v.games_per_slot)) {
      s.phase = NZ_PHASE_IDLE;
      // manual mode keeps the final position's node as the root (a view of it stays readable)
      rc.hi.z = 1u;
      rc.hi.w = 0u;
      if (t.tl == 0) st_raw(v, nb, rc);
      s.root_N0 = (uint32_t)rc.N();
    } else {  // fresh game in the same slot: Node(0) root, new game object
      uid += (uint32_t)v.G;
      move = 0;
      s.pool_top = (uint32_t)v.g0;
      s.half = 0;
      Game::reset(rootS, v, (int)s.map, t);
      if (t.tl == 0) clear_node(v, nb);
    }
  } else {
    // keep_subtree: the chosen child becomes node 0, its children move to nodes 1..K' (their own children stay where
    // they are: links point from parent to child only, so nothing else has to be rewritten)
    const int nk = (int)rc.K();
    for (int i0 = 0; i0 < nk; i0 += TILE) {
      const int i = i0 + t.tl;
      if (i < nk) st_raw(v, nb + 1 + i, ld_raw(v, nb + rc.base() + i));
    }
    rc.hi.z = 1u;
    if (t.tl == 0) st_raw(v, nb, rc);
    s.root_K = (uint32_t)nk;
    s.root_N0 = (uint32_t)rc.N();
    t.sync();
    if (v.compact) {
      // lazy: only when the current half may not hold another move's growth (<= sims * max_children new
      // nodes).  Short games (Tic-Tac-Toe) never pay for it; long SCS games compact every few moves.
      const uint32_t need = min((uint32_t)v.half_nodes >> 1, (uint32_t)v.sims * (uint32_t)(v.max_children + 1));
      if (pool_end(v, s) - s.pool_top < need) compact_subtree<Game>(v, s, nb, t);
    }
    if (v.training && s.phase == NZ_PHASE_READY) add_root_noise<Game>(v, s, move, uid, g, nb, t);
  }
  t.sync();  // every lane has read the old ctl words
  if (t.tl == 0) {
    ctl[NZ_CTL_MOVE] = move;
    ctl[NZ_CTL_UID] = uid;
    ctl[NZ_CTL_GAMES_DONE] = games_done;
    atomicAdd(ctl + NZ_CTL_N_MOVES, 1u);
  }
  t.sync();
}

// ---- in-kernel inference cache (Explorer.evaluate's cache.get, Explorer.py:146-155) --------------------------------------
// key = the leaf's compact state words (in shared memory) + the slot's scenario map, compared word for word.  Entry states
// (cache_meta): 0 empty, 1 a slot is writing the key, 3 PENDING (key written; the state waits for the network in dense row
// cache_row[p] of this launch), 2 ready (nz_cache_insert_dense stored the network's output).  Outcomes of a probe:
//   HIT    a ready entry: expand from the stored row, the game goes on;
//   SHARE  a pending entry with the same key: another game of this launch already sends this state to the network — wait for
//          ITS row instead of adding a duplicate (thousands of games on one scenario reach the same states together);
//   OWN    nobody has it: the first empty slot of the probe sequence is claimed, the leaf takes the next dense row.
// An entry that is being written while we look (state 1) is re-read a few times; if it stays busy it is skipped, which at
// worst evaluates a state twice.  OWN / SHARE return row | lane << 30 (the lane whose network call produces the row: a pending
// entry may belong to the previous launch, whose forward runs while this launch searches).
enum { NZ_PROBE_OWN = 0, NZ_PROBE_HIT = 1, NZ_PROBE_SHARE = 2 };
template <int TILE>
__device__ __forceinline__ int cache_probe(const View& v, uint32_t* key, uint32_t map, int g, const Tl<TILE>& t, uint32_t& out,
                                           uint32_t& slot1) {
  const int kw = v.cache_kw;
  slot1 = 0u;  // table slot + 1 of the entry this leaf belongs to (0: none)
  if (t.tl == 0) key[kw - 1] = map;
  t.sync();
  uint32_t p = cache_hash_tile<TILE>(key, kw, t.tl, t.mask) & v.cache_mask;
  for (int probe = 0; probe < 64; ++probe) {
    int m = 0;
    if (t.tl == 0) {
      m = ld_acquire_s32(v.cache_meta + p);
      for (int spin = 0; m == 1 && spin < 64; ++spin) {
        __nanosleep(40);
        m = ld_acquire_s32(v.cache_meta + p);
      }
      if (m == 0) m = atomicCAS(v.cache_meta + p, 0, 1) == 0 ? -1 : ld_acquire_s32(v.cache_meta + p);  // -1: claimed by us
    }
    t.sync();
    m = t.bcast(m, 0);
    if (m == -1) {  // OWN: publish the key and the dense row
      for (int i = t.tl; i < kw; i += TILE) v.cache_keys[(size_t)p * kw + i] = key[i];
      __threadfence();
      t.sync();
      uint32_t idx = 0u;
      if (t.tl == 0) {
        idx = atomicAdd(v.dense_count + 4 * v.lane, 1u);
        v.dense_rows[(size_t)v.lane * v.G + idx] = g;
        v.cache_row[p] = (int32_t)(idx | ((uint32_t)v.lane << 30));
        __threadfence();
        atomicExch(v.cache_meta + p, 3);
      }
      out = t.bcast(idx, 0) | ((uint32_t)v.lane << 30);
      slot1 = p + 1u;
      return NZ_PROBE_OWN;
    }
    if (m == 2 || m == 3) {
      bool same = true;
      for (int i = t.tl; i < kw; i += TILE) same &= __ldcg(v.cache_keys + (size_t)p * kw + i) == key[i];
      if (t.ballot(!same) == 0u) {
        slot1 = p + 1u;
        if (m == 2) { out = p; return NZ_PROBE_HIT; }
        out = (uint32_t)__ldcg(v.cache_row + p);
        return NZ_PROBE_SHARE;
      }
    }
    p = (p + 1) & v.cache_mask;
  }
  // no room in the neighbourhood: an ordinary dense row without a table entry
  uint32_t idx = 0u;
  if (t.tl == 0) {
    idx = atomicAdd(v.dense_count + 4 * v.lane, 1u);
    v.dense_rows[(size_t)v.lane * v.G + idx] = g;
  }
  out = t.bcast(idx, 0) | ((uint32_t)v.lane << 30);
  return NZ_PROBE_OWN;
}

// ---- one simulation's descent (Explorer.py:54-58 + select_child :99-101) --------------------------
// Returns the leaf node; path[0..depth] filled; scratch stepped to the leaf.  Starts at `node` (child range base / K,
// visit count Np) with path[0..depth] already filled — the root with depth 0, or the node where the previous launch ran
// out of its level budget.  `*paused` is set when this launch's level budget ends before a leaf is reached.
template <class Game>
__device__ __forceinline__ uint32_t descend(const View& v, Slot& s, size_t nb, typename Game::Scratch& scr,
                                            uint32_t* path, uint32_t node, uint32_t cbase, int K, int Np, int& depth,
                                            int& levels_left, bool* paused, const typename Game::T& t, int* last_action = nullptr,
                                            uint32_t* parent_base = nullptr) {
  constexpr int TILE = Game::TILE;
  const bool stateless = Game::NODE_STATE && v.nstate != nullptr;  // select only; leaf_state() steps the game once afterwards
  *paused = false;
  if (t.tl == 0) path[depth] = node;
  while (K != 0) {
    if (levels_left <= 0) { *paused = true; break; }
    levels_left -= 1;
    const uint32_t base = cbase;
    if (depth + 1 >= v.max_depth) {
      s.err |= NZ_ERR_DEPTH;
      s.phase = NZ_PHASE_ERROR;
      break;
    }
    const bool flip = Game::MAY_FLIP && Game::to_play(scr) == 2;  // literal `parent.to_play == 2` (Explorer.py:124)
    const double2 cs = bias_sqrt(v, Np, s.err);
    unsigned long long best_key = 0ull;
    int best_i = -1, best_n = 0;
    uint32_t best_base = 0u, best_ka = 0u;
    for (int i = t.tl; i < K; i += TILE) {
      const NodeRec h = ld_node(v, nb + base + i);  // one 32-byte sector per child, one 256-bit load
      const int n = h.N;
      const double u = __ddiv_rn(cs.y, (double)(n + 1));       // sqrt(N_parent) / (n + 1)  (Explorer.py:110-112)
      double q = (n == 0) ? 0.0 : __ddiv_rn(h.W, (double)n);   // child.value() (Search/Node.py:17-20)
      if (flip) q = -q;
      q = __dmul_rn(q, v.value_factor);
      double sc;
      if (Game::PRIOR_F64 || (h.flags & 1u)) sc = score_f64(h.prior, u, cs.x, q);
      else sc = score_f32((float)h.prior, u, cs.x, q);
      const unsigned long long key = order_key(sc + 0.0);
      if (key >= best_key) {  // later index wins ties
        best_key = key; best_i = i; best_n = n; best_base = h.base; best_ka = (h.flags & 0xffff0000u) | h.K;
      }
    }
    // arg-max over the tile: REDUX.MAX on the key's high word; exact ties (python max over
    // (score, action): the HIGHEST action wins) fall back to the low word and the child index
    const uint32_t hi = (uint32_t)(best_key >> 32), lo = (uint32_t)best_key;
    const uint32_t mh = t.rmax(hi);
    const unsigned top = t.ballot(hi == mh);
    int wi;
    if ((top & (top - 1u)) == 0u) {  // a single lane holds the maximal high word: done after one REDUX
      wi = t.bcast(best_i, __ffs(top) - 1);
    } else {
      const uint32_t ml = t.rmax(hi == mh ? lo : 0u);
      wi = (int)t.rmax((hi == mh && lo == ml) ? (uint32_t)(best_i + 1) : 0u) - 1;
    }
    const int owner = wi & (TILE - 1);
    Np = t.bcast(best_n, owner);
    cbase = t.bcast(best_base, owner);
    const uint32_t ka = t.bcast(best_ka, owner);
    s.d_levels += 1;
    s.d_scanned += (uint32_t)K;
    K = (int)(ka & 0xffffu);
    if (stateless) { if (last_action) { *last_action = (int)(ka >> 16); *parent_base = base; } }
    else Game::step_descend(scr, v, (int)s.map, (int)(ka >> 16), t);
    node = base + (uint32_t)wi;
    depth += 1;
    if (t.tl == 0) path[depth] = node;
  }
  t.sync();
  return node;
}

// node_state_cache: the game state at the leaf = the stored state of the leaf's parent (the root's is `rootS`) stepped once
// by the action that leads to the leaf.  `action` / `parent_base`: the last selected action and the child run it was selected
// from (action < 0: the descent selected nothing in this launch — depth 0, or a paused descent that stopped on its leaf).
template <class Game>
__device__ __forceinline__ void leaf_state(const View& v, const Slot& s, size_t nb, typename Game::Scratch& scr,
                                           const typename Game::Scratch& rootS, const uint32_t* path, int depth, int action,
                                           uint32_t parent_base, const typename Game::T& t) {
  if (!(Game::NODE_STATE && v.nstate != nullptr)) return;
  if (depth == 0) {  // the root itself is the leaf
    Game::copy(scr, rootS, v, t);
    return;
  }
  if (action < 0) {
    action = (int)(ld_flags(v, nb + path[depth]) >> 16);
    parent_base = path[depth - 1] == 0u ? 1u : __ldcg(node_base_ptr(v, nb + path[depth - 1]));
  }
  if (parent_base == 1u) Game::copy(scr, rootS, v, t);  // nodes 1..K are the root's children
  else Game::load(scr, nstate_row(v, nb + parent_base), v, (int)s.map, t);
  Game::step(scr, v, (int)s.map, action, t);
}

// per-tile slab: path[max_depth] | legal-mask words | state tmp (+ 1: the map word of a cache key) | scratch | root scratch
// (sizes are computed once on the host: View::slab_bytes / slab_words_bytes)
template <class Game>
__host__ __device__ __forceinline__ size_t tile_slab_words_bytes(const View& v) {
  const int nwords = (v.A + 31) >> 5;
  const size_t words = (size_t)v.max_depth + nwords + v.state_words + 1;
  return (words * 4 + 15) & ~(size_t)15;
}
template <class Game>
__host__ __device__ __forceinline__ size_t tile_slab_bytes(const View& v) {
  const size_t scr = Game::SMEM ? Game::scratch_bytes(v) : 0;
  return tile_slab_words_bytes<Game>(v) + 2 * scr;
}

// DENSE: the instantiation with dense leaf rows and the in-kernel inference cache (nz_engine_attach_cache); the plain one
// keeps the register budget of the hot path untouched.
template <class Game, bool DENSE>
__global__ void __launch_bounds__(NZ_CTA_THREADS, Game::MIN_CTAS)
advance_kernel(const __grid_constant__ View v, void* leaf_out, const void* policy_in, const float* value_in, int leaf_dtype,
               int policy_dtype) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int TILE = Game::TILE;
  const typename Game::T t;
  const int tile_in_cta = threadIdx.x / TILE;
  const int g = blockIdx.x * (NZ_CTA_THREADS / TILE) + tile_in_cta;
  if (g >= v.G) return;
  // per-tile slab: path[max_depth] | legal-mask words | state tmp | scratch | root scratch
  const int nwords = (v.A + 31) >> 5;
  const size_t words_bytes = (size_t)v.slab_words_bytes;
  unsigned char* slab = smem_raw + tile_in_cta * v.slab_bytes;
  uint32_t* path = (uint32_t*)slab;
  uint32_t* words = path + v.max_depth;
  uint32_t* state_tmp = words + nwords;
  // small games (TTT) keep both game states in registers; SCS stages them in shared memory
  const size_t scr_bytes = Game::scratch_bytes(v);
  typename Game::Scratch scr_reg, root_reg;
  typename Game::Scratch& scr = Game::SMEM ? *(typename Game::Scratch*)(slab + words_bytes) : scr_reg;
  typename Game::Scratch& rootS = Game::SMEM ? *(typename Game::Scratch*)(slab + words_bytes + scr_bytes) : root_reg;

  uint32_t* ctl = v.ctl + (size_t)g * NZ_CTL_WORDS;
  uint32_t* gs_root = v.gstate + (size_t)g * (1 + v.V) * v.state_words;
  uint32_t* gs_leaf = gs_root + v.state_words;
  uint32_t* gpath = v.path + (size_t)g * v.V * v.max_depth;
  const size_t nb = (size_t)g * v.P;
  if (!Game::SMEM) Game::load(rootS, gs_root, v, 0, t);  // register-resident games: goes out together with the control block
  Slot s;
  slot_load(s, ctl);
  if (s.phase >= NZ_PHASE_MOVE_READY && s.phase != NZ_PHASE_DESCENDING) return;  // waiting for the host, idle, or faulted
  // dense rows: a leaf that waits for the OTHER lane's network call is not ready yet (that call may still be running)
  if (DENSE && s.phase == NZ_PHASE_LEAF_PENDING && (s.leaf >> 31) != (uint32_t)v.lane) return;
  if (Game::SMEM) Game::load(rootS, gs_root, v, (int)s.map, t);
  t.sync();
  bool root_dirty = false;

  if (s.phase == NZ_PHASE_LEAF_PENDING) {
    Game::load(scr, gs_leaf, v, (int)s.map, t);
    const int n_path = (int)s.path_len;
    const uint32_t leaf = DENSE ? (s.leaf & 0x7fffffffu) : s.leaf;
    for (int i = t.tl; i < n_path; i += TILE) path[i] = gpath[i];
    t.sync();
    s.phase = NZ_PHASE_READY;
    uint32_t new_base;
    int new_k;
    // dense rows: the network's answer sits in the row the leaf was written to, not in row g
    const size_t row = DENSE ? (size_t)ctl[NZ_CTL_LEAF_ROW] : (size_t)g;
    const int exp_slot = (DENSE && v.cache_exp_meta != nullptr) ? (int)ctl[NZ_CTL_LEAF_SLOT] - 1 : -1;
    const double value = expand<Game>(v, s, ctl, row, nb, leaf, scr, words, policy_in, policy_dtype, value_in, t, new_base, new_k, exp_slot);
    if (s.phase == NZ_PHASE_READY) {
      backup<TILE>(v, nb, path, n_path, value, t);
      s.sims_done += 1;
      s.d_sims += 1;
    }
  }

  int budget = v.max_sims_per_launch;
  int levels_left = v.max_levels;
  uint32_t d_hits = 0u, d_shared = 0u;
  bool resume = false;
  const bool stateless = Game::NODE_STATE && v.nstate != nullptr;
  if (s.phase == NZ_PHASE_DESCENDING) {  // pick up the descent the previous launch had to pause
    if (!stateless) Game::load(scr, gs_leaf, v, (int)s.map, t);
    const int n_path = (int)s.path_len;
    for (int i = t.tl; i < n_path; i += TILE) path[i] = gpath[i];
    t.sync();
    resume = true;
    s.phase = NZ_PHASE_READY;
  }
  while (s.phase == NZ_PHASE_READY) {
    if ((int)s.sims_done >= v.sims) {
      if (!v.auto_advance) {
        const int K = (int)s.root_K;
        if (K == 0) { s.err |= NZ_ERR_ILLEGAL; s.phase = NZ_PHASE_ERROR; break; }
        const int ch = choose_child<Game>(v, ctl[NZ_CTL_MOVE], ctl[NZ_CTL_UID], g, nb, 1u, K, Game::length(rootS), t);
        if (t.tl == 0) ctl[NZ_CTL_CHOSEN] = (uint32_t)ch;
        s.phase = NZ_PHASE_MOVE_READY;
        break;
      }
      {
        Slot tmp = s;  // the out-of-line helper takes the address of its Slot: keep `s` in registers
        commit_move<Game>(v, tmp, ctl, g, nb, rootS, state_tmp, -1, t);
        s = tmp;
      }
      root_dirty = true;
      continue;
    }
    if (budget <= 0) break;
    // dense rows: the launch ends for everybody once enough leaves wait for the network to fill a batch — games whose leaves
    // keep hitting the cache run on, nobody idles behind a fixed number of simulations (the read races with the other
    // slots' increments; only the launch boundary depends on it, never a result)
    if (DENSE && (__ldcg(v.dense_count + 4 * v.lane) >= v.dense_target || __ldcg(v.dense_count + 4 * v.lane + 1) >= v.park_target)) break;
    budget -= 1;
    int depth = 0;
    uint32_t start = 0u, sbase = 1u;
    int sK = (int)s.root_K, sN = (int)(s.root_N0 + s.sims_done);
    if (resume) {
      depth = (int)s.path_len - 1;
      start = path[depth];
      const NodeRec h = ld_node(v, nb + start);
      sbase = h.base; sK = (int)h.K; sN = h.N;
      resume = false;
    } else if (!stateless) {
      Game::copy(scr, rootS, v, t);  // game.shallow_clone() (Explorer.py:51)
    }
    bool paused;
    int last_action = -1;
    uint32_t parent_base = 0u;
    const uint32_t node = descend<Game>(v, s, nb, scr, path, start, sbase, sK, sN, depth, levels_left, &paused, t, &last_action, &parent_base);
    if (s.phase != NZ_PHASE_READY) break;
    if (paused) {  // out of levels for this launch: park the half-finished descent
      if (!stateless) Game::save(scr, gs_leaf, v, t);
      for (int i = t.tl; i <= depth; i += TILE) gpath[i] = path[i];
      s.path_len = (uint32_t)(depth + 1);
      s.phase = NZ_PHASE_DESCENDING;
      break;
    }
    leaf_state<Game>(v, s, nb, scr, rootS, path, depth, last_action, parent_base, t);
    Game::settle(scr, v, (int)s.map, t);
    if (Game::terminal(scr)) {  // Explorer.py:140-142: terminal leaves return the game's value
      const double tv = (double)Game::terminal_value(scr);
      backup<TILE>(v, nb, path, depth + 1, tv, t);
      s.sims_done += 1;
      s.d_sims += 1;
      s.d_terminal += 1;
      continue;
    }
    size_t row = (size_t)g;
    bool shared_row = false;
    uint32_t row_lane = 0u, slot1 = 0u;
    if (DENSE) {
      uint32_t idx = 0u;
      if (v.cache_keys != nullptr) {
        // Explorer.evaluate asks the cache first (Explorer.py:146-155): a state that was evaluated before is expanded from
        // the stored network output and the game goes on with its next simulation in this launch
        Game::save(scr, state_tmp, v, t);
        const int what = cache_probe<TILE>(v, state_tmp, s.map, g, t, idx, slot1);
        if (what == NZ_PROBE_HIT) {
          uint32_t new_base;
          int new_k;
          const double value = expand<Game>(v, s, ctl, (size_t)idx, nb, node, scr, words, v.cache_pol, policy_dtype, v.cache_val, t, new_base, new_k,
                                            v.cache_exp_meta != nullptr ? (int)idx : -1);
          if (s.phase != NZ_PHASE_READY) break;
          backup<TILE>(v, nb, path, depth + 1, value, t);
          s.sims_done += 1;
          s.d_sims += 1;
          d_hits += 1u;
          continue;
        }
        shared_row = what == NZ_PROBE_SHARE;
        if (shared_row) d_shared += 1u;
      } else if (t.tl == 0) {
        idx = atomicAdd(v.dense_count + 4 * v.lane, 1u);
        v.dense_rows[(size_t)v.lane * v.G + idx] = g;
        idx |= (uint32_t)v.lane << 30;
      }
      idx = t.bcast(idx, 0);
      row_lane = idx >> 30;
      idx &= 0x3fffffffu;
      if (t.tl == 0) {
        ctl[NZ_CTL_LEAF_ROW] = idx;
        ctl[NZ_CTL_LEAF_SLOT] = slot1;
        atomicAdd(v.dense_count + 4 * v.lane + 1, 1u);  // games of this launch that wait for the network (own row or somebody else's)
      }
      row = (size_t)idx;
    }
    // non-terminal leaf: hand its encoded state to the network (Explorer.py:145) — unless another game's row already carries it
    if (!shared_row) Game::encode(scr, v, (int)s.map, leaf_out, leaf_dtype, row, t);
    Game::save(scr, gs_leaf, v, t);
    for (int i = t.tl; i <= depth; i += TILE) gpath[i] = path[i];
    s.path_len = (uint32_t)(depth + 1);
    s.leaf = DENSE ? (node | (row_lane << 31)) : node;
    s.phase = NZ_PHASE_LEAF_PENDING;
  }
  if (root_dirty) Game::save(rootS, gs_root, v, t);
  if (d_hits && t.tl == 0) atomicAdd(ctl + NZ_CTL_N_CACHE_HITS, d_hits);
  if (d_shared && t.tl == 0) atomicAdd(ctl + NZ_CTL_N_CACHE_SHARED, d_shared);
  slot_store(s, ctl, t.tl);
}

// ---- throughput mode: up to V leaves of one game wait at the network (nz_config.virtual_loss_width > 1) -------------
// Not the reference's sequencing (Explorer.py:49-62 runs one simulation at a time), so results are NOT bit-identical to
// it; every invariant of the tree still holds once nothing is pending (N(node) = 1 + sum N(children), root N = carried +
// simulations).  A descent that ends in a fresh non-terminal leaf adds a virtual visit (N += 1, no value yet) along its
// path, parks leaf state / path in slot j and encodes row g * V + j; the game then starts the next descent, which the
// virtual visits steer elsewhere.  A descent that reaches a leaf already parked ends the game's launch without effect.
// The next launch first consumes every parked leaf (expand, W += value along the parked path).
template <class Game>
__global__ void __launch_bounds__(NZ_CTA_THREADS, Game::MIN_CTAS)
advance_vl_kernel(const __grid_constant__ View v, void* leaf_out, const void* policy_in, const float* value_in, int leaf_dtype,
                  int policy_dtype) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int TILE = Game::TILE;
  constexpr int VMAX = 8;
  const typename Game::T t;
  const int tile_in_cta = threadIdx.x / TILE;
  const int g = blockIdx.x * (NZ_CTA_THREADS / TILE) + tile_in_cta;
  if (g >= v.G) return;
  const int nwords = (v.A + 31) >> 5;
  const size_t words_bytes = (size_t)v.slab_words_bytes;
  unsigned char* slab = smem_raw + tile_in_cta * v.slab_bytes;
  uint32_t* path = (uint32_t*)slab;
  uint32_t* words = path + v.max_depth;
  uint32_t* state_tmp = words + nwords;
  const size_t scr_bytes = Game::scratch_bytes(v);
  typename Game::Scratch scr_reg, root_reg;
  typename Game::Scratch& scr = Game::SMEM ? *(typename Game::Scratch*)(slab + words_bytes) : scr_reg;
  typename Game::Scratch& rootS = Game::SMEM ? *(typename Game::Scratch*)(slab + words_bytes + scr_bytes) : root_reg;

  uint32_t* ctl = v.ctl + (size_t)g * NZ_CTL_WORDS;
  Slot s;
  slot_load(s, ctl);
  if (s.phase >= NZ_PHASE_MOVE_READY) return;  // waiting for the host, idle, or faulted
  const int V = v.V;
  const size_t nb = (size_t)g * v.P;
  uint32_t* gs_root = v.gstate + (size_t)g * (1 + V) * v.state_words;
  uint32_t* gpath = v.path + (size_t)g * V * v.max_depth;
  uint32_t* pend = v.pend + (size_t)g * V * 2;
  Game::load(rootS, gs_root, v, (int)s.map, t);
  t.sync();
  bool root_dirty = false;
  int n_pend = (int)ctl[NZ_CTL_N_PENDING];

  if (s.phase == NZ_PHASE_LEAF_PENDING) {
    s.phase = NZ_PHASE_READY;
    for (int j = 0; j < n_pend && s.phase == NZ_PHASE_READY; ++j) {
      Game::load(scr, gs_root + (size_t)(1 + j) * v.state_words, v, (int)s.map, t);
      const uint32_t leaf = pend[2 * j];
      const int n_path = (int)pend[2 * j + 1];
      for (int i = t.tl; i < n_path; i += TILE) path[i] = gpath[(size_t)j * v.max_depth + i];
      t.sync();
      uint32_t new_base;
      int new_k;
      const double value = expand<Game>(v, s, ctl, (size_t)g * V + j, nb, leaf, scr, words, policy_in, policy_dtype, value_in, t, new_base, new_k);
      if (s.phase == NZ_PHASE_READY) {
        backup<TILE>(v, nb, path, n_path, value, t, 0, true);  // the visit was counted when the leaf was parked
        s.sims_done += 1;
        s.d_sims += 1;
      }
    }
    n_pend = 0;
  }

  uint32_t parked[VMAX];  // leaves parked by this launch (duplicate test)
  int budget = v.max_sims_per_launch;
  while (s.phase == NZ_PHASE_READY) {
    if (n_pend == 0 && (int)s.sims_done >= v.sims) {
      if (!v.auto_advance) {
        const int K = (int)s.root_K;
        if (K == 0) { s.err |= NZ_ERR_ILLEGAL; s.phase = NZ_PHASE_ERROR; break; }
        const int ch = choose_child<Game>(v, ctl[NZ_CTL_MOVE], ctl[NZ_CTL_UID], g, nb, 1u, K, Game::length(rootS), t);
        if (t.tl == 0) ctl[NZ_CTL_CHOSEN] = (uint32_t)ch;
        s.phase = NZ_PHASE_MOVE_READY;
        break;
      }
      {
        Slot tmp = s;
        commit_move<Game>(v, tmp, ctl, g, nb, rootS, state_tmp, -1, t);
        s = tmp;
      }
      root_dirty = true;
      continue;
    }
    if (budget <= 0 || n_pend >= V || (int)s.sims_done + n_pend >= v.sims) break;
    budget -= 1;
    int depth = 0, levels_left = 0x7fffffff;
    bool paused;
    const bool stateless = Game::NODE_STATE && v.nstate != nullptr;
    if (!stateless) Game::copy(scr, rootS, v, t);
    int last_action = -1;
    uint32_t parent_base = 0u;
    // the root's visit count includes the virtual visits of the leaves parked so far
    const uint32_t node = descend<Game>(v, s, nb, scr, path, 0u, 1u, (int)s.root_K, (int)(s.root_N0 + s.sims_done) + n_pend, depth,
                                        levels_left, &paused, t, &last_action, &parent_base);
    if (s.phase != NZ_PHASE_READY) break;
    leaf_state<Game>(v, s, nb, scr, rootS, path, depth, last_action, parent_base, t);
    Game::settle(scr, v, (int)s.map, t);
    if (Game::terminal(scr)) {
      backup<TILE>(v, nb, path, depth + 1, (double)Game::terminal_value(scr), t);
      s.sims_done += 1;
      s.d_sims += 1;
      s.d_terminal += 1;
      continue;
    }
    bool dup = false;
#pragma unroll
    for (int j = 0; j < VMAX; ++j) dup |= (j < n_pend) && parked[j] == node;
    if (dup) break;  // this leaf is already waiting for the network: nothing was changed, try again next launch
    backup<TILE>(v, nb, path, depth + 1, 0.0, t, 1, false);  // virtual visit
    Game::encode(scr, v, (int)s.map, leaf_out, leaf_dtype, (size_t)g * V + n_pend, t);
    Game::save(scr, gs_root + (size_t)(1 + n_pend) * v.state_words, v, t);
    for (int i = t.tl; i <= depth; i += TILE) gpath[(size_t)n_pend * v.max_depth + i] = path[i];
    if (t.tl == 0) {
      pend[2 * n_pend] = node;
      pend[2 * n_pend + 1] = (uint32_t)(depth + 1);
    }
#pragma unroll
    for (int j = 0; j < VMAX; ++j)
      if (j == n_pend) parked[j] = node;
    n_pend += 1;
  }
  if (s.phase == NZ_PHASE_READY && n_pend > 0) s.phase = NZ_PHASE_LEAF_PENDING;
  if (t.tl == 0) ctl[NZ_CTL_N_PENDING] = (uint32_t)n_pend;
  if (root_dirty) Game::save(rootS, gs_root, v, t);
  slot_store(s, ctl, t.tl);
}

// Manual mode: commit the chosen (or forced) move of every MOVE_READY slot.
template <class Game>
__global__ void __launch_bounds__(NZ_CTA_THREADS) commit_kernel(const __grid_constant__ View v, const int32_t* actions) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int TILE = Game::TILE;
  const typename Game::T t;
  const int tile_in_cta = threadIdx.x / TILE;
  const int g = blockIdx.x * (NZ_CTA_THREADS / TILE) + tile_in_cta;
  if (g >= v.G) return;
  const size_t scr_bytes = Game::scratch_bytes(v);
  const size_t tmp_bytes = ((size_t)v.state_words * 4 + 15) & ~(size_t)15;
  unsigned char* slab = smem_raw + tile_in_cta * (scr_bytes + tmp_bytes);
  uint32_t* state_tmp = (uint32_t*)slab;
  typename Game::Scratch root_reg;
  typename Game::Scratch& rootS = Game::SMEM ? *(typename Game::Scratch*)(slab + tmp_bytes) : root_reg;
  uint32_t* ctl = v.ctl + (size_t)g * NZ_CTL_WORDS;
  Slot s;
  slot_load(s, ctl);
  if (s.phase != NZ_PHASE_MOVE_READY) return;
  uint32_t* gs_root = v.gstate + (size_t)g * (1 + v.V) * v.state_words;
  Game::load(rootS, gs_root, v, (int)s.map, t);
  t.sync();
  commit_move<Game>(v, s, ctl, g, (size_t)g * v.P, rootS, state_tmp, actions ? actions[g] : -1, t);
  Game::save(rootS, gs_root, v, t);
  slot_store(s, ctl, t.tl);
}

template <class Game>
__global__ void __launch_bounds__(NZ_CTA_THREADS) reset_kernel(const __grid_constant__ View v) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int TILE = Game::TILE;
  const typename Game::T t;
  const int tile_in_cta = threadIdx.x / TILE;
  const int g = blockIdx.x * (NZ_CTA_THREADS / TILE) + tile_in_cta;
  if (g >= v.G) return;
  const size_t scr_bytes = Game::scratch_bytes(v);
  typename Game::Scratch root_reg;
  typename Game::Scratch& rootS = Game::SMEM ? *(typename Game::Scratch*)(smem_raw + tile_in_cta * scr_bytes) : root_reg;
  uint32_t* ctl = v.ctl + (size_t)g * NZ_CTL_WORDS;
  const uint32_t map = ctl[NZ_CTL_MAP];  // the host assigns scenario maps before reset
  t.sync();
  for (int i = t.tl; i < NZ_CTL_WORDS; i += TILE) ctl[i] = 0u;
  t.sync();
  Slot s = {};
  s.phase = NZ_PHASE_READY;
  s.pool_top = (uint32_t)v.g0;
  s.map = map;
  const size_t nb = (size_t)g * v.P;
  if (t.tl == 0) {
    clear_node(v, nb);
    ctl[NZ_CTL_MAP] = map;
    ctl[NZ_CTL_UID] = (uint32_t)g;
  }
  Game::reset(rootS, v, (int)map, t);
  t.sync();
  Game::save(rootS, v.gstate + (size_t)g * (1 + v.V) * v.state_words, v, t);
  slot_store(s, ctl, t.tl);
}

}  // namespace nz
