// Warp-per-game PUCT search over structure-of-arrays node pools (sm_100a).
//
// One warp owns one game slot for the whole launch: it consumes the pending network row (expand +
// backup), then runs simulations — select with a shuffle arg-max, game step, terminal backup —
// until a leaf needs the network, and commits finished moves itself in auto mode.  Per-game
// sequencing is exactly the reference's (one simulation in flight per game, Search/Explorer.py:49-62);
// the batch comes from the number of games, so results are bit-identical to the reference.
#pragma once
#include "common.cuh"

namespace nz {

#define NZ_REC_HDR 12

struct Slot {  // warp-uniform registers mirroring the ctl words
  uint32_t phase, root, pool_top, sims_done, move, uid, games_done, path_len, err, leaf, chosen, noised;
  uint32_t n_sims, n_levels, n_scanned, n_expand, n_created, n_moves, n_terminal, map;
};

__device__ __forceinline__ void slot_load(Slot& s, const uint32_t* ctl, int lane) {
  uint32_t w = ctl[lane];
  s.phase = __shfl_sync(NZ_FULL, w, NZ_CTL_PHASE);
  s.root = __shfl_sync(NZ_FULL, w, NZ_CTL_ROOT);
  s.pool_top = __shfl_sync(NZ_FULL, w, NZ_CTL_POOL_TOP);
  s.sims_done = __shfl_sync(NZ_FULL, w, NZ_CTL_SIMS_DONE);
  s.move = __shfl_sync(NZ_FULL, w, NZ_CTL_MOVE);
  s.uid = __shfl_sync(NZ_FULL, w, NZ_CTL_UID);
  s.games_done = __shfl_sync(NZ_FULL, w, NZ_CTL_GAMES_DONE);
  s.path_len = __shfl_sync(NZ_FULL, w, NZ_CTL_PATH_LEN);
  s.err = __shfl_sync(NZ_FULL, w, NZ_CTL_ERROR);
  s.leaf = __shfl_sync(NZ_FULL, w, NZ_CTL_LEAF);
  s.chosen = __shfl_sync(NZ_FULL, w, NZ_CTL_CHOSEN);
  s.noised = __shfl_sync(NZ_FULL, w, NZ_CTL_NOISED);
  s.n_sims = __shfl_sync(NZ_FULL, w, NZ_CTL_N_SIMS);
  s.n_levels = __shfl_sync(NZ_FULL, w, NZ_CTL_N_LEVELS);
  s.n_scanned = __shfl_sync(NZ_FULL, w, NZ_CTL_N_SCANNED);
  s.n_expand = __shfl_sync(NZ_FULL, w, NZ_CTL_N_EXPAND);
  s.n_created = __shfl_sync(NZ_FULL, w, NZ_CTL_N_CREATED);
  s.n_moves = __shfl_sync(NZ_FULL, w, NZ_CTL_N_MOVES);
  s.n_terminal = __shfl_sync(NZ_FULL, w, NZ_CTL_N_TERMINAL);
  s.map = __shfl_sync(NZ_FULL, w, NZ_CTL_MAP);
}

__device__ __forceinline__ void slot_store(const Slot& s, uint32_t* ctl, int lane) {
  uint32_t w = 0;
  switch (lane) {
    case NZ_CTL_PHASE: w = s.phase; break;
    case NZ_CTL_ROOT: w = s.root; break;
    case NZ_CTL_POOL_TOP: w = s.pool_top; break;
    case NZ_CTL_SIMS_DONE: w = s.sims_done; break;
    case NZ_CTL_MOVE: w = s.move; break;
    case NZ_CTL_UID: w = s.uid; break;
    case NZ_CTL_GAMES_DONE: w = s.games_done; break;
    case NZ_CTL_PATH_LEN: w = s.path_len; break;
    case NZ_CTL_ERROR: w = s.err; break;
    case NZ_CTL_LEAF: w = s.leaf; break;
    case NZ_CTL_CHOSEN: w = s.chosen; break;
    case NZ_CTL_NOISED: w = s.noised; break;
    case NZ_CTL_N_SIMS: w = s.n_sims; break;
    case NZ_CTL_N_LEVELS: w = s.n_levels; break;
    case NZ_CTL_N_SCANNED: w = s.n_scanned; break;
    case NZ_CTL_N_EXPAND: w = s.n_expand; break;
    case NZ_CTL_N_CREATED: w = s.n_created; break;
    case NZ_CTL_N_MOVES: w = s.n_moves; break;
    case NZ_CTL_N_TERMINAL: w = s.n_terminal; break;
    case NZ_CTL_MAP: w = s.map; break;
    default: break;
  }
  ctl[lane] = w;
}

// exploration bias c(N) = log((N + base + 1) / base) + init  (Explorer.py:103-108).  The table is
// computed on the host with the same libm `log` the reference uses; beyond it the device log is
// used and the slot is flagged.
__device__ __forceinline__ double bias_of(const View& v, int n, uint32_t& err) {
  if (n < v.ctable_len) return v.ctable[n];
  err |= NZ_ERR_CTABLE;
  return log(((double)n + v.pb_c_base + 1.0) / v.pb_c_base) + v.pb_c_init;
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
__device__ __forceinline__ void backup(const View& v, size_t nb, const uint32_t* path, int n_path, double value,
                                       int lane) {
  for (int i = lane; i < n_path; i += 32) {
    size_t idx = nb + path[i];
    v.node_N[idx] += 1;
    v.node_W[idx] = __dadd_rn(v.node_W[idx], value);
  }
  __syncwarp();
}

// iterate the set bits of a warp-shared bit set; body(a, rank) runs on the lane that owns bit a
template <class F>
__device__ __forceinline__ int for_each_valid(const uint32_t* words, int nwords, int lane, F body) {
  int running = 0;
  for (int wb = 0; wb < nwords; wb += 32) {
    uint32_t myw = (wb + lane < nwords) ? words[wb + lane] : 0u;
    uint32_t nzm = __ballot_sync(NZ_FULL, myw != 0u);
    while (nzm) {
      int j = __ffs(nzm) - 1;
      nzm &= nzm - 1;
      uint32_t bits = __shfl_sync(NZ_FULL, myw, j);
      if ((bits >> lane) & 1u) body((wb + j) * 32 + lane, running + __popc(bits & ((1u << lane) - 1u)));
      running += __popc(bits);
    }
  }
  return running;
}

// ---- expand (Explorer.evaluate, Explorer.py:137-181) --------------------------------------------
// Returns the network value; creates one child per legal action, ascending action order.
template <class Game>
__device__ __forceinline__ double expand(const View& v, Slot& s, int g, size_t nb, uint32_t leaf,
                                         typename Game::Scratch& scr, uint32_t* words, const void* policy_in,
                                         int policy_dtype, const float* value_in, int lane) {
  using PriorT = typename Game::PriorT;
  const int A = v.A, nwords = (A + 31) >> 5;
  double value = (double)value_in[g];  // predicted_value.item() (Explorer.py:162)
  Game::legal(scr, v, (int)s.map, words, lane);
  const size_t prow = (size_t)g * A;

  // softmax over ALL actions when the network emits logits (Explorer.py:152/159), in f32
  float smax = 0.f, ssum = 1.f;
  if (!v.policy_is_prob) {
    float m = -INFINITY;
    for (int a = lane; a < A; a += 32) m = fmaxf(m, load_policy(policy_in, policy_dtype, prow + a));
    m = warp_max(m);
    float e = 0.f;
    for (int a = lane; a < A; a += 32) e += __expf(load_policy(policy_in, policy_dtype, prow + a) - m);
    smax = m;
    ssum = warp_sum(e);
  }
  auto prob_of = [&](int a) -> float {
    float x = load_policy(policy_in, policy_dtype, prow + a);
    return v.policy_is_prob ? x : __fdiv_rn(__expf(x - smax), ssum);
  };

  // total over the legal actions (np.sum(probs), Explorer.py:169)
  PriorT part = (PriorT)0;
  int K = for_each_valid(words, nwords, lane, [&](int a, int) { part += (PriorT)prob_of(a); });
  PriorT total = warp_sum(part);
  bool uniform = (total == (PriorT)0);  // "network predicted zero valid actions" workaround (:171-174)
  if (uniform) total = (PriorT)K;
  if (K == 0) return value;  // no legal action: node stays childless and is re-evaluated each visit
  if (K > v.max_children || s.pool_top + (uint32_t)K > (uint32_t)v.P) {
    s.err |= NZ_ERR_POOL_FULL;
    s.phase = NZ_PHASE_ERROR;
    return value;
  }
  const uint32_t base = s.pool_top;
  s.pool_top += (uint32_t)K;
  PriorT* prior = (PriorT*)v.node_prior;
  for_each_valid(words, nwords, lane, [&](int a, int rank) {
    size_t idx = nb + base + rank;
    PriorT p = uniform ? (PriorT)1 : (PriorT)prob_of(a);
    v.node_N[idx] = 0;
    v.node_W[idx] = 0.0;
    prior[idx] = p / total;  // IEEE division, f64 or f32 like the reference's numpy scalar
    v.node_link[idx] = make_uint2(0u, (uint32_t)a << 16);
  });
  if (lane == 0) {
    uint2 lk = v.node_link[nb + leaf];
    v.node_link[nb + leaf] = make_uint2(base, (lk.y & 0xffff0000u) | (uint32_t)K);
  }
  s.n_expand += 1;
  s.n_created += (uint32_t)K;
  __syncwarp();
  return value;
}

// ---- root noise (Explorer.add_exploration_noise, Explorer.py:201-210) ----------------------------
template <class Game>
__device__ __forceinline__ void add_root_noise(const View& v, Slot& s, int g, size_t nb, int lane) {
  using PriorT = typename Game::PriorT;
  uint2 lk = v.node_link[nb + s.root];
  int K = (int)(lk.y & 0xffffu);
  s.noised = 0;
  if (K == 0) return;
  const double frac = v.noise_frac;
  PriorT* prior = (PriorT*)v.node_prior;
  for (int i = lane; i < K; i += 32) {
    double n;
    if (v.tape_moves > 0) {
      int m = min((int)s.move, v.tape_moves - 1);
      n = (i < v.tape_width) ? v.gamma_tape[((size_t)g * v.tape_moves + m) * v.tape_width + i] : 0.0;
    } else {
      n = philox_gamma(v.seed ^ ((unsigned long long)s.uid * 0x9E3779B97F4A7C15ull), s.move, 1u, (uint32_t)i,
                       v.noise_alpha, v.noise_beta);
    }
    size_t idx = nb + lk.x + i;
    double nf = __dmul_rn(n, frac);
    if (Game::PRIOR_F64) {
      double p = (double)prior[idx];
      prior[idx] = (PriorT)__dadd_rn(__dmul_rn(p, 1.0 - frac), nf);
    } else {
      // np.float32 * python float -> float32 ; + np.float64 -> float64 (kept in the side array)
      float p = (float)prior[idx];
      float scaled = __fmul_rn(p, __double2float_rn(1.0 - frac));
      v.root_prior64[(size_t)g * v.max_children + i] = __dadd_rn((double)scaled, nf);
    }
  }
  s.noised = 1;
  __syncwarp();
}

// ---- final action choice (Explorer.select_action, Explorer.py:70-97) -> child index --------------
template <class Game>
__device__ __forceinline__ int choose_child(const View& v, const Slot& s, int g, size_t nb, uint32_t base, int K,
                                            int game_length, int lane) {
  // max_action (:183-185): python max with key -> first maximum -> LOWEST action on ties
  long long key = -1;
  int idx = -1;
  for (int i = lane; i < K; i += 32) {
    long long n = v.node_N[nb + base + i];
    if (idx < 0 || n > key) { key = n; idx = i; }
  }
  warp_argmax_lo(key, idx);
  if (!v.training) return idx;

  double u0, u1, u2;
  if (v.tape_moves > 0) {
    int m = min((int)s.move, v.tape_moves - 1);
    const double* row = v.unif_tape + ((size_t)g * v.tape_moves + m) * 3;
    u0 = row[0]; u1 = row[1]; u2 = row[2];
  } else {
    uint32_t r[4], r2[4];
    unsigned long long k = v.seed ^ ((unsigned long long)s.uid * 0x9E3779B97F4A7C15ull);
    Philox::gen(k, s.move, 2u, 0u, 0u, r);
    Philox::gen(k, s.move, 2u, 1u, 0u, r2);
    u0 = u01(r[0], r[1]); u1 = u01(r[2], r[3]); u2 = u01(r2[0], r2[1]);
  }
  int mode = 0;  // 0 max, 1 softmax over visit counts, 2 uniform over legal actions
  if (game_length < v.n_softmax_moves) mode = 1;
  else if (u0 < v.eps_softmax) mode = 1;
  else if (u1 < v.eps_random) mode = 2;
  if (mode == 0) return idx;

  // sequential on lane 0, in the reference's order of operations (rare: a few % of moves)
  int pick = 0;
  if (lane == 0) {
    if (mode == 1) {
      // softmax_action (:187-199): scipy softmax of the raw counts, renormalise, np.random.choice
      double mx = (double)key, sum = 0.0;
      for (int i = 0; i < K; ++i) sum += exp((double)v.node_N[nb + base + i] - mx);
      double sum2 = 0.0;
      for (int i = 0; i < K; ++i) sum2 += exp((double)v.node_N[nb + base + i] - mx) / sum;
      double tot = 0.0;
      for (int i = 0; i < K; ++i) tot += (exp((double)v.node_N[nb + base + i] - mx) / sum) / sum2;
      double cdf = 0.0;
      for (int i = 0; i < K; ++i) {
        cdf += (exp((double)v.node_N[nb + base + i] - mx) / sum) / sum2;
        if (cdf / tot <= u2) pick = i + 1;
      }
    } else {
      double each = 1.0 / (double)K, tot = 0.0, cdf = 0.0;
      for (int i = 0; i < K; ++i) tot += each;
      for (int i = 0; i < K; ++i) {
        cdf += each;
        if (cdf / tot <= u2) pick = i + 1;
      }
    }
    pick = min(pick, K - 1);
  }
  return __shfl_sync(NZ_FULL, pick, 0);
}

// ---- move record ---------------------------------------------------------------------------------
template <class Game>
__device__ __forceinline__ void write_record(const View& v, Slot& s, int g, size_t nb, uint32_t base, int K,
                                             int child, int action, int player, const uint32_t* state_before,
                                             bool game_end, int tv, int length_after, int lane) {
  using PriorT = typename Game::PriorT;
  const int SW = v.state_words;
  const int len = NZ_REC_HDR + SW + 2 * K + (v.record_detail ? 4 * K : 0);
  uint32_t off = 0;
  if (lane == 0) off = atomicAdd(v.arena_top, (uint32_t)len);
  off = __shfl_sync(NZ_FULL, off, 0);
  if (off + (uint32_t)len > (uint32_t)v.arena_words) {
    if (lane == 0) atomicAdd(v.arena_top + 1, 1u);
    s.err |= NZ_ERR_ARENA_FULL;
    return;
  }
  uint32_t* r = v.arena + off;
  const int rootN = v.node_N[nb + s.root];
  const double rootW = v.node_W[nb + s.root];
  uint32_t dummy = 0;
  const double bias = bias_of(v, rootN, dummy);
  if (lane == 0) {
    r[0] = (uint32_t)len;
    r[1] = s.uid;
    r[2] = s.move | ((uint32_t)K << 16);
    r[3] = (uint32_t)action | ((uint32_t)player << 16) |
           (((v.record_detail ? 1u : 0u) | (game_end ? 2u : 0u) | ((uint32_t)(tv + 1) << 2)) << 24);
    r[4] = (uint32_t)rootN;
    r[5] = (uint32_t)__double_as_longlong(rootW);
    r[6] = (uint32_t)(__double_as_longlong(rootW) >> 32);
    r[7] = (uint32_t)g;
    r[8] = (uint32_t)__double_as_longlong(bias);
    r[9] = (uint32_t)(__double_as_longlong(bias) >> 32);
    r[10] = (uint32_t)length_after;
    r[11] = (uint32_t)child;
  }
  for (int i = lane; i < SW; i += 32) r[NZ_REC_HDR + i] = state_before[i];
  uint32_t* c = r + NZ_REC_HDR + SW;
  const PriorT* prior = (const PriorT*)v.node_prior;
  for (int i = lane; i < K; i += 32) {
    size_t idx = nb + base + i;
    c[2 * i] = v.node_link[idx].y >> 16;
    c[2 * i + 1] = (uint32_t)v.node_N[idx];
    if (v.record_detail) {
      long long w = __double_as_longlong(v.node_W[idx]);
      double pd = (!Game::PRIOR_F64 && s.noised) ? v.root_prior64[(size_t)g * v.max_children + i] : (double)prior[idx];
      long long p = __double_as_longlong(pd);
      uint32_t* d = c + 2 * K + 4 * i;
      d[0] = (uint32_t)w; d[1] = (uint32_t)(w >> 32); d[2] = (uint32_t)p; d[3] = (uint32_t)(p >> 32);
    }
  }
}

// ---- commit a move: Training/Gamer.py:74-79 (+ restart, Gamer.play_game called again) -------------
template <class Game>
__device__ __forceinline__ void commit_move(const View& v, Slot& s, int g, size_t nb, typename Game::Scratch& rootS,
                                            uint32_t* state_tmp, int forced_action, int lane) {
  uint2 lk = v.node_link[nb + s.root];
  const int K = (int)(lk.y & 0xffffu);
  const uint32_t base = lk.x;
  if (K == 0) {  // the reference would raise on max() of an empty sequence
    s.err |= NZ_ERR_ILLEGAL;
    s.phase = NZ_PHASE_ERROR;
    return;
  }
  int child;
  if (forced_action >= 0) {
    int found = -1;
    for (int i = lane; i < K; i += 32)
      if ((int)(v.node_link[nb + base + i].y >> 16) == forced_action) found = i;
    for (int off = 16; off > 0; off >>= 1) found = max(found, __shfl_xor_sync(NZ_FULL, found, off));
    if (found < 0) {
      s.err |= NZ_ERR_ILLEGAL;
      s.phase = NZ_PHASE_ERROR;
      return;
    }
    child = found;
  } else if (s.phase == NZ_PHASE_MOVE_READY) {
    child = (int)s.chosen;
  } else {
    child = choose_child<Game>(v, s, g, nb, base, K, Game::length(rootS), lane);
  }
  const int action = (int)(v.node_link[nb + base + child].y >> 16);
  const int player = Game::to_play(rootS);
  Game::save(rootS, state_tmp, lane);  // state before the move, for the record
  __syncwarp();
  Game::step(rootS, v, (int)s.map, action, lane);
  const bool over = Game::terminal(rootS);
  write_record<Game>(v, s, g, nb, base, K, child, action, player, state_tmp, over, over ? Game::terminal_value(rootS) : 0,
                     Game::length(rootS), lane);
  s.n_moves += 1;
  s.move += 1;
  s.sims_done = 0;
  s.root = base + (uint32_t)child;  // keep_subtree: the chosen child becomes the root
  s.noised = 0;
  s.phase = NZ_PHASE_READY;
  if (over) {
    s.games_done += 1;
    if (!v.auto_advance || (v.games_per_slot > 0 && (int)s.games_done >= v.games_per_slot)) {
      s.phase = NZ_PHASE_IDLE;
    } else {  // fresh game in the same slot: Node(0) root, new game object
      s.uid += (uint32_t)v.G;
      s.move = 0;
      s.root = 0;
      s.pool_top = 1;
      Game::reset(rootS, v, (int)s.map, lane);
      if (lane == 0) {
        v.node_N[nb] = 0;
        v.node_W[nb] = 0.0;
        v.node_link[nb] = make_uint2(0u, 0u);
      }
      __syncwarp();
    }
  } else if (v.training) {
    add_root_noise<Game>(v, s, g, nb, lane);
  }
}

// ---- one simulation's descent (Explorer.py:54-58 + select_child :99-101) --------------------------
// Returns the leaf node; path[0..depth] filled; scratch stepped to the leaf.
template <class Game>
__device__ __forceinline__ uint32_t descend(const View& v, Slot& s, int g, size_t nb, typename Game::Scratch& scr,
                                            uint32_t* path, int& depth, int lane) {
  using PriorT = typename Game::PriorT;
  const PriorT* prior = (const PriorT*)v.node_prior;
  uint32_t node = s.root;
  uint2 link = v.node_link[nb + node];
  int Np = v.node_N[nb + node];
  depth = 0;
  if (lane == 0) path[0] = node;
  while ((link.y & 0xffffu) != 0u) {
    const int K = (int)(link.y & 0xffffu);
    const uint32_t base = link.x;
    if (depth + 1 >= v.max_depth) {
      s.err |= NZ_ERR_DEPTH;
      s.phase = NZ_PHASE_ERROR;
      break;
    }
    const bool flip = Game::to_play(scr) == 2;  // literal `parent.to_play == 2` (Explorer.py:124)
    const double c = bias_of(v, Np, s.err);
    const double sq = __dsqrt_rn((double)Np);
    const bool noised_root = (!Game::PRIOR_F64) && depth == 0 && s.noised;
    double best_s = 0.0;
    int best_i = -1, best_n = 0;
    uint2 best_lk = make_uint2(0u, 0u);
    for (int i = lane; i < K; i += 32) {
      const size_t idx = nb + base + i;
      const int n = v.node_N[idx];
      const double w = v.node_W[idx];
      const uint2 lk = v.node_link[idx];
      const double u = __ddiv_rn(sq, (double)(n + 1));
      double q = (n == 0) ? 0.0 : __ddiv_rn(w, (double)n);
      if (flip) q = -q;
      q = __dmul_rn(q, v.value_factor);
      double sc;
      if (Game::PRIOR_F64) sc = score_f64((double)prior[idx], u, c, q);
      else if (noised_root) sc = score_f64(v.root_prior64[(size_t)g * v.max_children + i], u, c, q);
      else sc = score_f32((float)prior[idx], u, c, q);
      if (best_i < 0 || sc >= best_s) { best_s = sc; best_i = i; best_n = n; best_lk = lk; }
    }
    warp_argmax_hi(best_s, best_i);
    const int owner = best_i & 31;
    Np = __shfl_sync(NZ_FULL, best_n, owner);
    link.x = __shfl_sync(NZ_FULL, best_lk.x, owner);
    link.y = __shfl_sync(NZ_FULL, best_lk.y, owner);
    Game::step(scr, v, (int)s.map, (int)(link.y >> 16), lane);
    node = base + (uint32_t)best_i;
    depth += 1;
    if (lane == 0) path[depth] = node;
    s.n_levels += 1;
    s.n_scanned += (uint32_t)K;
  }
  __syncwarp();
  return node;
}

template <class Game>
__global__ void __launch_bounds__(32 * NZ_WARPS_PER_CTA)
advance_kernel(const View v, void* leaf_out, const void* policy_in, const float* value_in, int leaf_dtype,
               int policy_dtype) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = blockIdx.x * NZ_WARPS_PER_CTA + warp;
  if (g >= v.G) return;
  // per-warp slab: path[max_depth] | mask words | state tmp | scratch | root scratch
  const int nwords = (v.A + 31) >> 5;
  const size_t slab_words = (size_t)v.max_depth + nwords + v.state_words;
  const size_t scr_bytes = (sizeof(typename Game::Scratch) + 15) & ~(size_t)15;
  const size_t slab_bytes = ((slab_words * 4 + 15) & ~(size_t)15) + 2 * scr_bytes;
  unsigned char* slab = smem_raw + warp * slab_bytes;
  uint32_t* path = (uint32_t*)slab;
  uint32_t* words = path + v.max_depth;
  uint32_t* state_tmp = words + nwords;
  // small games (TTT) keep both game states in registers; SCS stages them in shared memory
  typename Game::Scratch scr_reg, root_reg;
  typename Game::Scratch& scr =
      Game::SMEM ? *(typename Game::Scratch*)(slab + ((slab_words * 4 + 15) & ~(size_t)15)) : scr_reg;
  typename Game::Scratch& rootS =
      Game::SMEM ? *(typename Game::Scratch*)(slab + ((slab_words * 4 + 15) & ~(size_t)15) + scr_bytes) : root_reg;

  uint32_t* ctl = v.ctl + (size_t)g * NZ_CTL_WORDS;
  Slot s;
  slot_load(s, ctl, lane);
  if (s.phase >= NZ_PHASE_MOVE_READY) return;  // waiting for the host, idle, or faulted
  const size_t nb = (size_t)g * v.P;
  uint32_t* gs_root = v.gstate + (size_t)g * 2 * v.state_words;
  uint32_t* gs_leaf = gs_root + v.state_words;
  Game::load(rootS, gs_root, lane);
  __syncwarp();
  bool root_dirty = false;

  if (s.phase == NZ_PHASE_LEAF_PENDING) {
    Game::load(scr, gs_leaf, lane);
    const int n_path = (int)s.path_len;
    for (int i = lane; i < n_path; i += 32) path[i] = v.path[(size_t)g * v.max_depth + i];
    __syncwarp();
    s.phase = NZ_PHASE_READY;
    const double value = expand<Game>(v, s, g, nb, s.leaf, scr, words, policy_in, policy_dtype, value_in, lane);
    if (s.phase == NZ_PHASE_READY) {
      backup(v, nb, path, n_path, value, lane);
      s.sims_done += 1;
      s.n_sims += 1;
    }
  }

  int budget = v.max_sims_per_launch;
  while (s.phase == NZ_PHASE_READY) {
    if ((int)s.sims_done >= v.sims) {
      if (!v.auto_advance) {
        uint2 lk = v.node_link[nb + s.root];
        const int K = (int)(lk.y & 0xffffu);
        if (K == 0) { s.err |= NZ_ERR_ILLEGAL; s.phase = NZ_PHASE_ERROR; break; }
        s.chosen = (uint32_t)choose_child<Game>(v, s, g, nb, lk.x, K, Game::length(rootS), lane);
        s.phase = NZ_PHASE_MOVE_READY;
        break;
      }
      commit_move<Game>(v, s, g, nb, rootS, state_tmp, -1, lane);
      root_dirty = true;
      continue;
    }
    if (budget <= 0) break;
    budget -= 1;
    Game::copy(scr, rootS, lane);  // game.shallow_clone() (Explorer.py:51)
    int depth;
    const uint32_t node = descend<Game>(v, s, g, nb, scr, path, depth, lane);
    if (s.phase != NZ_PHASE_READY) break;
    if (Game::terminal(scr)) {  // Explorer.py:140-142: terminal leaves return the game's value
      backup(v, nb, path, depth + 1, (double)Game::terminal_value(scr), lane);
      s.sims_done += 1;
      s.n_sims += 1;
      s.n_terminal += 1;
      continue;
    }
    // non-terminal leaf: hand its encoded state to the network (Explorer.py:145)
    Game::encode(scr, v, (int)s.map, leaf_out, leaf_dtype, (size_t)g, lane);
    Game::save(scr, gs_leaf, lane);
    for (int i = lane; i <= depth; i += 32) v.path[(size_t)g * v.max_depth + i] = path[i];
    s.path_len = (uint32_t)(depth + 1);
    s.leaf = node;
    s.phase = NZ_PHASE_LEAF_PENDING;
  }
  if (root_dirty) Game::save(rootS, gs_root, lane);
  slot_store(s, ctl, lane);
}

// Manual mode: commit the chosen (or forced) move of every MOVE_READY slot.
template <class Game>
__global__ void __launch_bounds__(32 * NZ_WARPS_PER_CTA) commit_kernel(const View v, const int32_t* actions) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = blockIdx.x * NZ_WARPS_PER_CTA + warp;
  if (g >= v.G) return;
  const size_t scr_bytes = (sizeof(typename Game::Scratch) + 15) & ~(size_t)15;
  const size_t tmp_bytes = ((size_t)v.state_words * 4 + 15) & ~(size_t)15;
  unsigned char* slab = smem_raw + warp * (scr_bytes + tmp_bytes);
  uint32_t* state_tmp = (uint32_t*)slab;
  typename Game::Scratch root_reg;
  typename Game::Scratch& rootS = Game::SMEM ? *(typename Game::Scratch*)(slab + tmp_bytes) : root_reg;
  uint32_t* ctl = v.ctl + (size_t)g * NZ_CTL_WORDS;
  Slot s;
  slot_load(s, ctl, lane);
  if (s.phase != NZ_PHASE_MOVE_READY) return;
  uint32_t* gs_root = v.gstate + (size_t)g * 2 * v.state_words;
  Game::load(rootS, gs_root, lane);
  __syncwarp();
  commit_move<Game>(v, s, g, (size_t)g * v.P, rootS, state_tmp, actions ? actions[g] : -1, lane);
  Game::save(rootS, gs_root, lane);
  slot_store(s, ctl, lane);
}

template <class Game>
__global__ void reset_kernel(const View v, int n_maps) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = blockIdx.x * NZ_WARPS_PER_CTA + warp;
  if (g >= v.G) return;
  const size_t scr_bytes = (sizeof(typename Game::Scratch) + 15) & ~(size_t)15;
  typename Game::Scratch root_reg;
  typename Game::Scratch& rootS = Game::SMEM ? *(typename Game::Scratch*)(smem_raw + warp * scr_bytes) : root_reg;
  uint32_t* ctl = v.ctl + (size_t)g * NZ_CTL_WORDS;
  const uint32_t map = ctl[NZ_CTL_MAP];  // the host assigns scenario maps before reset
  __syncwarp();
  Slot s = {};
  s.phase = NZ_PHASE_READY;
  s.pool_top = 1;
  s.uid = (uint32_t)g;
  s.map = map;
  const size_t nb = (size_t)g * v.P;
  if (lane == 0) {
    v.node_N[nb] = 0;
    v.node_W[nb] = 0.0;
    v.node_link[nb] = make_uint2(0u, 0u);
  }
  Game::reset(rootS, v, (int)map, lane);
  __syncwarp();
  Game::save(rootS, v.gstate + (size_t)g * 2 * v.state_words, lane);
  slot_store(s, ctl, lane);
}

}  // namespace nz
