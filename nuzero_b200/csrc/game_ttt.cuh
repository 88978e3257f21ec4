// Tic-Tac-Toe on the device: the whole game is one 32-bit word held in a register.
// Behaviour follows Games/Tic_Tac_Toe/tic_tac_toe.py (players 1/2, state = [P1 stones, P2 stones]).
#pragma once
#include "common.cuh"

namespace nz {

struct TTT {
  // bits 0-8 P1 stones, 9-17 P2 stones, 18-21 length, 22 terminal, 23-24 terminal_value+1
  static constexpr int STATE_WORDS = 1;
  static constexpr int A = 9, PLANES = 1, R = 3, CC = 3, C = 2;
  static constexpr bool PRIOR_F64 = true;  // float64 mask -> float64 priors (tic_tac_toe.py:122)
  static constexpr bool NODE_STATE = false;  // the whole game is one register: replaying the path costs a few bit operations
  static constexpr bool MAY_FLIP = true;
  using PriorT = double;
#ifndef NZ_TTT_TILE
#define NZ_TTT_TILE 8
#endif
#ifndef NZ_TTT_MIN_CTAS
#define NZ_TTT_MIN_CTAS 7
#endif
  static constexpr int TILE = NZ_TTT_TILE;          // 8 lanes per game: 4 games share a warp
  using T = Tl<TILE>;
  static constexpr int MIN_CTAS = NZ_TTT_MIN_CTAS;  // <= 72 registers: 16384 games resident in one wave

  static constexpr bool SMEM = false;  // per-lane register copy, every lane computes the same word
  struct Scratch { uint32_t s; };
  __host__ __device__ static size_t scratch_bytes(const View&) { return 16; }
  __device__ static __forceinline__ void copy(Scratch& d, const Scratch& s, const View&, const T&) { d.s = s.s; }

  __device__ static __forceinline__ uint32_t initial() { return 1u << 23; }  // value 0 -> code 1

  __device__ static __forceinline__ void load(Scratch& sc, const uint32_t* g, const View&, int, const T&) { sc.s = g[0]; }
  __device__ static __forceinline__ void save(const Scratch& sc, uint32_t* g, const View&, const T& t) {
    if (t.tl == 0) g[0] = sc.s;
  }
  __device__ static __forceinline__ void reset(Scratch& sc, const View&, int, const T&) { sc.s = initial(); }

  __device__ static __forceinline__ int length(const Scratch& sc) { return (sc.s >> 18) & 15; }
  // get_current_player(): length % 2 + 1 (tic_tac_toe.py:165, also after the terminal move)
  __device__ static __forceinline__ int to_play(const Scratch& sc) { return (length(sc) & 1) + 1; }
  __device__ static __forceinline__ bool terminal(const Scratch& sc) { return (sc.s >> 22) & 1; }
  __device__ static __forceinline__ int terminal_value(const Scratch& sc) { return (int)((sc.s >> 23) & 3) - 1; }

  __device__ static __forceinline__ bool has_line(uint32_t b) {
    // three in a row <=> some cell has both neighbours along one of the 4 directions:
    // rows (shift 1, middle column), columns (shift 3), diagonals (shift 4 / 2 through the centre)
    uint32_t rows = b & (b >> 1) & (b >> 2) & 0111u;  // cells 0,3,6 start a row
    uint32_t cols = b & (b >> 3) & (b >> 6) & 0007u;  // cells 0,1,2 start a column
    uint32_t d1 = b & (b >> 4) & (b >> 8) & 0001u;    // 0,4,8
    uint32_t d2 = b & (b >> 2) & (b >> 4) & 0004u;    // 2,4,6
    return (rows | cols | d1 | d2) != 0u;
  }

  // Descent step: place the mover's stone, length += 1.  The terminal test is deferred to settle():
  // the search only descends through expanded (hence non-terminal) nodes.
  __device__ static __forceinline__ void step_descend(Scratch& sc, const View&, int, int action, const T&) {
    uint32_t s = sc.s;
    uint32_t me = (s >> 18) & 1u;
    sc.s = (s | (1u << (action + 9 * me))) + (1u << 18);
  }
  // check_terminal (tic_tac_toe.py:198-262): P1 lines are tested first, a full board ends the game
  // with the value found so far.
  __device__ static __forceinline__ void settle(Scratch& sc, const View&, int, const T&) {
    uint32_t s = sc.s & 0x3fffffu;
    uint32_t len = (s >> 18) & 15;
    uint32_t p1 = s & 0x1ffu, p2 = (s >> 9) & 0x1ffu;
    int v = has_line(p1) ? 1 : (has_line(p2) ? -1 : 0);
    bool done = (v != 0) || (len == 9);
    sc.s = s | ((done ? 1u : 0u) << 22) | ((uint32_t)(v + 1) << 23);
  }
  // step (tic_tac_toe.py:161-167)
  __device__ static __forceinline__ void step(Scratch& sc, const View& v, int m, int action, const T& t) {
    step_descend(sc, v, m, action, t);
    settle(sc, v, m, t);
  }

  // possible_actions (tic_tac_toe.py:121-129) as a bit set in per-tile shared memory
  __device__ static __forceinline__ void legal(const Scratch& sc, const View&, int, uint32_t* words, const T& t) {
    uint32_t occ = (sc.s | (sc.s >> 9)) & 0x1ffu;
    if (t.tl == 0) words[0] = (~occ) & 0x1ffu;
    t.sync();
  }

  // generate_state_image (tic_tac_toe.py:135-159): 18 values, plane-major
  __device__ static __forceinline__ void encode(const Scratch& sc, const View&, int, void* out, int dtype, size_t row,
                                                const T& t) {
    for (int i = t.tl; i < 18; i += TILE) store_leaf(out, dtype, row * 18 + i, ((sc.s >> i) & 1u) ? 1.0f : 0.0f);
  }
};

}  // namespace nz
