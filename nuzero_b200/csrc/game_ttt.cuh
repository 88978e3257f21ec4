// Tic-Tac-Toe on the device: the whole game is one 32-bit word held in a register.
// Behaviour follows Games/Tic_Tac_Toe/tic_tac_toe.py (players 1/2, state = [P1 stones, P2 stones]).
#pragma once
#include "common.cuh"

namespace nz {

struct TTT {
  // bits 0-8 P1 stones, 9-17 P2 stones, 18-21 length, 22 terminal, 23-24 terminal_value+1
  static constexpr int STATE_WORDS = 1;
  static constexpr int A = 9, PLANES = 1, R = 3, CC = 3, C = 2;
  static constexpr int MASK_WORDS = 1;
  static constexpr bool PRIOR_F64 = true;  // float64 mask -> float64 priors (tic_tac_toe.py:122)
  using PriorT = double;

  static constexpr bool SMEM = false;  // per-lane register copy, every lane computes the same word
  struct Scratch { uint32_t s; };
  __device__ static __forceinline__ void copy(Scratch& d, const Scratch& s, int) { d.s = s.s; }

  __device__ static __forceinline__ uint32_t initial() { return 1u << 23; }  // value 0 -> code 1

  __device__ static __forceinline__ void load(Scratch& sc, const uint32_t* g, int) { sc.s = g[0]; }
  __device__ static __forceinline__ void save(const Scratch& sc, uint32_t* g, int lane) {
    if (lane == 0) g[0] = sc.s;
  }
  __device__ static __forceinline__ void reset(Scratch& sc, const View&, int, int) { sc.s = initial(); }

  __device__ static __forceinline__ int length(const Scratch& sc) { return (sc.s >> 18) & 15; }
  // get_current_player(): length % 2 + 1 (tic_tac_toe.py:165, also after the terminal move)
  __device__ static __forceinline__ int to_play(const Scratch& sc) { return (length(sc) & 1) + 1; }
  __device__ static __forceinline__ bool terminal(const Scratch& sc) { return (sc.s >> 22) & 1; }
  __device__ static __forceinline__ int terminal_value(const Scratch& sc) { return (int)((sc.s >> 23) & 3) - 1; }

  __device__ static __forceinline__ bool has_line(uint32_t b) {
    // rows 0007 0070 0700, columns 0111 0222 0444, diagonals 0421 0124 (octal)
    return ((b & 0007u) == 0007u) | ((b & 0070u) == 0070u) | ((b & 0700u) == 0700u) | ((b & 0111u) == 0111u) |
           ((b & 0222u) == 0222u) | ((b & 0444u) == 0444u) | ((b & 0421u) == 0421u) | ((b & 0124u) == 0124u);
  }

  // step (tic_tac_toe.py:161-167) + check_terminal (:198-262): P1 lines are tested first, a full
  // board ends the game with the value found so far.  Returns false on an occupied cell.
  __device__ static __forceinline__ bool step(Scratch& sc, const View&, int, int action, int) {
    uint32_t s = sc.s;
    uint32_t len = (s >> 18) & 15;
    uint32_t occ = (s | (s >> 9)) & 0x1ffu;
    bool ok = action >= 0 && action < 9 && !((occ >> action) & 1u);
    uint32_t me = len & 1u;
    s |= 1u << (action + 9 * me);
    len += 1;
    uint32_t p1 = s & 0x1ffu, p2 = (s >> 9) & 0x1ffu;
    int v = has_line(p1) ? 1 : (has_line(p2) ? -1 : 0);
    bool done = (v != 0) || (len == 9);
    s = (s & 0x3ffffu) | (len << 18) | ((done ? 1u : 0u) << 22) | ((uint32_t)(v + 1) << 23);
    sc.s = s;
    return ok;
  }

  // possible_actions (tic_tac_toe.py:121-129) as a bit set
  __device__ static __forceinline__ void legal(const Scratch& sc, const View&, int, uint32_t* words, int lane) {
    uint32_t occ = (sc.s | (sc.s >> 9)) & 0x1ffu;
    if (lane == 0) words[0] = (~occ) & 0x1ffu;
    __syncwarp();  // `words` is per-warp shared memory
  }

  // generate_state_image (tic_tac_toe.py:135-159): 18 values, plane-major
  __device__ static __forceinline__ void encode(const Scratch& sc, const View& v, int, void* out, int dtype,
                                                size_t row, int lane) {
    if (lane < 18) store_leaf(out, dtype, row * 18 + lane, ((sc.s >> lane) & 1u) ? 1.0f : 0.0f);
  }
};

}  // namespace nz
