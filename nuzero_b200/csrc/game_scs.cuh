// SCS hex wargame on the device (Games/SCS/SCS_Game.py + Unit.py / Tile.py / Terrain.py).
//
// One warp owns one game.  The dynamic state — stage machine, unit words, ordered attacker list and
// the per-tile stack table — is staged in that warp's shared memory; the scenario (terrain, schedule,
// arrival sets, victory points) is a read-only image in global memory shared by all games.
// A unit is one 32-bit word: pos (12 bits) | status (3) | stack level (2) | movement points (8).
// The three per-player status lists of the reference become the status field.
#pragma once
#include <cstdio>
#include <cstring>
#include <vector>

#include "common.cuh"

namespace nz {

#define SCS_MAX_UNITS 32
#define SCS_MAX_ATT 20
#define SCS_QUEUED 3
#define SCS_AVAILABLE 0
#define SCS_MOVED 1
#define SCS_ATTACKED 2
#define SCS_DEAD 4
#define SCS_HDR_WORDS 8

struct ScsStatic {  // header of the scenario image; offsets in bytes from the image start
  int R, Cc, RC, turns, S, planes, A, C;
  int n_types, n_units, n_maps, n_arrsets;
  int n_vp0, n_vp1, count0, count1;
  int first1, arr_words, map_stride, pad0;
  int off_types;    // double[n_types][3]: attack_modifier, defense_modifier, cost
  int off_units;    // int[n_units][8]: player, turn, attack, defense, movement, arrival set, -, -
  int off_cum;      // int[2][turns + 2]: units scheduled strictly before turn t
  int off_arr;      // uint32[n_arrsets][arr_words]: bit set over tiles
  int off_terrain;  // uint8[n_maps][map_stride]: terrain type of each tile
  int off_vpown;    // uint8[n_maps][map_stride]: 0 none, 1 victory point of P1, 2 of P2
  int off_vplist;   // int[n_maps][n_vp0 + n_vp1]
  int off_planes;   // int[C]: decoded state-plane descriptors (kind | a0 << 4 | a1 << 12 | a2 << 20)
  uint32_t magic_rc, magic_s;  // floor(2^32 / d) + 1: n / d == umulhi(n, magic) for n < 2^16 (action and plane decoding)
  int off_nbr;      // int16[RC][8]: neighbour of a tile in directions n, ne, se, s, sw, nw (-1 off the board), two pad entries
  int total_bytes;
};

// check_tiles (SCS_Game.py:1048-1094): neighbour of `tile` in direction d (n, ne, se, s, sw, nw) or -1 off the board; even
// columns are shifted up (:1199-1243).  Evaluated once per tile on the host (the table above); the kernels look it up —
// the division and the six-way switch were ~10 % of the search kernel's instructions.
__host__ __device__ inline int scs_neighbour(int R, int C, int tile, int d) {
  const int r = tile / C, c = tile - r * C;
  const bool even = (c & 1) == 0;
  switch (d) {
    case 0: return r == 0 ? -1 : tile - C;
    case 3: return r == R - 1 ? -1 : tile + C;
    case 5: return (c == 0 || (r == 0 && even)) ? -1 : (even ? tile - C - 1 : tile - 1);
    case 4: return (c == 0 || (r == R - 1 && !even)) ? -1 : (even ? tile - 1 : tile + C - 1);
    case 1: return (c == C - 1 || (r == 0 && even)) ? -1 : (even ? tile - C + 1 : tile + 1);
    default: return (c == C - 1 || (r == R - 1 && !even)) ? -1 : (even ? tile + 1 : tile + C + 1);
  }
}

// ---- host side: parse the flat int32 description produced by nuzero_b200/games/scs_config.py -------
struct ScsHost {
  int A = 0, C = 0, R = 0, CC = 0, planes = 0, state_words = 1, S = 0, RC = 0, n_units = 0;
  std::vector<unsigned char> image;
  size_t static_bytes() const { return image.size() < 8 ? 8 : image.size(); }
  void write_image(void* out) const { memcpy(out, image.data(), image.size()); }
  size_t occ_bytes() const { return (((size_t)RC * S) + 15) & ~(size_t)15; }
};

inline int scs_parse(const int32_t* d, int n, ScsHost& h, char* err, size_t errlen) {
  auto bad = [&](const char* m) { snprintf(err, errlen, "scs_desc: %s", m); return -1; };
  if (n < 16 || d[0] != 0x53435331) return bad("bad magic");
  ScsStatic st;
  memset(&st, 0, sizeof(st));
  st.R = d[1]; st.Cc = d[2]; st.turns = d[3]; st.S = d[4]; st.n_types = d[5]; st.n_units = d[6];
  st.n_maps = d[7]; st.n_vp0 = d[8]; st.n_vp1 = d[9]; st.count0 = d[10]; st.count1 = d[11]; st.n_arrsets = d[12];
  st.RC = st.R * st.Cc;
  if (st.R < 1 || st.Cc < 1 || st.RC > 4095) return bad("board must have 1..4095 tiles");
  if (st.S < 1 || st.S > 3) return bad("stacking limit must be 1..3");
  if (st.turns < 1 || st.turns > 250) return bad("turns must be 1..250");
  if (st.n_units < 1 || st.n_units > SCS_MAX_UNITS) return bad("at most 32 units per scenario");
  if (st.count0 + st.count1 != st.n_units) return bad("unit counts do not add up");
  if (st.n_vp0 < 1 || st.n_vp1 < 1) return bad("each player needs a victory point");
  if (st.n_maps < 1 || st.n_types < 1 || st.n_arrsets < 1) return bad("empty tables");
  st.planes = 3 + 9 * st.S;
  st.A = st.planes * st.RC;
  st.C = 48 + 19 * st.S;
  st.first1 = st.count0;
  st.arr_words = (st.RC + 31) / 32;
  st.map_stride = (st.RC + 3) & ~3;
  const int nvp = st.n_vp0 + st.n_vp1;
  const long need = 16L + 6L * st.n_types + 6L * st.n_units + (long)st.n_arrsets * st.arr_words +
                    (long)st.n_maps * st.RC + (long)st.n_maps * nvp;
  if (n != need) return bad("length mismatch");
  size_t off = (sizeof(ScsStatic) + 15) & ~(size_t)15;
  auto take = [&](size_t bytes) { size_t o = off; off = (off + bytes + 15) & ~(size_t)15; return (int)o; };
  st.off_types = take((size_t)st.n_types * 3 * 8);
  st.off_units = take((size_t)st.n_units * 8 * 4);
  st.off_cum = take((size_t)2 * (st.turns + 2) * 4);
  st.off_arr = take((size_t)st.n_arrsets * st.arr_words * 4);
  st.off_terrain = take((size_t)st.n_maps * st.map_stride);
  st.off_vpown = take((size_t)st.n_maps * st.map_stride);
  st.off_vplist = take((size_t)st.n_maps * nvp * 4);
  st.off_planes = take((size_t)st.C * 4);
  st.off_nbr = take((size_t)st.RC * 8 * 2);
  st.magic_rc = st.RC > 1 ? (uint32_t)(0x100000000ull / (unsigned long long)st.RC) + 1u : 0u;  // 0: divisor 1
  st.magic_s = st.S > 1 ? (uint32_t)(0x100000000ull / (unsigned long long)st.S) + 1u : 0u;
  st.total_bytes = (int)off;
  h.image.assign(off, 0);
  unsigned char* img = h.image.data();
  const int32_t* p = d + 16;
  memcpy(img + st.off_types, p, (size_t)st.n_types * 24);  // doubles stored as (lo, hi) int pairs
  p += 6 * st.n_types;
  int* units = (int*)(img + st.off_units);
  int* cum = (int*)(img + st.off_cum);
  int seen[2] = {0, 0};
  for (int u = 0; u < st.n_units; ++u, p += 6) {
    const int pl = p[0], turn = p[1];
    if (pl != (u >= st.first1 ? 1 : 0)) return bad("units must list player 1 first, then player 2");
    if (turn < 0 || turn > st.turns) return bad("unit turn out of range");
    if (p[4] < 0 || p[4] > 255) return bad("movement allowance must be 0..255");
    if (p[5] < 0 || p[5] >= st.n_arrsets) return bad("arrival set out of range");
    if (u > 0 && units[(u - 1) * 8] == pl && units[(u - 1) * 8 + 1] > turn) return bad("units must be in schedule order");
    for (int k = 0; k < 6; ++k) units[u * 8 + k] = p[k];
    seen[pl]++;
  }
  for (int pl = 0; pl < 2; ++pl)
    for (int t = 0; t < st.turns + 2; ++t) {
      int c = 0;
      for (int u = 0; u < st.n_units; ++u)
        if (units[u * 8] == pl && units[u * 8 + 1] < t) c++;
      cum[pl * (st.turns + 2) + t] = c;
    }
  memcpy(img + st.off_arr, p, (size_t)st.n_arrsets * st.arr_words * 4);
  p += st.n_arrsets * st.arr_words;
  for (int m = 0; m < st.n_maps; ++m)
    for (int tl = 0; tl < st.RC; ++tl) {
      const int ty = p[m * st.RC + tl];
      if (ty < 0 || ty >= st.n_types) return bad("terrain id out of range");
      img[st.off_terrain + m * st.map_stride + tl] = (unsigned char)ty;
    }
  p += st.n_maps * st.RC;
  int* vplist = (int*)(img + st.off_vplist);
  for (int m = 0; m < st.n_maps; ++m)
    for (int k = 0; k < nvp; ++k) {
      const int tl = p[m * nvp + k];
      if (tl < 0 || tl >= st.RC) return bad("victory point off the board");
      vplist[m * nvp + k] = tl;
      img[st.off_vpown + m * st.map_stride + tl] = (unsigned char)(k < st.n_vp0 ? 1 : 2);
    }
  {
    short* nb = (short*)(img + st.off_nbr);
    for (int tl = 0; tl < st.RC; ++tl)
      for (int dd = 0; dd < 8; ++dd) nb[tl * 8 + dd] = dd < 6 ? (short)scs_neighbour(st.R, st.Cc, tl, dd) : (short)-1;
  }
  const double* types = (const double*)(img + st.off_types);
  for (int ty = 0; ty < st.n_types; ++ty) {
    const double c = types[ty * 3 + 2];
    if (c < 0 || c > 255 || c != (double)(int)c) return bad("terrain cost must be an integer 0..255");
  }
  {  // state-plane descriptors (generate_state, SCS_Game.py:1348-1505), decoded once on the host
    int* pt = (int*)(img + st.off_planes);
    const int S = st.S, ub = 41, fb = 41 + 18 * S;
    for (int plane = 0; plane < st.C; ++plane) {
      int kind, a0 = 0, a1 = 0, a2 = 0;
      if (plane < 3) { kind = 0; a0 = plane; }                      // terrain attack / defense / cost
      else if (plane < 5) { kind = 1; a0 = plane - 2; }             // victory points of P1 / P2
      else if (plane < ub) {                                        // next three reinforcements of each player
        const int p = (plane - 5) / 18, j = (plane - 5) % 18, k = j / 6, f = j % 6;
        kind = f < 3 ? 2 : 6; a0 = p; a1 = k; a2 = f % 3;
      } else if (plane < fb) {                                      // units: player, status, stack level, stat
        const int q = plane - ub, p = q / (9 * S), r = q % (9 * S), stt = r / (3 * S), r2 = r % (3 * S);
        kind = 3; a0 = p | (stt << 1) | ((r2 / 3) << 3); a1 = r2 % 3;
      } else if (plane == fb) { kind = 4; a0 = 15; }                // target tile
      else if (plane < fb + 1 + S) { kind = 4; a0 = plane - fb - 1; }   // attackers by stack level
      else if (plane < fb + 5 + S) { kind = 5; a0 = plane - (fb + 1 + S); }  // sub-phase one-hot
      else if (plane == fb + 5 + S) { kind = 7; }                   // turn fraction
      else { kind = 8; }                                            // player plane
      pt[plane] = kind | (a0 << 4) | (a1 << 12) | (a2 << 20);
    }
  }
  memcpy(img, &st, sizeof(st));
  h.A = st.A; h.C = st.C; h.R = st.R; h.CC = st.Cc; h.planes = st.planes; h.S = st.S; h.RC = st.RC;
  h.n_units = st.n_units;
  h.state_words = SCS_HDR_WORDS + st.n_units;
  return 0;
}

// ---- device side -----------------------------------------------------------------------------------
struct SCS {
  static constexpr bool PRIOR_F64 = false;  // int8 mask -> float32 priors (SCS_Game.py:399-408)
  static constexpr bool NODE_STATE = true;  // nodes may keep their game state (View::nstate): one game step per simulation
  static constexpr bool MAY_FLIP = false;   // players are 0 / 1: `to_play == 2` (Explorer.py:124) never holds
  using PriorT = float;
  static constexpr int TILE = 32;
  // <= 72 registers (7 CTAs of 4 warps per SM): all 4096 games of the SCS-5 workload are resident in ONE wave (148 x 28
  // warps); at 128 registers the second wave was 73 % full and the launch took 474 instead of 320 us, spills included
#ifndef NZ_SCS_MIN_CTAS
#define NZ_SCS_MIN_CTAS 7
#endif
  static constexpr int MIN_CTAS = NZ_SCS_MIN_CTAS;
  using T = Tl<TILE>;
  static constexpr bool SMEM = true;

  struct Scratch {
    int stage, turn, length, terminal, tv, target, n_att, sub_phase, player;
    int placed[2];
    int pad;
    unsigned char att[SCS_MAX_ATT];
    uint32_t unit[SCS_MAX_UNITS];
    // followed by the stack table occ[RC * S] (unit id + 1 by stack level, 0 = empty slot)
  };
  // fixed part + stack table (View::scratch_extra = RC * S rounded up to 16 bytes)
  __host__ __device__ static size_t scratch_bytes(const View& v) {
    return ((sizeof(Scratch) + 15) & ~(size_t)15) + (size_t)v.scratch_extra;
  }

  struct Ctx {
    const ScsStatic* st;
    const unsigned char* img;
    const double* types;
    const int* units;
    const unsigned char* terrain;
    const unsigned char* vpown;
    const int* vplist;
    const short* nbrtab;
    __device__ __forceinline__ Ctx(const View& v, int map) {
      img = (const unsigned char*)v.gstatic;
      st = (const ScsStatic*)img;
      types = (const double*)(img + st->off_types);
      units = (const int*)(img + st->off_units);
      terrain = img + st->off_terrain + (size_t)map * st->map_stride;
      vpown = img + st->off_vpown + (size_t)map * st->map_stride;
      vplist = (const int*)(img + st->off_vplist) + (size_t)map * (st->n_vp0 + st->n_vp1);
      nbrtab = (const short*)(img + st->off_nbr);
    }
    __device__ __forceinline__ int cost(int tile) const { return (int)types[terrain[tile] * 3 + 2]; }
    __device__ __forceinline__ int uplayer(int u) const { return u >= st->first1 ? 1 : 0; }
    __device__ __forceinline__ int ustat(int u, int k) const { return units[u * 8 + 2 + k]; }  // attack, defense, movement
    __device__ __forceinline__ int cum(int p, int t) const { return ((const int*)(img + st->off_cum))[p * (st->turns + 2) + t]; }
    __device__ __forceinline__ bool arrbit(int set, int tile) const {
      return (((const uint32_t*)(img + st->off_arr))[set * st->arr_words + (tile >> 5)] >> (tile & 31)) & 1u;
    }
    // neighbour of `tile` in direction d (n, ne, se, s, sw, nw) or -1 off the board: host-built table (scs_neighbour)
    __device__ __forceinline__ int neighbour(int tile, int d) const { return (int)nbrtab[tile * 8 + d]; }
  };

  __device__ static __forceinline__ unsigned char* occ(Scratch& sc) { return (unsigned char*)(&sc + 1); }
  __device__ static __forceinline__ const unsigned char* occ(const Scratch& sc) { return (const unsigned char*)(&sc + 1); }
  __device__ static __forceinline__ uint32_t pack(int pos, int status, int level, int mov) {
    return (uint32_t)pos | ((uint32_t)status << 12) | ((uint32_t)level << 15) | ((uint32_t)mov << 17);
  }
  __device__ static __forceinline__ int upos(uint32_t w) { return (int)(w & 0xfffu); }
  __device__ static __forceinline__ int ustatus(uint32_t w) { return (int)((w >> 12) & 7u); }
  __device__ static __forceinline__ int ulevel(uint32_t w) { return (int)((w >> 15) & 3u); }
  __device__ static __forceinline__ int umov(uint32_t w) { return (int)((w >> 17) & 0xffu); }
  __device__ static __forceinline__ bool on_board(uint32_t w) { return ustatus(w) <= SCS_ATTACKED; }

  __device__ static __forceinline__ int count_at(const Scratch& sc, int S, int tile) {
    const unsigned char* o = occ(sc) + tile * S;
    int c = 0;
    for (int l = 0; l < S; ++l) c += o[l] != 0;
    return c;
  }
  // tile.player (Tile.py:28-36): owner of the stack, -1 when empty
  __device__ static __forceinline__ int owner_at(const Scratch& sc, const Ctx& cx, int tile) {
    const int u = occ(sc)[tile * cx.st->S];
    return u ? cx.uplayer(u - 1) : -1;
  }

  // ---- interface ----------------------------------------------------------------------------------
  __device__ static __forceinline__ int length(const Scratch& sc) { return sc.length; }
  __device__ static __forceinline__ int to_play(const Scratch& sc) { return sc.player; }  // players are 0 / 1 (:93)
  __device__ static __forceinline__ bool terminal(const Scratch& sc) { return sc.terminal != 0; }
  __device__ static __forceinline__ int terminal_value(const Scratch& sc) { return sc.tv; }

  // game.shallow_clone() (:1782-1793): word copy of the fixed part and the stack table
  __device__ static void copy(Scratch& d, const Scratch& s, const View& v, const T& t) {
    const uint32_t* src = (const uint32_t*)&s;
    uint32_t* dst = (uint32_t*)&d;
    const int n = (int)(sizeof(Scratch) / 4) + v.scratch_extra / 4;
    t.sync();
    for (int i = t.tl; i < n; i += TILE) dst[i] = src[i];
    t.sync();
  }

  __device__ static void rebuild_occ(Scratch& sc, const Ctx& cx, const T& t) {
    const int S = cx.st->S, n = (cx.st->RC * S + 15) & ~15;
    uint32_t* o32 = (uint32_t*)occ(sc);
    for (int i = t.tl; i < n / 4; i += TILE) o32[i] = 0u;
    t.sync();
    if (t.tl < cx.st->n_units) {
      const uint32_t w = sc.unit[t.tl];
      if (on_board(w)) occ(sc)[upos(w) * S + ulevel(w)] = (unsigned char)(t.tl + 1);
    }
    t.sync();
  }

  __device__ static void save(const Scratch& sc, uint32_t* g, const View& v, const T& t) {
    if (t.tl < ((const ScsStatic*)v.gstatic)->n_units) g[SCS_HDR_WORDS + t.tl] = sc.unit[t.tl];
    if (t.tl == 0) {
      g[0] = (uint32_t)(sc.stage + 2) | ((uint32_t)sc.turn << 4) | ((uint32_t)sc.terminal << 12) |
             ((uint32_t)(sc.tv + 1) << 13) | ((uint32_t)sc.n_att << 15) | ((uint32_t)sc.sub_phase << 20) |
             ((uint32_t)sc.player << 22);
      g[1] = (uint32_t)sc.length | ((uint32_t)sc.placed[0] << 16) | ((uint32_t)sc.placed[1] << 24);
      g[2] = (uint32_t)(sc.target + 1);
    }
    if (t.tl < SCS_MAX_ATT / 4) g[3 + t.tl] = ((const uint32_t*)sc.att)[t.tl];
  }
  __device__ static void load_hdr(Scratch& sc, const uint32_t* g, int n_units, const T& t) {
    if (t.tl == 0) {
      const uint32_t a = g[0], b = g[1], c = g[2];
      sc.stage = (int)(a & 15u) - 2;
      sc.turn = (int)((a >> 4) & 0xffu);
      sc.terminal = (int)((a >> 12) & 1u);
      sc.tv = (int)((a >> 13) & 3u) - 1;
      sc.n_att = (int)((a >> 15) & 31u);
      sc.sub_phase = (int)((a >> 20) & 3u);
      sc.player = (int)((a >> 22) & 1u);
      sc.length = (int)(b & 0xffffu);
      sc.placed[0] = (int)((b >> 16) & 0xffu);
      sc.placed[1] = (int)((b >> 24) & 0xffu);
      sc.target = (int)c - 1;
    }
    if (t.tl < SCS_MAX_ATT / 4) ((uint32_t*)sc.att)[t.tl] = g[3 + t.tl];
    sc.unit[t.tl] = t.tl < n_units ? g[SCS_HDR_WORDS + t.tl] : pack(0, SCS_DEAD, 0, 0);
  }

  // ---- rules --------------------------------------------------------------------------------------
  // remove unit u from its tile: higher stack levels shift down (Tile.remove_unit, Tile.py:32-35)
  __device__ static void remove_from_tile(Scratch& sc, int S, int u) {  // one lane
    const uint32_t w = sc.unit[u];
    unsigned char* o = occ(sc) + upos(w) * S;
    for (int l = ulevel(w); l < S; ++l) {
      const int nxt = (l + 1 < S) ? o[l + 1] : 0;
      o[l] = (unsigned char)nxt;
      if (nxt) {
        const uint32_t x = sc.unit[nxt - 1];
        sc.unit[nxt - 1] = pack(upos(x), ustatus(x), l, umov(x));
      }
    }
  }
  __device__ static void set_status(Scratch& sc, int u, int status) {  // one lane
    const uint32_t x = sc.unit[u];
    sc.unit[u] = pack(upos(x), status, ulevel(x), umov(x));
  }

  // end_movement (:927-940): status 1, and straight to status 2 when no enemy unit is adjacent
  __device__ static void end_movement(Scratch& sc, const Ctx& cx, int u, const T& t) {
    const int pos = upos(sc.unit[u]);
    const int enemy = cx.uplayer(u) ^ 1;
    bool adj = false;
    if (t.tl < 6) {
      const int nt = cx.neighbour(pos, t.tl);
      adj = nt >= 0 && owner_at(sc, cx, nt) == enemy;
    }
    const bool any = t.ballot(adj) != 0u;
    if (t.tl == 0) set_status(sc, u, any ? SCS_MOVED : SCS_ATTACKED);
    t.sync();
  }

  // get_strongest_attacker / _defender (:1253-1285): strict improvements only -> first in list order
  __device__ static bool stronger(const Ctx& cx, int u, int best, int k1, int k2) {
    const int a = cx.ustat(u, k1), b = cx.ustat(best, k1);
    if (a != b) return a > b;
    const int a2 = cx.ustat(u, k2), b2 = cx.ustat(best, k2);
    if (a2 != b2) return a2 > b2;
    return cx.ustat(u, 2) > cx.ustat(best, 2);
  }

  // resolve_combat (:997-1044), sequential on one lane in the reference's order of operations
  __device__ static void resolve_combat(Scratch& sc, const Ctx& cx) {
    const int S = cx.st->S, tt = sc.target;
    const unsigned char* o = occ(sc) + tt * S;
    int dsum = 0;
    for (int l = 0; l < S; ++l)
      if (o[l]) dsum += cx.ustat(o[l] - 1, 1);
    const double total_def = (double)dsum * cx.types[cx.terrain[tt] * 3 + 1];
    double total_att = 0.0;
    for (int i = 0; i < sc.n_att; ++i) {
      const int a = sc.att[i];
      total_att += (double)cx.ustat(a, 0) * cx.types[cx.terrain[upos(sc.unit[a])] * 3 + 0];
      set_status(sc, a, SCS_ATTACKED);
    }
    if (total_att <= total_def) {  // attacker loses its strongest unit
      int best = sc.att[0];
      for (int i = 1; i < sc.n_att; ++i)
        if (stronger(cx, sc.att[i], best, 0, 1)) best = sc.att[i];
      remove_from_tile(sc, S, best);
      set_status(sc, best, SCS_DEAD);
    }
    if (total_att >= total_def) {  // defender loses its strongest unit (live stack of the target tile)
      int best = o[0] - 1;
      for (int l = 1; l < S; ++l)
        if (o[l] && stronger(cx, o[l] - 1, best, 1, 0)) best = o[l] - 1;
      if (best >= 0) {
        remove_from_tile(sc, S, best);
        set_status(sc, best, SCS_DEAD);
      }
    }
  }

  // update_game_env (:687-831): the stage machine; every lane evaluates it on the shared state
  __device__ static void update_env(Scratch& sc, const Ctx& cx, const T& t) {
    const ScsStatic* st = cx.st;
    t.sync();
    int stage = sc.stage, turn = sc.turn;
    bool done = false;
    const int placed0 = sc.placed[0], placed1 = sc.placed[1], target = sc.target;
    uint32_t w = sc.unit[t.tl];
    const int mine = (t.tl < st->n_units) ? cx.uplayer(t.tl) : -1;
    while (true) {
      if (stage == -2) {
        if (placed0 >= cx.cum(0, turn + 1)) { stage = -1; continue; }
      } else if (stage == -1) {
        if (placed1 >= cx.cum(1, turn + 1)) { turn += 1; stage = 0; continue; }
      } else if (stage == 0 || stage == 4) {
        if ((stage ? placed1 : placed0) >= cx.cum(stage >> 2, turn + 1)) { stage += 1; continue; }
      } else if (stage == 1 || stage == 5) {
        if (t.ballot(mine == (stage >> 2) && ustatus(w) == SCS_AVAILABLE) == 0u) { stage += 1; continue; }
      } else if (stage == 2 || stage == 6) {
        if (t.ballot(mine == (stage >> 2) && ustatus(w) == SCS_MOVED) == 0u) {
          if (stage == 2) { stage = 4; continue; }
          if (turn + 1 > st->turns) { done = true; break; }
          turn += 1;
          stage = 0;
          if (ustatus(w) == SCS_ATTACKED)  // new_turn (:845-855)
            w = pack(upos(w), SCS_AVAILABLE, ulevel(w), cx.ustat(t.tl, 2));
          continue;
        } else if (target >= 0) { stage += 1; continue; }
      } else {  // 3 / 7
        if (target < 0) { stage -= 1; continue; }
      }
      break;
    }
    sc.unit[t.tl] = w;
    int tv = 0;
    if (done) {  // check_termination (:857-894)
      const int nvp = st->n_vp0 + st->n_vp1;
      int cap0 = 0, cap1 = 0;  // cap1: P1 victory points held by P2, cap0: P2 victory points held by P1
      for (int k = t.tl; k < nvp; k += TILE) {
        const int own = owner_at(sc, cx, cx.vplist[k]);
        if (k < st->n_vp0) cap1 += own == 1;
        else cap0 += own == 0;
      }
      cap0 = t.isum(cap0);
      cap1 = t.isum(cap1);
      const double a = (double)cap0 / (double)st->n_vp1, b = (double)cap1 / (double)st->n_vp0;
      tv = a > b ? 1 : (a < b ? -1 : 0);
    }
    t.sync();
    if (t.tl == 0) {
      sc.stage = stage;
      sc.turn = turn;
      sc.player = (stage == -2 || (stage >= 0 && stage <= 3)) ? 0 : 1;
      sc.sub_phase = stage < 0 ? 0 : ((stage & 3) == 0 ? 0 : (stage & 3));
      if (done) { sc.terminal = 1; sc.tv = tv; }
    }
    t.sync();
  }

  __device__ static void reset(Scratch& sc, const View& v, int map, const T& t) {
    const Ctx cx(v, map);
    if (t.tl == 0) {
      sc.stage = -2; sc.turn = 0; sc.length = 0; sc.terminal = 0; sc.tv = 0; sc.target = -1; sc.n_att = 0;
      sc.sub_phase = 0; sc.player = 0; sc.placed[0] = 0; sc.placed[1] = 0; sc.pad = 0;
    }
    if (t.tl < SCS_MAX_ATT / 4) ((uint32_t*)sc.att)[t.tl] = 0u;
    sc.unit[t.tl] = t.tl < cx.st->n_units ? pack(0, SCS_QUEUED, 0, cx.ustat(t.tl, 2)) : pack(0, SCS_DEAD, 0, 0);
    t.sync();
    rebuild_occ(sc, cx, t);
    update_env(sc, cx, t);
  }

  __device__ static void load(Scratch& sc, const uint32_t* g, const View& v, int map, const T& t) {
    const Ctx cx(v, map);
    t.sync();
    load_hdr(sc, g, cx.st->n_units, t);
    t.sync();
    rebuild_occ(sc, cx, t);
  }

  // step (:375-391) -> play_action (:569-633) -> update_game_env.  The action is assumed legal
  // (the search only plays actions taken from the legal mask; nz_env_step checks first).
  __device__ static void step(Scratch& sc, const View& v, int map, int action, const T& t) {
    const Ctx cx(v, map);
    const ScsStatic* st = cx.st;
    const int S = st->S, RC = st->RC;
    const int plane = st->magic_rc ? (int)__umulhi((uint32_t)action, st->magic_rc) : action, tile = action - plane * RC;  // action / RC
    t.sync();
    if (plane < 1) {  // place the next reinforcement (:572-580)
      if (t.tl == 0) {
        const int p = sc.player;
        const int u = (p ? st->first1 : 0) + sc.placed[p];
        const int lvl = count_at(sc, S, tile);
        sc.placed[p] += 1;
        sc.unit[u] = pack(tile, SCS_AVAILABLE, lvl, cx.ustat(u, 2));
        occ(sc)[tile * S + lvl] = (unsigned char)(u + 1);
      }
    } else if (plane < 1 + 6 * S) {  // movement (:582-599): plane = dir * S + stack level
      const int d = st->magic_s ? (int)__umulhi((uint32_t)(plane - 1), st->magic_s) : plane - 1, s = (plane - 1) - d * S;  // (plane - 1) / S
      const int u = occ(sc)[tile * S + s] - 1;
      const int dest = cx.neighbour(tile, d);
      const int left = umov(sc.unit[u]) - cx.cost(dest);
      t.sync();
      if (t.tl == 0) {
        remove_from_tile(sc, S, u);
        const int lvl = count_at(sc, S, dest);
        sc.unit[u] = pack(dest, SCS_AVAILABLE, lvl, left);
        occ(sc)[dest * S + lvl] = (unsigned char)(u + 1);
      }
      t.sync();
      bool can = false;  // check_mobility(unit, consider_other_units=False) (:1096-1111)
      if (t.tl < 6) {
        const int nt = cx.neighbour(dest, t.tl);
        can = nt >= 0 && left - cx.cost(nt) >= 0;
      }
      if (t.ballot(can) == 0u) end_movement(sc, cx, u, t);
    } else if (plane < 2 + 6 * S) {  // choose target (:606-608)
      if (t.tl == 0) sc.target = tile;
    } else if (plane < 2 + 7 * S) {  // choose attacker (:610-613)
      if (t.tl == 0) {
        sc.att[sc.n_att] = (unsigned char)(occ(sc)[tile * S + (plane - (2 + 6 * S))] - 1);
        sc.n_att += 1;
      }
    } else if (plane < 3 + 7 * S) {  // confirm attack (:615-618)
      if (t.tl == 0) {
        resolve_combat(sc, cx);
        sc.target = -1;
        sc.n_att = 0;
      }
    } else if (plane < 3 + 8 * S) {  // no move (:620-623)
      end_movement(sc, cx, occ(sc)[tile * S + (plane - (3 + 7 * S))] - 1, t);
    } else {  // no fight (:625-628)
      if (t.tl == 0) set_status(sc, occ(sc)[tile * S + (plane - (3 + 8 * S))] - 1, SCS_ATTACKED);
    }
    if (t.tl == 0) sc.length += 1;
    update_env(sc, cx, t);
  }
  __device__ static __forceinline__ void step_descend(Scratch& sc, const View& v, int map, int action, const T& t) {
    step(sc, v, map, action, t);
  }
  __device__ static __forceinline__ void settle(Scratch&, const View&, int, const T&) {}

  // possible_actions (:395-484) as a bit set over the flat action index plane * RC + tile
  __device__ static void legal(const Scratch& sc, const View& v, int map, uint32_t* words, const T& t) {
    const Ctx cx(v, map);
    const ScsStatic* st = cx.st;
    const int S = st->S, RC = st->RC, nwords = (st->A + 31) >> 5;
    for (int i = t.tl; i < nwords; i += TILE) words[i] = 0u;
    t.sync();
    auto set_bit = [&](int plane, int tile) {
      const int a = plane * RC + tile;
      atomicOr(&words[a >> 5], 1u << (a & 31));
    };
    const int p = sc.player, enemy = p ^ 1;
    const uint32_t w = sc.unit[t.tl];
    const bool mine = t.tl < st->n_units && cx.uplayer(t.tl) == p && on_board(w);
    if (sc.sub_phase == 0) {
      const int u = min((p ? st->first1 : 0) + sc.placed[p], st->n_units - 1);  // get_next_reinforcement (:1331-1332)
      const int set = cx.units[u * 8 + 5];
      for (int tile = t.tl; tile < RC; tile += TILE)
        if (cx.arrbit(set, tile) && !(owner_at(sc, cx, tile) == enemy || count_at(sc, S, tile) == S)) set_bit(0, tile);
    } else if (sc.sub_phase == 1) {
      if (mine && ustatus(w) == SCS_AVAILABLE) {
        const int pos = upos(w), lvl = ulevel(w);
        set_bit(3 + 7 * S + lvl, pos);
        for (int d = 0; d < 6; ++d) {  // check_mobility(unit, consider_other_units=True)
          const int nt = cx.neighbour(pos, d);
          if (nt >= 0 && umov(w) - cx.cost(nt) >= 0 && count_at(sc, S, nt) != S && owner_at(sc, cx, nt) != enemy)
            set_bit(1 + d * S + lvl, pos);
        }
      }
    } else if (sc.sub_phase == 2) {
      if (mine && ustatus(w) == SCS_MOVED) {
        const int pos = upos(w);
        set_bit(3 + 8 * S + ulevel(w), pos);
        for (int d = 0; d < 6; ++d) {
          const int nt = cx.neighbour(pos, d);
          if (nt >= 0 && owner_at(sc, cx, nt) == enemy) set_bit(1 + 6 * S, nt);
        }
      }
    } else {
      if (mine && ustatus(w) != SCS_ATTACKED) {
        bool adj = false;
        for (int d = 0; d < 6; ++d) adj |= cx.neighbour(sc.target, d) == upos(w);
        bool chosen = false;
        for (int i = 0; i < sc.n_att; ++i) chosen |= sc.att[i] == t.tl;
        if (adj && !chosen) set_bit(2 + 6 * S + ulevel(w), upos(w));
      }
      if (t.tl == 0 && sc.n_att > 0) set_bit(2 + 7 * S, sc.target);
    }
    t.sync();
  }

  // generate_state (:1348-1505): [C, R, Cc] planes.  Most of the row is zero, so the row is first
  // cleared with wide stores and then only the non-zero content is written: terrain and victory-point
  // planes per tile, the queued reinforcements' arrival sets and duration fills, three values per unit
  // on the board (lane = unit), target / attacker marks, and the three feature fills.
  template <typename E>
  __device__ static void encode_t(const Scratch& sc, const Ctx& cx, E* __restrict__ row, const T& t) {
    const ScsStatic* st = cx.st;
    const int S = st->S, RC = st->RC, C = st->C;
    const int n = C * RC;
    {  // clear: 2/4-byte elements up to the first 16-byte boundary, 16-byte stores, then the tail
      const uintptr_t p0 = (uintptr_t)row;
      int head = (int)(((16 - (p0 & 15)) & 15) / sizeof(E));
      head = head < n ? head : n;
      for (int i = t.tl; i < head; i += TILE) row[i] = E(0.f);
      uint4* mid = (uint4*)(row + head);
      const int per = 16 / (int)sizeof(E), nmid = (n - head) / per;
      for (int i = t.tl; i < nmid; i += TILE) mid[i] = make_uint4(0u, 0u, 0u, 0u);
      for (int i = head + nmid * per + t.tl; i < n; i += TILE) row[i] = E(0.f);
    }
    t.sync();
    const int ub = 41, fb = 41 + 18 * S;
    for (int tile = t.tl; tile < RC; tile += TILE) {
      const int ty = cx.terrain[tile] * 3;
      row[0 * RC + tile] = E((float)cx.types[ty + 0]);  // attack modifier
      row[1 * RC + tile] = E((float)cx.types[ty + 1]);  // defense modifier
      row[2 * RC + tile] = E((float)cx.types[ty + 2]);  // movement cost
      const int own = cx.vpown[tile];
      if (own) row[(2 + own) * RC + tile] = E(1.f);
      row[(fb + 1 + S + sc.sub_phase) * RC + tile] = E(1.f);
      row[(fb + 5 + S) * RC + tile] = E((float)((double)sc.turn / (double)st->turns));
      row[(fb + 6 + S) * RC + tile] = E(sc.player == 1 ? -1.f : 1.f);
    }
    for (int p = 0; p < 2; ++p)  // the next three queued units of each player, in schedule order
      for (int k = 0; k < 3; ++k) {
        const int idx = sc.placed[p] + k;
        if (idx >= (p ? st->count1 : st->count0)) break;
        const int u = (p ? st->first1 : 0) + idx, set = cx.units[u * 8 + 5];
        const int turns_left = cx.units[u * 8 + 1] - sc.turn;
        const E dur = E((float)((double)((st->turns + 1) - turns_left) / (double)(st->turns + 1)));
        const E a = E((float)cx.ustat(u, 0)), d = E((float)cx.ustat(u, 1)), m = E((float)cx.ustat(u, 2));
        E* blk = row + (size_t)(5 + 18 * p + 6 * k) * RC;
        for (int tile = t.tl; tile < RC; tile += TILE) {
          if (cx.arrbit(set, tile)) { blk[tile] = a; blk[RC + tile] = d; blk[2 * RC + tile] = m; }
          blk[3 * RC + tile] = dur; blk[4 * RC + tile] = dur; blk[5 * RC + tile] = dur;
        }
      }
    if (t.tl < st->n_units) {  // units on the board: plane = player | status | stack level | stat
      const uint32_t w = sc.unit[t.tl];
      if (on_board(w)) {
        E* blk = row + (size_t)(ub + cx.uplayer(t.tl) * 9 * S + ustatus(w) * 3 * S + ulevel(w) * 3) * RC + upos(w);
        blk[0] = E((float)cx.ustat(t.tl, 0));
        blk[RC] = E((float)cx.ustat(t.tl, 1));
        blk[2 * RC] = E((float)umov(w));
      }
    }
    if (t.tl == 0 && sc.target >= 0) row[(size_t)fb * RC + sc.target] = E(1.f);
    if (t.tl < sc.n_att) {
      const uint32_t w = sc.unit[sc.att[t.tl]];
      row[(size_t)(fb + 1 + ulevel(w)) * RC + upos(w)] = E(1.f);
    }
  }
  __device__ static void encode(const Scratch& sc, const View& v, int map, void* out, int dtype, size_t row,
                                const T& t) {
    const Ctx cx(v, map);
    const size_t base = row * (size_t)cx.st->C * cx.st->RC;
    if (dtype == NZ_BF16) encode_t<__nv_bfloat16>(sc, cx, (__nv_bfloat16*)out + base, t);
    else encode_t<float>(sc, cx, (float*)out + base, t);
  }
};

}  // namespace nz
