// SCS hex wargame on the device — placeholder interface, filled in by the SCS milestone.
#pragma once
#include <cstdio>
#include <vector>

#include "common.cuh"

namespace nz {

struct ScsHost {
  int A = 0, C = 0, R = 0, CC = 0, planes = 0, state_words = 1;
  size_t static_bytes() const { return 8; }
  void write_image(void*) const {}
};

inline int scs_parse(const int32_t*, int, ScsHost&, char* err, size_t errlen) {
  snprintf(err, errlen, "SCS kernels are not built yet");
  return -1;
}

struct SCS {
  static constexpr bool PRIOR_F64 = false;
  using PriorT = float;
  static constexpr int TILE = 32;
  static constexpr int MIN_CTAS = 4;
  using T = Tl<TILE>;
  static constexpr bool SMEM = true;
  struct Scratch { uint32_t w[4]; };
  __device__ static void copy(Scratch& d, const Scratch& s, const T& t) { if (t.tl < 4) d.w[t.tl] = s.w[t.tl]; t.sync(); }
  __device__ static void load(Scratch& sc, const uint32_t* g, const T& t) { if (t.tl < 1) sc.w[0] = g[0]; }
  __device__ static void save(const Scratch& sc, uint32_t* g, const T& t) { if (t.tl == 0) g[0] = sc.w[0]; }
  __device__ static void reset(Scratch& sc, const View&, int, const T& t) { if (t.tl == 0) sc.w[0] = 0; }
  __device__ static int length(const Scratch&) { return 0; }
  __device__ static int to_play(const Scratch&) { return 0; }
  __device__ static bool terminal(const Scratch&) { return true; }
  __device__ static int terminal_value(const Scratch&) { return 0; }
  __device__ static void step(Scratch&, const View&, int, int, const T&) {}
  __device__ static void step_descend(Scratch&, const View&, int, int, const T&) {}
  __device__ static void settle(Scratch&, const View&, int, const T&) {}
  __device__ static void legal(const Scratch&, const View&, int, uint32_t*, const T&) {}
  __device__ static void encode(const Scratch&, const View&, int, void*, int, size_t, const T&) {}
};

}  // namespace nz
