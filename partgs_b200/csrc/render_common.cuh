// Constants and launch heuristics shared by the forward (render.cu) and backward (render_bwd.cu) blend kernels.
#pragma once
#include "common.cuh"
#include <cstdlib>

namespace pgs {

constexpr unsigned RFULL = 0xffffffffu;
constexpr int NWARP = TILE_PIX / 32;
constexpr int CHUNK = 32;  // candidates per warp step (one per lane)

inline int render_env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}

// Warps per CTA.  A whole tile (8 warps) per CTA is the default; images with few tiles (e.g. the
// 400x300 DTU training resolution: 475 tiles for 592 CTA slots) are launched in finer units so that
// the block scheduler can balance the warps of heavy tiles over all SMs.
inline int warps_per_cta(int ntiles) {
  static const int forced = render_env_int("PGS_WARPS_PER_CTA", 0);
  if (forced == 1 || forced == 2 || forced == 4 || forced == 8) return forced;
  if (ntiles >= 3000) return 8;
  if (ntiles >= 1500) return 4;
  if (ntiles >= 750) return 2;
  return 1;
}
}  // namespace pgs
