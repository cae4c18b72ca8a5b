// sweep_select.cuh -- maps (key bytes, tile geometry, staging depth, chunk widths) to an instantiation of
// onesweep_kernel.  The instantiations of each key width live in their own translation unit
// (sweep_kb{1,2,4,8}.cu) so that they compile in parallel.
#pragma once
#include "kernels.cuh"

namespace b200sort {

using SweepFn = void (*)(const SweepArgs);

struct TileCfg { int threads, ipt, minb; };
// tile geometries of the scatter kernel (option "tile_cfg"); minb = CTAs per SM the kernel is compiled for
constexpr TileCfg kTileCfgs[] = {{512, 16, 1}, {256, 16, 3}, {256, 8, 4}, {256, 16, 4}};
constexpr int kNumTileCfgs = sizeof(kTileCfgs) / sizeof(kTileCfgs[0]);

template <int KB, int THREADS, int IPT, int MINB>
inline SweepFn sweep_variant(int nstage, bool any, bool lut, bool fix) {
  if (lut) return onesweep_kernel<KB, THREADS, IPT, MINB, 1, true, true>;  // multi-GPU partition pass
  if constexpr (KB == 8 && THREADS == 256 && IPT == 16 && MINB == 3) {
    // last pass of the MSB hybrid plan (8-byte keys, default geometry): orders the final segments a tile holds
    if (fix) return any ? onesweep_kernel<KB, THREADS, IPT, MINB, 1, true, false, true> : onesweep_kernel<KB, THREADS, IPT, MINB, 1, false, false, true>;
  }
  if (nstage == 2) return any ? onesweep_kernel<KB, THREADS, IPT, MINB, 2, true, false> : onesweep_kernel<KB, THREADS, IPT, MINB, 2, false, false>;
  return any ? onesweep_kernel<KB, THREADS, IPT, MINB, 1, true, false> : onesweep_kernel<KB, THREADS, IPT, MINB, 1, false, false>;
}

template <int KB>
inline SweepFn sweep_fn(int cfg, int nstage, bool any, bool lut, bool fix) {
  switch (cfg) {
    case 0: return sweep_variant<KB, 512, 16, 1>(nstage, any, lut, fix);
    case 1: return sweep_variant<KB, 256, 16, 3>(nstage, any, lut, fix);
    case 2: return sweep_variant<KB, 256, 8, 4>(nstage, any, lut, fix);
    default: return sweep_variant<KB, 256, 16, 4>(nstage, any, lut, fix);
  }
}

SweepFn sweep_fn_kb1(int cfg, int nstage, bool any, bool lut, bool fix);
SweepFn sweep_fn_kb2(int cfg, int nstage, bool any, bool lut, bool fix);
SweepFn sweep_fn_kb4(int cfg, int nstage, bool any, bool lut, bool fix);
SweepFn sweep_fn_kb8(int cfg, int nstage, bool any, bool lut, bool fix);

}  // namespace b200sort
