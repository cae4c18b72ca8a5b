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
inline SweepFn sweep_variant(int nstage, bool any, bool lut) {
  if (lut) return onesweep_kernel<KB, THREADS, IPT, MINB, 1, true, true>;  // multi-GPU partition pass
  if (nstage == 2) return any ? onesweep_kernel<KB, THREADS, IPT, MINB, 2, true, false> : onesweep_kernel<KB, THREADS, IPT, MINB, 2, false, false>;
  return any ? onesweep_kernel<KB, THREADS, IPT, MINB, 1, true, false> : onesweep_kernel<KB, THREADS, IPT, MINB, 1, false, false>;
}

template <int KB>
inline SweepFn sweep_fn(int cfg, int nstage, bool any, bool lut) {
  switch (cfg) {
    case 0: return sweep_variant<KB, 512, 16, 1>(nstage, any, lut);
    case 1: return sweep_variant<KB, 256, 16, 3>(nstage, any, lut);
    case 2: return sweep_variant<KB, 256, 8, 4>(nstage, any, lut);
    default: return sweep_variant<KB, 256, 16, 4>(nstage, any, lut);
  }
}

SweepFn sweep_fn_kb1(int cfg, int nstage, bool any, bool lut);
SweepFn sweep_fn_kb2(int cfg, int nstage, bool any, bool lut);
SweepFn sweep_fn_kb4(int cfg, int nstage, bool any, bool lut);
SweepFn sweep_fn_kb8(int cfg, int nstage, bool any, bool lut);

}  // namespace b200sort
