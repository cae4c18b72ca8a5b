// sweep_select.cuh -- maps (key bytes, tile geometry, staging depth, chunk widths, ranking method) to an instantiation of onesweep_kernel.  The instantiations of each (key width, geometry) pair
// live in their own translation unit (sweep_inst.cu compiled with -DSWEEP_KB=.. -DSWEEP_CFG=..) so that
// they compile in parallel.
#pragma once
#include "kernels.cuh"

namespace b200sort {

using SweepFn = void (*)(const SweepArgs);

struct TileCfg { int threads, ipt, minb; };
// tile geometries of the scatter kernel (option "tile_cfg"); minb = CTAs per SM the kernel is compiled for.
// (512x16x1 was measured slower in round 1, 256x16x4 -- four CTAs per SM at 64 registers -- no faster in either
//  round: 8.13 vs 8.18 ms per pass at 1e9 records; both are gone: profiles/README.md)
constexpr TileCfg kTileCfgs[] = {{256, 16, 3}, {256, 8, 4}, {256, 32, 3}};
// (index 2: 8192-key tiles for 4-byte keys whose every stream moves in chunks of at most 4 bytes -- the same 32 KB
//  of staging as 4096 8-byte keys, half the per-tile work (look-back, scans, tickets) per key; kb = 4 only)
constexpr int kWideTileCfg = 2;
constexpr int kNumTileCfgs = sizeof(kTileCfgs) / sizeof(kTileCfgs[0]);
constexpr int kDefaultTileCfg = 0;

struct SweepSel {
  int nstage;   // 1 or 2 staging buffers
  bool any;     // a stream with 1- or 2-byte chunks takes part
  bool lut;     // partition pass of the multi-GPU sort
  bool fix;     // last pass of the MSB hybrid plan (8-byte keys, default geometry)
  int rank;     // RANK_BALLOT / RANK_ATOMIC
  bool bytewise;  // the host knows the plan: no range reduction, no left shift (digits are bytes of the raw key)
};

template <int KB, int THREADS, int IPT, int MINB, int RANK, bool BW>
inline SweepFn sweep_variant2(const SweepSel &s) {
  if (s.nstage == 2) return s.any ? onesweep_kernel<KB, THREADS, IPT, MINB, 2, true, false, false, RANK, BW> : onesweep_kernel<KB, THREADS, IPT, MINB, 2, false, false, false, RANK, BW>;
  return s.any ? onesweep_kernel<KB, THREADS, IPT, MINB, 1, true, false, false, RANK, BW> : onesweep_kernel<KB, THREADS, IPT, MINB, 1, false, false, false, RANK, BW>;
}

template <int KB, int THREADS, int IPT, int MINB, bool BW>
inline SweepFn sweep_variant1(const SweepSel &s) {
  if constexpr (KB == 8 && THREADS == 256 && IPT == 16) {
    // last pass of the MSB hybrid plan (8-byte keys, default geometry): orders the final segments a tile holds
    if (s.fix) return s.any ? onesweep_kernel<KB, THREADS, IPT, MINB, 1, true, false, true, RANK_BALLOT, BW> : onesweep_kernel<KB, THREADS, IPT, MINB, 1, false, false, true, RANK_BALLOT, BW>;
  }
  if (s.rank == RANK_ATOMIC) return sweep_variant2<KB, THREADS, IPT, MINB, RANK_ATOMIC, BW>(s);
  return sweep_variant2<KB, THREADS, IPT, MINB, RANK_BALLOT, BW>(s);
}

template <int KB, int THREADS, int IPT, int MINB>
inline SweepFn sweep_variant(const SweepSel &s) {
  if (s.lut) return onesweep_kernel<KB, THREADS, IPT, MINB, 1, true, true, false, RANK_BALLOT, false>;  // multi-GPU partition pass
  return s.bytewise ? sweep_variant1<KB, THREADS, IPT, MINB, true>(s) : sweep_variant1<KB, THREADS, IPT, MINB, false>(s);
}

// defined in sweep_inst.cu, one translation unit per (key bytes, geometry)
template <int KB, int CFG> SweepFn sweep_fn_inst(const SweepSel &s);

inline SweepFn sweep_fn(int kb, int cfg, const SweepSel &s);

}  // namespace b200sort
