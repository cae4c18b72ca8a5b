// onesweep_kernel instantiations of one (key width, tile geometry) pair: compiled once per pair with
// -DSWEEP_KB=<1|2|4|8> -DSWEEP_CFG=<index into kTileCfgs> (see build.py and sweep_select.cuh)
#include "sweep_select.cuh"
namespace b200sort {
template <> SweepFn sweep_fn_inst<SWEEP_KB, SWEEP_CFG>(const SweepSel &s) {
  constexpr TileCfg c = kTileCfgs[SWEEP_CFG];
  return sweep_variant<SWEEP_KB, c.threads, c.ipt, c.minb>(s);
}
}  // namespace b200sort
