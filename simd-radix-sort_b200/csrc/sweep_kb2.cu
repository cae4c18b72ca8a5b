// onesweep_kernel instantiations for 2-byte keys (see sweep_select.cuh)
#include "sweep_select.cuh"
namespace b200sort {
SweepFn sweep_fn_kb2(int cfg, int nstage, bool any, bool lut, bool fix) { return sweep_fn<2>(cfg, nstage, any, lut, fix); }
}  // namespace b200sort
