// fdes_b200 -- the sweeps of ONE grid size (compile with -DFDES_SWEEP_N=<N>; see sweep_vtable.h).
#include "sweep_kernels.cuh"

#ifndef FDES_SWEEP_N
#error "compile with -DFDES_SWEEP_N=<grid size>"
#endif
#define FDES_VT_NAME_(n) sweep_vtable_##n
#define FDES_VT_NAME(n) FDES_VT_NAME_(n)

namespace fdes {
const SweepVTable* FDES_VT_NAME(FDES_SWEEP_N)() { return make_sweep_vtable<FDES_SWEEP_N>(); }
}  // namespace fdes
