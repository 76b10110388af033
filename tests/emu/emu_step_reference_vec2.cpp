// TEST INFRASTRUCTURE -- not product code.  emu_step_reference.cpp with two cells per thread: step_cells<LBM_COMPAT_REFERENCE, MODE_BULK,
// ..., VEC = 2> -- 64-bit loads / stores, the u16 flag pair, and collide_reference_t<P2> on host stand-ins of the packed f32x2
// primitives (lbm_phys.cuh; mul0 = fma(a, b, +0) as on the device).
#define EMU_VEC 2
#include "emu_step_reference.cpp"
