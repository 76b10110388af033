// TEST INFRASTRUCTURE -- not product code.  emu_step_reference.cpp with the one-cell legacy collision routed through the V-generic
// template, collide_reference_t<float> (lbm_step_kernel.cuh): the statement sequence the packed VEC = 2 kernel instantiates with V = P2.
#define LBM_REF_GENERIC_COLLISION 1
#include "emu_step_reference.cpp"
