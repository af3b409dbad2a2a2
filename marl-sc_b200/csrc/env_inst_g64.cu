// env_inst_g64.cu - K1 instantiations for teams of 64 lanes = two warps (SKUs per lane: 2 4 8), lean only.
#include "env_split.cuh"
#define STEP_CASES \
  MARLSC_SPL_CASE(64, 2, launch_step_t, a, io, t, s) \
  MARLSC_SPL_CASE(64, 4, launch_step_t, a, io, t, s) \
  MARLSC_SPL_CASE(64, 8, launch_step_t, a, io, t, s) \

#define RESET_CASES   /* reset runs through the 32-lane kernel (env_step.cu) */

MARLSC_DEFINE_G(64, STEP_CASES, RESET_CASES)
MARLSC_DEFINE_SPLIT(64, MARLSC_SPLIT_CASE(64, 2) MARLSC_SPLIT_CASE(64, 4) MARLSC_SPLIT_CASE(64, 8))
