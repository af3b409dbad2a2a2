// env_inst_g1.cu - K1 instantiations for teams of 1 lanes (SKUs per lane: 1 2 4 8).
#include "env_kernels.cuh"
#define STEP_CASES \
  MARLSC_SPL_CASE(1, 1, launch_step_t, a, io, t, s) \
  MARLSC_SPL_CASE(1, 2, launch_step_t, a, io, t, s) \
  MARLSC_SPL_CASE(1, 4, launch_step_t, a, io, t, s) \
  MARLSC_SPL_CASE(1, 8, launch_step_t, a, io, t, s) \

#define RESET_CASES \
  MARLSC_SPL_CASE(1, 1, launch_reset_t, a, init, per_env, obs, s) \
  MARLSC_SPL_CASE(1, 2, launch_reset_t, a, init, per_env, obs, s) \
  MARLSC_SPL_CASE(1, 4, launch_reset_t, a, init, per_env, obs, s) \
  MARLSC_SPL_CASE(1, 8, launch_reset_t, a, init, per_env, obs, s) \

MARLSC_DEFINE_G(1, STEP_CASES, RESET_CASES)
