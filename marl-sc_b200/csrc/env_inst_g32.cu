// env_inst_g32.cu - K1 instantiations for teams of 32 lanes (SKUs per lane: 1 4 8 16).
#include "env_split.cuh"
#define STEP_CASES \
  MARLSC_SPL_CASE(32, 1, launch_step_t, a, io, t, s) \
  MARLSC_SPL_CASE(32, 4, launch_step_t, a, io, t, s) \
  MARLSC_SPL_CASE(32, 8, launch_step_t, a, io, t, s) \
  MARLSC_SPL_CASE(32, 16, launch_step_t, a, io, t, s) \

#define RESET_CASES \
  MARLSC_SPL_CASE(32, 1, launch_reset_t, a, init, per_env, obs, s) \
  MARLSC_SPL_CASE(32, 4, launch_reset_t, a, init, per_env, obs, s) \
  MARLSC_SPL_CASE(32, 8, launch_reset_t, a, init, per_env, obs, s) \
  MARLSC_SPL_CASE(32, 16, launch_reset_t, a, init, per_env, obs, s) \

MARLSC_DEFINE_G(32, STEP_CASES, RESET_CASES)
MARLSC_DEFINE_SPLIT(32, MARLSC_SPLIT_CASE(32, 1) MARLSC_SPLIT_CASE(32, 4) MARLSC_SPLIT_CASE(32, 8) MARLSC_SPLIT_CASE(32, 16))
