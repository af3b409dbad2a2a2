// env_inst_g16.cu - K1 instantiations for teams of 16 lanes (SKUs per lane: 1 4 8).
#include "env_split.cuh"
#define STEP_CASES \
  MARLSC_SPL_CASE(16, 1, launch_step_t, a, io, t, s) \
  MARLSC_SPL_CASE(16, 4, launch_step_t, a, io, t, s) \
  MARLSC_SPL_CASE(16, 8, launch_step_t, a, io, t, s) \

#define RESET_CASES \
  MARLSC_SPL_CASE(16, 1, launch_reset_t, a, init, per_env, obs, s) \
  MARLSC_SPL_CASE(16, 4, launch_reset_t, a, init, per_env, obs, s) \
  MARLSC_SPL_CASE(16, 8, launch_reset_t, a, init, per_env, obs, s) \

MARLSC_DEFINE_G(16, STEP_CASES, RESET_CASES)
MARLSC_DEFINE_SPLIT(16, MARLSC_SPLIT_CASE(16, 1) MARLSC_SPLIT_CASE(16, 4))
