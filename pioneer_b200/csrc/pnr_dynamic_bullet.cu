// K2 with the opt-in Bullet-like substep (PNR_STEPPING_BULLET, pnr_dynamics.cuh::pnr_bullet_substep): its own translation
// unit so that these instantiations compile beside the explicit ones.  The substep runs on the general ABA: run-time axis
// codes, or the shipped robot's axis structure (the isotropic-link specialisation maps to the latter).
#include "pnr_dynamic_kernel.cuh"

PnrDynKernel pnr_dynamic_bullet_kernel(int chain, int obs_mode, int obst) {
#define PNR_BULLET_ROW(CH) \
    {{pnr_step_dynamic_kernel<PNR_OBS_TERMINAL, false, CH, PNR_STEPPING_BULLET>, \
      pnr_step_dynamic_kernel<PNR_OBS_TERMINAL, true, CH, PNR_STEPPING_BULLET>}, \
     {pnr_step_dynamic_kernel<PNR_OBS_AUTORESET, false, CH, PNR_STEPPING_BULLET>, \
      pnr_step_dynamic_kernel<PNR_OBS_AUTORESET, true, CH, PNR_STEPPING_BULLET>}}
    static PnrDynKernel kernels[2][2][2] = {PNR_BULLET_ROW(PNR_CHAIN_GENERIC), PNR_BULLET_ROW(PNR_CHAIN_PIONEER)};
    return kernels[chain == PNR_CHAIN_GENERIC ? 0 : 1][obs_mode][obst];
}
