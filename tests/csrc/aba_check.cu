// Host-side check of pioneer_b200/csrc/pnr_dynamics.cuh (compiled with nvcc, run on the CPU): the float32 articulated-
// body algorithm the dynamic-mode kernel runs, in its three variants, for tests/test_dynamics_host.py to compare with
// the float64 oracle.  `params` is the PnrParams block from pnr_debug_build_params (libpioneer_b200.so).
//   variant 0: pnr_aba_general<GENERIC> (run-time axis codes)   1: pnr_aba_general<PIONEER>   2: pnr_aba_pioneer<false>
//           3: pnr_aba_pioneer<true> (isotropic stub inertials; only valid when PnrParams::dyn_iso_links is set)
#include <cstring>
#include "../../pioneer_b200/csrc/pnr_dynamics.cuh"

extern "C" long aba_params_size() { return (long)sizeof(PnrParams); }

extern "C" int aba_check(const void* params, int variant, long n, const float* q, const float* qd, const float* tau,
                         float* qdd) {
    PnrParams p;
    std::memcpy(&p, params, sizeof(p));
    for (long e = 0; e < n; ++e) {
        float q_[PNR_DOF], qd_[PNR_DOF], tau_[PNR_DOF], out[PNR_DOF];
        for (int i = 0; i < PNR_DOF; ++i) { q_[i] = q[e * PNR_DOF + i]; qd_[i] = qd[e * PNR_DOF + i]; tau_[i] = tau[e * PNR_DOF + i]; }
        if (variant == 0) pnr_aba_general<PNR_CHAIN_GENERIC>(p, q_, qd_, tau_, out);
        else if (variant == 1) pnr_aba_general<PNR_CHAIN_PIONEER>(p, q_, qd_, tau_, out);
        else if (variant == 2) pnr_aba_pioneer<false>(p, q_, qd_, tau_, out);
        else if (variant == 3 && p.dyn_iso_links) pnr_aba_pioneer<true>(p, q_, qd_, tau_, out);
        else return -1;
        for (int i = 0; i < PNR_DOF; ++i) qdd[e * PNR_DOF + i] = out[i];
    }
    return 0;
}

// frame_skip substeps in place (control, ABA, semi-implicit Euler, limit stops): what one env step of the kernel does
extern "C" int aba_substeps(const void* params, int chain, long n, float* q, float* qd, const float* action) {
    PnrParams p;
    std::memcpy(&p, params, sizeof(p));
    for (long e = 0; e < n; ++e) {
        float q_[PNR_DOF], qd_[PNR_DOF], a_[PNR_DOF];
        for (int i = 0; i < PNR_DOF; ++i) { q_[i] = q[e * PNR_DOF + i]; qd_[i] = qd[e * PNR_DOF + i]; a_[i] = action[e * PNR_DOF + i]; }
        if (p.dyn_stepping == PNR_STEPPING_BULLET) {            // the Bullet-like substep runs on the general ABA
            if (chain == PNR_CHAIN_GENERIC) pnr_dynamic_substeps<PNR_CHAIN_GENERIC, PNR_STEPPING_BULLET>(p, q_, qd_, a_);
            else pnr_dynamic_substeps<PNR_CHAIN_PIONEER, PNR_STEPPING_BULLET>(p, q_, qd_, a_);
        }
        else if (chain == PNR_CHAIN_PIONEER_ISO && p.dyn_iso_links) pnr_dynamic_substeps<PNR_CHAIN_PIONEER_ISO>(p, q_, qd_, a_);
        else if (chain == PNR_CHAIN_PIONEER) pnr_dynamic_substeps<PNR_CHAIN_PIONEER>(p, q_, qd_, a_);
        else if (chain == PNR_CHAIN_GENERIC) pnr_dynamic_substeps<PNR_CHAIN_GENERIC>(p, q_, qd_, a_);
        else return -1;
        for (int i = 0; i < PNR_DOF; ++i) { q[e * PNR_DOF + i] = q_[i]; qd[e * PNR_DOF + i] = qd_[i]; }
    }
    return 0;
}
