// TEST INFRASTRUCTURE: runs the statements of csrc/vel_profile_core.cuh (the arithmetic of vel_profile_kernel) on the
// host so that the CPU test-suite can compare them with the numpy oracle without a GPU.  Built by
// tests/test_velprofile_host.py with g++ into a temporary directory; never part of libmincurv_b200.so.
#include <vector>
#include "../../global_racetrajectory_optimization_b200/csrc/vel_profile_core.cuh"

using namespace mc::vp;

extern "C" int vp_host_profile(int n, const double *kappa, const double *el, const double *mu, double scale, double v_max,
                               int n_ggv, const double *ggv, int n_mach, const double *mach, double dyn_model_exp,
                               double drag_coeff, double m_veh, int filt_window, int stride, double *vx, double *ax,
                               double *t, double *laptime) {
    std::vector<double> gv(n_ggv), gax(n_ggv), gay(n_ggv), mv(n_mach), ma(n_mach);
    for (int k = 0; k < n_ggv; ++k) { gv[k] = ggv[3 * k]; gax[k] = ggv[3 * k + 1]; gay[k] = ggv[3 * k + 2]; }
    for (int k = 0; k < n_mach; ++k) { mv[k] = mach[2 * k]; ma[k] = mach[2 * k + 1]; }
    Tables tb{gv.data(), gax.data(), gay.data(), n_ggv, mv.data(), ma.data(), n_mach};
    Params pr{dyn_model_exp, drag_coeff, m_veh, filt_window, VP_DECEL_SLICE_UPPER};
    // same addressing as the kernel: vector k of the profile at ws + k * n * stride + p, element stride `stride`
    const size_t P = (size_t)stride, p = (size_t)(stride - 1), vec = (size_t)n * P;
    std::vector<double> ws(5 * vec, -777.0);
    Strided R{ws.data() + p, P}, EL{ws.data() + vec + p, P}, MU{ws.data() + 2 * vec + p, P}, V{ws.data() + 3 * vec + p, P},
        W{ws.data() + 4 * vec + p, P};
    return profile_thread(n, kappa, el, mu, scale, v_max, tb, pr, R, EL, MU, V, W, vx, ax, t, laptime);
}

extern "C" void vp_host_ax_t(int n, const double *vx, const double *el, const double *ax_in, double t_start, double *ax_out,
                             double *t_out) {
    ax_t_thread(n, vx, el, ax_in, t_start, ax_out, t_out);
}

extern "C" double vp_host_interp(double x, int n, const double *xp, const double *fp, double s, int hint) {
    return interp(x, xp, fp, n, s, hint);
}
