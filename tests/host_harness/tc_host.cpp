// TEST INFRASTRUCTURE: the statements of csrc/traj_check_core.cuh (arithmetic of the K6 kernels) run on the host in the
// same order as the kernels run them, for comparison with the fixtures produced by the reference's own helpers
// (tests/golden/refback_*.npz).  Built by tests/test_refback_host.py with g++; never part of libmincurv_b200.so.
#include <cmath>
#include <vector>
#include "../../global_racetrajectory_optimization_b200/csrc/traj_check_core.cuh"

using namespace mc::tc;

// interp_track_kernel, one track
extern "C" int tc_host_interp_track(int n, const double *pts, int stride, const double *normvec, double sign, int width_col,
                                    double stepsize, int n_out_max, double *out) {
    std::vector<double> dc(n + 1);
    ClosedCol cx{pts, (size_t)stride, normvec, 2, pts + width_col, (size_t)stride, sign, n};
    ClosedCol cy{pts + 1, (size_t)stride, normvec ? normvec + 1 : nullptr, 2, pts + width_col, (size_t)stride, sign, n};
    for (int i = 0; i < n; ++i) dc[i + 1] = closed_el_length(cx, cy, i);
    double acc = 0.0;
    dc[0] = 0.0;
    for (int i = 1; i <= n; ++i) { acc = add(acc, dc[i]); dc[i] = acc; }
    const double total = dc[n];
    const int num = resample_count(total, stepsize);
    const int m = num - 1;
    if (!(total > 0.0)) return 0;
    if (m > n_out_max) return -m;
    const bool plain4 = (normvec == nullptr && stride == 4);
    ClosedCol c2{pts + 2, (size_t)stride, nullptr, 0, nullptr, 0, 0.0, n};
    ClosedCol c3{pts + 3, (size_t)stride, nullptr, 0, nullptr, 0, 0.0, n};
    for (int j = 0; j < m; ++j) {
        const double d = linspace0(j, num, total);
        out[4 * j] = interp_closed(d, dc.data(), cx);
        out[4 * j + 1] = interp_closed(d, dc.data(), cy);
        out[4 * j + 2] = plain4 ? interp_closed(d, dc.data(), c2) : 0.0;
        out[4 * j + 3] = plain4 ? interp_closed(d, dc.data(), c3) : 0.0;
    }
    return m;
}

// min_bound_dists_kernel, one track
extern "C" void tc_host_min_bound_dists(int nt, const double *xy, const double *psi, int nb1, const double *b1, int nb2,
                                        const double *b2, int bstride, double length_veh, double width_veh, double *out) {
    for (int i = 0; i < nt; ++i) {
        double c[8];
        vehicle_corners(xy[2 * i], xy[2 * i + 1], psi[i], length_veh, width_veh, c);
        double best = INFINITY;
        for (int side = 0; side < 2; ++side) {
            const double *bd = side ? b2 : b1;
            const int nb = side ? nb2 : nb1;
            for (int k = 0; k < nb; ++k) {
                const double bx = bd[(size_t)k * bstride], by = bd[(size_t)k * bstride + 1];
                const double d0 = dist2(bx, by, c[0], c[1]), d1 = dist2(bx, by, c[2], c[3]);
                const double d2 = dist2(bx, by, c[4], c[5]), d3 = dist2(bx, by, c[6], c[7]);
                best = fmin(best, fmin(fmin(d0, d1), fmin(d2, d3)));
            }
        }
        out[i] = sqrt(best);
    }
}

// traj_extrema_kernel, one track
extern "C" void tc_host_extrema(int n, const double *kappa, const double *vx, const double *ax, const double *min_dists,
                                double dragcoeff, double mass_veh, double *e) {
    double mn_d = INFINITY, mx_k = -INFINITY, mx_ay = -INFINITY, mx_ax = -INFINITY, mn_ax = INFINITY, mx_at = -INFINITY,
           mx_v = -INFINITY;
    for (int i = 0; i < n; ++i) {
        const PointChecks q = point_checks(kappa[i], vx[i], ax[i], dragcoeff, mass_veh);
        if (min_dists) mn_d = fmin(mn_d, min_dists[i]);
        mx_k = fmax(mx_k, q.kappa_abs); mx_ay = fmax(mx_ay, q.ay); mx_ax = fmax(mx_ax, q.ax_wo_drag);
        mn_ax = fmin(mn_ax, q.ax_wo_drag); mx_at = fmax(mx_at, q.a_tot); mx_v = fmax(mx_v, q.v);
    }
    e[0] = mn_d; e[1] = mx_k; e[2] = mx_ay; e[3] = mx_ax; e[4] = mn_ax; e[5] = mx_at; e[6] = mx_v; e[7] = (double)n;
}

// normals_crossing_kernel, one track
extern "C" int tc_host_normals_crossing(int n, const double *track, const double *nrm, int horizon) {
    for (int i = 0; i < n; ++i)
        if (normals_cross_point(i, n, horizon, track, nrm)) return 1;
    return 0;
}
