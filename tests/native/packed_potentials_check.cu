// Host-side check of the packed fp32x2 force functors (csrc/potentials.cuh, LJCForce2) against the float64
// closed forms (LJCPot<..., double>) they restate.  Compiled by tests/test_packed_potentials.py with nvcc as a
// HOST program (no GPU needed): the packed value type falls back to component-wise fmaf on the host.
// Prints one line per variant:  name  rms_rel  max_rel   (errors relative to the RMS of the reference).
#include <cmath>
#include <cstdio>
#include <cstring>
#include <random>

#include "../../atomsmm_b200/csrc/potentials.cuh"

template <typename T>
static PotParams<T> params(double rs, double rc, int degree) {
    PotParams<T> p;
    memset(&p, 0, sizeof(p));
    p.kc = T(138.935456);
    p.sign = T(-1);
    p.rs = T(rs);
    p.iw = T(1.0/(rc - rs));
    p.degree = degree;
    p.rsd = T(pow(rs, degree));
    p.iwd = T(1.0/(pow(rc, degree) - pow(rs, degree)));
    p.krf = T(0.4948);
    p.crf = T(1.4922);
    p.alpha = T(2.9);
    p.tasp = T(2.0*2.9/sqrt(M_PI));
    p.inv_rc0 = T(1.0/rc);
    return p;
}

template <int COUL, int LJ, int SW, int SWF, int VAR>
static int check(const char* name, double rs, double rc, int degree, bool zone = false) {
    LJCPot<COUL, LJ, SW, SWF, VAR, double> ref{params<double>(rs, rc, degree)};
    LJCForce2<COUL, LJ, SW, SWF, VAR> packed{params<float>(rs, rc, degree)};
    std::mt19937_64 gen(12345);
    std::uniform_real_distribution<double> ur(zone ? rs : 0.26, rc), uq(-1.0, 1.0), us(0.1, 0.36), ue(0.0, 1.2);
    double sum_ref = 0, sum_err = 0, max_err = 0;
    const int N = 20000;
    for (int k = 0; k < N; k++) {
        double r[2], qq[2], sig[2], eps[2], want[2];
        for (int h = 0; h < 2; h++) {
            r[h] = ur(gen); qq[h] = uq(gen)*uq(gen); sig[h] = us(gen); eps[h] = ue(gen);
            if (sig[h] > 0.9*r[h]) sig[h] = 0.9*r[h];          // stay off the r^-12 wall, as real configurations do
            double rF, e, rinv2;
            ref.template operator()<false>(r[h]*r[h], qq[h], sig[h], eps[h], rF, e, rinv2);
            want[h] = rF*rinv2;
        }
        const float kc = packed.charge_scale(), es = decltype(packed)::EPS_SCALE;
        const F2 got = packed(f2((float)(r[0]*r[0]), (float)(r[1]*r[1])), f2(kc*(float)qq[0], kc*(float)qq[1]),
                              f2((float)sig[0], (float)sig[1]), f2(es*(float)eps[0], es*(float)eps[1]));
        const double g[2] = {got.v.x, got.v.y};
        for (int h = 0; h < 2; h++) {
            sum_ref += want[h]*want[h];
            const double d = g[h] - want[h];
            sum_err += d*d;
            max_err = fmax(max_err, fabs(d));
        }
    }
    const double rms = sqrt(sum_ref/(2*N));
    printf("%s%s %.3e %.3e\n", name, zone ? "@switch" : "", sqrt(sum_err/(2*N))/rms, max_err/rms);
    if (!zone && SW != SW_NONE) check<COUL, LJ, SW, SWF, VAR>(name, rs, rc, degree, true);    // the switching zone alone
    return 0;
}

int main() {
    // the instantiations of DISPATCH in csrc/pair.cu
    check<COUL_PLAIN, LJ_STD, SW_ALL, SWF_LINEAR, VAR_NONE>("near_none", 0.5, 0.7, 1);
    check<COUL_PLAIN, LJ_STD, SW_ALL, SWF_LINEAR, VAR_SHIFT>("near_shift", 0.5, 0.7, 1);
    check<COUL_PLAIN, LJ_STD, SW_ALL, SWF_LINEAR, VAR_FSWITCH>("near_fswitch", 0.5, 0.7, 1);
    check<COUL_ERFC, LJ_STD, SW_ALL, SWF_LINEAR, VAR_NONE>("damped_1", 0.95, 1.0, 1);
    check<COUL_ERFC, LJ_STD, SW_ALL, SWF_POWER, VAR_NONE>("damped_2", 0.95, 1.0, 2);
    check<COUL_ERFC, LJ_STD, SW_ALL, SWF_POWER, VAR_NONE>("damped_3", 0.95, 1.0, 3);
    check<COUL_RF, LJ_STD, SW_NONE, SWF_LINEAR, VAR_NONE>("ljc_rf", 0.9, 1.0, 1);
    check<COUL_ERFC, LJ_STD, SW_NONE, SWF_LINEAR, VAR_NONE>("ljc_erfc", 0.9, 1.0, 1);
    check<COUL_RF, LJ_STD, SW_LJ, SWF_LINEAR, VAR_NONE>("ljc_rf_switch", 0.9, 1.0, 1);
    check<COUL_ERFC, LJ_STD, SW_LJ, SWF_LINEAR, VAR_NONE>("ljc_erfc_switch", 0.9, 1.0, 1);
    check<COUL_PLAIN, LJ_STD, SW_NONE, SWF_LINEAR, VAR_NONE>("ljc_plain", 0.9, 1.0, 1);
    check<COUL_NONE, LJ_VIRIAL, SW_ALL, SWF_LINEAR, VAR_NONE>("lj_virial", 0.9, 1.0, 1);
    return 0;
}
