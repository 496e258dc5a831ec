// Derivative-free adaptive rejection sampling (Gilks, Best & Tan 1995) as hydra's BayesW uses it
// (reference src/BayesW_arms.cpp:135-918 with ninit=4, npoint=100, nsamp=1, convex=1, dometrop=0:
// src/BayesW.cpp:1336-1343): log-concave densities only, no Metropolis step, one sample per call.
//
// Restated for device and host: the envelope is an index-linked array in the caller's frame (no malloc,
// no pointers), the log-density is a functor, the uniforms come from a functor (the reference calls libc
// rand(): u = (rand()+0.5)/(RAND_MAX+1), :913-918).  The arithmetic of every step follows the reference
// expression by expression, so that the same uniforms give the same sample.
#pragma once
#include <math.h>
#include <stdint.h>

#ifndef __CUDACC__
#define HB_HD
#else
#define HB_HD __host__ __device__
#endif

namespace hb {

constexpr int kArmsPoints = 100;       // npoint
constexpr double kArmsXEPS = 0.00001;  // critical relative x-value difference      (:42)
constexpr double kArmsYEPS = 0.1;      // critical y-value difference               (:43)
constexpr double kArmsEYEPS = 0.001;   // critical relative exp(y) difference       (:44)
constexpr double kArmsYCEIL = 50.0;    // maximum y avoiding overflow in exp(y)     (:45)

enum ArmsError : int {
    ARMS_OK = 0, ARMS_FEW_INIT = 1001, ARMS_MANY_INIT = 1002, ARMS_BOUNDS = 1003, ARMS_ORDER = 1004,
    ARMS_VIOLATION = 2000,  // envelope violation without metropolis: the density is not log-concave
    ARMS_INTERNAL = 3000    // the reference calls exit() in these places
};

struct ArmsEnvelope {
    double x[kArmsPoints], y[kArmsPoints], ey[kArmsPoints], cum[kArmsPoints];
    int8_t f[kArmsPoints];            // is y an evaluated point of the log-density
    int8_t pl[kArmsPoints], pr[kArmsPoints];  // neighbours, -1 = none
    int cpoint;
    double ymax;
    int neval;
};

HB_HD inline double arms_expshift(double y, double y0) { return (y - y0 > -2.0 * kArmsYCEIL) ? exp(y - y0 + kArmsYCEIL) : 0.0; }  // :836-844
HB_HD inline double arms_logshift(double y, double y0) { return log(y) + y0 - kArmsYCEIL; }                                        // :846-854

// intersection of the chords around point q (:693-806). Returns 0, ARMS_VIOLATION or ARMS_INTERNAL.
HB_HD inline int arms_meet(ArmsEnvelope &e, int q) {
    if (e.f[q]) return ARMS_INTERNAL;
    double gl = 0, gr = 0, grl = 0, dl = 0, dr = 0;
    const int l = e.pl[q], r = e.pr[q];
    const int lll = (l >= 0 && e.pl[l] >= 0) ? e.pl[e.pl[l]] : -1;
    const int rrr = (r >= 0 && e.pr[r] >= 0) ? e.pr[e.pr[r]] : -1;
    const bool il = (l >= 0) && (lll >= 0), ir = (r >= 0) && (rrr >= 0), irl = (l >= 0) && (r >= 0);
    if (il) gl = (e.y[l] - e.y[lll]) / (e.x[l] - e.x[lll]);
    if (ir) gr = (e.y[r] - e.y[rrr]) / (e.x[r] - e.x[rrr]);
    if (irl) grl = (e.y[r] - e.y[l]) / (e.x[r] - e.x[l]);
    if (irl && il && (gl < grl)) return ARMS_VIOLATION;  // convexity on the left, no metropolis
    if (irl && ir && (gr > grl)) return ARMS_VIOLATION;
    if (il && irl) {
        dr = (gl - grl) * (e.x[r] - e.x[l]);
        if (dr < kArmsYEPS) dr = kArmsYEPS;
    }
    if (ir && irl) {
        dl = (grl - gr) * (e.x[r] - e.x[l]);
        if (dl < kArmsYEPS) dl = kArmsYEPS;
    }
    if (il && ir && irl) {
        e.x[q] = (dl * e.x[r] + dr * e.x[l]) / (dl + dr);
        e.y[q] = (dl * e.y[r] + dr * e.y[l] + dl * dr) / (dl + dr);
    } else if (il && irl) {
        e.x[q] = e.x[r];
        e.y[q] = e.y[r] + dr;
    } else if (ir && irl) {
        e.x[q] = e.x[l];
        e.y[q] = e.y[l] + dl;
    } else if (il) {
        e.y[q] = e.y[l] + gl * (e.x[q] - e.x[l]);
    } else if (ir) {
        e.y[q] = e.y[r] - gr * (e.x[r] - e.x[q]);
    } else {
        return ARMS_INTERNAL;
    }
    if ((l >= 0 && e.x[q] < e.x[l]) || (r >= 0 && e.x[q] > e.x[r])) return ARMS_INTERNAL;
    return 0;
}

// exponentiate and integrate the envelope (:659-689, area :810-832)
HB_HD inline void arms_cumulate(ArmsEnvelope &e) {
    int lm = 0;
    while (e.pl[lm] >= 0) lm = e.pl[lm];
    e.ymax = e.y[lm];
    for (int q = e.pr[lm]; q >= 0; q = e.pr[q])
        if (e.y[q] > e.ymax) e.ymax = e.y[q];
    for (int q = lm; q >= 0; q = e.pr[q]) e.ey[q] = arms_expshift(e.y[q], e.ymax);
    e.cum[lm] = 0.0;
    for (int q = e.pr[lm]; q >= 0; q = e.pr[q]) {
        const int l = e.pl[q];
        double a;
        if (e.x[l] == e.x[q]) a = 0.0;
        else if (fabs(e.y[q] - e.y[l]) < kArmsYEPS) a = 0.5 * (e.ey[q] + e.ey[l]) * (e.x[q] - e.x[l]);
        else a = ((e.ey[q] - e.ey[l]) / (e.y[q] - e.y[l])) * (e.x[q] - e.x[l]);
        e.cum[q] = e.cum[l] + a;
    }
}

struct ArmsSerialCumulate {  // default integration of the envelope: the reference's serial loops
    HB_HD void operator()(ArmsEnvelope &e) const { arms_cumulate(e); }
};

struct ArmsPoint {
    double x, y, ey;
    int pl, pr;
};

// x-value at cumulative probability prob under the envelope (:370-445)
HB_HD inline void arms_invert(const ArmsEnvelope &e, double prob, ArmsPoint &p) {
    int q = 0;
    while (e.pr[q] >= 0) q = e.pr[q];
    const double u = prob * e.cum[q];
    while (e.cum[e.pl[q]] > u) q = e.pl[q];
    const int l = e.pl[q];
    p.pl = l;
    p.pr = q;
    const double prop = (u - e.cum[l]) / (e.cum[q] - e.cum[l]);
    if (e.x[l] == e.x[q]) {
        p.x = e.x[q]; p.y = e.y[q]; p.ey = e.ey[q];
    } else {
        const double xl = e.x[l], xr = e.x[q], yl = e.y[l], yr = e.y[q], eyl = e.ey[l], eyr = e.ey[q];
        if (fabs(yr - yl) < kArmsYEPS) {
            if (fabs(eyr - eyl) > kArmsEYEPS * fabs(eyr + eyl))
                p.x = xl + ((xr - xl) / (eyr - eyl)) * (-eyl + sqrt((1. - prop) * eyl * eyl + prop * eyr * eyr));
            else
                p.x = xl + (xr - xl) * prop;
            p.ey = ((p.x - xl) / (xr - xl)) * (eyr - eyl) + eyl;
            p.y = arms_logshift(p.ey, e.ymax);
        } else {
            p.x = xl + ((xr - xl) / (yr - yl)) * (-yl + arms_logshift(((1. - prop) * eyl + prop * eyr), e.ymax));
            p.y = ((p.x - xl) / (xr - xl)) * (yr - yl) + yl;
            p.ey = arms_expshift(p.y, e.ymax);
        }
    }
}

// One sample from the density exp(logdens(x)) on [xl, xr], starting abscissae xinit[0..ninit) ascending.
// logdens: double(double); urand: double() in (0,1). Returns ARMS_OK or an error code; *xsamp holds the sample.
// cumulate: how the envelope is exponentiated and integrated after every change (the device passes a warp-cooperative
// functor that computes the same numbers, one envelope point per lane).
template <class LogDens, class URand, class Cumulate = ArmsSerialCumulate>
HB_HD inline int arms_sample(const double *xinit, int ninit, double xl, double xr, LogDens &logdens, URand &urand, double *xsamp,
                             ArmsEnvelope &e, Cumulate cumulate = Cumulate()) {
    // ---- initial envelope (:247-354)
    if (ninit < 3) return ARMS_FEW_INIT;
    const int mpoint = 2 * ninit + 1;
    if (kArmsPoints < mpoint) return ARMS_MANY_INIT;
    if ((xinit[0] <= xl) || (xinit[ninit - 1] >= xr)) return ARMS_BOUNDS;
    for (int i = 1; i < ninit; i++)
        if (xinit[i] <= xinit[i - 1]) return ARMS_ORDER;
    e.neval = 0;
    for (int j = 0, k = 0; j < mpoint; j++) {
        e.f[j] = 0; e.y[j] = 0.0; e.x[j] = 0.0;
        e.pl[j] = (int8_t)(j - 1);
        e.pr[j] = (int8_t)((j + 1 < mpoint) ? j + 1 : -1);
        if (j % 2) {  // point on the log density
            e.x[j] = xinit[k++];
            e.y[j] = logdens(e.x[j]);
            e.neval++;
            e.f[j] = 1;
        }
    }
    e.x[0] = xl;
    e.x[mpoint - 1] = xr;
    for (int j = 0; j < mpoint; j += 2) {
        const int rc = arms_meet(e, j);
        if (rc) return rc;
    }
    cumulate(e);
    e.cpoint = mpoint;
    // ---- adaptive rejection (:207-225): sample (:358-368), test (:449-548), update (:552-655)
    for (;;) {
        ArmsPoint p;
        arms_invert(e, urand(), p);
        const double u = urand() * p.ey;
        const double y = arms_logshift(u, e.ymax);
        if (e.pl[p.pl] >= 0 && e.pr[p.pr] >= 0) {  // squeezing test
            const int ql = e.f[p.pl] ? p.pl : e.pl[p.pl];
            const int qr = e.f[p.pr] ? p.pr : e.pr[p.pr];
            const double ysqueez = (e.y[qr] * (p.x - e.x[ql]) + e.y[ql] * (e.x[qr] - p.x)) / (e.x[qr] - e.x[ql]);
            if (y <= ysqueez) { *xsamp = p.x; return ARMS_OK; }
        }
        const double ynew = logdens(p.x);
        e.neval++;
        // update the envelope with the new point on the log density
        if (!(e.cpoint > kArmsPoints - 2)) {
            const int q = e.cpoint++, m = e.cpoint++;
            e.x[q] = p.x; e.y[q] = ynew; e.f[q] = 1;
            e.f[m] = 0; e.x[m] = 0.0; e.y[m] = 0.0;
            if (e.f[p.pl] && !e.f[p.pr]) {         // new intersection between p.pl and p
                e.pl[m] = (int8_t)p.pl; e.pr[m] = (int8_t)q;
                e.pl[q] = (int8_t)m; e.pr[q] = (int8_t)p.pr;
                e.pr[p.pl] = (int8_t)m; e.pl[p.pr] = (int8_t)q;
            } else if (!e.f[p.pl] && e.f[p.pr]) {  // new intersection between p and p.pr
                e.pr[m] = (int8_t)p.pr; e.pl[m] = (int8_t)q;
                e.pr[q] = (int8_t)m; e.pl[q] = (int8_t)p.pl;
                e.pl[p.pr] = (int8_t)m; e.pr[p.pl] = (int8_t)q;
            } else {
                return ARMS_INTERNAL;
            }
            // adjust the position of q if too close to an end point
            const int ql = (e.pl[e.pl[q]] >= 0) ? e.pl[e.pl[q]] : e.pl[q];
            const int qr = (e.pr[e.pr[q]] >= 0) ? e.pr[e.pr[q]] : e.pr[q];
            if (e.x[q] < (1. - kArmsXEPS) * e.x[ql] + kArmsXEPS * e.x[qr]) {
                e.x[q] = (1. - kArmsXEPS) * e.x[ql] + kArmsXEPS * e.x[qr];
                e.y[q] = logdens(e.x[q]);
                e.neval++;
            } else if (e.x[q] > kArmsXEPS * e.x[ql] + (1. - kArmsXEPS) * e.x[qr]) {
                e.x[q] = kArmsXEPS * e.x[ql] + (1. - kArmsXEPS) * e.x[qr];
                e.y[q] = logdens(e.x[q]);
                e.neval++;
            }
            int rc = arms_meet(e, e.pl[q]);
            if (rc) return rc;
            rc = arms_meet(e, e.pr[q]);
            if (rc) return rc;
            if (e.pl[e.pl[q]] >= 0) {
                rc = arms_meet(e, e.pl[e.pl[e.pl[q]]]);
                if (rc) return rc;
            }
            if (e.pr[e.pr[q]] >= 0) {
                rc = arms_meet(e, e.pr[e.pr[e.pr[q]]]);
                if (rc) return rc;
            }
            cumulate(e);
        }
        if (!(y >= ynew)) { *xsamp = p.x; return ARMS_OK; }  // accepted at the rejection step
    }
}

}  // namespace hb
