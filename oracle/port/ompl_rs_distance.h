// TEST INFRASTRUCTURE ONLY — CPU oracle. Never included, linked or executed by the product path.
//
// Restatement of ompl::base::ReedsSheppStateSpace::distance (OMPL 1.5.x, un-vendored and un-pinned
// third-party dependency of the reference: mpros_utils/include/mpros_utils/rs_statespace.h:23,34,52;
// called from hybrid_a_star.cpp:145,258-260). Algorithm: Reeds & Shepp 1990 formulas 8.1-8.11 as OMPL
// organises them (src/ompl/base/spaces/src/ReedsSheppStateSpace.cpp: CSC, CCC, CCCC, CCSC, CCSCC),
// see SURVEY.md Appendix A.
//
// PARITY UNPINNED: OMPL is absent from /root/reference and from this image, the reference holds no
// test vector for it. The only available pin is a cross-check against the reference's own RS::RSCurve
// with zero penalties (tests/test_oracle_port.py (test_ompl_standin_*)).
#pragma once
#include <cfloat>
#include <cmath>

namespace ompl_rs
{
    static const double kPi = 3.14159265358979323846;
    static const double kTwoPi = 2.0 * 3.14159265358979323846;
    static const double kZero = 10.0 * DBL_EPSILON;

    inline double mod2pi(double x)
    {
        double v = std::fmod(x, kTwoPi);
        if (v < -kPi)
            v += kTwoPi;
        else if (v > kPi)
            v -= kTwoPi;
        return v;
    }
    inline void polar(double x, double y, double &r, double &theta)
    {
        r = std::sqrt(x * x + y * y);
        theta = std::atan2(y, x);
    }
    inline void tauOmega(double u, double v, double xi, double eta, double phi, double &tau, double &omega)
    {
        double delta = mod2pi(u - v), A = std::sin(u) - std::sin(delta), B = std::cos(u) - std::cos(delta) - 1.;
        double t1 = std::atan2(eta * A - xi * B, xi * A + eta * B), t2 = 2. * (std::cos(delta) - std::cos(v) - std::cos(u)) + 3;
        tau = (t2 < 0) ? mod2pi(t1 + kPi) : mod2pi(t1);
        omega = mod2pi(tau - u + v - phi);
    }
    // formula 8.1
    inline bool LpSpLp(double x, double y, double phi, double &t, double &u, double &v)
    {
        polar(x - std::sin(phi), y - 1. + std::cos(phi), u, t);
        if (t >= -kZero)
        {
            v = mod2pi(phi - t);
            if (v >= -kZero)
                return true;
        }
        return false;
    }
    // formula 8.2
    inline bool LpSpRp(double x, double y, double phi, double &t, double &u, double &v)
    {
        double t1, u1;
        polar(x + std::sin(phi), y - 1. - std::cos(phi), u1, t1);
        u1 = u1 * u1;
        if (u1 >= 4.)
        {
            u = std::sqrt(u1 - 4.);
            double theta = std::atan2(2., u);
            t = mod2pi(t1 + theta);
            v = mod2pi(t - phi);
            return t >= -kZero && v >= -kZero;
        }
        return false;
    }
    // formula 8.3 / 8.4
    inline bool LpRmL(double x, double y, double phi, double &t, double &u, double &v)
    {
        double xi = x - std::sin(phi), eta = y - 1. + std::cos(phi), u1, theta;
        polar(xi, eta, u1, theta);
        if (u1 <= 4.)
        {
            u = -2. * std::asin(.25 * u1);
            t = mod2pi(theta + .5 * u + kPi);
            v = mod2pi(phi - t + u);
            return t >= -kZero && u <= kZero;
        }
        return false;
    }
    // formula 8.7
    inline bool LpRupLumRm(double x, double y, double phi, double &t, double &u, double &v)
    {
        double xi = x + std::sin(phi), eta = y - 1. - std::cos(phi), rho = .25 * (2. + std::sqrt(xi * xi + eta * eta));
        if (rho <= 1.)
        {
            u = std::acos(rho);
            tauOmega(u, -u, xi, eta, phi, t, v);
            return t >= -kZero && v <= kZero;
        }
        return false;
    }
    // formula 8.8
    inline bool LpRumLumRp(double x, double y, double phi, double &t, double &u, double &v)
    {
        double xi = x + std::sin(phi), eta = y - 1. - std::cos(phi), rho = (20. - xi * xi - eta * eta) / 16.;
        if (rho >= 0 && rho <= 1)
        {
            u = -std::acos(rho);
            if (u >= -.5 * kPi)
            {
                tauOmega(u, u, xi, eta, phi, t, v);
                return t >= -kZero && v >= -kZero;
            }
        }
        return false;
    }
    // formula 8.9
    inline bool LpRmSmLm(double x, double y, double phi, double &t, double &u, double &v)
    {
        double xi = x - std::sin(phi), eta = y - 1. + std::cos(phi), rho, theta;
        polar(xi, eta, rho, theta);
        if (rho >= 2.)
        {
            double r = std::sqrt(rho * rho - 4.);
            u = 2. - r;
            t = mod2pi(theta + std::atan2(r, -2.));
            v = mod2pi(phi - .5 * kPi - t);
            return t >= -kZero && u <= kZero && v <= kZero;
        }
        return false;
    }
    // formula 8.10
    inline bool LpRmSmRm(double x, double y, double phi, double &t, double &u, double &v)
    {
        double xi = x + std::sin(phi), eta = y - 1. - std::cos(phi), rho, theta;
        polar(-eta, xi, rho, theta);
        if (rho >= 2.)
        {
            t = theta;
            u = 2. - rho;
            v = mod2pi(t + .5 * kPi - phi);
            return t >= -kZero && u <= kZero && v <= kZero;
        }
        return false;
    }
    // formula 8.11
    inline bool LpRmSLmRp(double x, double y, double phi, double &t, double &u, double &v)
    {
        double xi = x + std::sin(phi), eta = y - 1. - std::cos(phi), rho, theta;
        polar(xi, eta, rho, theta);
        if (rho >= 2.)
        {
            u = 4. - std::sqrt(rho * rho - 4.);
            if (u <= kZero)
            {
                t = mod2pi(std::atan2((4. - u) * xi - 2. * eta, -2. * xi + (u - 4.) * eta));
                v = mod2pi(t - phi);
                return t >= -kZero && v >= -kZero;
            }
        }
        return false;
    }

    typedef bool (*Formula)(double, double, double, double &, double &, double &);

    // try one base formula on (x,y,phi) and its timeflip / reflect / both images; L = |t|+|u|*umul+|v|
    inline void tryFour(Formula f, double x, double y, double phi, double umul, double &Lmin)
    {
        double t, u, v, L;
        if (f(x, y, phi, t, u, v) && Lmin > (L = std::fabs(t) + umul * std::fabs(u) + std::fabs(v)))
            Lmin = L;
        if (f(-x, y, -phi, t, u, v) && Lmin > (L = std::fabs(t) + umul * std::fabs(u) + std::fabs(v)))
            Lmin = L;
        if (f(x, -y, -phi, t, u, v) && Lmin > (L = std::fabs(t) + umul * std::fabs(u) + std::fabs(v)))
            Lmin = L;
        if (f(-x, -y, phi, t, u, v) && Lmin > (L = std::fabs(t) + umul * std::fabs(u) + std::fabs(v)))
            Lmin = L;
    }

    // shortest Reeds-Shepp length for the normalised goal (x, y, phi), unit turning radius
    inline double reedsSheppLength(double x, double y, double phi)
    {
        double best = DBL_MAX; // length of OMPL's default-constructed path
        double xb = x * std::cos(phi) + y * std::sin(phi), yb = x * std::sin(phi) - y * std::cos(phi);
        // CSC
        {
            double Lmin = best;
            tryFour(LpSpLp, x, y, phi, 1., Lmin);
            tryFour(LpSpRp, x, y, phi, 1., Lmin);
            best = Lmin;
        }
        // CCC
        {
            double Lmin = best;
            tryFour(LpRmL, x, y, phi, 1., Lmin);
            tryFour(LpRmL, xb, yb, phi, 1., Lmin);
            best = Lmin;
        }
        // CCCC
        {
            double Lmin = best;
            tryFour(LpRupLumRm, x, y, phi, 2., Lmin);
            tryFour(LpRumLumRp, x, y, phi, 2., Lmin);
            best = Lmin;
        }
        // CCSC (fixed quarter turn)
        {
            double Lmin = best - .5 * kPi, L0 = Lmin;
            tryFour(LpRmSmLm, x, y, phi, 1., Lmin);
            tryFour(LpRmSmRm, x, y, phi, 1., Lmin);
            tryFour(LpRmSmLm, xb, yb, phi, 1., Lmin);
            tryFour(LpRmSmRm, xb, yb, phi, 1., Lmin);
            if (Lmin < L0)
                best = Lmin + .5 * kPi;
        }
        // CCSCC (two fixed quarter turns)
        {
            double Lmin = best - kPi, L0 = Lmin;
            tryFour(LpRmSLmRp, x, y, phi, 1., Lmin);
            if (Lmin < L0)
                best = Lmin + kPi;
        }
        return best;
    }

    // ompl::base::ReedsSheppStateSpace(rho).distance(s1, s2)
    inline double distance(double rho, double x1, double y1, double th1, double x2, double y2, double th2)
    {
        double dx = x2 - x1, dy = y2 - y1, dth = th2 - th1;
        double c = std::cos(th1), s = std::sin(th1);
        double x = c * dx + s * dy, y = -s * dx + c * dy;
        return rho * reedsSheppLength(x / rho, y / rho, dth);
    }
}
