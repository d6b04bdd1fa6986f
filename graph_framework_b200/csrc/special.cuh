//------------------------------------------------------------------------------
//  special.cuh -- erfi(x) for real x, the only special function on the absorption path.
//
//  The reference evaluates the plasma dispersion function through erfi
//  (/root/reference/graph_framework/dispersion.hpp:289-305) with the Faddeeva-package algorithms in
//  /root/reference/graph_framework/special_functions.hpp.  Ray state is real, so only the branch
//  for purely real arguments is ever taken (special_functions.hpp:1504-1512 via :1583-1586):
//      erfi(x) = x^2 > 720 ? +-DBL_MAX : exp(x^2) w_im(x),
//  with w_im(x) = Im w(x) = exp(-x^2) erfi(x) (special_functions.hpp:545-562): continued-fraction
//  tails for |x| > 45, a Taylor series for |x| < 0.0309 and, in between, one polynomial per unit
//  interval of y100 = 100/(1 + |x|).  The polynomials here are this repository's own fits
//  (tools/make_wim_table.py), not the reference's coefficients; both approximate the same function
//  of the same y100 to ~1e-16.
//
//  The file compiles in three places: appended to the skeleton under NVRTC when a kernel calls
//  gfb::erfi, in the g++ CPU harness, and in the host front end (constant folding).
//------------------------------------------------------------------------------
#ifndef gfb_special_cuh
#define gfb_special_cuh

#if defined(__CUDACC_RTC__) || defined(__CUDACC__)
#define GFB_SPECIAL_FN __device__ __forceinline__
#define GFB_SPECIAL_TABLE __device__ const
#else
#include <cmath>
#define GFB_SPECIAL_FN inline
#define GFB_SPECIAL_TABLE static const
#endif

namespace gfb {
    GFB_SPECIAL_TABLE double wim_table[97][10] = {
#include "wim_table.inc"
    };

    GFB_SPECIAL_FN double w_im(const double x) {
        const double ax = fabs(x);
        const double x2 = ax*ax;
        double r;
        if (!(ax <= 45.0)) {            // also taken by NaN, which must not reach the table index
//  1/sqrt(pi) times the 1-term or 5-term continued fraction (special_functions.hpp:547-557).
            r = ax > 5.0e7 ? 0.56418958354775628695/ax :
                0.56418958354775628695*(x2*(x2 - 4.5) + 2.0)/(ax*(x2*(x2 - 5.0) + 3.75));
        } else {
            const double y100 = 100.0/(1.0 + ax);
            if (y100 >= 97.0) {
//  (2/sqrt(pi)) (x - 2/3 x^3 + 4/15 x^5 - 8/105 x^7 + 16/945 x^9), special_functions.hpp:496-510.
                r = ax*(1.1283791670955125739 - x2*(0.75225277806367504925 - x2*(0.30090111122547001970 -
                        x2*(0.085971746064420005629 - x2*0.019104832458760001251))));
            } else {
                const int k = static_cast<int> (y100);
                const double t = 2.0*y100 - static_cast<double> (2*k + 1);
                const double *c = wim_table[k];
                r = c[9];
#pragma unroll
                for (int j = 8; j >= 0; j--) r = fma(r, t, c[j]);
            }
        }
        return copysign(r, x);
    }

    GFB_SPECIAL_FN double erfi(const double x) {
        const double x2 = x*x;
        return x2 > 720.0 ? copysign(1.7976931348623157e308, x) : exp(x2)*w_im(x);
    }
}
#endif /* gfb_special_cuh */
