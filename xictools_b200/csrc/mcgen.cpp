// mcgen.cpp -- Monte-Carlo trial generation of the dispatcher (host code):
// the deviates a WRspice Monte-Carlo deck draws with gauss() / unif() come from sRnd
// (xt_base/include/miscutil/random.h:47-105, xt_base/miscutil/random.cc:51-72), whose
// uniform integers are the additive-feedback generator of BSD random(3), "TYPE_3"
// (x**31 + x**3 + 1), as xt_base/miscutil/randval.cc:207-262 runs it: seeding fills the 31
// state words with the Park-Miller minimal-standard LCG (16807 mod 2^31-1, Schrage's form),
// discards 310 values, and every draw adds the rear word into the front word and returns the
// sum without its low bit.  Restated here so a batched run draws EXACTLY the values the
// reference's serial loop would (tests/test_refrand.py pins it against oracle/_ref).
#include <cmath>
#include <cstdint>

#include "b200spice.h"

namespace {

struct Random3 {
    uint32_t state[31];
    int f, r;                                   // front / rear index
    static uint32_t park_miller(int32_t x) {    // randval.cc:167-182 (non-weak seeding)
        if (x == 0) x = 123459876;
        int32_t hi = x/127773, lo = x % 127773;
        x = 16807*lo - 2836*hi;
        if (x < 0) x += 0x7fffffff;
        return (uint32_t)x;
    }
    void seed(uint32_t x) {                     // randval.cc:196-211
        state[0] = x;
        for (int i = 1; i < 31; i++) state[i] = park_miller((int32_t)state[i-1]);
        f = 3; r = 0;
        for (int i = 0; i < 310; i++) next();
    }
    uint32_t next() {                           // randval.cc:230-258
        state[f] += state[r];
        uint32_t v = (state[f] >> 1) & 0x7fffffffu;
        if (++f >= 31) { f = 0; ++r; }
        else if (++r >= 31) r = 0;
        return v;
    }
    double uniform() { return (double)(next() & 0x7fffffffu)/(0x7fffffff + 1.0); }   // random.h:76-79
};

} // namespace

extern "C" int b200_mc_uniform(int seed, long long n, double *out)
{
    if (n < 0 || !out) return B200_E_BADPARM;
    Random3 g; g.seed((uint32_t)seed);
    for (long long i = 0; i < n; i++) out[i] = g.uniform();
    return B200_OK;
}

extern "C" int b200_mc_gauss(int seed, long long n, double *out)
{
    // sRnd::gauss, random.cc:51-72: polar Box-Muller; the v2 deviate is returned first, the v1
    // deviate is kept for the next call
    if (n < 0 || !out) return B200_E_BADPARM;
    Random3 g; g.seed((uint32_t)seed);
    bool have = false; double kept = 0.0;
    for (long long i = 0; i < n; i++) {
        if (!have) {
            double r, v1, v2;
            do {
                v1 = 2.0*g.uniform() - 1.0;
                v2 = 2.0*g.uniform() - 1.0;
                r = v1*v1 + v2*v2;
            } while (r >= 1.0);
            double fac = sqrt(-2.0*log(r)/r);
            kept = v1*fac; have = true;
            out[i] = v2*fac;
        }
        else { have = false; out[i] = kept; }
    }
    return B200_OK;
}
