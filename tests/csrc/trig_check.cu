// Host-side accuracy check of pioneer_b200/csrc/pnr_trig.cuh (compiled with nvcc, run on the CPU).
// Prints: name, samples, max |sin err|, max |cos err| against float64 libm.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include "../../pioneer_b200/csrc/pnr_trig.cuh"

template <typename F>
static void sweep(const char* name, double lo, double hi, long n, F f) {
    double es = 0, ec = 0;
    for (long i = 0; i <= n; ++i) {
        const float x = (float)(lo + (hi - lo) * (double)i / (double)n);
        float s, c;
        f(x, s, c);
        es = fmax(es, fabs((double)s - sin((double)x)));
        ec = fmax(ec, fabs((double)c - cos((double)x)));
    }
    printf("%s %ld %.4e %.4e\n", name, n, es, ec);
}

int main() {
    auto b = [](float x, float& s, float& c) { pnr_sincos_bounded(x, s, c); };
    auto f = [](float x, float& s, float& c) { pnr_sincos_fast(x, s, c); };
    sweep("bounded_pi", -3.1416, 3.1416, 4000000, b);
    sweep("bounded_2pi", 0.0, 6.2832, 4000000, b);
    sweep("bounded_4pi", -12.5664, 12.5664, 4000000, b);
    sweep("bounded_64", -64.0, 64.0, 4000000, b);
    sweep("fast_126", -125.664, 125.664, 4000000, f);
    sweep("fast_1e4", -1.0e4, 1.0e4, 4000000, f);
    sweep("fast_limit", -105615.0, 105615.0, 8000000, f);
    // exact values the reset state produces: sin(0) = 0, cos(0) = 1
    float s, c;
    pnr_sincos_bounded(0.f, s, c);
    printf("zero %d %.9g %.9g\n", 1, (double)s, (double)c);
    pnr_sincos_fast(0.f, s, c);
    printf("zero_fast %d %.9g %.9g\n", 1, (double)s, (double)c);
    return 0;
}
