// TEST TOOL: accuracy of csrc/b2c_softmath.h against libm in double.   g++ -O2 -ffp-contract=off -o /tmp/softmath_check tools/softmath_check.cpp && /tmp/softmath_check
#include <cstdio>
#include <cmath>
#include <cstring>
#include <cstdlib>
#include "../pyctcdecode_b200/csrc/b2c_softmath.h"
static double ulp_err(float got, double want) {
    float w = (float)want; int ex; frexpf(w, &ex); double ulp = ldexp(1.0, ex - 24); return fabs((double)got - want) / ulp;
}
int main() {
    double worst = 0, worst_x = 0; 
    for (int i = 0; i <= 2000000; ++i) { float x = -87.0f * (float)i / 2000000.0f; double e = ulp_err(b2c_sm_expf(x), exp((double)x)); if (e > worst) { worst = e; worst_x = x; } }
    printf("expf max ulp err %.3f at %g; exp(0)=%.9g\n", worst, worst_x, b2c_sm_expf(0.0f));
    worst = 0;
    for (int i = 0; i <= 4000000; ++i) { float x = 1.0f + 1100.0f * (float)i / 4000000.0f; double e = ulp_err(b2c_sm_logf(x), log((double)x)); if (e > worst) { worst = e; worst_x = x; } }
    printf("logf max ulp err %.3f at %g; log(1)=%.9g\n", worst, worst_x, b2c_sm_logf(1.0f));
    for (int i = 0; i <= 1000000; ++i) { float x = 1.0f + 3e-3f * (float)i / 1000000.0f; double a = fabs((double)b2c_sm_logf(x) - log((double)x)); if (a > worst) { worst = a; worst_x = x; } }
    printf("logf near 1: max abs err %.3g at %.9g\n", worst, worst_x);
    return 0;
}
