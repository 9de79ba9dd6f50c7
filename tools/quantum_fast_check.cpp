// TEST TOOL: b2c_sm_quantum_fast(d) == b2c_sm_quantum(b2c_sm_expf(d)) for EVERY float32 d <= 0 (finite, -0.0, -inf).
//   g++ -O2 -ffp-contract=off -fopenmp -o /tmp/quantum_fast_check tools/quantum_fast_check.cpp && /tmp/quantum_fast_check [stride]
// stride 1 = exhaustive (2^31 values, about a minute on 16 threads); tests/test_host_logic.py runs a strided sweep.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include "../pyctcdecode_b200/csrc/b2c_softmath.h"
static uint64_t reference_quantum(float d) {      // b2c_prepare.h: b2c_sm_quantum(b2c_sm_expf(d)) on rows without special values
    const float e = b2c_sm_expf(d);
    return static_cast<uint64_t>(llrintf(e * 4294967296.0f));
}
int main(int argc, char** argv) {
    const uint64_t stride = argc > 1 ? std::strtoull(argv[1], nullptr, 10) : 1;
    uint64_t bad = 0, n = 0;
#pragma omp parallel for reduction(+ : bad, n) schedule(static)
    for (int64_t k = 0; k <= static_cast<int64_t>(0x7F800000ull / stride); ++k) {
        const uint32_t mag = static_cast<uint32_t>(static_cast<uint64_t>(k) * stride);       // 0 .. 0x7F800000 (-0.0 .. -inf)
        const float d = b2c_sm_from_bits(0x80000000u | mag);
        if (b2c_sm_quantum_fast(d) != reference_quantum(d)) {
            if (bad < 5) std::printf("mismatch at d=%.9g (bits %08x): fast %llu definition %llu\n", d, 0x80000000u | mag,
                                     (unsigned long long)b2c_sm_quantum_fast(d), (unsigned long long)reference_quantum(d));
            ++bad;
        }
        ++n;
    }
    const float zero = 0.0f;
    if (b2c_sm_quantum_fast(zero) != reference_quantum(zero)) ++bad;
    std::printf("checked %llu values, mismatches %llu\n", (unsigned long long)n + 1, (unsigned long long)bad);
    return bad ? 1 : 0;
}
