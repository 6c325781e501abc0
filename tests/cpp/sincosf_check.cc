// TEST INFRASTRUCTURE: the host build of og::sincosf_glibc (the function the descriptor kernel uses for cos / sin of the key-point
// angle) against this machine's glibc sincosf — what the compiled reference calls (ORBextractor.cc:113) — on every `stride`-th
// float of [0, 2*pi].  stride 1 = exhaustive (1.09 G arguments, a few seconds).
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "og_math.cuh"
int main(int argc, char** argv) {
    const uint32_t stride = argc > 1 ? (uint32_t)atoi(argv[1]) : 64;
    long bad = 0, tot = 0;
    for (uint32_t u = 0; u <= 0x40c91000u; u += stride) {
        float y, s0, c0, s1, c1;
        memcpy(&y, &u, 4);
        sincosf(y, &s0, &c0);
        og::sincosf_glibc(y, &s1, &c1);
        if (memcmp(&s0, &s1, 4) || memcmp(&c0, &c1, 4)) {
            if (bad < 5) printf("%a: glibc %a %a, og %a %a\n", y, s0, c0, s1, c1);
            ++bad;
        }
        ++tot;
    }
    printf("sincosf_glibc vs glibc sincosf: %ld arguments, %ld mismatches\n", tot, bad);
    return bad != 0;
}
