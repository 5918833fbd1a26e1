// Host-side exhaustive-ish check of gspn::FastDiv (csrc/common.cuh), the multiply-shift divider the element-indexed kernels
// use instead of a 64-bit divide.  Built and run by tests/test_host_helpers.py with nvcc (no GPU needed).
#include <cstdio>
#include <cstdlib>
#include "../../gspn_b200/csrc/common.cuh"

int main() {
    using gspn::FastDiv;
    long bad = 0;
    const unsigned ds[] = {1, 2, 3, 5, 7, 8, 16, 24, 32, 48, 64, 67, 100, 128, 131, 192, 512, 2048, 32768, 32769, 65535, 1000003, 0x7fffffffu};
    for (unsigned d : ds) {
        FastDiv f(d);
        for (unsigned long long x = 0; x < (1ull << 31); x += (x < 200000 ? 1 : 99991))
            if (f.div((unsigned)x) != (unsigned)x / d) ++bad;
        const unsigned top = 0x7fffffffu;
        if (f.div(top) != top / d) ++bad;
        for (unsigned k = 1; k < 64 && (unsigned long long)k * d < (1ull << 31); ++k) {  // around multiples of d
            const unsigned x = k * d;
            if (f.div(x) != k || f.div(x - 1) != (x - 1) / d) ++bad;
        }
    }
    srand(1);
    for (int i = 0; i < 2000000; ++i) {
        const unsigned d = (unsigned)(rand() % 100000) + 1;
        const unsigned x = ((unsigned)rand() * 7919u + (unsigned)rand()) & 0x7fffffffu;
        if (FastDiv(d).div(x) != x / d) ++bad;
    }
    printf("bad=%ld\n", bad);
    return bad != 0;
}
