// tests/host/host_check.cpp -- exercises the GPU-independent parts of the C++ host tree
// (CACODE, SearchCode, SearchEnable) and prints results for tests/test_host.py.
#include <stdio.h>
#include "gps_offline.h"
#include "cacode.h"

double FC = 4.092e6, FS = 5.456e6, max_fo = 5000.0;

int main()
{
    // first 10 chips of PRN 1 (taps 2,6) and PRN 32 (taps 4,9) as octal, 1023-chip ones count
    const int taps[2][2] = {{2, 6}, {4, 9}};
    for (int k = 0; k < 2; k++) {
        CACODE ca(taps[k][0], taps[k][1]);
        unsigned first = 0; int ones = 0;
        for (int i = 0; i < 1023; i++) { int c = ca.Chip(); if (i < 10) first = (first << 1) | c; ones += c; ca.Clock(); }
        CACODE again(taps[k][0], taps[k][1]);
        printf("first10 %o ones %d g1_after_period %u g1_seed %u\n", first, ones, ca.GetG1(), again.GetG1());
    }
    printf("SearchCode %d %d %d %d\n", SearchCode(0, 0x3FF), SearchCode(3, 0x2AA), SearchCode(5, 0), SearchCode(40, 1));
    SearchEnable(3);
    SearchEnable(99);
    return 0;
}
