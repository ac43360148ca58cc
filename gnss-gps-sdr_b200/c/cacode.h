// cacode.h -- GPS C/A (Gold) code generator with the reference's CACODE interface
// (c/cacode.h:9-35: constructor taking the two G2 tap positions, Chip(), Clock(), GetG1()).
//
// The two 10-stage shift registers are kept as bit masks: bit k-1 holds stage k, stage 10 is
// the output end, the feedback bit enters stage 1.
//   G1 = 1 + x^3 + x^10            G2 = 1 + x^2 + x^3 + x^6 + x^8 + x^9 + x^10
#ifndef CACODE_B200_H
#define CACODE_B200_H

struct CACODE {
    unsigned g1, g2;
    int tap0, tap1;                       // G2 stages XORed into the output (1..10)

    CACODE(int t0, int t1) : g1(0x3FFu), g2(0x3FFu), tap0(t0), tap1(t1) {}

    static unsigned stage(unsigned reg, int k) { return (reg >> (k - 1)) & 1u; }

    int Chip() const { return (int)(stage(g1, 10) ^ stage(g2, tap0) ^ stage(g2, tap1)); }

    void Clock() {
        const unsigned f1 = stage(g1, 3) ^ stage(g1, 10);
        const unsigned f2 = stage(g2, 2) ^ stage(g2, 3) ^ stage(g2, 6) ^ stage(g2, 8) ^ stage(g2, 9) ^ stage(g2, 10);
        g1 = ((g1 << 1) | f1) & 0x3FFu;
        g2 = ((g2 << 1) | f2) & 0x3FFu;
    }

    // stage 10 is the most significant of the 10 returned bits, stage 1 the least
    unsigned GetG1() const {
        unsigned r = 0;
        for (int k = 10; k >= 1; k--) r = (r << 1) | stage(g1, k);
        return r;
    }
};

#endif
