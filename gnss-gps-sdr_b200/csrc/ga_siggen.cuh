// ga_siggen.cuh -- synthetic 1-bit GPS L1 C/A IF capture generator on the GPU.
// What the reference does in MATLAB for one satellite (gps_sig_gen.m:8-41 with cacode.m:65-134: code x NAV
// bits -> carrier at IF -> sign -> 'ubit1' LSB-first), generalised to several satellites with Doppler, code
// phase, carrier phase and additive Gaussian noise, directly at the target sampling rate (code NCO).
//   x[n] = sigma*g[n] + sum_k amp_k * nav_k(floor(t*bps)) * ca_k(floor((t*rate_k + phase_k) mod 1023)) * cos(2*pi*((fc+fd_k)*t) + phi_k)
//   bit[n] = x[n] < 0        t = n/fs,  rate_k = 1.023e6*(1 + fd_k/1575.42e6)
// Everything that decides a bit is computed in double; noise and NAV bits come from a counter-based
// generator (splitmix64 of (seed, stream, index)), so any sample can be produced independently -- the numpy
// restatement in siggen.py (synth_capture_counter) reproduces it bit for bit up to |x| ~ 1e-15 ties.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ga {

struct SynthSat { int prn; int t0, t1; double amp, doppler_hz, code_phase_chips, carrier_phase_cycles; };
constexpr int SYNTH_MAX_SATS = 16;
struct SynthParams {
    SynthSat sat[SYNTH_MAX_SATS];
    int n_sats;
    double fs, fc, sigma, nav_bps;
    unsigned long long seed;
};

__host__ __device__ inline unsigned long long splitmix64(unsigned long long x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
// standard normal for (seed, index): Box-Muller on two 53-bit uniforms
__device__ inline double synth_gauss(unsigned long long seed, unsigned long long n)
{
    const unsigned long long a = splitmix64(seed ^ splitmix64(2 * n)), b = splitmix64(seed ^ splitmix64(2 * n + 1));
    const double u1 = ((double)(a >> 11) + 1.0) * (1.0 / 9007199254740993.0);   // (0,1)
    const double u2 = (double)(b >> 11) * (1.0 / 9007199254740992.0);           // [0,1)
    return sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
}

__global__ void synth_bits_kernel(SynthParams p, size_t n_samples, size_t n0, unsigned char *__restrict__ bits)
{
    __shared__ signed char chips[SYNTH_MAX_SATS][1024];
    if (threadIdx.x < p.n_sats) {          // C/A code of each satellite: G1/G2 LFSRs (cacode.m:103-120, c/cacode.h:15-28)
        unsigned g1 = 0x3FF, g2 = 0x3FF;
        const int t0 = p.sat[threadIdx.x].t0, t1 = p.sat[threadIdx.x].t1;
        for (int i = 0; i < 1023; i++) {
            const unsigned c = ((g1 >> 9) ^ (g2 >> (t0 - 1)) ^ (g2 >> (t1 - 1))) & 1u;
            chips[threadIdx.x][i] = c ? -1 : 1;
            const unsigned f1 = ((g1 >> 2) ^ (g1 >> 9)) & 1u;
            const unsigned f2 = ((g2 >> 1) ^ (g2 >> 2) ^ (g2 >> 5) ^ (g2 >> 7) ^ (g2 >> 8) ^ (g2 >> 9)) & 1u;
            g1 = ((g1 << 1) | f1) & 0x3FF;
            g2 = ((g2 << 1) | f2) & 0x3FF;
        }
    }
    __syncthreads();
    const size_t byte = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (byte * 8 >= n_samples) return;
    unsigned out = 0;
    for (int k = 0; k < 8; k++) {
        const size_t n = n0 + byte * 8 + k;
        if (byte * 8 + k >= n_samples) break;
        const double t = (double)n / p.fs;
        double x = p.sigma > 0.0 ? p.sigma * synth_gauss(p.seed, n) : 0.0;
        for (int s = 0; s < p.n_sats; s++) {
            const SynthSat &q = p.sat[s];
            const double chipf = fmod(t * (1.023e6 * (1.0 + q.doppler_hz / 1575.42e6)) + q.code_phase_chips, 1023.0);
            const int chip = (int)chipf;
            const unsigned long long nb = (unsigned long long)(t * p.nav_bps);
            const double nav = (splitmix64(p.seed ^ splitmix64(0xA5A5000000000000ull + ((unsigned long long)s << 40) + nb)) & 1ull) ? -1.0 : 1.0;
            const double cyc = (p.fc + q.doppler_hz) * t + q.carrier_phase_cycles;
            x += q.amp * nav * (double)chips[s][chip] * cospi(2.0 * (cyc - floor(cyc)));
        }
        out |= (x < 0.0 ? 1u : 0u) << k;
    }
    bits[byte] = (unsigned char)out;
}

}  // namespace ga
