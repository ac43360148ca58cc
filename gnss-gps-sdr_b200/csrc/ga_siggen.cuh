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

// ---------------------------------------------------------------------------------------------------------------
// gps_sig_gen.m, literally (gps_sig_gen.m:8-41): the chain that wrote the reference's bundled gps_sig_tmp.bin.
//   g = 1 - 2*cacode(sv)            chips +-1                                   (:15)
//   g = upsample(g, 8)              zero-stuffed to 8.184 Msps                  (:16)
//   data = kron(nav, kron(ones(1,20), g))      20 code periods per NAV bit      (:17-20)
//   data = conv(data, rcosine(1, 8))           49-tap raised cosine (R = 0.5, delay 3), +48 samples   (:23,:35)
//   y = real(data .* exp(1i*2*pi*fc*(0:N-1)*(1/ca_rate))),  fc = ca_rate/4      (:34,:36)
//   y = (1 - sign(y))/2 ; fwrite(fid, y, 'ubit1')                               (:37-41)
// Everything in double like MATLAB, with the operation ORDER that decides the rounding: conv sums data(j)*num(n-j)
// over ascending j (an exactly cancelling set of terms leaves a rounding residue whose sign becomes the bit), the
// phase is ((2*pi*fc)*n)*(1/ca_rate), no fused multiply-add.  sign(0) = 0 gives y = 0.5, which fwrite rounds to 1.
// With the NAV bits of the file (the script draws them with rand) this reproduces gps_sig_tmp.bin bit for bit.
// The taps are rcosine(1,8) evaluated in double: sinc(n/8)*cos(pi*n/16)/(1-(n/8)^2), pi/4*sinc(1) at n = +-8.
// ---------------------------------------------------------------------------------------------------------------
__constant__ double c_rcos[49] = {
    0x1.2972f529d570dp-110, 0x1.2a3b74d882f6ap-10, 0x1.38ca36608bcdap-8, 0x1.5a3aace0dc09dp-7,
    0x1.18f7a0110173ep-6, 0x1.6b7d46d0781cep-6, 0x1.74baeb06b1e48p-6, 0x1.06033a3318282p-6,
    -0x1.df63f92f267c1p-57, -0x1.9efd294c27da5p-6, -0x1.d7f6a7ef8ad01p-5, -0x1.77ac1861056bfp-4,
    -0x1.ebb1581dc28abp-4, -0x1.113c3cf2ada5ep-3, -0x1.f5c461e58aedbp-4, -0x1.45bc139efcd69p-4,
    0x1.1a62633145c07p-55, 0x1.daa4571ade237p-4, 0x1.0ccdc6baf8240p-2, 0x1.b7473e8a172cap-2,
    0x1.334ed7129996cp-1, 0x1.847ab07b99f87p-1, 0x1.c643ce7028ce9p-1, 0x1.f11f44233a675p-1,
    0x1.0000000000000p+0, 0x1.f11f44233a675p-1, 0x1.c643ce7028ce9p-1, 0x1.847ab07b99f87p-1,
    0x1.334ed7129996cp-1, 0x1.b7473e8a172cap-2, 0x1.0ccdc6baf8240p-2, 0x1.daa4571ade237p-4,
    0x1.1a62633145c07p-55, -0x1.45bc139efcd69p-4, -0x1.f5c461e58aedbp-4, -0x1.113c3cf2ada5ep-3,
    -0x1.ebb1581dc28abp-4, -0x1.77ac1861056bfp-4, -0x1.d7f6a7ef8ad01p-5, -0x1.9efd294c27da5p-6,
    -0x1.df63f92f267c1p-57, 0x1.06033a3318282p-6, 0x1.74baeb06b1e48p-6, 0x1.6b7d46d0781cep-6,
    0x1.18f7a0110173ep-6, 0x1.5a3aace0dc09dp-7, 0x1.38ca36608bcdap-8, 0x1.2a3b74d882f6ap-10,
    0x1.2972f529d570dp-110};

constexpr int SIGLIT_OV = 8, SIGLIT_PERIODS = 20, SIGLIT_TAPS = 49;
constexpr long long SIGLIT_PER_BIT = 1023LL * SIGLIT_OV * SIGLIT_PERIODS;        // 163,680 samples per NAV bit

__global__ void sig_gen_literal_kernel(int t0, int t1, const unsigned char *__restrict__ nav01, long long n_data,
                                       long long n_out, double two_pi_fc, double inv_rate, unsigned char *__restrict__ bits)
{
    __shared__ signed char chips[1024];
    if (threadIdx.x == 0) {
        unsigned g1 = 0x3FF, g2 = 0x3FF;
        for (int i = 0; i < 1023; i++) {
            const unsigned c = ((g1 >> 9) ^ (g2 >> (t0 - 1)) ^ (g2 >> (t1 - 1))) & 1u;
            chips[i] = c ? -1 : 1;                                               // 1 - 2*chip
            const unsigned f1 = ((g1 >> 2) ^ (g1 >> 9)) & 1u;
            const unsigned f2 = ((g2 >> 1) ^ (g2 >> 2) ^ (g2 >> 5) ^ (g2 >> 7) ^ (g2 >> 8) ^ (g2 >> 9)) & 1u;
            g1 = ((g1 << 1) | f1) & 0x3FF;
            g2 = ((g2 << 1) | f2) & 0x3FF;
        }
    }
    __syncthreads();
    const long long byte = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (byte * 8 >= n_out) return;
    unsigned out = 0;
    for (int k = 0; k < 8; k++) {
        const long long n = byte * 8 + k;
        if (n >= n_out) break;
        // conv(data, num)(n) = sum over ascending j of data(j)*num(n-j); data is non-zero only where j is a multiple of 8
        long long j = n - (SIGLIT_TAPS - 1);
        if (j < 0) j = 0;
        j = (j + SIGLIT_OV - 1) / SIGLIT_OV * SIGLIT_OV;
        double x = 0.0;
        for (; j <= n && j < n_data; j += SIGLIT_OV) {
            const long long nb = j / SIGLIT_PER_BIT;
            const int chip = (int)((j % (1023LL * SIGLIT_OV)) / SIGLIT_OV);
            const double d = (nav01[nb] ? -1.0 : 1.0) * (double)chips[chip];     // data = 1 - 2*round(rand): bit 1 -> -1
            x = __dadd_rn(x, __dmul_rn(d, c_rcos[n - j]));
        }
        const double ph = __dmul_rn(__dmul_rn(two_pi_fc, (double)n), inv_rate);
        const double y = __dmul_rn(x, cos(ph));
        out |= (y > 0.0 ? 0u : 1u) << k;                                         // y < 0 -> 1; y == 0 -> 0.5 -> ubit1 writes 1
    }
    bits[byte] = (unsigned char)out;
}

}  // namespace ga
