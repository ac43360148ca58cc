// conv_1bit_bin_to_hackrf_bin.cpp -- the reference's 1-bit IF -> HackRF int8 IQ converter (c/conv_1bit_bin_to_hackrf_bin.cpp)
// on top of the libgpsacq C ABI (gpsacq_bits_to_iq8, hand-written sm_100a kernel; include/gpsacq.h).
//
// Behaviour kept from the reference (line numbers in /root/reference/c/conv_1bit_bin_to_hackrf_bin.cpp):
//   * no arguments: reads gps.samples.1bit.I.fs5456.if4092.bin, writes gps.samples.8bit.IQinterleave.fs5456.if0.bin (:91)
//   * the file is consumed in packets of 871744*64 bytes (:25); a trailing partial packet is dropped after
//     "seems run out!" (:55-58); the packet count is printed as it goes (:59)
//   * FC / FS are the receiver's constants of c/gps.h:23-24 (2.6 MHz / 10 MHz) -- that is what the reference program is
//     compiled with, whatever the file name says; amplitude +-30 (:17-19); the phase NCO runs on across packets (:35)
//   * "can not open file for read!/write!" and exit code 0 in every case (:41-50)
// Optional arguments (absent = reference behaviour):  conv_1bit_bin_to_hackrf_bin [in [out [FC FS]]]
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "../../include/gpsacq.h"

int main(int argc, char **argv)
{
    const char *in = argc > 1 ? argv[1] : "gps.samples.1bit.I.fs5456.if4092.bin";
    const char *out = argc > 2 ? argv[2] : "gps.samples.8bit.IQinterleave.fs5456.if0.bin";
    const double fc = argc > 4 ? atof(argv[3]) : 2.6e6, fs = argc > 4 ? atof(argv[4]) : 10e6;
    const size_t PACKET = (size_t)871744 * 64;

    FILE *fp_in = fopen(in, "rb");
    if (!fp_in) { printf("can not open file for read!\n"); return 0; }
    FILE *fp_out = fopen(out, "wb");
    if (!fp_out) { printf("can not open file for write!\n"); fclose(fp_in); return 0; }

    std::vector<unsigned char> rx(PACKET);
    std::vector<int8_t> iq(16 * PACKET);
    const char *dev = getenv("GPSACQ_DEVICE");
    unsigned run_count = 0;
    size_t first_sample = 0;
    for (;;) {
        const size_t got = fread(rx.data(), 1, PACKET, fp_in);
        if (got != PACKET) { printf("seems run out!\n"); break; }
        printf("%d\n", ++run_count);
        const int rc = gpsacq_bits_to_iq8(dev ? atoi(dev) : 0, rx.data(), PACKET, first_sample, fc, fs, 30, iq.data());
        if (rc) { fprintf(stderr, "gpsacq_bits_to_iq8: %s\n", gpsacq_last_error(NULL)); break; }
        fwrite(iq.data(), 1, iq.size(), fp_out);
        first_sample += 8 * PACKET;
    }
    fclose(fp_in);
    fclose(fp_out);
    return 0;
}
