// test_search_offline.cpp -- the `gps_test` program: same command line and banner as the
// reference's c/test_search_offline.cpp (usage text :24-29, argument rules :31-38, defaults
// :19-22; the 4th argument is accepted and, like there, not used: max_fo stays 5000 unless the
// opt-in GPSACQ_HONOR_MAX_FO=1 is set).
#include <stdio.h>
#include <stdlib.h>
#include <string>

double FC = 4.092e6, FS = 5.456e6, max_fo = 5000.0;
#include "gps_offline.h"

static const char *const kBanner[] = {
    "GPS CA code offline search. Extract from http://www.aholme.co.uk/GPS/Main.htm",
    "Jiao Xianjun (putaoshu@gmail.com). 2014-05.",
    "usage:",
    "gps_test   filename_of_1bit_IF_cap   carrier_freq   sampling_rate   max_freq_offset",
    "or",
    "gps_test (Make sure gps.samples.1bit.I.fs5456.if4092.bin can be found. Download http://www.jks.com/gps/gps.html)",
};

int main(int argc, char **argv)
{
    std::string file = "gps.samples.1bit.I.fs5456.if4092.bin";
    for (const char *line : kBanner) puts(line);

    if (argc == 5) {
        file = argv[1];
        FC = atof(argv[2]);
        FS = atof(argv[3]);
        const char *honor = getenv("GPSACQ_HONOR_MAX_FO");
        if (honor && honor[0] == '1') max_fo = atof(argv[4]);
    } else if (argc != 1) {
        puts("Please run with 3 arguments or without argument!");
        return 0;
    }

    const int rc = SearchInit();
    if (rc) {
        printf("SearchInit() returned %d\n", rc);
        return rc;
    }
    SearchTask(&file[0]);
    SearchFree();
    return 0;
}
