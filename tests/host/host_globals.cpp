// tests/host/host_globals.cpp -- the reference reads FS in SearchInit() (replicas, c/search_offline.cpp:76) and FC, FS,
// max_fo again on every Sample() / Correlate() (:127,:176,:190).  This caller changes the globals BETWEEN SearchInit() and
// SearchTask(), as a reference caller may: FC is wrong (0) while the replicas are built and right when the file is searched
// (argv[2] = "fc"); or max_fo is 1000 at SearchInit() and 5000 at SearchTask() (argv[2] = "max_fo").  Either way the
// unmodified reference prints exactly what it prints for constant globals.  Used by tests/test_gpu_parity.py.
#include <stdio.h>
#include <string.h>
#include "gps_offline.h"

double FC = 4.092e6, FS = 5.456e6, max_fo = 5000.0;

int main(int argc, char **argv)
{
    if (argc != 3) { fprintf(stderr, "usage: host_globals <capture> fc|max_fo\n"); return 2; }
    if (!strcmp(argv[2], "fc")) FC = 0.0; else max_fo = 1000.0;
    const int rc = SearchInit();
    if (rc) { printf("SearchInit() returned %d\n", rc); return rc; }
    FC = 4.092e6; max_fo = 5000.0;
    SearchTask(argv[1]);
    SearchFree();
    return 0;
}
