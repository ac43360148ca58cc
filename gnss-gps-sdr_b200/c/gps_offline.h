// gps_offline.h -- public header of the offline GPS L1 C/A search, B200 edition.
//
// Keeps the part of the reference's c/gps_offline.h that gps_test and any other caller of
// the search path needs: the size macros (:15-16), the chip rate (:30), the three globals
// the CALLER defines (:23-25) and the five Search* entry points (:87-91), with C++ linkage
// like the reference, so that an object file compiled against the reference header links
// against this implementation unchanged.  The on-target declarations (SPI, channels,
// ephemeris, UI) are out of scope and deliberately absent.
#ifndef GPS_OFFLINE_B200_H
#define GPS_OFFLINE_B200_H

#define FFT_LEN  40000     // coherent window in samples
#define NUM_SATS 32        // PRN 1..32, index sv = PRN-1
#define CPS      1.023e6   // C/A chip rate

extern double FC;          // carrier frequency at IF, Hz   -- defined by the caller
extern double FS;          // sampling rate, Hz             -- defined by the caller
extern double max_fo;      // Doppler search half-span, Hz  -- defined by the caller

int  SearchInit();                        // 0 = ok; builds the 32 replica spectra on the GPU
void SearchFree();
void SearchTask(char *filename_1bit_bin); // whole-file search, report on stdout
void SearchEnable(int sv);
int  SearchCode(int sv, int g1);          // chips until the G1 register reads g1

#endif
