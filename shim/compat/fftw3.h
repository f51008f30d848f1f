/* shim/compat/fftw3.h -- only what the reference's util.c needs from <fftw3.h> when FFTW3 itself
 * is not installed: the two wisdom entry points (util.c:484,495).  See shim/fftw_compat.c. */
#ifndef DSPB200_COMPAT_FFTW3_H
#define DSPB200_COMPAT_FFTW3_H
int fftw_import_wisdom_from_filename(const char *filename);
int fftw_export_wisdom_to_filename(const char *filename);
#endif
