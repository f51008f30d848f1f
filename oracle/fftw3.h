/*
 * oracle/fftw3.h -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Declaration-compatible subset of the FFTW3 API, so that the reference's
 * unmodified fir.c / fir_p.c / resample.c / util.c / matrix4_mb.c compile here
 * without FFTW3 (which is un-vendored and absent; /root/reference/configure:137).
 * Only the ten entry points those files use are provided:
 *   fir.c:127,132,329-330,344,352   fir_p.c:72,87,469-470,485,492
 *   resample.c:113,133,336,344-345,366   util.c:484,495
 * The reference includes <complex.h> before <fftw3.h>, so fftw_complex is the
 * C99 double _Complex (same rule as the real header).
 */
#ifndef ORACLE_FFTW3_SHIM_H
#define ORACLE_FFTW3_SHIM_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_Complex_I) && defined(complex) && defined(I)
typedef double _Complex fftw_complex;
#else
typedef double fftw_complex[2];
#endif

typedef struct oracle_fft_plan_s *fftw_plan;

#define FFTW_MEASURE  (0U)
#define FFTW_ESTIMATE (1U << 6)

void *fftw_malloc(size_t n);
void fftw_free(void *p);

fftw_plan fftw_plan_dft_r2c_1d(int n, double *in, fftw_complex *out, unsigned flags);
fftw_plan fftw_plan_dft_c2r_1d(int n, fftw_complex *in, double *out, unsigned flags);
void fftw_execute(const fftw_plan p);
void fftw_execute_dft_r2c(const fftw_plan p, double *in, fftw_complex *out);
void fftw_execute_dft_c2r(const fftw_plan p, fftw_complex *in, double *out);
void fftw_destroy_plan(fftw_plan p);

int fftw_import_wisdom_from_filename(const char *filename);
int fftw_export_wisdom_to_filename(const char *filename);

#ifdef __cplusplus
}
#endif

#endif
