/*
 * shim/fftw_compat.c -- the two FFTW symbols the reference's util.c references when built with
 * -DHAVE_FFTW3 (util.c:484,495: wisdom import/export).  -DHAVE_FFTW3 is what makes effect.c list
 * fir / fir_p / hilbert / resample at all (fir.h:22-35 etc.); with the GPU objects nothing plans
 * an FFTW transform any more, so "wisdom" is a no-op.
 */
int fftw_import_wisdom_from_filename(const char *filename) { (void) filename; return 0; }
int fftw_export_wisdom_to_filename(const char *filename) { (void) filename; return 0; }
