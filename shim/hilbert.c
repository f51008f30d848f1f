/*
 * shim/hilbert.c -- drop-in replacement object for the reference's hilbert.o (hilbert.c:28-92):
 * Blackman-windowed odd-length Hilbert FIR with optional phase angle, handed to fir / fir_p.
 * Options: -p (fir_p), -z (zita_convolver; falls back to fir_p as in hilbert.c:82-87 when zita is
 * not built), -c (reference point = taps/2 for the chain's alignment), -a angle (degrees).
 */
#include <stdlib.h>
#include <math.h>
#include "hilbert.h"
#include "fir.h"
#include "fir_p.h"
#include "util.h"
#include "dsp_b200.h"

struct effect * hilbert_effect_init(const struct effect_info *ei, const struct stream_info *istream, const char *channel_selector, const char *dir, int argc, const char *const *argv)
{
	struct dsp_getopt_state g = DSP_GETOPT_STATE_INITIALIZER;
	int partitioned = 0, centre = 0, opt;
	double angle = -M_PI_2;
	char *endptr;

	while ((opt = dsp_getopt(&g, argc - 1, argv, "pzca:")) != -1) {
		switch (opt) {
		case 'p': partitioned = 1; break;
		case 'z':
			LOG_FMT(LL_ERROR, "%s: warning: zita_convolver not available; using fir_p instead", argv[0]);
			partitioned = 1;
			break;
		case 'c': centre = 1; break;
		case 'a':
			angle = strtod(g.arg, &endptr) / 180.0 * M_PI;
			CHECK_ENDPTR(g.arg, endptr, "angle", return NULL);
			break;
		default:
			dsp_getopt_print_error(&g, opt, argv[0]);
			print_effect_usage(ei);
			return NULL;
		}
	}
	if (g.ind != argc - 1) {
		print_effect_usage(ei);
		return NULL;
	}
	const ssize_t taps = strtol(argv[g.ind], &endptr, 10);
	CHECK_ENDPTR(argv[g.ind], endptr, "taps", return NULL);
	if (taps <= 3) {
		LOG_FMT(LL_ERROR, "%s: error: taps must be > 3", argv[0]);
		return NULL;
	}
	if (taps % 2 == 0) {
		LOG_FMT(LL_ERROR, "%s: error: taps must be odd", argv[0]);
		return NULL;
	}
	sample_t *h = calloc(taps, sizeof(sample_t));
	if (check_alloc(ei->name, h)) return NULL;
	dspb200_hilbert_taps(taps, angle, h);
	const ssize_t ref = (centre) ? taps / 2 : 0;
	struct effect *e = (partitioned)
		? fir_p_effect_init_with_filter(ei, istream, channel_selector, h, 1, taps, ref, 0)
		: fir_effect_init_with_filter(ei, istream, channel_selector, h, 1, taps, ref, 0);
	free(h);
	return e;
}
