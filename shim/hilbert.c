/*
 * shim/hilbert.c -- drop-in replacement object for the reference's hilbert.o (hilbert.c:28-92):
 * Blackman-windowed odd-length Hilbert FIR with optional phase angle, handed to fir / fir_p.
 * Options: -p (fir_p), -z (zita_convolver; falls back to fir_p as in hilbert.c:82-87 when zita is
 * not built), -c (reference point = taps/2 for the chain's alignment), -a angle (degrees).
 * The taps come from the library (dspb200_hilbert_taps, same arithmetic as hilbert.c:65-77); the
 * effect itself is whatever the GPU fir / fir_p shim builds from them.
 */
#include <stdlib.h>
#include <math.h>
#include "hilbert.h"
#include "fir.h"
#include "fir_p.h"
#include "util.h"
#include "dsp_b200.h"

struct hilbert_request {
	int use_fir_p;      /* -p, or -z without zita */
	int centre_ref;     /* -c */
	double angle;       /* radians */
	ssize_t taps;
};

/* 0 on success; on failure the message has been logged (same texts as the reference) */
static int hilbert_parse(const struct effect_info *ei, int argc, const char *const *argv, struct hilbert_request *rq)
{
	struct dsp_getopt_state g = DSP_GETOPT_STATE_INITIALIZER;
	const char *name = argv[0];
	char *end;
	int opt, bad = 0;

	rq->use_fir_p = rq->centre_ref = 0;
	rq->angle = -M_PI_2;
	while (!bad && (opt = dsp_getopt(&g, argc - 1, argv, "pzca:")) != -1) {
		if (opt == 'p') rq->use_fir_p = 1;
		else if (opt == 'c') rq->centre_ref = 1;
		else if (opt == 'z') {
			LOG_FMT(LL_ERROR, "%s: warning: zita_convolver not available; using fir_p instead", name);
			rq->use_fir_p = 1;
		}
		else if (opt == 'a') {
			const double deg = strtod(g.arg, &end);
			CHECK_ENDPTR(g.arg, end, "angle", return -1);
			rq->angle = deg / 180.0 * M_PI;
		}
		else {
			dsp_getopt_print_error(&g, opt, name);
			bad = 1;
		}
	}
	if (bad || g.ind != argc - 1) {
		print_effect_usage(ei);
		return -1;
	}
	rq->taps = strtol(argv[g.ind], &end, 10);
	CHECK_ENDPTR(argv[g.ind], end, "taps", return -1);
	const char *why = (rq->taps <= 3) ? "taps must be > 3" : (rq->taps % 2 == 0) ? "taps must be odd" : NULL;
	if (why) {
		LOG_FMT(LL_ERROR, "%s: error: %s", name, why);
		return -1;
	}
	return 0;
}

struct effect * hilbert_effect_init(const struct effect_info *ei, const struct stream_info *istream, const char *channel_selector, const char *dir, int argc, const char *const *argv)
{
	struct hilbert_request rq;
	if (hilbert_parse(ei, argc, argv, &rq)) return NULL;
	sample_t *h = calloc(rq.taps, sizeof(sample_t));
	if (check_alloc(ei->name, h)) return NULL;
	dspb200_hilbert_taps(rq.taps, rq.angle, h);
	struct effect * (*make)(const struct effect_info *, const struct stream_info *, const char *, sample_t *, int, ssize_t, ssize_t, int) =
		rq.use_fir_p ? fir_p_effect_init_with_filter : fir_effect_init_with_filter;
	struct effect *e = make(ei, istream, channel_selector, h, 1, rq.taps, rq.centre_ref ? rq.taps / 2 : 0, 0);
	free(h);
	return e;
}
