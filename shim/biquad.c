/*
 * shim/biquad.c -- drop-in replacement object for the reference's biquad.o.
 *
 * Exports the reference's symbols with unchanged signatures (biquad.h:71-74):
 *   biquad_init, biquad_reset, biquad_init_using_type   (also used by crossfeed.c:143-146,
 *       matrix4.c:402-424, matrix4_common.c:409, matrix4_mb.c:135-136, reverse_iir.c:540 together with
 *       the inline biquad() of biquad.h, which stays)
 *   biquad_effect_init                                  (the 20 effect names of biquad.h:97-117)
 * The coefficient design is dspb200_biquad_design() (bit-identical to biquad.c:111-294, see
 * tests/test_abi.py); the per-block work (biquad.c:296-315) runs as the fused scan kernel K1.
 */
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include "biquad.h"
#include "reverse_iir.h"
#include "util.h"
#include "gpu_effect.h"

struct effect * gpu_riir_effect_new(const struct effect_info *, const struct stream_info *, const char *, const struct biquad_state *, double);

void biquad_reset(struct biquad_state *state)
{
	state->m0 = state->m1 = 0.0;
}

void biquad_init(struct biquad_state *state, double b0, double b1, double b2, double a0, double a1, double a2)
{
	state->c0 = b0 / a0;
	state->c1 = b1 / a0;
	state->c2 = b2 / a0;
	state->c3 = a1 / a0;
	state->c4 = a2 / a0;
	biquad_reset(state);
}

void biquad_init_using_type(struct biquad_state *b, int type, double fs, double arg0, double arg1, double arg2, double arg3, int width_type)
{
	double c[5] = { 1.0, 0.0, 0.0, 0.0, 0.0 };
	dspb200_biquad_design(type, fs, arg0, arg1, arg2, arg3, width_type, c);
	b->c0 = c[0]; b->c1 = c[1]; b->c2 = c[2]; b->c3 = c[3]; b->c4 = c[4];
	biquad_reset(b);
}

/* width[q|s|d|o|h|k] or bw<order>[.<index>] (Butterworth pole-pair Q), cf. biquad.c:27-89 */
static int parse_width_arg(const char *s, double *w, int *type)
{
	char *end;
	*type = BIQUAD_WIDTH_Q;
	if (strncmp(s, "bw", 2) == 0 && s[2] != '\0') {
		const long order = strtol(s + 2, &end, 10);
		long idx = 0;
		if (end == s + 2 || (*end != '\0' && *end != '.')) return 1;
		if (order < 2) {
			LOG_FMT(LL_ERROR, "%s(): filter order must be >= 2", __func__);
			return 1;
		}
		if (*end == '.') {
			const char *p = end + 1;
			idx = strtol(p, &end, 10);
			if (end == p || *end != '\0') return 1;
		}
		if (idx < 0 || idx >= order / 2) {
			LOG_FMT(LL_ERROR, "%s(): filter index out of range", __func__);
			return 1;
		}
		*w = 1.0 / (2.0 * sin(M_PI / order * ((order / 2 - idx) - 0.5)));
		return 0;
	}
	*w = strtod(s, &end);
	if (end == s) return 1;
	switch (*end) {
	case 'q': *type = BIQUAD_WIDTH_Q; ++end; break;
	case 's': *type = BIQUAD_WIDTH_SLOPE; ++end; break;
	case 'd': *type = BIQUAD_WIDTH_SLOPE_DB; ++end; break;
	case 'o': *type = BIQUAD_WIDTH_BW_OCT; ++end; break;
	case 'k': *w *= 1000.0; /* fall through */
	case 'h': *type = BIQUAD_WIDTH_BW_HZ; ++end; break;
	}
	return (*end != '\0');
}

/* positional argument kinds per effect: f = frequency, w = width, g = plain number */
static const char * arg_signature(int effect_number)
{
	switch (effect_number) {
	case BIQUAD_LOWPASS_1: case BIQUAD_HIGHPASS_1: case BIQUAD_ALLPASS_1: case BIQUAD_LOWPASS_1P: return "f";
	case BIQUAD_LOWSHELF_1: case BIQUAD_HIGHSHELF_1: return "fg";
	case BIQUAD_LOWPASS: case BIQUAD_HIGHPASS: case BIQUAD_BANDPASS_SKIRT: case BIQUAD_BANDPASS_PEAK:
	case BIQUAD_NOTCH: case BIQUAD_ALLPASS: return "fw";
	case BIQUAD_PEAK: case BIQUAD_LOWSHELF: case BIQUAD_HIGHSHELF: return "fwg";
	case BIQUAD_LOWPASS_TRANSFORM: case BIQUAD_HIGHPASS_TRANSFORM: return "fwfw";
	case BIQUAD_DEEMPH: return "";
	case BIQUAD_BIQUAD: return "gggggg";
	}
	return NULL;
}

static const char * arg_name(const char *sig, int i)
{
	static const char *const bq[] = { "b0", "b1", "b2", "a0", "a1", "a2" };
	static const char *const tr[] = { "fz", "width_z", "fp", "width_p" };
	if (sig[0] == 'g') return bq[i];
	if (strlen(sig) == 4) return tr[i];
	return (sig[i] == 'f') ? "f0" : (sig[i] == 'w') ? "width" : "gain";
}

struct effect * biquad_effect_init(const struct effect_info *ei, const struct stream_info *istream, const char *channel_selector, const char *dir, int argc, const char *const *argv)
{
	const char *sig = arg_signature(ei->effect_number);
	if (!sig) {
		dsp_perror(DSP_ENOEFFNUM, __FILE__, NULL);
		return NULL;
	}
	const int n_args = (int) strlen(sig);
	int reverse = 0, opt;
	double thresh = 80.0;
	char *endptr;
	struct dsp_getopt_state g = DSP_GETOPT_STATE_INITIALIZER;
	/* options are only looked for in front of the positional arguments (negative gains are not options) */
	while ((opt = dsp_getopt(&g, argc - n_args, argv, "r::")) != -1) {
		if (opt != 'r') {
			dsp_getopt_print_error(&g, opt, argv[0]);
			print_effect_usage(ei);
			return NULL;
		}
		reverse = 1;
		if (g.arg) {
			thresh = strtol(g.arg, &endptr, 10);
			CHECK_ENDPTR(g.arg, endptr, "thresh", return NULL);
			CHECK_RANGE(thresh >= 10.0 && thresh <= 200.0, "thresh", return NULL);
		}
	}
	if (argc - g.ind != n_args) {
		print_effect_usage(ei);
		return NULL;
	}

	double v[6] = { 0.0, 0.0, 0.0, 0.0, 0.0, 0.0 };
	int width_type = BIQUAD_WIDTH_Q, n_widths = 0;
	for (int i = 0; i < n_args; ++i) {
		const char *s = argv[g.ind + i];
		const char *what = arg_name(sig, i);
		switch (sig[i]) {
		case 'f':
			v[i] = parse_freq(s, &endptr);
			CHECK_ENDPTR(s, endptr, what, return NULL);
			CHECK_FREQ(v[i], istream->fs, what, return NULL);
			break;
		case 'w': {
			int wt;
			if (parse_width_arg(s, &v[i], &wt)) {
				dsp_perror(DSP_ETRCHAR, argv[0], what);
				return NULL;
			}
			CHECK_RANGE(v[i] > 0.0, what, return NULL);
			width_type = wt;
			++n_widths;
			/* biquad.c:466-489: slope widths only for the shelves, Q only for the transforms */
			const int is_slope = (wt == BIQUAD_WIDTH_SLOPE || wt == BIQUAD_WIDTH_SLOPE_DB);
			const int shelf = (ei->effect_number == BIQUAD_LOWSHELF || ei->effect_number == BIQUAD_HIGHSHELF);
			if ((is_slope && !shelf) || (n_args == 4 && wt != BIQUAD_WIDTH_Q)) {
				LOG_FMT(LL_ERROR, "%s: error: invalid width type", argv[0]);
				return NULL;
			}
			break;
		}
		default:
			v[i] = strtod(s, &endptr);
			CHECK_ENDPTR(s, endptr, what, return NULL);
		}
	}
	(void) n_widths;

	struct biquad_state b = { 0 };
	if (ei->effect_number == BIQUAD_BIQUAD)
		biquad_init(&b, v[0], v[1], v[2], v[3], v[4], v[5]);
	else if (ei->effect_number == BIQUAD_DEEMPH) {
		/* biquad.c:497-515 */
		double f0, slope, gain;
		if (istream->fs == 44100) { f0 = 5283; slope = 0.4845; gain = -9.477; }
		else if (istream->fs == 48000) { f0 = 5356; slope = 0.479; gain = -9.62; }
		else {
			LOG_FMT(LL_ERROR, "%s: error: sample rate must be 44100 or 48000", argv[0]);
			return NULL;
		}
		biquad_init_using_type(&b, BIQUAD_HIGHSHELF, istream->fs, f0, slope, gain, 0.0, BIQUAD_WIDTH_SLOPE);
	}
	else {
		/* map the positional values onto (arg0 = f0|fz, arg1 = width|qz, arg2 = gain|fp, arg3 = qp) */
		double a0 = v[0], a1 = 0.0, a2 = 0.0, a3 = 0.0;
		if (strcmp(sig, "fg") == 0) a2 = v[1];
		else if (strcmp(sig, "fw") == 0) a1 = v[1];
		else if (strcmp(sig, "fwg") == 0) { a1 = v[1]; a2 = v[2]; }
		else if (strcmp(sig, "fwfw") == 0) { a1 = v[1]; a2 = v[2]; a3 = v[3]; }
		biquad_init_using_type(&b, ei->effect_number, istream->fs, a0, a1, a2, a3, width_type);
	}

	if (reverse)
		return gpu_riir_effect_new(ei, istream, channel_selector, &b, thresh);   /* shim/riir.c: the reference's design, the device's convolution */

	const int C = istream->channels;
	dspb200_chain *chain = gpu_chain_new(ei->name, istream);
	struct gpu_part *part = gpu_part_new(GPU_PART_BIQUAD, channel_selector, C);
	double *coefs = calloc((size_t) C * 5, sizeof(double));
	if (!chain || !part || !coefs || !(part->bq = calloc(C, sizeof(struct biquad_state)))) {
		if (chain && (!part || !coefs)) dsp_perror(DSP_ENOMEM, ei->name, NULL);
		goto fail;
	}
	for (int k = 0; k < C; ++k) {
		double *c = &coefs[(size_t) k * 5];
		if (GET_BIT(channel_selector, k)) {
			part->bq[k] = b;
			c[0] = b.c0; c[1] = b.c1; c[2] = b.c2; c[3] = b.c3; c[4] = b.c4;
		}
		else c[0] = 1.0;   /* identity section: the channel passes through, biquad.c:301-303 */
	}
	if (dspb200_chain_add_biquad(chain, 1, coefs) != 0) {
		LOG_FMT(LL_ERROR, "%s: error: %s", ei->name, dspb200_last_error());
		goto fail;
	}
	free(coefs);
	struct effect *e = gpu_effect_new(ei, istream, chain, part);
	if (!e) {
		coefs = NULL;
		goto fail;
	}
	return e;

	fail:
	free(coefs);
	gpu_part_free(part);
	dspb200_chain_destroy(chain);
	return NULL;
}
