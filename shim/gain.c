/*
 * shim/gain.c -- drop-in replacement object for the reference's gain.o (gain, mult, add; gain.h:33-36).
 *
 * Not one of the heavy kernels, but `gain` is what real chains start with (config 2: `gain -12` in front
 * of ten `eq`).  As a reference CPU effect it would cost a pass over the block on the host AND split the
 * run of GPU effects; as a GPU part it merges into the neighbouring device chain (one H2D/D2H per block).
 * Semantics of gain.c:80-141: `gain dB` -> 10^(dB/20), `mult x`, `add x`; unselected channels carry the
 * neutral element; gain/mult are flagged reorderable, add is not; plot as gain.c:43-49 / effect.c:96-100.
 */
#include <stdlib.h>
#include <math.h>
#include "gain.h"
#include "util.h"
#include "gpu_effect.h"

struct effect * gain_effect_init(const struct effect_info *ei, const struct stream_info *istream, const char *channel_selector, const char *dir, int argc, const char *const *argv)
{
	char *endptr;
	if (argc != 2) {
		print_effect_usage(ei);
		return NULL;
	}
	const char *arg = argv[1];
	double v = strtod(arg, &endptr);
	const char *what = "value";
	switch (ei->effect_number) {
	case GAIN_EFFECT_NUMBER_GAIN: what = "gain"; break;
	case GAIN_EFFECT_NUMBER_MULT: what = "multiplier"; break;
	case GAIN_EFFECT_NUMBER_ADD: break;
	default:
		dsp_perror(DSP_ENOEFFNUM, __FILE__, NULL);
		return NULL;
	}
	CHECK_ENDPTR(arg, endptr, what, return NULL);
	if (ei->effect_number == GAIN_EFFECT_NUMBER_GAIN) v = pow(10.0, v / 20.0);
	const int is_add = (ei->effect_number == GAIN_EFFECT_NUMBER_ADD);

	const int C = istream->channels;
	dspb200_chain *chain = gpu_chain_new(ei->name, istream);
	struct gpu_part *part = gpu_part_new(GPU_PART_GAIN, channel_selector, C);
	double *mult = calloc(C, sizeof(double)), *add = calloc(C, sizeof(double));
	if (!chain || !part || !mult || !add || !(part->gain = calloc(C, sizeof(sample_t)))) {
		if (chain) dsp_perror(DSP_ENOMEM, ei->name, NULL);
		goto fail;
	}
	part->is_add = is_add;
	for (int k = 0; k < C; ++k) {
		const int sel = GET_BIT(channel_selector, k) ? 1 : 0;
		mult[k] = (!is_add && sel) ? v : 1.0;
		add[k] = (is_add && sel) ? v : 0.0;
		part->gain[k] = is_add ? add[k] : mult[k];
	}
	if (dspb200_chain_add_gain(chain, mult, add) != 0) {
		LOG_FMT(LL_ERROR, "%s: error: %s", ei->name, dspb200_last_error());
		goto fail;
	}
	free(mult);
	free(add);
	mult = add = NULL;
	struct effect *e = gpu_effect_new(ei, istream, chain, part);
	if (!e) goto fail;
	if (is_add) e->flags &= ~EFFECT_FLAG_OPT_REORDERABLE;   // gain.c:120-131: only gain/mult may be reordered
	return e;

	fail:
	free(mult);
	free(add);
	gpu_part_free(part);
	dspb200_chain_destroy(chain);
	return NULL;
}
