/*
 * shim/resample.c -- drop-in replacement object for the reference's resample.o (CLI build only,
 * resample.h:22: !SYMMETRIC_IO).
 *
 * Exports resample_effect_init (resample.h:26).  Argument grammar and checks as in
 * resample.c:213-252: `resample [bandwidth] fs[k]|x{mult}|/{div}`, bandwidth in [0.7, 0.999];
 * equal rates return an effect without run() which the chain drops (resample.c:254-259,
 * effects_chain.c:586-590).  Hooks as in resample.c:260-267: run writes obuf and changes *frames,
 * reset, drain2, destroy; flags = CH_DEPS_IDENTITY only; no plot, no merge (the rate changes).
 */
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include "resample.h"
#include "util.h"
#include "gpu_effect.h"

static void gpu_resample_reset(struct effect *e)
{
	struct gpu_effect_state *state = (struct gpu_effect_state *) e->data;
	if (state->head) return;   /* a passenger: the head of the run resets the whole device chain */
	dspb200_chain_reset(state->chain);
}

static sample_t * gpu_resample_drain2(struct effect *e, ssize_t *frames, sample_t *buf1, sample_t *buf2)
{
	struct gpu_effect_state *state = (struct gpu_effect_state *) e->data;
	/* as a passenger the resampler lives in the head's device chain: drain that one (it runs what follows, too) */
	const long r = dspb200_chain_drain_host(state->head ? state->head->chain : state->chain, *frames, buf2);
	if (r < 0) {
		*frames = -1;
		return buf1;
	}
	*frames = r;
	return buf2;
}

static void gpu_resample_destroy(struct effect *e)
{
	struct gpu_effect_state *state = (struct gpu_effect_state *) e->data;
	if (!state) return;
	dspb200_chain_destroy(state->chain);
	gpu_part_free(state->parts);
	free(state);
}

struct effect * resample_effect_init(const struct effect_info *ei, const struct stream_info *istream, const char *channel_selector, const char *dir, int argc, const char *const *argv)
{
	char *endptr;
	double bw = 0.0;
	int rate;

	if (argc < 2 || argc > 3) {
		print_effect_usage(ei);
		return NULL;
	}
	const char *rate_arg = argv[argc - 1];
	if (argc == 3) {
		bw = strtod(argv[1], &endptr);
		CHECK_ENDPTR(argv[1], endptr, "bandwidth", return NULL);
		CHECK_RANGE(bw >= 0.7 && bw <= 0.999, "bandwidth", return NULL);
	}
	if (rate_arg[0] == 'x') {
		rate = istream->fs * strtol(rate_arg + 1, &endptr, 10);
		CHECK_ENDPTR(rate_arg, endptr, "fs multiplier", return NULL);
	}
	else if (rate_arg[0] == '/') {
		const int rate_div = strtol(rate_arg + 1, &endptr, 10);
		CHECK_ENDPTR(rate_arg, endptr, "fs divisor", return NULL);
		if (rate_div == 0 || istream->fs % rate_div != 0) {
			LOG_FMT(LL_ERROR, "%s: error: %d is not a factor of %d", argv[0], rate_div, istream->fs);
			return NULL;
		}
		rate = istream->fs / rate_div;
	}
	else {
		rate = lround(parse_freq(rate_arg, &endptr));
		CHECK_ENDPTR(rate_arg, endptr, "fs", return NULL);
	}
	CHECK_RANGE(rate > 0, "rate", return NULL);

	struct effect *e = calloc(1, sizeof(struct effect));
	if (check_alloc(ei->name, e)) return NULL;
	if (rate == istream->fs) {
		LOG_FMT(LL_VERBOSE, "%s: info: sample rates match; no proccessing will be done", argv[0]);
		return e;   /* no run(): the chain discards it */
	}
	struct gpu_effect_state *state = calloc(1, sizeof(struct gpu_effect_state));
	if (check_alloc(ei->name, state)) goto fail;
	state->chain = gpu_chain_new(ei->name, istream);
	if (!state->chain) goto fail;
	if (dspb200_chain_add_resample(state->chain, rate, bw) != 0) {
		LOG_FMT(LL_ERROR, "%s: error: %s", ei->name, dspb200_last_error());
		goto fail;
	}
	long p[8];
	if (dspb200_resample_params(istream->fs, rate, bw, p) == 0)
		LOG_FMT(LL_VERBOSE, "%s: info: ratio=%ld/%ld filter_len=%ld in_len=%ld out_len=%ld (B200 polyphase, %ld taps/phase)",
			argv[0], p[0], p[1], p[2] + 1, p[3], p[4], p[7]);
	e->name = ei->name;
	e->istream.fs = istream->fs;
	e->ostream.fs = rate;
	e->istream.channels = e->ostream.channels = istream->channels;
	e->flags |= EFFECT_FLAG_CH_DEPS_IDENTITY;
	gpu_register_effect_kind(gpu_resample_destroy);
	e->run = gpu_linked_run;   /* out of place: the chain holds a rate-changing operator */
	e->reset = gpu_resample_reset;
	e->drain2 = gpu_resample_drain2;
	e->destroy = gpu_resample_destroy;
	e->data = state;
	return e;

	fail:
	if (state) dspb200_chain_destroy(state->chain);
	free(state);
	free(e);
	return NULL;
}
