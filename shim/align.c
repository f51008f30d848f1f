/*
 * shim/align.c -- the chain's channel-alignment effect on the device (replaces the reference's align.o in the drop-in
 * builds; the reference object stays linked under the name ref_align_effect_insert for chains without GPU effects).
 *
 * Exports align_effect_insert (align.h:27), which effects_chain.c:744-864 calls after every effect whose channels
 * have drifted apart.  The arithmetic of the insertion is the reference's (align.c:95-162):
 *   - nothing to do when every offsets[k] already equals its target (align_refs[k], or 0 without refs)
 *   - target of channel k: align_refs[k], or the largest offset (at least 0 at the end of the chain, so that
 *     negative offsets are zeroed there); channel k is delayed by target - offsets[k] frames
 *   - the smallest target, if positive, is latency every channel shares: it is subtracted from all offsets and that
 *     many frames are dropped from the head of the stream (CLI build; not under SYMMETRIC_IO)
 * The delays and the dropped frames then happen in ONE device operator (dsp_b200/csrc/delay.cu), and -- because
 * the effect is a GPU effect -- the device chain of a latency-bearing `fir` in front of it and of whatever GPU effect
 * follows continues through it without a host copy (gpu_link_neighbours).
 */
#include <stdlib.h>
#include <string.h>
#include "align.h"
#include "util.h"
#include "list_util.h"
#include "gpu_effect.h"

int ref_align_effect_insert(struct effects_chain *, struct effect *, ssize_t *, ssize_t *);   /* reference align.o, symbol renamed at link time */

static void gpu_align_reset(struct effect *e)
{
	struct gpu_effect_state *state = (struct gpu_effect_state *) e->data;
	if (state->head) return;
	dspb200_chain_reset(state->chain);
}

static void gpu_align_drain_samples(struct effect *e, ssize_t *drain_samples)
{
	struct gpu_effect_state *state = (struct gpu_effect_state *) e->data;
	for (int k = 0; k < e->istream.channels; ++k) drain_samples[k] += state->align_len[k];   /* align.c:78-83 */
}

static void gpu_align_destroy(struct effect *e)
{
	struct gpu_effect_state *state = (struct gpu_effect_state *) e->data;
	if (!state) return;
	dspb200_chain_destroy(state->chain);
	free(state->align_len);
	free(state);
}

/* per-channel targets and the largest offset (*top); returns whether any channel has to move */
static int align_targets(const struct effect *prev, const ssize_t *offsets, const ssize_t *align_refs, ssize_t *target, ssize_t *top_out)
{
	const int C = prev->ostream.channels;
	ssize_t top = (prev->next) ? offsets[0] : 0;
	for (int k = 0; k < C; ++k) top = MAXIMUM(top, offsets[k]);
	int moves = 0;
	for (int k = 0; k < C; ++k) {
		target[k] = (align_refs) ? align_refs[k] : top;
		if ((align_refs ? align_refs[k] : 0) != offsets[k]) moves = 1;
	}
	*top_out = top;
	return moves;
}

int align_effect_insert(struct effects_chain *chain, struct effect *prev, ssize_t *offsets, ssize_t *align_refs)
{
	/* a chain without a GPU effect next to this spot keeps the reference's host implementation */
	if (getenv("DSP_B200_CPU_ALIGN") || !(gpu_effect_is(prev) || gpu_effect_is(prev->next)))
		return ref_align_effect_insert(chain, prev, offsets, align_refs);

	const int C = prev->ostream.channels;
	const char *next_name = (prev->next) ? prev->next->name : "[end of chain]";
	ssize_t *target = calloc(C, sizeof(ssize_t));
	long *delay = calloc(C, sizeof(long));
	struct effect *e = NULL;
	struct gpu_effect_state *state = NULL;
	if (!target || !delay) goto nomem;
	ssize_t top = 0;
	if (!align_targets(prev, offsets, align_refs, target, &top)) {
		LOG_FMT(LL_VERBOSE, "info: no alignment needed: %s", next_name);
		free(target); free(delay);
		return 0;
	}
	e = calloc(1, sizeof(struct effect));
	state = calloc(1, sizeof(struct gpu_effect_state));
	if (!e || !state || !(state->align_len = calloc(C, sizeof(ssize_t)))) goto nomem;

	ssize_t shared = top;
	for (int k = 0; k < C; ++k) {
		shared = MINIMUM(shared, target[k]);
		if (target[k] < offsets[k]) {
			LOG_FMT(LL_ERROR, "align (%s): error: channel %d would need a negative delay", next_name, k);
			goto fail;
		}
		delay[k] = (long) (target[k] - offsets[k]);
		state->align_len[k] = target[k] - offsets[k];
		if (delay[k] != 0) LOG_FMT(LL_VERBOSE, "align (%s): info: channel %d: %ld", next_name, k, delay[k]);
		offsets[k] = target[k];
	}
	long discard = 0;
	if (shared > 0) {
		for (int k = 0; k < C; ++k) offsets[k] -= shared;
#ifndef SYMMETRIC_IO
		discard = (long) shared;
#endif
		LOG_FMT(LL_VERBOSE, "align (%s): info: discarding %zd frames", next_name, shared);
	}

	e->name = "align";
	e->istream.fs = e->ostream.fs = prev->ostream.fs;
	e->istream.channels = e->ostream.channels = C;
	e->flags |= EFFECT_FLAG_CH_DEPS_IDENTITY;
	state->chain = gpu_chain_new(e->name, &e->istream);
	if (!state->chain) goto fail;
	if (dspb200_chain_add_align(state->chain, delay, discard) != 0) {
		LOG_FMT(LL_ERROR, "align: error: %s", dspb200_last_error());
		goto fail;
	}
	gpu_register_effect_kind(gpu_align_destroy);
	e->run = gpu_linked_run;
	e->reset = gpu_align_reset;
	e->plot = effect_plot_noop;
	e->drain_samples = gpu_align_drain_samples;
	e->destroy = gpu_align_destroy;
	e->data = state;
	LIST_INSERT(chain, e, prev);
	free(target); free(delay);
	return 0;

	nomem:
	dsp_perror(DSP_ENOMEM, "align", NULL);
	fail:
	if (state) {
		dspb200_chain_destroy(state->chain);
		free(state->align_len);
	}
	free(state); free(e); free(target); free(delay);
	return 1;
}
