/*
 * shim/gpu_effect.c -- hooks shared by every GPU-backed effect.
 *
 * Mirrors, hook by hook, what the reference's effects do (file:line in the reference):
 *   run             biquad.c:296-315, fir.c:109-149, fir_p.c:127-181  -> one dspb200_chain_run_host()
 *   reset           biquad.c:317-323, fir.c:151-161, fir_p.c:183-207   -> dspb200_chain_reset()
 *   plot            biquad.c:325-337, fir.c:163-177, fir_p.c:209-233   (product of the parts' responses)
 *   drain_samples   fir.c:180-187, fir_p.c:235-240
 *   channel_offsets fir.c:208-217, fir_p.c:283-288
 *   merge           biquad.c:344-376 widened: ANY two zero-latency GPU effects of the same stream become
 *                   one device chain (effects_chain.c:605-641 offers the pairs), so a run of N effects costs
 *                   one H2D + one D2H per block, and cascaded biquads fuse into one kernel pass.
 *   destroy         frees chain and parts; the host free()s the struct itself (effect.c:78-85)
 */
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include "gpu_effect.h"
#include "util.h"

dspb200_chain * gpu_chain_new(const char *name, const struct stream_info *istream)
{
	int devices[64], n_devices = 0, slabs = 1;
	const char *env = getenv("DSP_B200_DEVICES");
	if (env) {
		char *end;
		while (*env != '\0' && n_devices < 64) {
			const long v = strtol(env, &end, 10);
			if (end == env) break;
			devices[n_devices++] = (int) v;
			env = (*end == ',') ? end + 1 : end;
		}
	}
	env = getenv("DSP_B200_SLABS");
	if (env) slabs = atoi(env);
	else {
		/* wide blocks over page-locked buffers (DSP_B200_PIN=1): copy-in, kernels and copy-out of channel slabs overlap.
		 * Pageable buffers stay in one piece: strided copies out of pageable memory are staged row by row (measured
		 * 3.1 ms against 1.3 ms per 8 MiB block). */
		const char *pin = getenv("DSP_B200_PIN");
		if (pin && pin[0] == '1') slabs = (istream->channels >= 128) ? 4 : (istream->channels >= 32) ? 2 : 1;
	}
	dspb200_chain *chain = dspb200_chain_create(istream->fs, istream->channels, (n_devices) ? devices : NULL, n_devices, slabs);
	if (!chain) LOG_FMT(LL_ERROR, "%s: error: %s", name, dspb200_last_error());
	return chain;
}

struct gpu_part * gpu_part_new(int kind, const char *channel_selector, int channels)
{
	struct gpu_part *p = calloc(1, sizeof(struct gpu_part));
	if (!p) return NULL;
	p->kind = kind;
	p->selector = NEW_SELECTOR(channels);
	if (!p->selector) {
		free(p);
		return NULL;
	}
	COPY_SELECTOR(p->selector, channel_selector, channels);
	return p;
}

void gpu_part_free(struct gpu_part *p)
{
	while (p) {
		struct gpu_part *next = p->next;
		free(p->selector);
		free(p->bq);
		free(p->gain);
		free(p->taps);
		free(p);
		p = next;
	}
}

/* ---- device hand-off between neighbouring GPU effects --------------------------------------------------- */
static void (*gpu_kinds[8])(struct effect *);
static int n_gpu_kinds;

void gpu_register_effect_kind(void (*destroy)(struct effect *))
{
	for (int i = 0; i < n_gpu_kinds; ++i)
		if (gpu_kinds[i] == destroy) return;
	if (n_gpu_kinds < (int) LENGTH(gpu_kinds)) gpu_kinds[n_gpu_kinds++] = destroy;
}

int gpu_effect_is(const struct effect *e)
{
	if (!e || !e->data || !e->destroy) return 0;
	for (int i = 0; i < n_gpu_kinds; ++i)
		if (gpu_kinds[i] == e->destroy) return 1;
	return 0;
}

static sample_t * gpu_passenger_run(struct effect *e, ssize_t *frames, sample_t *ibuf, sample_t *obuf)
{
	return ibuf;   /* this effect's operators already ran inside the device chain of the head of its run */
}

void gpu_link_neighbours(struct effect *e)
{
	struct gpu_effect_state *st = (struct gpu_effect_state *) e->data;
	if (st->link_checked) return;
	st->link_checked = 1;
	if (!st->head && !getenv("DSP_B200_NO_LINK")) {
		for (struct effect *n = e->next; gpu_effect_is(n); n = n->next) {
			struct gpu_effect_state *ns = (struct gpu_effect_state *) n->data;
			if (ns->head || !ns->chain || n->istream.channels != e->istream.channels) break;
			if (dspb200_chain_absorb(st->chain, ns->chain) != 0) break;
			LOG_FMT(LL_VERBOSE, "%s: info: device chain continues through %s (no host copy in between)", e->name, n->name);
			ns->head = st;
			ns->link_checked = 1;
			n->run = gpu_passenger_run;
		}
	}
	st->inplace = dspb200_chain_inplace_ok(st->chain);
}

/* run() of the head of a run of GPU effects (and of every GPU effect that stands alone) */
sample_t * gpu_linked_run(struct effect *e, ssize_t *frames, sample_t *ibuf, sample_t *obuf)
{
	struct gpu_effect_state *state = (struct gpu_effect_state *) e->data;
	gpu_link_neighbours(e);
	sample_t *dst = (state->inplace) ? ibuf : obuf;
	const long r = dspb200_chain_run_host(state->chain, *frames, ibuf, dst);
	if (r < 0) {
		if (!state->failed) {
			/* run() has no error channel: say so once; the block's content is whatever the copies left */
			state->failed = 1;
			LOG_FMT(LL_ERROR, "%s: error: device run failed, this block is not processed: %s", e->name, dspb200_last_error());
		}
		if (!state->inplace) *frames = 0;
		return dst;
	}
	*frames = r;
	return dst;
}

static void gpu_effect_reset(struct effect *e)
{
	struct gpu_effect_state *state = (struct gpu_effect_state *) e->data;
	if (state->head) return;   /* the head of the run resets the whole device chain */
	dspb200_chain_reset(state->chain);
}

/* one part's transfer function on channel k, as a gnuplot expression in w */
static void part_plot_expr(const struct gpu_part *p, int k)
{
	if (!GET_BIT(p->selector, k) || (p->kind == GPU_PART_GAIN && p->is_add)) {
		fputs("1.0", stdout);   /* `add` plots as unity, effect.c:96-100 */
		return;
	}
	if (p->kind == GPU_PART_GAIN) {
		printf("%.15e", p->gain[k]);
	}
	else if (p->kind == GPU_PART_BIQUAD) {
		printf("(" BIQUAD_PLOT_FMT ")", BIQUAD_PLOT_FMT_ARGS(&p->bq[k]));
	}
	else {
		int col = 0;
		if (p->fc > 1)
			for (int i = 0; i < k; ++i) col += (GET_BIT(p->selector, i)) ? 1 : 0;
		printf("exp(-j*w*%zd)*(0.0", -p->ref);
		for (ssize_t i = 0; i < p->frames; ++i)
			printf("+exp(-j*w*%zd)*%.15e", i, p->taps[i * p->fc + col]);
		putchar(')');
	}
}

static void gpu_effect_plot(struct effect *e, int i)
{
	struct gpu_effect_state *state = (struct gpu_effect_state *) e->data;
	for (int k = 0; k < e->ostream.channels; ++k) {
		int any = 0;
		for (struct gpu_part *p = state->parts; p; p = p->next) any |= (GET_BIT(p->selector, k)) ? 1 : 0;
		if (!any) {
			printf("H%d_%d(w)=1.0\n", k, i);
			continue;
		}
		printf("H%d_%d(w)=(abs(w)<=pi)?", k, i);
		for (struct gpu_part *p = state->parts; p; p = p->next) {
			if (p != state->parts) putchar('*');
			part_plot_expr(p, k);
		}
		puts(":0/0");
	}
}

static void gpu_effect_drain_samples(struct effect *e, ssize_t *drain_samples)
{
	struct gpu_effect_state *state = (struct gpu_effect_state *) e->data;
	for (struct gpu_part *p = state->parts; p; p = p->next) {
		if (p->kind != GPU_PART_FIR) continue;
		for (int k = 0; k < e->ostream.channels; ++k)
			if (GET_BIT(p->selector, k)) drain_samples[k] += p->latency + p->frames - 1;
	}
}

static void gpu_effect_channel_offsets(struct effect *e, ssize_t *latency, ssize_t *req_delay)
{
	struct gpu_effect_state *state = (struct gpu_effect_state *) e->data;
	for (struct gpu_part *p = state->parts; p; p = p->next) {
		if (p->kind != GPU_PART_FIR) continue;
		for (int k = 0; k < e->istream.channels; ++k) {
			if (GET_BIT(p->selector, k)) {
				latency[k] += p->latency;
				req_delay[k] -= p->ref;
			}
		}
	}
}

static void gpu_effect_destroy(struct effect *e)
{
	struct gpu_effect_state *state = (struct gpu_effect_state *) e->data;
	if (!state) return;
	dspb200_chain_destroy(state->chain);
	gpu_part_free(state->parts);
	free(state);
}

static int parts_have_latency(const struct gpu_part *p)
{
	for (; p; p = p->next)
		if (p->kind == GPU_PART_FIR && p->latency != 0) return 1;
	return 0;
}

static int gpu_effect_merge(struct effect *dest, struct effect *src)
{
	if (dest->merge != src->merge) return 0;
	struct gpu_effect_state *d = (struct gpu_effect_state *) dest->data, *s = (struct gpu_effect_state *) src->data;
	/* an effect that reports latency keeps its own slot so the chain's align pass sees it (effects_chain.c:727) */
	if (parts_have_latency(d->parts) || parts_have_latency(s->parts)) return 0;
	if (getenv("DSP_B200_NO_MERGE")) return 0;
	/* the optimizer also offers pairs with skipped effects in between; only hop over effects that
	 * declare themselves reorderable (LTI and channel-wise), so the signal path is unchanged */
	if (dest->next != src && !(src->flags & EFFECT_FLAG_OPT_REORDERABLE)) return 0;
	for (struct effect *between = dest->next; between && between != src; between = between->next)
		if (!(between->flags & EFFECT_FLAG_OPT_REORDERABLE)) return 0;
	if (dspb200_chain_absorb(d->chain, s->chain) != 0) return 0;
	struct gpu_part **tail = &d->parts;
	while (*tail) tail = &(*tail)->next;
	*tail = s->parts;
	s->parts = NULL;
	if (!(src->flags & EFFECT_FLAG_OPT_REORDERABLE)) dest->flags &= ~EFFECT_FLAG_OPT_REORDERABLE;
	if (!dest->drain_samples) dest->drain_samples = src->drain_samples;
	if (!dest->channel_offsets) dest->channel_offsets = src->channel_offsets;
	return 1;
}

struct effect * gpu_effect_new(const struct effect_info *ei, const struct stream_info *istream, dspb200_chain *chain, struct gpu_part *part)
{
	struct effect *e = calloc(1, sizeof(struct effect));
	struct gpu_effect_state *state = calloc(1, sizeof(struct gpu_effect_state));
	if (!e || !state) {
		dsp_perror(DSP_ENOMEM, ei->name, NULL);
		free(e);
		free(state);
		return NULL;
	}
	state->chain = chain;
	state->parts = part;
	e->name = ei->name;
	e->istream.fs = e->ostream.fs = istream->fs;
	e->istream.channels = e->ostream.channels = istream->channels;
	e->flags |= EFFECT_FLAG_OPT_REORDERABLE;
	e->flags |= EFFECT_FLAG_CH_DEPS_IDENTITY;
	gpu_register_effect_kind(gpu_effect_destroy);
	e->run = gpu_linked_run;
	e->reset = gpu_effect_reset;
	e->plot = gpu_effect_plot;
	e->destroy = gpu_effect_destroy;
	e->merge = gpu_effect_merge;
	if (part->kind == GPU_PART_FIR) {
		e->drain_samples = gpu_effect_drain_samples;
		e->channel_offsets = gpu_effect_channel_offsets;
	}
	e->data = state;
	return e;
}
