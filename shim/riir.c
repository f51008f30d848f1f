/*
 * shim/riir.c -- `biquad -r` (reverse IIR, reverse_iir.c) on the device.
 *
 * The reference realises a time-reversed IIR section as Vicanek's truncated cascade: per pole a product of stages
 * y[n] = p^(2^j) x[n] + x[n - 2^j], j = 0..N-1 (reverse_iir.c:79-92), sections in parallel form plus a short FIR
 * part (:104-139), possibly a cascade of such blocks (:143-150).  Every stage is a finite delay line, so the whole
 * effect is a LINEAR, TIME-INVARIANT filter with a FINITE impulse response of exactly latency + 1 frames per
 * channel (latency = sum over the cascade of 2^N + fir.n - 1, :621-623) -- a plain FIR, which is what K2 computes.
 *
 * So the design (partial fractions, residues, pole sorting, cascade splitting: reverse_iir.c:381-630) stays the
 * reference's own code, unmodified and unduplicated: this wrapper owns the reference effect ("inner"), lets the
 * chain merge and prepare it exactly as before (merge appends sections per channel, :632-643; prepare builds the
 * filter), then measures its impulse response once -- by running the reference's run() on a unit impulse -- and
 * hands the taps (one column per active channel) to the device convolution engine.  plot / drain_samples /
 * channel_offsets keep coming from the inner effect, so alignment and drain bookkeeping are the reference's.
 */
#include <stdlib.h>
#include <string.h>
#include "reverse_iir.h"
#include "util.h"
#include "gpu_effect.h"

struct riir_wrap {
	struct gpu_effect_state g;  /* g.chain: device FIR, built in prepare() (first member: the common GPU-effect state) */
	struct effect *inner;       /* the reference's reverse_iir effect: design, plot, offsets */
};

static void riir_wrap_reset(struct effect *e)
{
	struct riir_wrap *w = (struct riir_wrap *) e->data;
	if (w->g.chain && !w->g.head) dspb200_chain_reset(w->g.chain);
}

static void riir_wrap_plot(struct effect *e, int i)
{
	struct riir_wrap *w = (struct riir_wrap *) e->data;
	if (w->inner->plot) w->inner->plot(w->inner, i);
}

static void riir_wrap_drain_samples(struct effect *e, ssize_t *drain_samples)
{
	struct riir_wrap *w = (struct riir_wrap *) e->data;
	if (w->inner->drain_samples) w->inner->drain_samples(w->inner, drain_samples);
}

static void riir_wrap_channel_offsets(struct effect *e, ssize_t *latency, ssize_t *req_delay)
{
	struct riir_wrap *w = (struct riir_wrap *) e->data;
	if (w->inner->channel_offsets) w->inner->channel_offsets(w->inner, latency, req_delay);
}

static void riir_wrap_destroy(struct effect *e)
{
	struct riir_wrap *w = (struct riir_wrap *) e->data;
	if (!w) return;
	dspb200_chain_destroy(w->g.chain);
	if (w->inner) {
		if (w->inner->destroy) w->inner->destroy(w->inner);
		free(w->inner);
	}
	free(w);
}

static int riir_wrap_merge(struct effect *dest, struct effect *src)
{
	if (dest->merge != src->merge) return 0;
	struct riir_wrap *d = (struct riir_wrap *) dest->data, *s = (struct riir_wrap *) src->data;
	if (d->g.chain || s->g.chain || !d->inner->merge) return 0;   /* only before prepare(), as in the reference */
	return d->inner->merge(d->inner, s->inner);
}

static int riir_wrap_prepare(struct effect *e)
{
	struct riir_wrap *w = (struct riir_wrap *) e->data;
	struct effect *in = w->inner;
	const int C = e->istream.channels;
	if (in->prepare && in->prepare(in)) return 1;

	/* active channels and their response lengths: req_delay[k] = -latency_k (reverse_iir.c:250-255) */
	ssize_t *lat = calloc(C, sizeof(ssize_t)), *req = calloc(C, sizeof(ssize_t));
	char *sel = NEW_SELECTOR(C);
	if (!lat || !req || !sel) {
		dsp_perror(DSP_ENOMEM, e->name, NULL);
		goto fail;
	}
	if (in->channel_offsets) in->channel_offsets(in, lat, req);
	ssize_t len = 0;
	int n_act = 0;
	for (int k = 0; k < C; ++k) {
		if (req[k] < 0) {
			SET_BIT(sel, k);
			++n_act;
			if (-req[k] + 1 > len) len = -req[k] + 1;
		}
	}
	w->g.chain = gpu_chain_new(e->name, &e->istream);
	if (!w->g.chain) goto fail;
	if (n_act > 0) {
		/* impulse response of every active channel: one pass of the reference's own run() over a unit impulse */
		sample_t *buf = calloc((size_t) len * C, sizeof(sample_t)), *scratch = calloc((size_t) len * C, sizeof(sample_t));
		sample_t *taps = calloc((size_t) len * n_act, sizeof(sample_t));
		if (!buf || !scratch || !taps) {
			free(buf); free(scratch); free(taps);
			dsp_perror(DSP_ENOMEM, e->name, NULL);
			goto fail;
		}
		for (int k = 0; k < C; ++k) buf[k] = 1.0;
		ssize_t f = len;
		const sample_t *y = in->run(in, &f, buf, scratch);
		for (ssize_t i = 0; i < len; ++i) {
			int col = 0;
			for (int k = 0; k < C; ++k)
				if (GET_BIT(sel, k)) taps[i * n_act + col++] = y[i * C + k];
		}
		if (in->reset) in->reset(in);
		const int rc = dspb200_chain_add_fir(w->g.chain, sel, taps, n_act, len, 0, 0);
		free(buf); free(scratch); free(taps);
		if (rc != 0) {
			LOG_FMT(LL_ERROR, "%s: error: %s", e->name, dspb200_last_error());
			goto fail;
		}
		LOG_FMT(LL_VERBOSE, "%s: info: reverse IIR as a %zd-tap FIR on %d channel(s) (B200 partitioned convolution)", e->name, len, n_act);
	}
	free(lat); free(req); free(sel);
	return 0;

	fail:
	free(lat); free(req); free(sel);
	return 1;
}

struct effect * gpu_riir_effect_new(const struct effect_info *ei, const struct stream_info *istream, const char *channel_selector, const struct biquad_state *b, double thresh)
{
	struct effect *inner = reverse_iir_effect_init_from_biquad(ei, istream, channel_selector, b, thresh);
	if (!inner) return NULL;
	struct effect *e = calloc(1, sizeof(struct effect));
	struct riir_wrap *w = calloc(1, sizeof(struct riir_wrap));
	if (!e || !w) {
		dsp_perror(DSP_ENOMEM, ei->name, NULL);
		if (inner->destroy) inner->destroy(inner);
		free(inner); free(e); free(w);
		return NULL;
	}
	w->inner = inner;
	e->name = ei->name;
	e->istream = inner->istream;
	e->ostream = inner->ostream;
	e->flags = inner->flags;
	gpu_register_effect_kind(riir_wrap_destroy);
	e->run = gpu_linked_run;
	e->reset = riir_wrap_reset;
	e->plot = riir_wrap_plot;
	e->drain_samples = riir_wrap_drain_samples;
	e->channel_offsets = riir_wrap_channel_offsets;
	e->merge = riir_wrap_merge;
	e->prepare = riir_wrap_prepare;
	e->destroy = riir_wrap_destroy;
	e->data = w;
	return e;
}
