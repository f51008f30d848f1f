/*
 * oracle/ref_driver.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Thin C-ABI harness around the UNMODIFIED reference objects (compiled from
 * /root/reference by oracle/Makefile into oracle/_ref/libdspref.so).  It plays the
 * role of a frontend (cf. dsp.c:1418-1431 / ladspa_dsp.c:316-355): it owns
 * `dsp_globals` and the log lock, builds a chain from a chain string
 * (build_effects_chain_from_string, effects_chain.h:42), and drives
 * run_effects_chain / drain_effects_chain (effects_chain.h:47,52) block by block.
 * It also exposes the reference's `sgen` codec (sgen.c) so parity inputs are the
 * reference's own arithmetic.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include "dsp.h"
#include "effect.h"
#include "effects_chain.h"
#include "codec.h"
#include "util.h"

struct dsp_globals dsp_globals = { LL_ERROR, "dspref" };
static pthread_mutex_t log_lock = PTHREAD_MUTEX_INITIALIZER;

void dsp_log_acquire(void) { pthread_mutex_lock(&log_lock); }
void dsp_log_release(void) { pthread_mutex_unlock(&log_lock); }
#ifdef DSP_STATUSLINES
void dsp_statuslines_acquire(void) { pthread_mutex_lock(&log_lock); }
void dsp_statuslines_release(void) { pthread_mutex_unlock(&log_lock); }
void dsp_statusline_register(struct statusline_state *s) { (void) s; }
void dsp_statusline_unregister(struct statusline_state *s) { (void) s; }
void dsp_get_term_size(int *rows, int *cols) { *rows = 24; *cols = 80; }
#endif

struct ref_chain {
	struct effects_chain chain;
	struct stream_info in, out;
	sample_t *buf1, *buf2;
	ssize_t buf_frames;   /* input frames the buffers are sized for */
	ssize_t buf_len;
};

void dspref_set_loglevel(int l) { dsp_globals.loglevel = l; }

static int ensure_bufs(struct ref_chain *rc, ssize_t frames)
{
	if (frames <= rc->buf_frames) return 0;
	const ssize_t len = get_effects_chain_buffer_len(&rc->chain, frames, rc->in.channels);
	sample_t *b1 = realloc(rc->buf1, len * sizeof(sample_t));
	if (b1) rc->buf1 = b1;
	sample_t *b2 = realloc(rc->buf2, len * sizeof(sample_t));
	if (b2) rc->buf2 = b2;
	if (!b1 || !b2) return 1;
	rc->buf_frames = frames;
	rc->buf_len = len;
	return 0;
}

void * dspref_chain_new(const char *chain_str, int fs, int channels, const char *dir)
{
	struct ref_chain *rc = calloc(1, sizeof(*rc));
	if (!rc) return NULL;
	rc->in.fs = fs;
	rc->in.channels = channels;
	struct stream_info stream = rc->in;
	char *mask = NEW_SELECTOR(channels);
	SET_SELECTOR(mask, channels);
	const int err = build_effects_chain_from_string(chain_str, dir, &rc->chain, &stream, mask, (dir) ? dir : ".");
	free(mask);
	if (err) {
		destroy_effects_chain(&rc->chain);
		free(rc);
		return NULL;
	}
	rc->out = stream;
	return rc;
}

int dspref_chain_out_fs(void *h) { return ((struct ref_chain *) h)->out.fs; }
int dspref_chain_out_channels(void *h) { return ((struct ref_chain *) h)->out.channels; }
int dspref_chain_n_effects(void *h)
{
	int n = 0;
	for (struct effect *e = ((struct ref_chain *) h)->chain.head; e; e = e->next) ++n;
	return n;
}
const char * dspref_chain_effect_name(void *h, int i)
{
	struct effect *e = ((struct ref_chain *) h)->chain.head;
	while (e && i-- > 0) e = e->next;
	return (e) ? e->name : NULL;
}
long dspref_chain_max_out_frames(void *h, long in_frames)
{
	return get_effects_chain_max_out_frames(&((struct ref_chain *) h)->chain, in_frames);
}
long dspref_chain_buffer_len(void *h, long in_frames)
{
	struct ref_chain *rc = h;
	return get_effects_chain_buffer_len(&rc->chain, in_frames, rc->in.channels);
}
double dspref_chain_delay(void *h) { return get_effects_chain_delay(&((struct ref_chain *) h)->chain, 0); }
long dspref_chain_drain_frames(void *h) { return ((struct ref_chain *) h)->chain.drain_frames; }

/* One block through run_effects_chain(); `out` must hold max_out_frames(frames)*out_channels. */
long dspref_chain_run(void *h, long frames, const double *in, double *out)
{
	struct ref_chain *rc = h;
	if (frames < 1) return 0;
	if (ensure_bufs(rc, frames)) return -2;
	memcpy(rc->buf1, in, (size_t) frames * rc->in.channels * sizeof(sample_t));
	ssize_t f = frames;
	sample_t *r = run_effects_chain(&rc->chain, &f, rc->buf1, rc->buf2);
	if (f > 0) memcpy(out, r, (size_t) f * rc->out.channels * sizeof(sample_t));
	return f;
}

/* Same, but without the copies: timing loops call this on the chain's own buffers. */
long dspref_chain_run_inplace(void *h, long frames, int refill)
{
	struct ref_chain *rc = h;
	if (ensure_bufs(rc, frames)) return -2;
	if (refill) {
		/* deterministic non-trivial content, cheap to produce */
		uint32_t s = 12345u;
		const ssize_t n = (ssize_t) frames * rc->in.channels;
		for (ssize_t i = 0; i < n; ++i)
			rc->buf1[i] = (double) pm_rand1_r(&s) / PM_RAND_MAX - 0.5;
	}
	ssize_t f = frames;
	run_effects_chain(&rc->chain, &f, rc->buf1, rc->buf2);
	return f;
}

/* drain_effects_chain(); returns -1 when dry. `out` must hold max_out_frames(frames)*out_channels. */
long dspref_chain_drain(void *h, long frames, double *out)
{
	struct ref_chain *rc = h;
	if (ensure_bufs(rc, frames)) return -2;
	ssize_t f = frames;
	sample_t *r = drain_effects_chain(&rc->chain, &f, rc->buf1, rc->buf2);
	if (f > 0) memcpy(out, r, (size_t) f * rc->out.channels * sizeof(sample_t));
	return f;
}

void dspref_chain_reset(void *h) { reset_effects_chain(&((struct ref_chain *) h)->chain); }

void dspref_chain_free(void *h)
{
	struct ref_chain *rc = h;
	if (!rc) return;
	destroy_effects_chain(&rc->chain);
	free(rc->buf1);
	free(rc->buf2);
	free(rc);
}

/* The reference's signal generator: `spec` is an sgen path such as
 * "sine:freq=20-20k+10s" (sgen.c:85-180); writes frames*channels doubles. */
long dspref_sgen(const char *spec, int fs, int channels, long frames, double *out)
{
	struct codec_params p = CODEC_PARAMS_AUTO(spec, CODEC_MODE_READ);
	p.type = "sgen";
	p.fs = fs;
	p.channels = channels;
	struct codec *c = init_codec(&p);
	if (!c) return -1;
	long done = 0;
	while (done < frames) {
		const ssize_t r = c->read(c, out + (size_t) done * channels, frames - done);
		if (r <= 0) break;
		done += r;
	}
	destroy_codec(c);
	return done;
}

/* Direct access to the reference's coefficient design, biquad.c:111-294. */
#include "biquad.h"
void dspref_biquad_design(int type, double fs, double arg0, double arg1, double arg2, double arg3, int width_type, double *c5)
{
	struct biquad_state b;
	biquad_init_using_type(&b, type, fs, arg0, arg1, arg2, arg3, width_type);
	c5[0] = b.c0; c5[1] = b.c1; c5[2] = b.c2; c5[3] = b.c3; c5[4] = b.c4;
}

long dspref_next_fast_fftw_len(long n) { return next_fast_fftw_len(n); }
