/*
 * shim/frontend.c -- a library frontend over the reference's chain runtime + the GPU effects.
 *
 * Plays the role dsp.c:1418-1431 and ladspa_dsp.c:316-355 play: owns `dsp_globals` and the log lock, builds a
 * `struct effects_chain` from a chain string with the reference's own parser/optimizer
 * (build_effects_chain_from_string, effects_chain.h:42 -- reference code, linked unmodified), allocates the two
 * block buffers exactly as the frontends do (plain calloc memory of get_effects_chain_buffer_len() samples,
 * dsp.c:1067-1081) and calls run_effects_chain() block by block.  Linked with the shim's replacement objects it is
 * the drop-in as a shared library (shim/_build/libdsp_b200_frontend.so): what bench.py times as `e2e_dropin`
 * (clock_gettime around run_effects_chain(), SURVEY.md 8d mode A) and what an embedding application would call.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <pthread.h>
#include "dsp.h"
#include "effect.h"
#include "effects_chain.h"
#include "util.h"

struct dsp_globals dsp_globals = { LL_ERROR, "dsp_b200" };
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

struct front_chain {
	struct effects_chain chain;
	struct stream_info in, out;
	sample_t *buf1, *buf2;
	ssize_t buf_frames;
};

void dspfront_set_loglevel(int l) { dsp_globals.loglevel = l; }

static int front_bufs(struct front_chain *fc, ssize_t frames)
{
	if (frames <= fc->buf_frames) return 0;
	const ssize_t len = get_effects_chain_buffer_len(&fc->chain, frames, fc->in.channels);
	free(fc->buf1);
	free(fc->buf2);
	fc->buf1 = calloc(len, sizeof(sample_t));
	fc->buf2 = calloc(len, sizeof(sample_t));
	if (!fc->buf1 || !fc->buf2) return 1;
	fc->buf_frames = frames;
	return 0;
}

void * dspfront_chain_new(const char *chain_str, int fs, int channels, const char *dir)
{
	struct front_chain *fc = calloc(1, sizeof(*fc));
	if (!fc) return NULL;
	fc->in.fs = fs;
	fc->in.channels = channels;
	struct stream_info stream = fc->in;
	char *mask = NEW_SELECTOR(channels);
	if (!mask) {
		free(fc);
		return NULL;
	}
	SET_SELECTOR(mask, channels);
	const int err = build_effects_chain_from_string(chain_str, dir, &fc->chain, &stream, mask, (dir) ? dir : ".");
	free(mask);
	if (err) {
		destroy_effects_chain(&fc->chain);
		free(fc);
		return NULL;
	}
	fc->out = stream;
	return fc;
}

void dspfront_chain_free(void *h)
{
	struct front_chain *fc = h;
	if (!fc) return;
	destroy_effects_chain(&fc->chain);
	free(fc->buf1);
	free(fc->buf2);
	free(fc);
}

int dspfront_chain_out_fs(void *h) { return ((struct front_chain *) h)->out.fs; }
int dspfront_chain_out_channels(void *h) { return ((struct front_chain *) h)->out.channels; }

int dspfront_chain_n_effects(void *h)
{
	int n = 0;
	for (struct effect *e = ((struct front_chain *) h)->chain.head; e; e = e->next) ++n;
	return n;
}

const char * dspfront_chain_effect_name(void *h, int i)
{
	struct effect *e = ((struct front_chain *) h)->chain.head;
	while (e && i-- > 0) e = e->next;
	return (e) ? e->name : NULL;
}

long dspfront_chain_max_out_frames(void *h, long in_frames)
{
	return get_effects_chain_max_out_frames(&((struct front_chain *) h)->chain, in_frames);
}

/* One block: in -> the chain's own buffer -> run_effects_chain() -> out (max_out_frames(frames) * out channels). */
long dspfront_chain_run(void *h, long frames, const double *in, double *out)
{
	struct front_chain *fc = h;
	if (frames < 1) return 0;
	if (front_bufs(fc, frames)) return -2;
	memcpy(fc->buf1, in, (size_t) frames * fc->in.channels * sizeof(sample_t));
	ssize_t f = frames;
	sample_t *r = run_effects_chain(&fc->chain, &f, fc->buf1, fc->buf2);
	if (f > 0) memcpy(out, r, (size_t) f * fc->out.channels * sizeof(sample_t));
	return f;
}

static double now_s(void)
{
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return (double) ts.tv_sec + 1e-9 * (double) ts.tv_nsec;
}

/*
 * Timing loop of mode A: `warm` untimed + `blocks` timed calls of run_effects_chain() on `frames`-frame blocks taken
 * round-robin from pool[n_pool][frames*channels]; the block is copied into the chain's buffer (what a codec read
 * does) OUTSIDE the timed spans, clock_gettime brackets each run_effects_chain() call.  Returns the frames of the
 * last call; *seconds = summed call time; checksum = sum |sample| of the last result.
 */
long dspfront_chain_time(void *h, long frames, int warm, int blocks, const double *pool, int n_pool, double *seconds, double *checksum)
{
	struct front_chain *fc = h;
	if (frames < 1 || n_pool < 1 || front_bufs(fc, frames)) return -2;
	const size_t n = (size_t) frames * fc->in.channels;
	double total = 0.0;
	ssize_t f = 0;
	sample_t *r = NULL;
	for (int i = 0; i < warm + blocks; ++i) {
		memcpy(fc->buf1, pool + (size_t) (i % n_pool) * n, n * sizeof(sample_t));
		f = frames;
		const double t0 = now_s();
		r = run_effects_chain(&fc->chain, &f, fc->buf1, fc->buf2);
		const double t1 = now_s();
		if (i >= warm) total += t1 - t0;
	}
	if (seconds) *seconds = total;
	if (checksum) {
		double s = 0.0;
		for (size_t i = 0; r && i < (size_t) f * fc->out.channels; ++i) s += (r[i] < 0.0) ? -r[i] : r[i];
		*checksum = s;
	}
	return f;
}
