/*
 * shim/fir.c -- drop-in replacement object for the reference's fir.o.
 *
 * Exports fir_effect_init and fir_effect_init_with_filter (fir.h:27-28; the latter is also called by
 * hilbert.c:89, fir_p.c:365 and matrix4_mb.c:776).  Behaviour kept from fir.c:219-388:
 *   - filter_channels must be 1 or the number of selected channels (:221-225); filter_frames >= 1
 *   - <= 16 taps or force_direct: zero-latency direct form (:240-283)
 *   - otherwise: convolution with a reported latency of len = next_fast_fftw_len(taps) frames
 *     (:296-298, channel_offsets :208-217, drain_samples :180-187) -- the engine itself is
 *     zero-latency (K2), the latency is reproduced with a delay ring so the chain's align pass and
 *     the plot/drain bookkeeping see exactly what they saw with the reference object
 *   - filter data and argv are borrowed (:385-386)
 */
#include <stdlib.h>
#include <string.h>
#include "fir.h"
#include "util.h"
#include "gpu_effect.h"

#define MAX_DIRECT_TAPS 16   /* fir.c:29 */

struct effect * gpu_fir_effect_new(const struct effect_info *ei, const struct stream_info *istream, const char *channel_selector,
	const sample_t *filter_data, int filter_channels, ssize_t filter_frames, ssize_t ref, ssize_t latency)
{
	const int n_channels = num_bits_set(channel_selector, istream->channels);
	if (filter_channels != 1 && filter_channels != n_channels) {
		LOG_FMT(LL_ERROR, "%s: error: channels mismatch: channels=%d filter_channels=%d", ei->name, n_channels, filter_channels);
		return NULL;
	}
	if (filter_frames < 1) {
		LOG_FMT(LL_ERROR, "%s: error: filter length must be >= 1", ei->name);
		return NULL;
	}
	LOG_FMT(LL_VERBOSE, "%s: info: filter_frames=%zd latency=%zd (B200 partitioned convolution)", ei->name, filter_frames, latency);
	dspb200_chain *chain = gpu_chain_new(ei->name, istream);
	struct gpu_part *part = gpu_part_new(GPU_PART_FIR, channel_selector, istream->channels);
	if (!chain || !part) goto fail;
	part->fc = filter_channels;
	part->frames = filter_frames;
	part->ref = ref;
	part->latency = latency;
	part->taps = malloc((size_t) filter_frames * filter_channels * sizeof(sample_t));
	if (!part->taps) {
		dsp_perror(DSP_ENOMEM, ei->name, NULL);
		goto fail;
	}
	memcpy(part->taps, filter_data, (size_t) filter_frames * filter_channels * sizeof(sample_t));
	if (dspb200_chain_add_fir(chain, channel_selector, filter_data, filter_channels, filter_frames, latency, 0) != 0) {
		LOG_FMT(LL_ERROR, "%s: error: %s", ei->name, dspb200_last_error());
		goto fail;
	}
	struct effect *e = gpu_effect_new(ei, istream, chain, part);
	if (e) return e;

	fail:
	gpu_part_free(part);
	dspb200_chain_destroy(chain);
	return NULL;
}

struct effect * fir_effect_init_with_filter(const struct effect_info *ei, const struct stream_info *istream, const char *channel_selector, sample_t *filter_data, int filter_channels, ssize_t filter_frames, ssize_t ref, int force_direct)
{
	const ssize_t latency = (filter_frames <= MAX_DIRECT_TAPS || force_direct) ? 0 : next_fast_fftw_len(filter_frames);
	return gpu_fir_effect_new(ei, istream, channel_selector, filter_data, filter_channels, filter_frames, ref, latency);
}

struct effect * fir_effect_init(const struct effect_info *ei, const struct stream_info *istream, const char *channel_selector, const char *dir, int argc, const char *const *argv)
{
	int filter_channels;
	ssize_t filter_frames;
	struct fir_config config;
	struct dsp_getopt_state g = DSP_GETOPT_STATE_INITIALIZER;

	if (fir_parse_opts(ei, istream, &config, &g, argc, argv, NULL, NULL, NULL) || g.ind != argc - 1) {
		print_effect_usage(ei);
		return NULL;
	}
	config.p.path = argv[g.ind];
	sample_t *filter_data = fir_read_filter(ei, istream, channel_selector, dir, &config.p, &filter_channels, &filter_frames);
	if (!filter_data) return NULL;
	const ssize_t ref = fir_get_offset(&config, filter_data, filter_channels, filter_frames);
	struct effect *e = fir_effect_init_with_filter(ei, istream, channel_selector, filter_data, filter_channels, filter_frames, ref, 0);
	free(filter_data);
	return e;
}
