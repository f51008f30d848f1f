/*
 * shim/fir_p.c -- drop-in replacement object for the reference's fir_p.o.
 *
 * Exports fir_p_effect_init and fir_p_effect_init_with_filter (fir_p.h:27-28; the latter is also
 * called by hilbert.c:80,86).  fir_p is zero-latency partitioned convolution: its output is exactly
 * x*h (fir_p.c:127-181), which is what K2 computes with uniform device-side partitions.  The
 * reference's `max_part_len` argument (fir_p.c:38,376-385,554-558) tunes its CPU partition plan; it is
 * parsed and validated identically but has no effect on the device plan.
 */
#include <stdlib.h>
#include <limits.h>
#include "fir_p.h"
#include "fir.h"
#include "util.h"
#include "gpu_effect.h"

#define DIRECT_TAPS 32   /* fir_p.c:34,364: at most this many taps -> the direct-form fir */

struct effect * gpu_fir_effect_new(const struct effect_info *, const struct stream_info *, const char *, const sample_t *, int, ssize_t, ssize_t, ssize_t);

struct effect * fir_p_effect_init_with_filter(const struct effect_info *ei, const struct stream_info *istream, const char *channel_selector, sample_t *filter_data, int filter_channels, ssize_t filter_frames, ssize_t ref, int max_part_len)
{
	if (filter_frames <= DIRECT_TAPS)
		return fir_effect_init_with_filter(ei, istream, channel_selector, filter_data, filter_channels, filter_frames, ref, 1);
	if (max_part_len == 0) max_part_len = 1 << 14;
	if (!IS_POWER_OF_2(max_part_len)) {
		LOG_FMT(LL_ERROR, "%s: error: max_part_len must be a power of two", ei->name);
		return NULL;
	}
	if (max_part_len < DIRECT_TAPS) {
		LOG_FMT(LL_ERROR, "%s: error: max_part_len must be within [%d,%d] or 0 for default", ei->name, DIRECT_TAPS, INT_MAX);
		return NULL;
	}
	return gpu_fir_effect_new(ei, istream, channel_selector, filter_data, filter_channels, filter_frames, ref, 0);
}

struct effect * fir_p_effect_init(const struct effect_info *ei, const struct stream_info *istream, const char *channel_selector, const char *dir, int argc, const char *const *argv)
{
	int filter_channels;
	ssize_t filter_frames, max_part_len = 0;
	struct fir_config config;
	struct dsp_getopt_state g = DSP_GETOPT_STATE_INITIALIZER;
	char *endptr;

	if (fir_parse_opts(ei, istream, &config, &g, argc, argv, NULL, NULL, NULL) || g.ind < argc - 2 || g.ind > argc - 1) {
		print_effect_usage(ei);
		return NULL;
	}
	if (g.ind == argc - 2) {
		max_part_len = strtol(argv[g.ind], &endptr, 10);
		CHECK_ENDPTR(argv[g.ind], endptr, "max_part_len", return NULL);
		++g.ind;
	}
	config.p.path = argv[g.ind];
	sample_t *filter_data = fir_read_filter(ei, istream, channel_selector, dir, &config.p, &filter_channels, &filter_frames);
	if (!filter_data) return NULL;
	const ssize_t ref = fir_get_offset(&config, filter_data, filter_channels, filter_frames);
	struct effect *e = fir_p_effect_init_with_filter(ei, istream, channel_selector, filter_data, filter_channels, filter_frames, ref, (int) max_part_len);
	free(filter_data);
	return e;
}
