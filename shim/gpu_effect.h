/*
 * shim/gpu_effect.h -- the reference's `struct effect` surface (effect.h:24-59) on top of
 * libdspb200.so (include/dsp_b200.h).
 *
 * Every effect this shim creates is ONE device chain plus a host-side list of "parts" that
 * remember what the effect is made of (for plot / drain_samples / channel_offsets / merge),
 * exactly the hooks the reference's own effects set.
 * Compiled against the reference's unchanged headers (-I<reference>).
 */
#ifndef DSPB200_SHIM_GPU_EFFECT_H
#define DSPB200_SHIM_GPU_EFFECT_H

#include "dsp.h"
#include "effect.h"
#include "biquad.h"
#include "dsp_b200.h"

enum gpu_part_kind { GPU_PART_BIQUAD = 1, GPU_PART_FIR, GPU_PART_RESAMPLE, GPU_PART_GAIN };

struct gpu_part {
	struct gpu_part *next;
	int kind;
	char *selector;                 /* [channels], owned */
	/* GPU_PART_BIQUAD */
	struct biquad_state *bq;        /* [channels] coefficients (selected channels only), owned */
	/* GPU_PART_GAIN */
	sample_t *gain;                 /* [channels] multiplier (or addend when is_add), owned */
	int is_add;
	/* GPU_PART_FIR */
	sample_t *taps;                 /* [frames][fc], owned (kept for plot) */
	int fc;
	ssize_t frames, ref, latency;
};

struct gpu_effect_state {
	dspb200_chain *chain;
	struct gpu_part *parts;
	int failed;                     /* a run() error was logged; audio passes through */
};

/* A chain sharded as DSP_B200_DEVICES / DSP_B200_SLABS ask (default: device 0, 1 slab). */
dspb200_chain * gpu_chain_new(const char *name, const struct stream_info *istream);
struct gpu_part * gpu_part_new(int kind, const char *channel_selector, int channels);
void gpu_part_free(struct gpu_part *);
/* Wrap chain + part into a calloc'd struct effect with all same-rate hooks set. */
struct effect * gpu_effect_new(const struct effect_info *ei, const struct stream_info *istream, dspb200_chain *chain, struct gpu_part *part);

#endif
