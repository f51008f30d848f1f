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

/* First member of the `data` of EVERY GPU-backed effect (same-rate effects, resample, align, biquad -r). */
struct gpu_effect_state {
	dspb200_chain *chain;
	struct gpu_part *parts;
	int failed;                     /* a run() error was logged */
	/* device hand-off between neighbouring GPU effects (gpu_link_neighbours): the first effect of a run of GPU
	 * effects absorbs the device chains of those behind it; they stay in the host list as passengers */
	struct gpu_effect_state *head;  /* non-NULL: this effect's operators run inside head's device chain */
	int link_checked, inplace;
	ssize_t *align_len;             /* align: per-channel delay (drain_samples) */
};

/* A chain sharded as DSP_B200_DEVICES / DSP_B200_SLABS ask (default: device 0, 1 slab). */
dspb200_chain * gpu_chain_new(const char *name, const struct stream_info *istream);
struct gpu_part * gpu_part_new(int kind, const char *channel_selector, int channels);
void gpu_part_free(struct gpu_part *);
/* Runs of consecutive GPU effects that the optimizer could not merge (a rate change, a latency-bearing fir, align,
 * biquad -r) are still ONE device chain per block: at its first run() the first of them absorbs the operators of
 * the following GPU effects (dspb200_chain_absorb) and turns those into passengers whose run() hands the buffer on
 * untouched -- one H2D and one D2H per block for the whole run. */
void gpu_register_effect_kind(void (*destroy)(struct effect *));
int gpu_effect_is(const struct effect *e);
void gpu_link_neighbours(struct effect *e);
sample_t * gpu_linked_run(struct effect *e, ssize_t *frames, sample_t *ibuf, sample_t *obuf);
/* Wrap chain + part into a calloc'd struct effect with all same-rate hooks set. */
struct effect * gpu_effect_new(const struct effect_info *ei, const struct stream_info *istream, dspb200_chain *chain, struct gpu_part *part);

#endif
