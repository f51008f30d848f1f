// ops.h -- operator factories (one instance per shard); each lives in its own .cu file.
#pragma once
#include "common.cuh"

namespace dspb200 {

// gain.c:25-33
Op *make_gain_op(int slab_channels, int fs, const double *mult, const double *add);
// biquad.c:296-315; coefs[stage][slab_channels][5]
Op *make_biquad_op(int slab_channels, int fs, int n_stages, const double *coefs);
// returns a new operator equivalent to a followed by b, or nullptr if they cannot be fused
Op *fuse_biquad_ops(Op *a, Op *b);
// fir.c / fir_p.c; taps[filter_frames][filter_channels]; taps_cols[k] = filter column of the
// k-th selected channel of this slab (ignored when filter_channels == 1)
Op *make_fir_op(int slab_channels, int fs, const char *slab_selector, const double *taps, int filter_channels,
                long filter_frames, const int *taps_cols, long latency, long block_hint, cudaStream_t st);
// align.c:35-64 / delay.c:47-63: per-channel whole-sample delays (+ frames dropped at the head of the stream)
Op *make_align_op(int slab_channels, int fs, const long *delay, long discard_frames);
// resample.c
Op *make_resample_op(int slab_channels, int fs_in, int fs_out, double bandwidth, cudaStream_t st);

struct ResampleParams {
	int n, d, m, in_len, out_len, sinc_len, out_delay, sinc_os, m_os;
	double fc_os;
};
int resample_params(int fs_in, int fs_out, double bw, ResampleParams *p);

void fir_debug_serialize(int on);
int test_rfft(int B, int n_ch, const double *d_in, double *d_spec, cudaStream_t st);
int test_irfft(int B, int n_ch, const double *d_spec, double *d_out2B, cudaStream_t st);

}  // namespace dspb200
