// delay.cu -- per-channel integer delay lines on the device: the chain's `align` effect and `delay`.
//
// Reference behaviour reproduced (/root/reference):
//   align_effect_run   align.c:35-64    every channel k with a buffer of len_k frames is delayed by len_k (ring swap per
//                                       sample, :35-44); in the CLI build the first `discard_frames` frames of the
//                                       STREAM are dropped: a call returns the last max(frames_so_far, 0) frames of its
//                                       block (:53-62; not under SYMMETRIC_IO)
//   delay_effect_run   delay.c:47-63    the same ring swap with one length for the selected channels (whole samples)
// On the device a block is delayed out of place: out[i][k] = ring_k[(pos + i) % len_k] for i < len_k, else
// in[i - len_k][k]; afterwards the ring keeps the last len_k input frames (slot of absolute frame a: a % len_k, the
// same slot arithmetic as the reference's running index p).  Channels with len_k = 0 pass through.  One gather kernel
// + one ring-update kernel per block; 16 algorithmic bytes per sample.
#include "common.cuh"
#include "ops.h"

namespace dspb200 {

// out[(i - skip) C + k], i in [skip, frames)
__global__ void __launch_bounds__(256) k_align_read(const double *__restrict__ in, double *__restrict__ out, const double *__restrict__ ring,
                                                    const long *__restrict__ len, const long *__restrict__ off, int C, long frames, long skip, long abs0)
{
	const long idx = (long) blockIdx.x * blockDim.x + threadIdx.x;
	if (idx >= (frames - skip) * C) return;
	const long i = skip + idx / C;
	const int k = (int) (idx % C);
	const long L = len[k];
	double v;
	if (L == 0) v = in[i * C + k];
	else if (i < L) v = ring[off[k] + (abs0 + i) % L];
	else v = in[(i - L) * C + k];
	out[(i - skip) * C + k] = v;
}

// ring_k[(abs0 + i) % len_k] = in[i][k] for the last min(frames, len_k) frames of the block
__global__ void __launch_bounds__(256) k_align_write(const double *__restrict__ in, double *__restrict__ ring, const long *__restrict__ len,
                                                     const long *__restrict__ off, int C, long frames, long max_len, long abs0)
{
	const long idx = (long) blockIdx.x * blockDim.x + threadIdx.x;
	const long span = (frames < max_len) ? frames : max_len;   // frames (counted from the end of the block) any ring may need
	if (idx >= span * C) return;
	const long i = frames - span + idx / C;
	const int k = (int) (idx % C);
	const long L = len[k];
	if (L == 0 || i < frames - L) return;
	ring[off[k] + (abs0 + i) % L] = in[i * C + k];
}

struct AlignOp : Op {
	std::vector<long> h_len;
	long *d_len = nullptr, *d_off = nullptr;
	double *d_ring = nullptr;
	long ring_total = 0, max_len = 0;
	long discard = 0;        // frames dropped at the head of the stream (align.c:53-62)
	long abs_pos = 0;        // frames seen so far
	long stream_frames = 0;  // the reference's state->frames: starts at -discard

	const char *name() const override { return "align"; }
	std::string describe() const override
	{
		char buf[128];
		snprintf(buf, sizeof(buf), "{\"op\":\"align\",\"max_delay\":%ld,\"discard\":%ld}", max_len, discard);
		return buf;
	}
	~AlignOp() override { dev_free(d_len); dev_free(d_off); dev_free(d_ring); }

	void reset(cudaStream_t st) override
	{
		abs_pos = 0;
		stream_frames = -discard;
		if (ring_total > 0) cudaMemsetAsync(d_ring, 0, (size_t) ring_total * sizeof(double), st);
	}

	long run(long frames, const double *in, double *out, cudaStream_t st) override
	{
		if (frames <= 0) return 0;
		long out_frames = frames;
		if (stream_frames < 0) {
			stream_frames += frames;
			out_frames = (stream_frames > 0) ? stream_frames : 0;
		}
		else stream_frames += frames;
		const long skip = frames - out_frames;
		const int C = channels;
		ProfScope prof("align", st);
		if (out_frames > 0)
			LAUNCH(k_align_read, ceil_div(out_frames * C, 256), 256, 0, st, in, out, d_ring, d_len, d_off, C, frames, skip, abs_pos);
		if (max_len > 0) {
			const long span = (frames < max_len) ? frames : max_len;
			LAUNCH(k_align_write, ceil_div(span * C, 256), 256, 0, st, in, d_ring, d_len, d_off, C, frames, max_len, abs_pos);
		}
		abs_pos += frames;
		return out_frames;
	}
};

Op *make_align_op(int slab_channels, int fs, const long *delay, long discard_frames)
{
	std::unique_ptr<AlignOp> op(new AlignOp());
	const int C = slab_channels;
	op->channels = C;
	op->fs_in = op->fs_out = fs;
	op->inplace_ok = false;   // a frame's value comes from an earlier row of the same buffer
	op->discard = (discard_frames > 0) ? discard_frames : 0;
	op->stream_frames = -op->discard;
	op->h_len.assign(delay, delay + C);
	std::vector<long> off(C, 0);
	for (int k = 0; k < C; ++k) {
		if (op->h_len[k] < 0) { set_error("align: negative delay on channel %d", k); return nullptr; }
		off[k] = op->ring_total;
		op->ring_total += op->h_len[k];
		if (op->h_len[k] > op->max_len) op->max_len = op->h_len[k];
	}
	op->d_len = dev_alloc<long>(C, false);
	op->d_off = dev_alloc<long>(C, false);
	op->d_ring = dev_alloc<double>(op->ring_total > 0 ? op->ring_total : 1, true);
	if (!op->d_len || !op->d_off || !op->d_ring) return nullptr;
	CUDA_TRY(cudaMemcpy(op->d_len, op->h_len.data(), C * sizeof(long), cudaMemcpyHostToDevice), return nullptr);
	CUDA_TRY(cudaMemcpy(op->d_off, off.data(), C * sizeof(long), cudaMemcpyHostToDevice), return nullptr);
	return op.release();
}

}  // namespace dspb200
