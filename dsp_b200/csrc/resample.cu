// resample.cu -- K3: rational resampler as a polyphase FIR.
//
// Reference behaviour reproduced (/root/reference/resample.c): parameters :274-316, Albrecht
// 9-term window :52-80, norm_sinc :82-87, windowed sinc :361-364 and its spectrum :366, the
// per-block spectral image/fold multiply + overlap-add :89-152, output pacing :93-108,144-147,
// reset :154-161, drain2 :163-188.
//
// The reference works in the frequency domain on blocks of in_len frames.  Because each block
// is zero-padded to 2*in_len and the filter is shorter than in_len, that algorithm is exactly
// the linear periodically-time-varying FIR (checked against the compiled
// reference to 3e-15 for 160/147, 147/160, 2/1, 1/2, 320/441 before this file was written)
//     y[m] = sum_{t=0}^{in_len-1} G[(m d) mod n][t] * x[floor(m d / n) - t]
//     G[ph][t] = g(ph + t n),
//     g(tau) = 1/(2 in_len) [ S_0 + 2 sum_{k=1}^{K-1} Re(S_k e^{i th k tau}) + Re(S_K e^{i th K tau}) ],
//     th = 2 pi / (2 in_len n),  K = sinc_len,  S = DFT_{2K}(windowed sinc)
// and the emitted stream is y[out_delay + q].  The n x in_len tap table G (0.75 MB for
// 44100->48000) is built on the device at init from the reference's own sinc samples; the hot
// kernel is a per-channel dot product with lanes along channels, so every load of the
// interleaved input is a coalesced row and the tap is a warp-uniform (broadcast) load.
// The reference's per-call frame pacing is replayed by an integer state machine on the host,
// so each call emits exactly the frame count resample_effect_run() would.
#include "common.cuh"
#include "ops.h"
#include <cmath>

namespace dspb200 {

static long next_smooth7(long n)
{
	for (;; ++n) {
		long m = n;
		for (int p : { 2, 3, 5, 7 })
			while (m % p == 0) m /= p;
		if (m == 1) return n;
	}
}

static int gcd_int(int a, int b)
{
	while (b) { const int c = a % b; a = b; b = c; }
	return a;
}

#define ALBRECHT_M_FACT 17.7822

// resample.c:274-316
int resample_params(int fs_in, int fs_out, double bw, ResampleParams *p)
{
	if (fs_in <= 0 || fs_out <= 0 || fs_in == fs_out) { set_error("resample: bad rates %d -> %d", fs_in, fs_out); return -1; }
	if (bw == 0.0) bw = 0.939;
	if (!(bw >= 0.7 && bw <= 0.999)) { set_error("resample: bandwidth out of range"); return -1; }
	const int max_rate = (fs_out > fs_in) ? fs_out : fs_in, min_rate = (fs_out > fs_in) ? fs_in : fs_out;
	const int g = gcd_int(fs_out, fs_in);
	p->n = fs_out / g;
	p->d = fs_in / g;
	const int max_factor = (p->n > p->d) ? p->n : p->d, min_factor = (p->n > p->d) ? p->d : p->n;
	p->m = (int) lround(2.0 * ALBRECHT_M_FACT * max_rate / (min_rate * (1.0 - bw)));
	const double width = ALBRECHT_M_FACT * max_rate / p->m;
	const double fc = (min_rate - width) / max_rate;
	p->sinc_os = (min_factor < 2) ? min_factor : 2;
	p->fc_os = fc / p->sinc_os;
	p->m_os = (p->m + 1) * p->sinc_os - 1;
	int len_mult = (p->m + 1) / max_factor;
	if ((p->m + 1) % max_factor != 0) len_mult += 1;
	if (len_mult > 16) {
		const int fast = (int) next_smooth7(len_mult);
		if (fast != len_mult && (p->n <= 16 || p->d <= 16 || next_smooth7(p->n) == p->n || next_smooth7(p->d) == p->d))
			len_mult = fast;
	}
	p->sinc_len = max_factor * len_mult * p->sinc_os;
	p->in_len = p->d * len_mult;
	p->out_len = p->n * len_mult;
	if (fs_out == max_rate) p->out_delay = p->m / 2;
	else p->out_delay = (int) lround(p->m / 2 * ((double) p->n / p->d));
	return 0;
}

// resample.c:52-80 (WINDOW_FUNCTION 3)
static double albrecht9(double x)
{
	static const double a[9] = {
		2.318028013590306028393e-1, 3.932575471789488615081e-1, 2.385434764970747429454e-1,
		1.014370437785239811268e-1, 2.911516061918003918645e-2, 5.280988177252078698806e-3,
		5.382909093381945363528e-4, 2.442086527507867730168e-5, 2.706153764205043532817e-7,
	};
	if (x >= 1.0 || x <= 0.0) return 0.0;
	double w = a[0];
	for (int i = 1; i < 9; ++i) {
		const double c = (i & 1) ? -a[i] : a[i];
		w += c * cos(2 * i * M_PI * x);
	}
	return w;
}

// S[k] = sum_i sinc[i] exp(-2 pi i i k / (2K)), k = 0..K   (what fftw r2c gives at resample.c:366)
__global__ void k_rs_sinc_dft(const double *sinc, int m_os, int K, double2 *S)
{
	const int k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k > K) return;
	double re = 0.0, im = 0.0;
	for (int i = 1; i < m_os; ++i) {
		const long r = ((long) i * k) % (2L * K);
		double sn, cs;
		sincospi((double) r / K, &sn, &cs);
		const double v = sinc[i];
		re = fma(v, cs, re);
		im = fma(-v, sn, im);
	}
	S[k] = make_double2(re, im);
}

// G[ph][t] = g(ph + t n)
__global__ void k_rs_table(const double2 *S, int K, int n, int in_len, double *G, int g_stride, int pad)
{
	const long idx = (long) blockIdx.x * blockDim.x + threadIdx.x;
	if (idx >= (long) n * in_len) return;
	const int ph = (int) (idx / in_len), t = (int) (idx % in_len);
	const long tau = ph + (long) t * n;
	const long Pd = 2L * in_len * n;
	double acc = 0.0;
	for (int k = 1; k <= K; ++k) {
		const long r = ((long) k * tau) % Pd;
		double sn, cs;
		sincospi(2.0 * (double) r / (double) Pd, &sn, &cs);
		const double2 s = S[k];
		const double term = s.x * cs - s.y * sn;
		acc += (k == K) ? 0.5 * term : term;
	}
	G[(long) ph * g_stride + pad + t] = (S[0].x + 2.0 * acc) / (2.0 * in_len);   // rows are zero-padded by `pad` on both sides
}

// out[q][c] = sum_t G[(m d) % n][t] * x[(m d) / n - t][c],  m = m0 + q;  x lives in a ring of rows.
// Register-blocked: a warp produces R consecutive output frames for 32*CH channels (lane = CH adjacent
// channels).  It walks the input rows i downwards once; every row is loaded once (CH doubles per lane,
// coalesced) and feeds all R outputs, each with its own tap G[ph_r][i_hi_r - i] (warp-uniform load).
// Tap rows are zero-padded by `pad` on both sides, so the ragged ends of the R tap windows need no
// branches.  2 R CH flops per (1 + R) loads.
template <int R, int CH>
__global__ void __launch_bounds__(128) k_rs_poly(const double *__restrict__ ring, long ring_len, int C, const double *__restrict__ G,
                                                 int g_stride, int pad, int n, int d, int in_len, long m0, long n_out, double *__restrict__ out)
{
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int c0 = (blockIdx.y * 32 + lane) * CH;
	const long q0 = ((long) blockIdx.x * 4 + warp) * R;
	if (q0 >= n_out) return;
	long ih[R];
	const double *g[R];
#pragma unroll
	for (int r = 0; r < R; ++r) {
		long q = q0 + r;
		if (q >= n_out) q = n_out - 1;   // duplicates the last frame's work; not stored
		const long md = (m0 + q) * d;
		ih[r] = md / n;
		g[r] = G + (md - ih[r] * n) * g_stride + pad;
	}
	const long i_top = ih[R - 1];
	long i_bot = ih[0] - in_len + 1;
	if (i_bot < 0) i_bot = 0;   // rows before the stream start are zero
#pragma unroll
	for (int r = 0; r < R; ++r) g[r] += ih[r] - i_top;   // tap index of row i_top (<= 0, inside the left pad)
	double acc[R][CH];
#pragma unroll
	for (int r = 0; r < R; ++r)
#pragma unroll
		for (int h = 0; h < CH; ++h) acc[r][h] = 0.0;
	const bool act = c0 < C;
	long row = i_top % ring_len;
	const double *col = ring + (act ? c0 : 0);
	for (long i = i_top; i >= i_bot; --i) {
		double x[CH];
		const double *p = col + row * C;
		if (CH == 4) {
			const double2 a = *reinterpret_cast<const double2 *>(p), b = *reinterpret_cast<const double2 *>(p + 2);
			x[0] = a.x; x[1] = a.y; x[2 % CH] = b.x; x[3 % CH] = b.y;
		}
		else if (CH == 2) {
			const double2 a = *reinterpret_cast<const double2 *>(p);
			x[0] = a.x; x[1 % CH] = a.y;
		}
		else x[0] = p[0];
#pragma unroll
		for (int r = 0; r < R; ++r) {
			const double w = __ldg(g[r]);
			++g[r];
#pragma unroll
			for (int h = 0; h < CH; ++h) acc[r][h] = fma(w, x[h], acc[r][h]);
		}
		row = (row == 0) ? ring_len - 1 : row - 1;
	}
	if (!act) return;
#pragma unroll
	for (int r = 0; r < R; ++r) {
		if (q0 + r >= n_out) break;
		double *o = out + (q0 + r) * C + c0;
		if (CH == 4) {
			*reinterpret_cast<double2 *>(o) = make_double2(acc[r][0], acc[r][1 % CH]);
			*reinterpret_cast<double2 *>(o + 2) = make_double2(acc[r][2 % CH], acc[r][3 % CH]);
		}
		else if (CH == 2) *reinterpret_cast<double2 *>(o) = make_double2(acc[r][0], acc[r][1 % CH]);
		else o[0] = acc[r][0];
	}
}

constexpr int RS_R = 8;   // output frames per warp

// move the live rows of the input ring into a bigger ring (absolute frame a lives at a % len)
__global__ void k_rs_ring_grow(const double *old_ring, long old_len, double *new_ring, long new_len, int C, long a0, long rows)
{
	const long idx = (long) blockIdx.x * blockDim.x + threadIdx.x;
	if (idx >= rows * C) return;
	const long a = a0 + idx / C;
	const int c = (int) (idx % C);
	new_ring[(a % new_len) * C + c] = old_ring[(a % old_len) * C + c];
}

struct ResampleOp : Op {
	ResampleParams p;
	double *d_G = nullptr, *d_ring = nullptr;
	int g_stride = 0, g_pad = 0;         // tap rows: [pad zeros | in_len taps | pad zeros]
	long ring_len = 0;
	// resample.c:39-50 bookkeeping (integers only; the sample data lives in the ring)
	long total_in = 0;          // frames appended since reset
	long emit_pos = 0;          // next raw output index to emit
	int in_buf_pos = 0, out_buf_pos = 0, has_output = 0;
	int is_draining = 0;
	long drain_pos = 0, drain_frames = 0;

	const char *name() const override { return "resample"; }
	~ResampleOp() override { dev_free(d_G); dev_free(d_ring); }

	long max_out_frames(long in_frames) const override
	{
		const long long r = (long long) in_frames * p.n;
		return (long) ((r % p.d != 0) ? r / p.d + 1 : r / p.d);
	}

	void reset(cudaStream_t st) override
	{
		// resample.c:154-161 (the drain counters are deliberately left alone, as there)
		in_buf_pos = out_buf_pos = 0;
		has_output = 0;
		total_in = 0;
		emit_pos = 0;
		if (d_ring) cudaMemsetAsync(d_ring, 0, (size_t) ring_len * channels * sizeof(double), st);
	}

	int ensure_ring(long frames, cudaStream_t st)
	{
		const long need = 3L * p.in_len + frames + 64;
		if (ring_len >= need) return 0;
		const long new_len = 3L * p.in_len + 2 * frames + 64;
		double *nr = dev_alloc<double>((size_t) new_len * channels, true);
		if (!nr) return -1;
		if (d_ring) {
			const long rows = (total_in < ring_len) ? total_in : ring_len;
			if (rows > 0)
				LAUNCH(k_rs_ring_grow, ceil_div(rows * channels, 256), 256, 0, st, d_ring, ring_len, nr, new_len, channels, total_in - rows, rows);
			CUDA_TRY(cudaStreamSynchronize(st), return -1);
			dev_free(d_ring);
		}
		d_ring = nr;
		ring_len = new_len;
		return 0;
	}

	long run(long frames, const double *in, double *out, cudaStream_t st) override
	{
		if (frames <= 0) return 0;
		const int C = channels;
		if (ensure_ring(frames, st)) return -1;
		// append the new rows (at most one wrap)
		{
			const long w = total_in % ring_len;
			const long first = (frames < ring_len - w) ? frames : ring_len - w;
			CUDA_TRY(cudaMemcpyAsync(d_ring + w * C, in, (size_t) first * C * sizeof(double), cudaMemcpyDeviceToDevice, st), return -1);
			if (first < frames)
				CUDA_TRY(cudaMemcpyAsync(d_ring, in + first * C, (size_t) (frames - first) * C * sizeof(double), cudaMemcpyDeviceToDevice, st), return -1);
			total_in += frames;
		}
		// replay resample_effect_run()'s pacing, resample.c:91-151
		const long max_oframes = max_out_frames(frames);
		long iframes = 0, oframes = 0;
		while (iframes < frames) {
			long take = p.in_len - in_buf_pos;
			if (take > frames - iframes) take = frames - iframes;
			in_buf_pos += (int) take;
			iframes += take;
			if (has_output) {
				long emit = p.out_len - out_buf_pos;
				if (emit > max_oframes - oframes) emit = max_oframes - oframes;
				out_buf_pos += (int) emit;
				oframes += emit;
			}
			if (in_buf_pos == p.in_len && (!has_output || out_buf_pos == p.out_len)) {
				in_buf_pos = out_buf_pos = 0;
				if (!has_output) {
					out_buf_pos = p.out_delay;
					emit_pos += p.out_delay;   // only ever happens before the first emitted frame
					has_output = 1;
				}
			}
			else if (take == 0) {
				set_error("resample: pacing stalled (in_buf_pos=%d out_buf_pos=%d)", in_buf_pos, out_buf_pos);
				return -1;
			}
		}
		const long first_m = emit_pos;   // emission is one contiguous raw range per call
		if (oframes > 0) {
			ProfScope prof("resample", st);
			// pick the register tile so that the grid still fills the chip: wide tiles (8 frames x 4 channels
			// per lane) have the best flop/load ratio, narrow ones more warps
			const long warps84 = (long) ceil_div(oframes, 8) * ceil_div(C, 128);
			const long warps82 = (long) ceil_div(oframes, 8) * ceil_div(C, 64);
			const long want = 148L * 24;
			if (C % 4 == 0 && warps84 >= want) {
				LAUNCH((k_rs_poly<8, 4>), dim3(ceil_div(oframes, 32), ceil_div(C, 128)), 128, 0, st, d_ring, ring_len, C, d_G, g_stride, g_pad, p.n, p.d, p.in_len, first_m, oframes, out);
			}
			else if (C % 2 == 0 && warps82 >= want) {
				LAUNCH((k_rs_poly<8, 2>), dim3(ceil_div(oframes, 32), ceil_div(C, 64)), 128, 0, st, d_ring, ring_len, C, d_G, g_stride, g_pad, p.n, p.d, p.in_len, first_m, oframes, out);
			}
			else if (C % 2 == 0) {
				LAUNCH((k_rs_poly<4, 2>), dim3(ceil_div(oframes, 16), ceil_div(C, 64)), 128, 0, st, d_ring, ring_len, C, d_G, g_stride, g_pad, p.n, p.d, p.in_len, first_m, oframes, out);
			}
			else {
				LAUNCH((k_rs_poly<4, 1>), dim3(ceil_div(oframes, 16), ceil_div(C, 32)), 128, 0, st, d_ring, ring_len, C, d_G, g_stride, g_pad, p.n, p.d, p.in_len, first_m, oframes, out);
			}
			emit_pos += oframes;
		}
		return oframes;
	}

	// resample.c:163-188
	long drain2(long frames, double *zeros, double *out, cudaStream_t st) override
	{
		if (!has_output && in_buf_pos == 0) return -1;
		if (!is_draining) {
			if (has_output) {
				drain_frames += p.out_delay;
				drain_frames += p.out_len - out_buf_pos;
			}
			drain_frames += max_out_frames(in_buf_pos);
			is_draining = 1;
		}
		if (drain_pos >= drain_frames) return -1;
		CUDA_TRY(cudaMemsetAsync(zeros, 0, (size_t) frames * channels * sizeof(double), st), return -2);
		long produced = run(frames, zeros, out, st);
		if (produced < 0) return -2;
		drain_pos += produced;
		if (drain_pos > drain_frames) produced -= drain_pos - drain_frames;
		return produced;
	}
};

Op *make_resample_op(int slab_channels, int fs_in, int fs_out, double bandwidth, cudaStream_t st)
{
	std::unique_ptr<ResampleOp> op(new ResampleOp());
	op->channels = slab_channels;
	op->fs_in = fs_in;
	op->fs_out = fs_out;
	op->inplace_ok = false;
	if (resample_params(fs_in, fs_out, bandwidth, &op->p)) return nullptr;
	const ResampleParams &p = op->p;

	// windowed sinc, resample.c:361-364 (host libm, like the reference)
	std::vector<double> sinc((size_t) p.m_os + 1, 0.0);
	for (int i = 1; i < p.m_os; ++i) {
		const double x = (i * 2 - p.m_os) / 2.0;
		const double s = (fabs(x) < 1e-9) ? p.fc_os : sin(M_PI * p.fc_os * x) / (M_PI * x);
		sinc[i] = s * albrecht9((double) i / p.m_os);
	}
	double *d_sinc = dev_alloc<double>(sinc.size(), false);
	double2 *d_S = dev_alloc<double2>((size_t) p.sinc_len + 1, false);
	// R consecutive outputs span at most (R-1) d/n + 1 input rows more than one output does
	op->g_pad = (int) (((long) (RS_R - 1) * p.d) / p.n + 2);
	op->g_stride = p.in_len + 2 * op->g_pad;
	op->d_G = dev_alloc<double>((size_t) p.n * op->g_stride, true);
	if (!d_sinc || !d_S || !op->d_G) return nullptr;
	CUDA_TRY(cudaMemcpyAsync(d_sinc, sinc.data(), sinc.size() * sizeof(double), cudaMemcpyHostToDevice, st), return nullptr);
	LAUNCH(k_rs_sinc_dft, ceil_div(p.sinc_len + 1, 128), 128, 0, st, d_sinc, p.m_os, p.sinc_len, d_S);
	LAUNCH(k_rs_table, ceil_div((long) p.n * p.in_len, 128), 128, 0, st, d_S, p.sinc_len, p.n, p.in_len, op->d_G, op->g_stride, op->g_pad);
	CUDA_TRY(cudaStreamSynchronize(st), return nullptr);
	dev_free(d_sinc);
	dev_free(d_S);
	return op.release();
}

}  // namespace dspb200
