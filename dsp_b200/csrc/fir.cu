// fir.cu -- K2: FFT convolution engine behind `fir`, `fir_p` and `hilbert`.
//
// Reference behaviour reproduced (file:line in /root/reference):
//   fir_p_effect_run  fir_p.c:127-181   out = in * h, zero latency, any frames per call
//   fir_effect_run    fir.c:109-149     out = (in * h) delayed by len frames
//   fir_direct_effect_run fir.c:43-62   out = in * h, zero latency (short filters)
// The reference's partition plan (32-tap direct head + <=4 FFT groups, fir_p.c:290-335) is a
// CPU latency device; its output is exactly the linear convolution, which is what is kept.
//
// B200 formulation: uniform partitions of B frames (B = power of two, 64..8192), overlap-add
// framing, frequency-domain delay line (FDL) per selected channel:
//   block j complete:  X_j = RFFT_2B([x_j | 0])                      -> FDL slot j mod P   (k_fir_fwd)
//                      S_j = sum_{p<P} X_{j-p} . H_p                 streams FDL + H      (k_fir_mac)
//                      s_j = IRFFT_2B(S_j); out_j = s_j[0:B) + carry; carry = s_j[B:2B)  (k_fir_inv)
// Spectra are stored "packed": B complex per row, bin 0 = (DC.re, Nyquist.re); rows are
// 16*B bytes, so every row is 128-byte aligned and a warp reads 512 contiguous bytes.
//
// Calls that are not whole aligned blocks take the general path, which is exact for ANY
// frames-per-call pattern: with R_j = sum_{1<=p<P} X_{j-p} . H_p (past blocks only)
//   out_j[m] = pre_j[m] + sum_{r<=m} x_j[r] h[m-r],   pre_j = IRFFT(R_j)[0:B) + carry_{j-1}
// i.e. a time-domain head over the samples of the still-incomplete block (k_fir_head) on top
// of a precomputed contribution of all completed blocks; when the block completes, X_j enters
// the FDL and carry_j is refreshed.  Both paths share {FDL, carry, block counter}.
#include "common.cuh"
#include "fft.cuh"
#include "ops.h"

namespace dspb200 {

// ------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------
struct FwdArgs {
	const double *in;       // time-domain source
	long stride;            // elements between consecutive frames
	const int *ch_map;      // selected-channel index -> element offset multiplier (NULL: s)
	long ch_mul;
	long valid;             // frames available (<= B); the rest is zero
	double2 *spec;          // destination rows
	long spec_ch_stride;    // double2 elements between channels
	int slot;               // row within the channel
	const double2 *tw;
	int s0, s1;             // selected-channel range handled by this launch
};

template <int N>
__global__ void __launch_bounds__(FftCfg<N>::THREADS) k_fir_fwd(FwdArgs a)
{
	extern __shared__ double2 smem[];
	constexpr int T = FftCfg<N>::T, CPB = FftCfg<N>::CPB;
	const int g = threadIdx.x / T, t = threadIdx.x % T;
	const int s = a.s0 + blockIdx.x * CPB + g;
	const bool active = s < a.s1;
	double2 *buf = smem + (size_t) g * N;

	if (active) {
		const long off = (long) (a.ch_map ? a.ch_map[s] : s) * a.ch_mul;
		const double *x = a.in + off;
		for (int n = t; n < N / 2; n += T) {
			const long f0 = 2L * n;
			const double x0 = (f0 < a.valid) ? x[f0 * a.stride] : 0.0;
			const double x1 = (f0 + 1 < a.valid) ? x[(f0 + 1) * a.stride] : 0.0;
			buf[n] = make_double2(x0, x1);
			buf[n + N / 2] = make_double2(0.0, 0.0);
		}
	}
	__syncthreads();
	fft_forward_smem<N>(buf, a.tw, t);
	if (active) {
		double2 *X = a.spec + (long) s * a.spec_ch_stride + (long) a.slot * N;
		for (int k = t; k <= N / 2; k += T) {
			if (k == 0) {
				const double2 z0 = buf[0];
				X[0] = make_double2(z0.x + z0.y, z0.x - z0.y);
			}
			else {
				const double2 zk = buf[k], zn = buf[N - k];
				const double2 e = make_double2(0.5 * (zk.x + zn.x), 0.5 * (zk.y - zn.y));
				const double2 o = make_double2(0.5 * (zk.y + zn.y), -0.5 * (zk.x - zn.x));
				const double2 wo = cmul(__ldg(&a.tw[k]), o);
				X[k] = cadd(e, wo);
				if (k != N / 2) X[N - k] = cconj(csub(e, wo));
			}
		}
	}
}

enum { INV_OUT = 1, INV_UPDATE_CARRY = 2 };

struct InvArgs {
	const double2 *Y;       // [s][N] packed spectra
	double *out;            // destination of first half + carry (INV_OUT)
	long stride;
	const int *ch_map;
	long ch_mul;
	double *carry;          // [s][B]
	int flags;
	const double2 *tw;
	int s0, s1;
};

template <int N>
__global__ void __launch_bounds__(FftCfg<N>::THREADS) k_fir_inv(InvArgs a)
{
	extern __shared__ double2 smem[];
	constexpr int T = FftCfg<N>::T, CPB = FftCfg<N>::CPB;
	const int g = threadIdx.x / T, t = threadIdx.x % T;
	const int s = a.s0 + blockIdx.x * CPB + g;
	const bool active = s < a.s1;
	double2 *buf = smem + (size_t) g * N;

	if (active) {
		const double2 *Y = a.Y + (long) s * N;
		for (int k = t; k <= N / 2; k += T) {
			if (k == 0) {
				const double2 y0 = Y[0];
				// Z[0] = E0 + i O0, stored conjugated
				buf[0] = make_double2(0.5 * (y0.x + y0.y), -0.5 * (y0.x - y0.y));
			}
			else {
				const double2 xk = Y[k], xn = Y[N - k];
				const double2 e = make_double2(0.5 * (xk.x + xn.x), 0.5 * (xk.y - xn.y));
				const double2 d = make_double2(0.5 * (xk.x - xn.x), 0.5 * (xk.y + xn.y));
				const double2 o = cmul(cconj(__ldg(&a.tw[k])), d);
				// Z[k] = E + iO, Z[N-k] = conj(E) + i conj(O); store conj(Z)
				buf[k] = make_double2(e.x - o.y, -(e.y + o.x));
				buf[N - k] = make_double2(e.x + o.y, -(o.x - e.y));
			}
		}
	}
	__syncthreads();
	fft_forward_smem<N>(buf, a.tw, t);
	if (active) {
		const double scale = 1.0 / N;
		const long off = (long) (a.ch_map ? a.ch_map[s] : s) * a.ch_mul;
		double *carry = a.carry + (long) s * N;
		for (int n = t; n < N / 2; n += T) {
			const double2 lo = buf[n], hi = buf[n + N / 2];
			if (a.flags & INV_OUT) {
				const double2 c = *reinterpret_cast<const double2 *>(carry + 2 * n);
				a.out[(2L * n) * a.stride + off] = fma(lo.x, scale, c.x);
				a.out[(2L * n + 1) * a.stride + off] = fma(-lo.y, scale, c.y);
			}
			if (a.flags & INV_UPDATE_CARRY)
				*reinterpret_cast<double2 *>(carry + 2 * n) = make_double2(hi.x * scale, -hi.y * scale);
		}
	}
}

// Y[s][k] = sum_{p in [p0,p1)} FDL[s][(slot0 - p) mod P][k] * H[s or 0][p][k]   (packed bin 0)
// This is the HBM-streaming kernel of the engine: 32 bytes in per complex MAC (16 with a
// shared IR, whose rows stay in L2), 4 DFMA.
struct MacArgs {
	const double2 *fdl;   // [s][P][N]
	const double2 *H;     // [s][P][N] or [P][N]
	double2 *Y;           // [s][N]
	int N, P;
	int slot0, p0, p1;
	long h_ch_stride;     // P*N or 0 (shared IR)
	int s0;
};

template <bool SHARED_H>
__global__ void __launch_bounds__(256) k_fir_mac(MacArgs a)
{
	const int k = blockIdx.x * blockDim.x + threadIdx.x;
	const int s = a.s0 + blockIdx.y;
	const double2 *fdl = a.fdl + (long) s * a.P * a.N + k;
	const double2 *H = a.H + (long) s * a.h_ch_stride + k;
	double2 acc0 = make_double2(0.0, 0.0), acc1 = acc0, acc2 = acc0, acc3 = acc0;
	const bool dc = (k == 0);
	int slot = a.slot0 - a.p0;
	slot %= a.P;
	if (slot < 0) slot += a.P;
	int p = a.p0;

#define LOAD_X(q) __ldcs(&fdl[(long) ((slot - (q) < 0) ? slot - (q) + a.P : slot - (q)) * a.N])
#define LOAD_H(q) (SHARED_H ? __ldg(&H[(long) (p + (q)) * a.N]) : __ldcs(&H[(long) (p + (q)) * a.N]))
#define CMAC(ACC, XV, HV)                                            \
	do {                                                             \
		if (dc) {                                                    \
			ACC.x = fma(XV.x, HV.x, ACC.x);                          \
			ACC.y = fma(XV.y, HV.y, ACC.y);                          \
		}                                                            \
		else {                                                       \
			ACC.x = fma(XV.x, HV.x, fma(-XV.y, HV.y, ACC.x));        \
			ACC.y = fma(XV.x, HV.y, fma(XV.y, HV.x, ACC.y));         \
		}                                                            \
	} while (0)

	for (; p + 4 <= a.p1; p += 4) {
		const double2 x0 = LOAD_X(0), x1 = LOAD_X(1), x2 = LOAD_X(2), x3 = LOAD_X(3);
		const double2 h0 = LOAD_H(0), h1 = LOAD_H(1), h2 = LOAD_H(2), h3 = LOAD_H(3);
		CMAC(acc0, x0, h0);
		CMAC(acc1, x1, h1);
		CMAC(acc2, x2, h2);
		CMAC(acc3, x3, h3);
		slot -= 4;
		if (slot < 0) slot += a.P;
	}
	for (; p < a.p1; ++p) {
		const double2 x0 = LOAD_X(0);
		const double2 h0 = LOAD_H(0);
		CMAC(acc0, x0, h0);
		slot -= 1;
		if (slot < 0) slot += a.P;
	}
#undef LOAD_X
#undef LOAD_H
#undef CMAC
	a.Y[(long) s * a.N + k] = make_double2((acc0.x + acc1.x) + (acc2.x + acc3.x), (acc0.y + acc1.y) + (acc2.y + acc3.y));
}

// general path: stash the new frames of the incomplete block, per-channel contiguous
__global__ void k_fir_stash(const double *in, long stride, const int *ch_map, double *xcur, int B, int pos, int seg, int s0)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	const int s = s0 + blockIdx.y;
	if (i < seg) xcur[(long) s * B + pos + i] = in[(long) i * stride + ch_map[s]];
}

// general path: out[m] = pre[m] + sum_{r<=m} xcur[r] h0[m-r], m = pos+i
__global__ void k_fir_head(const double *xcur, const double *pre, const double *h0, long h0_ch_stride,
                           double *out, long stride, const int *ch_map, int B, int pos, int seg, int s0)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	const int s = s0 + blockIdx.y;
	if (i >= seg) return;
	const int m = pos + i;
	const double *x = xcur + (long) s * B;
	const double *h = h0 + (long) s * h0_ch_stride;
	double acc0 = pre[(long) s * B + m], acc1 = 0.0;
	int r = 0;
	for (; r + 2 <= m + 1; r += 2) {
		acc0 = fma(x[r], h[m - r], acc0);
		acc1 = fma(x[r + 1], h[m - r - 1], acc1);
	}
	if (r <= m) acc0 = fma(x[r], h[m - r], acc0);
	out[(long) i * stride + ch_map[s]] = acc0 + acc1;
}

// fir.c latency: out[a] = y[a - L]; ring slot of absolute frame a is a % L
__global__ void k_delay_read(const double *y, long y_stride, const double *ring, double *out, long stride,
                             const int *ch_map, int n_sel, long frames, long L, long abs0)
{
	const long idx = (long) blockIdx.x * blockDim.x + threadIdx.x;
	if (idx >= frames * n_sel) return;
	const long i = idx / n_sel;
	const int s = (int) (idx - i * n_sel);
	const double v = (i < L) ? ring[((abs0 + i) % L) * n_sel + s] : y[(i - L) * y_stride + s];
	out[i * stride + ch_map[s]] = v;
}

__global__ void k_delay_write(const double *y, long y_stride, double *ring, int n_sel, long frames, long L, long abs0)
{
	const long idx = (long) blockIdx.x * blockDim.x + threadIdx.x;
	const long first = (frames > L) ? frames - L : 0;
	const long cnt = frames - first;
	if (idx >= cnt * n_sel) return;
	const long i = first + idx / n_sel;
	const int s = (int) (idx % n_sel);
	ring[((abs0 + i) % L) * n_sel + s] = y[i * y_stride + s];
}

// ------------------------------------------------------------------------------------------
// launch helpers (dispatch on the FFT size)
// ------------------------------------------------------------------------------------------
template <int N>
static int launch_fwd_n(const FwdArgs &a, cudaStream_t st)
{
	static std::atomic<int> configured[64];
	int dev = 0;
	cudaGetDevice(&dev);
	if (!configured[dev & 63].load()) {
		CUDA_TRY(cudaFuncSetAttribute(k_fir_fwd<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) FftCfg<N>::SMEM), return -1);
		CUDA_TRY(cudaFuncSetAttribute(k_fir_inv<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) FftCfg<N>::SMEM), return -1);
		configured[dev & 63].store(1);
	}
	const int n = a.s1 - a.s0;
	if (n <= 0) return 0;
	ProfScope prof("fir_fwd", st);
	LAUNCH(k_fir_fwd<N>, ceil_div(n, FftCfg<N>::CPB), FftCfg<N>::THREADS, FftCfg<N>::SMEM, st, a);
	return 0;
}

template <int N>
static int launch_inv_n(const InvArgs &a, cudaStream_t st)
{
	const int n = a.s1 - a.s0;
	if (n <= 0) return 0;
	ProfScope prof("fir_inv", st);
	LAUNCH(k_fir_inv<N>, ceil_div(n, FftCfg<N>::CPB), FftCfg<N>::THREADS, FftCfg<N>::SMEM, st, a);
	return 0;
}

#define DISPATCH_N(N_, FN, ...)                      \
	switch (N_) {                                    \
	case 64: return FN<64>(__VA_ARGS__);             \
	case 128: return FN<128>(__VA_ARGS__);           \
	case 256: return FN<256>(__VA_ARGS__);           \
	case 512: return FN<512>(__VA_ARGS__);           \
	case 1024: return FN<1024>(__VA_ARGS__);         \
	case 2048: return FN<2048>(__VA_ARGS__);         \
	case 4096: return FN<4096>(__VA_ARGS__);         \
	case 8192: return FN<8192>(__VA_ARGS__);         \
	default: set_error("unsupported FFT size %d", N_); return -1; \
	}

static int launch_fwd(int N, const FwdArgs &a, cudaStream_t st) { DISPATCH_N(N, launch_fwd_n, a, st) }
static int launch_inv(int N, const InvArgs &a, cudaStream_t st)
{
	// launch_fwd configures both kernels' shared-memory limits; make sure it ran for this N
	FwdArgs none = {};
	if (launch_fwd(N, none, st)) return -1;
	DISPATCH_N(N, launch_inv_n, a, st)
}

static void launch_mac(const MacArgs &a, int n_sel_range, bool shared_h, cudaStream_t st)
{
	if (n_sel_range <= 0) return;
	const int threads = (a.N < 256) ? a.N : 256;
	dim3 grid(a.N / threads, n_sel_range);
	ProfScope prof("fir_mac", st);
	if (shared_h) LAUNCH(k_fir_mac<true>, grid, threads, 0, st, a);
	else LAUNCH(k_fir_mac<false>, grid, threads, 0, st, a);
}

// ------------------------------------------------------------------------------------------
// unit-test hooks (tests/ only): packed real FFT round trip
// ------------------------------------------------------------------------------------------
int test_rfft(int B, int n_ch, const double *d_in, double *d_spec, cudaStream_t st)
{
	FwdArgs a = {};
	a.in = d_in; a.stride = 1; a.ch_map = nullptr; a.ch_mul = B; a.valid = B;
	a.spec = reinterpret_cast<double2 *>(d_spec); a.spec_ch_stride = B; a.slot = 0;
	a.tw = twiddles_2n(B);
	if (!a.tw) return -1;
	a.s0 = 0; a.s1 = n_ch;
	return launch_fwd(B, a, st);
}

int test_irfft(int B, int n_ch, const double *d_spec, double *d_out2B, cudaStream_t st)
{
	// first half -> out[ch][0:B) (carry is zeroed scratch), second half -> out[ch][B:2B)
	double *carry = dev_alloc<double>((size_t) n_ch * B);
	if (!carry) return -1;
	InvArgs a = {};
	a.Y = reinterpret_cast<const double2 *>(d_spec);
	a.out = d_out2B; a.stride = 1; a.ch_map = nullptr; a.ch_mul = 2L * B;
	a.carry = carry; a.flags = INV_OUT | INV_UPDATE_CARRY;
	a.tw = twiddles_2n(B);
	a.s0 = 0; a.s1 = n_ch;
	int r = (a.tw) ? launch_inv(B, a, st) : -1;
	if (r == 0) {
		for (int c = 0; c < n_ch && r == 0; ++c)
			if (cudaMemcpyAsync(d_out2B + (size_t) c * 2 * B + B, carry + (size_t) c * B, B * sizeof(double), cudaMemcpyDeviceToDevice, st) != cudaSuccess) r = -1;
	}
	cudaStreamSynchronize(st);
	dev_free(carry);
	return r;
}

// ------------------------------------------------------------------------------------------
// operator
// ------------------------------------------------------------------------------------------
struct FirOp : Op {
	// description (host)
	std::vector<int> h_ch_map;          // selected channel -> channel index in the slab
	std::vector<double> h_taps;         // [filter_frames][fc] (kept until planned)
	int fc = 1;                         // filter channels: 1 (shared) or n_sel
	long filter_frames = 0;
	long latency = 0;
	int n_sel = 0;

	// plan
	bool planned = false;
	int B = 0, P = 0;
	const double2 *tw = nullptr;

	// device state
	int *d_ch_map = nullptr;
	double2 *d_fdl = nullptr, *d_H = nullptr, *d_Y = nullptr;
	double *d_carry = nullptr, *d_xcur = nullptr, *d_pre = nullptr, *d_h0 = nullptr;
	double *d_ring = nullptr, *d_ytmp = nullptr;
	long ytmp_cap = 0;
	long blk = 0;        // completed blocks
	int pos = 0;         // frames of the current block already consumed
	bool pre_valid = false;
	long abs_frames = 0; // total frames seen (delay ring phase)

	const char *name() const override { return "fir"; }

	~FirOp() override
	{
		dev_free(d_ch_map); dev_free(d_fdl); dev_free(d_H); dev_free(d_Y); dev_free(d_carry);
		dev_free(d_xcur); dev_free(d_pre); dev_free(d_h0); dev_free(d_ring); dev_free(d_ytmp);
	}

	int plan(long hint, cudaStream_t st)
	{
		int b = 64;
		while (b * 2 <= hint && b < 8192) b *= 2;
		B = b;
		P = (int) ((filter_frames + B - 1) / B);
		tw = twiddles_2n(B);
		if (!tw) return -1;
		const size_t rows = (size_t) n_sel * P;
		d_fdl = dev_alloc<double2>(rows * B);
		d_H = dev_alloc<double2>((size_t) ((fc == 1) ? 1 : n_sel) * P * B);
		d_Y = dev_alloc<double2>((size_t) n_sel * B);
		d_carry = dev_alloc<double>((size_t) n_sel * B);
		d_xcur = dev_alloc<double>((size_t) n_sel * B);
		d_pre = dev_alloc<double>((size_t) n_sel * B);
		d_h0 = dev_alloc<double>((size_t) ((fc == 1) ? 1 : n_sel) * B);
		if (latency > 0) d_ring = dev_alloc<double>((size_t) latency * n_sel);
		if (!d_fdl || !d_H || !d_Y || !d_carry || !d_xcur || !d_pre || !d_h0 || (latency > 0 && !d_ring)) return -1;

		// filter spectra: H[c][p] = RFFT_2B(taps[pB : (p+1)B) of channel c), cf. fir_p.c:482-498
		const size_t n_taps = (size_t) filter_frames * fc;
		double *d_taps = dev_alloc<double>(n_taps, false);
		if (!d_taps) return -1;
		CUDA_TRY(cudaMemcpyAsync(d_taps, h_taps.data(), n_taps * sizeof(double), cudaMemcpyHostToDevice, st), return -1);
		const int nh = (fc == 1) ? 1 : n_sel;
		for (int p = 0; p < P; ++p) {
			FwdArgs a = {};
			a.in = d_taps + (size_t) p * B * fc;
			a.stride = fc; a.ch_map = nullptr; a.ch_mul = (fc == 1) ? 0 : 1;
			a.valid = filter_frames - (long) p * B;
			if (a.valid > B) a.valid = B;
			a.spec = d_H; a.spec_ch_stride = (long) P * B; a.slot = p; a.tw = tw;
			a.s0 = 0; a.s1 = nh;
			if (launch_fwd(B, a, st)) return -1;
		}
		// time-domain head taps h0[c][0:B)
		std::vector<double> h0((size_t) nh * B, 0.0);
		for (int c = 0; c < nh; ++c)
			for (long i = 0; i < B && i < filter_frames; ++i)
				h0[(size_t) c * B + i] = h_taps[(size_t) i * fc + c];
		CUDA_TRY(cudaMemcpyAsync(d_h0, h0.data(), h0.size() * sizeof(double), cudaMemcpyHostToDevice, st), return -1);
		CUDA_TRY(cudaStreamSynchronize(st), return -1);
		dev_free(d_taps);
		h_taps.clear();
		h_taps.shrink_to_fit();
		planned = true;
		return 0;
	}

	void reset(cudaStream_t st) override
	{
		blk = 0; pos = 0; pre_valid = false; abs_frames = 0;
		if (!planned) return;
		cudaMemsetAsync(d_fdl, 0, (size_t) n_sel * P * B * sizeof(double2), st);
		cudaMemsetAsync(d_carry, 0, (size_t) n_sel * B * sizeof(double), st);
		cudaMemsetAsync(d_xcur, 0, (size_t) n_sel * B * sizeof(double), st);
		if (d_ring) cudaMemsetAsync(d_ring, 0, (size_t) latency * n_sel * sizeof(double), st);
	}

	void mac(int p0, int p1, long slot_blk, cudaStream_t st)
	{
		MacArgs m = {};
		m.fdl = d_fdl; m.H = d_H; m.Y = d_Y; m.N = B; m.P = P;
		m.slot0 = (int) (slot_blk % P); m.p0 = p0; m.p1 = p1;
		m.h_ch_stride = (fc == 1) ? 0 : (long) P * B;
		m.s0 = 0;
		launch_mac(m, n_sel, fc == 1, st);
	}

	// one whole aligned block: src/dst are interleaved with `stride`, channel offsets via ch_map
	int fast_block(const double *src, long sstride, double *dst, long dstride, const int *dmap, long dmul, cudaStream_t st)
	{
		FwdArgs f = {};
		f.in = src; f.stride = sstride; f.ch_map = d_ch_map; f.ch_mul = 1; f.valid = B;
		f.spec = d_fdl; f.spec_ch_stride = (long) P * B; f.slot = (int) (blk % P); f.tw = tw;
		f.s0 = 0; f.s1 = n_sel;
		if (launch_fwd(B, f, st)) return -1;
		mac(0, P, blk, st);
		InvArgs v = {};
		v.Y = d_Y; v.out = dst; v.stride = dstride; v.ch_map = dmap; v.ch_mul = dmul;
		v.carry = d_carry; v.flags = INV_OUT | INV_UPDATE_CARRY; v.tw = tw; v.s0 = 0; v.s1 = n_sel;
		if (launch_inv(B, v, st)) return -1;
		++blk;
		pre_valid = false;
		return 0;
	}

	int ensure_pre(cudaStream_t st)
	{
		if (pre_valid) return 0;
		mac(1, P, blk, st);   // R_blk: completed blocks only
		InvArgs v = {};
		v.Y = d_Y; v.out = d_pre; v.stride = 1; v.ch_map = nullptr; v.ch_mul = B;
		v.carry = d_carry; v.flags = INV_OUT; v.tw = tw; v.s0 = 0; v.s1 = n_sel;
		if (launch_inv(B, v, st)) return -1;
		pre_valid = true;
		return 0;
	}

	long run(long frames, const double *in, double *out, cudaStream_t st) override
	{
		if (frames <= 0) return 0;
		if (!planned && plan(frames, st)) return -1;
		const long C = channels;
		if (in != out && n_sel < C)
			CUDA_TRY(cudaMemcpyAsync(out, in, (size_t) frames * C * sizeof(double), cudaMemcpyDeviceToDevice, st), return -1);
		if (n_sel == 0) return frames;

		// where the convolution result goes: straight to `out`, or to a compact temp when a
		// latency ring follows
		double *dst = out;
		long dstride = C;
		const int *dmap = d_ch_map;
		if (latency > 0) {
			if (ytmp_cap < frames) {
				dev_free(d_ytmp);
				d_ytmp = dev_alloc<double>((size_t) frames * n_sel, false);
				if (!d_ytmp) return -1;
				ytmp_cap = frames;
			}
			dst = d_ytmp; dstride = n_sel; dmap = nullptr;
		}

		long done = 0;
		while (done < frames) {
			const double *src = in + done * C;
			double *d = dst + done * dstride;
			if (pos == 0 && frames - done >= B) {
				if (fast_block(src, C, d, dstride, dmap, 1, st)) return -1;
				done += B;
				continue;
			}
			const int seg = (int) ((frames - done < B - pos) ? frames - done : B - pos);
			if (ensure_pre(st)) return -1;
			dim3 grid(ceil_div(seg, 128), n_sel);
			LAUNCH(k_fir_stash, grid, 128, 0, st, src, C, d_ch_map, d_xcur, B, pos, seg, 0);
			if (dmap) LAUNCH(k_fir_head, grid, 128, 0, st, d_xcur, d_pre, d_h0, (fc == 1) ? 0L : (long) B, d, dstride, dmap, B, pos, seg, 0);
			else LAUNCH(k_fir_head, grid, 128, 0, st, d_xcur, d_pre, d_h0, (fc == 1) ? 0L : (long) B, d, dstride, d_iota(), B, pos, seg, 0);
			pos += seg;
			done += seg;
			if (pos == B) {
				// block complete: X_blk into the FDL, carry_blk = IRFFT(S_blk)[B:2B)
				FwdArgs f = {};
				f.in = d_xcur; f.stride = 1; f.ch_map = nullptr; f.ch_mul = B; f.valid = B;
				f.spec = d_fdl; f.spec_ch_stride = (long) P * B; f.slot = (int) (blk % P); f.tw = tw;
				f.s0 = 0; f.s1 = n_sel;
				if (launch_fwd(B, f, st)) return -1;
				mac(0, P, blk, st);
				InvArgs v = {};
				v.Y = d_Y; v.out = nullptr; v.stride = 0; v.ch_map = nullptr; v.ch_mul = 0;
				v.carry = d_carry; v.flags = INV_UPDATE_CARRY; v.tw = tw; v.s0 = 0; v.s1 = n_sel;
				if (launch_inv(B, v, st)) return -1;
				++blk;
				pos = 0;
				pre_valid = false;
			}
		}

		if (latency > 0) {
			const long total = frames * n_sel;
			LAUNCH(k_delay_read, ceil_div(total, 256), 256, 0, st, d_ytmp, (long) n_sel, d_ring, out, C, d_ch_map, n_sel, frames, latency, abs_frames);
			const long cnt = (frames > latency) ? latency : frames;
			LAUNCH(k_delay_write, ceil_div(cnt * n_sel, 256), 256, 0, st, d_ytmp, (long) n_sel, d_ring, n_sel, frames, latency, abs_frames);
		}
		abs_frames += frames;
		return frames;
	}

	// identity map for the compact temp layout (head kernel wants a map)
	int *d_iota_buf = nullptr;
	const int *d_iota()
	{
		if (!d_iota_buf) {
			std::vector<int> v(n_sel);
			for (int i = 0; i < n_sel; ++i) v[i] = i;
			d_iota_buf = dev_alloc<int>(n_sel, false);
			cudaMemcpy(d_iota_buf, v.data(), n_sel * sizeof(int), cudaMemcpyHostToDevice);
		}
		return d_iota_buf;
	}
};

Op *make_fir_op(int slab_channels, int fs, const char *slab_selector, const double *taps, int filter_channels,
                long filter_frames, const int *taps_cols, long latency, long block_hint, cudaStream_t st)
{
	std::unique_ptr<FirOp> op(new FirOp());
	op->channels = slab_channels;
	op->fs_in = op->fs_out = fs;
	for (int c = 0; c < slab_channels; ++c)
		if (!slab_selector || slab_selector[c]) op->h_ch_map.push_back(c);
	op->n_sel = (int) op->h_ch_map.size();
	op->fc = (filter_channels == 1) ? 1 : op->n_sel;
	op->filter_frames = filter_frames;
	op->latency = latency;
	if (op->n_sel > 0) {
		// gather this slab's columns of the filter: taps_cols[k] = column of the k-th selected channel
		op->h_taps.resize((size_t) filter_frames * op->fc);
		for (long i = 0; i < filter_frames; ++i)
			for (int k = 0; k < op->fc; ++k)
				op->h_taps[(size_t) i * op->fc + k] = taps[(size_t) i * filter_channels + ((filter_channels == 1) ? 0 : taps_cols[k])];
		op->d_ch_map = dev_alloc<int>(op->n_sel, false);
		if (!op->d_ch_map) return nullptr;
		CUDA_TRY(cudaMemcpy(op->d_ch_map, op->h_ch_map.data(), op->n_sel * sizeof(int), cudaMemcpyHostToDevice), return nullptr);
		if (block_hint > 0 && op->plan(block_hint, st)) return nullptr;
	}
	return op.release();
}

}  // namespace dspb200
