// fft.cuh -- in-shared-memory FP64 Stockham FFT (power-of-two, radix 4 + one radix 2),
// and the split/merge steps that turn an N-point complex transform into a 2N-point real one.
//
// One FFT of N complex points is worked on by T = N/8 threads (FftCfg<N>::T); a CTA holds
// CPB independent FFTs side by side (thread group g = threadIdx.x / T works on s + g*N).
// Each pass: every thread pulls its butterflies into registers, the CTA synchronises, the
// butterflies are written back in Stockham (autosort) order, the CTA synchronises again --
// so a single N-point buffer suffices (N = 8192 complex doubles = 128 KB of the 227 KB).
//
// Twiddles come from a table W[t] = exp(-2 pi i t / (2N)), t in [0, 2N), built on the host
// in long double (twiddles_2n()); the same table serves the real-FFT split step.
#pragma once

#include <cuda_runtime.h>

namespace dspb200 {

template <int N>
struct FftCfg {
	static_assert(N >= 64 && N <= 8192 && (N & (N - 1)) == 0, "N must be a power of two in [64, 8192]");
	static constexpr int T = N / 8;                         // threads per FFT
	static constexpr int CPB = (T >= 256) ? 1 : 256 / T;    // FFTs per CTA
	static constexpr int THREADS = T * CPB;
	static constexpr size_t SMEM = (size_t) N * CPB * sizeof(double2);
};

__device__ __forceinline__ double2 cmul(double2 a, double2 b)
{
	return make_double2(fma(a.x, b.x, -a.y * b.y), fma(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 cconj(double2 a) { return make_double2(a.x, -a.y); }

// One radix-R Stockham pass over s[0..N) (sub-transform length so far: Ns).
template <int N, int R>
__device__ __forceinline__ void fft_pass(double2 *s, const double2 *__restrict__ tw, int t, int Ns)
{
	constexpr int T = FftCfg<N>::T;
	constexpr int BF = N / R;       // butterflies in this pass
	constexpr int PT = BF / T;      // butterflies per thread (2 for R = 4, 4 for R = 2)
	double2 v[PT][R];
#pragma unroll
	for (int b = 0; b < PT; ++b) {
		const int j = t + b * T;
#pragma unroll
		for (int r = 0; r < R; ++r) v[b][r] = s[j + r * BF];
	}
	__syncthreads();
#pragma unroll
	for (int b = 0; b < PT; ++b) {
		const int j = t + b * T;
		const int k = j & (Ns - 1);
		if (Ns > 1) {
			// W_{Ns*R}^{r k} = W_{2N}^{2 r k N/(Ns R)}
			const int step = 2 * (N / R / Ns) * k;
#pragma unroll
			for (int r = 1; r < R; ++r) v[b][r] = cmul(v[b][r], __ldg(&tw[r * step]));
		}
		const int j0 = (j - k) * R + k;
		if (R == 4) {
			const double2 a0 = cadd(v[b][0], v[b][2]), a1 = csub(v[b][0], v[b][2]);
			const double2 a2 = cadd(v[b][1], v[b][3]);
			const double2 d = csub(v[b][1], v[b][3]);
			const double2 a3 = make_double2(d.y, -d.x);  // d * (-i)
			s[j0] = cadd(a0, a2);
			s[j0 + Ns] = cadd(a1, a3);
			s[j0 + 2 * Ns] = csub(a0, a2);
			s[j0 + 3 * Ns] = csub(a1, a3);
		}
		else {
			s[j0] = cadd(v[b][0], v[b][1]);
			s[j0 + Ns] = csub(v[b][0], v[b][1]);
		}
	}
	__syncthreads();
}

// Forward complex DFT (exp(-2 pi i nk/N)) of s[0..N), natural order in and out.
// All FftCfg<N>::THREADS threads of the CTA must call it; t = thread index within the FFT.
template <int N>
__device__ __forceinline__ void fft_forward_smem(double2 *s, const double2 *__restrict__ tw, int t)
{
	int Ns = 1;
#pragma unroll
	for (int m = N; m >= 4; m >>= 2) {
		fft_pass<N, 4>(s, tw, t, Ns);
		Ns <<= 2;
	}
	if (Ns < N) fft_pass<N, 2>(s, tw, t, Ns);
}

}  // namespace dspb200
