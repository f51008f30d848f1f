// fft.cuh -- in-shared-memory FP64 Stockham FFT (power-of-two sizes 64..8192) with register-resident
// radix-16 / 8 / 4 butterflies, and the helpers of the packed real transform built on it.
//
// One FFT of N complex points is worked on by T = N/16 threads: in every pass a thread owns 16
// points (one radix-16 butterfly, two radix-8, or four radix-4), pulls them into registers, the CTA
// synchronises, the butterflies are written back in Stockham (autosort) order, the CTA synchronises
// again -- a single N-point buffer suffices (N = 8192 complex doubles = 128 KB of the 227 KB) and
// a 4096-point transform is 3 passes (16 x 16 x 16).  A CTA holds CPB independent FFTs side by side.
//
// Twiddles are built on the host in long double.  A butterfly loads only w^1, w^2, w^4, w^8 from
// a compact per-pass table (twiddles_pass(), see fft_pass_table_size()) and forms the other powers
// by multiplication (each power is a product of at most 4 table values); the real-transform split
// step uses W[k] = exp(-2 pi i k / (2N)) (twiddles_2n()).
#pragma once

#include <cuda_runtime.h>

namespace dspb200 {

template <int N>
struct FftCfg {
	static_assert(N >= 64 && N <= 8192 && (N & (N - 1)) == 0, "N must be a power of two in [64, 8192]");
	static constexpr int T = N / 16;                        // threads per FFT
	static constexpr int CPB = (T >= 256) ? 1 : 256 / T;    // FFTs per CTA
	static constexpr int THREADS = T * CPB;
	static constexpr int STRIDE = N + N / 16;                // padded points per FFT buffer (see spad())
	static constexpr size_t SMEM = (size_t) STRIDE * CPB * sizeof(double2);
};

// Radix of pass `pass` for an N-point transform (0 = no such pass): as many 16s as possible.
__host__ __device__ constexpr int fft_radix(int N, int pass)
{
	return (N == 64) ? ((pass == 0) ? 16 : (pass == 1) ? 4 : 0)
	     : (N == 128) ? ((pass == 0) ? 16 : (pass == 1) ? 8 : 0)
	     : (N == 256) ? ((pass < 2) ? 16 : 0)
	     : (N == 512) ? ((pass == 0) ? 16 : (pass == 1) ? 8 : (pass == 2) ? 4 : 0)
	     : (N == 1024) ? ((pass < 2) ? 16 : (pass == 2) ? 4 : 0)
	     : (N == 2048) ? ((pass < 2) ? 16 : (pass == 2) ? 8 : 0)
	     : (N == 4096) ? ((pass < 3) ? 16 : 0)
	     : (N == 8192) ? ((pass < 2) ? 16 : (pass == 2) ? 8 : (pass == 3) ? 4 : 0)
	     : 0;
}
// Per-pass twiddle tables (twiddles_pass()): for every pass with Ns > 1, 4*Ns entries
//   ptw[off + mi*Ns + k] = exp(-2 pi i k 2^mi / (Ns R)),  mi = 0..3, k = 0..Ns-1,
// laid out so that consecutive threads (consecutive k) read consecutive entries; the whole set is
// 17 KB for N = 4096 and stays in L1.  Size in entries:
__host__ __device__ constexpr int fft_pass_table_size(int N)
{
	int ns = 1, total = 0;
	for (int p = 0; p < 4 && fft_radix(N, p) != 0; ++p) {
		if (ns > 1) total += 4 * ns;
		ns *= fft_radix(N, p);
	}
	return total;
}

// Shared-memory index padding: one spare complex after every 16.  The first radix-16 pass writes
// 16 consecutive points per thread (thread stride 256 B = every lane on the same banks, an 8-way
// conflict for 128-bit stores); with the pad the thread stride is 272 B and the quarter-warps are
// conflict-free.  Every access to an FFT buffer goes through spad().
__device__ __forceinline__ int spad(int i) { return i + (i >> 4); }

__device__ __forceinline__ double2 cmul(double2 a, double2 b)
{
	return make_double2(fma(a.x, b.x, -a.y * b.y), fma(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 cconj(double2 a) { return make_double2(a.x, -a.y); }
__device__ __forceinline__ double2 mul_neg_i(double2 a) { return make_double2(a.y, -a.x); }   // a * (-i)

// ---- register butterflies: v <- DFT_R(v), natural order in and out (forward, exp(-2 pi i mr/R)) ----
__device__ __forceinline__ void dft4(double2 &v0, double2 &v1, double2 &v2, double2 &v3)
{
	const double2 a0 = cadd(v0, v2), a1 = csub(v0, v2), a2 = cadd(v1, v3), a3 = mul_neg_i(csub(v1, v3));
	v0 = cadd(a0, a2);
	v1 = cadd(a1, a3);
	v2 = csub(a0, a2);
	v3 = csub(a1, a3);
}

template <int R>
__device__ __forceinline__ void dft(double2 (&v)[R]);

template <>
__device__ __forceinline__ void dft<4>(double2 (&v)[4])
{
	dft4(v[0], v[1], v[2], v[3]);
}

template <>
__device__ __forceinline__ void dft<8>(double2 (&v)[8])
{
	constexpr double h = 0.70710678118654752440;   // sqrt(1/2)
	// even / odd 4-point transforms
	dft4(v[0], v[2], v[4], v[6]);   // E0..E3 in v0,v2,v4,v6
	dft4(v[1], v[3], v[5], v[7]);   // O0..O3 in v1,v3,v5,v7
	const double2 o0 = v[1];
	const double2 o1 = make_double2(h * (v[3].x + v[3].y), h * (v[3].y - v[3].x));     // * (1 - i)/sqrt2
	const double2 o2 = mul_neg_i(v[5]);
	const double2 o3 = make_double2(h * (v[7].y - v[7].x), -h * (v[7].x + v[7].y));    // * (-1 - i)/sqrt2
	const double2 e0 = v[0], e1 = v[2], e2 = v[4], e3 = v[6];
	v[0] = cadd(e0, o0); v[4] = csub(e0, o0);
	v[1] = cadd(e1, o1); v[5] = csub(e1, o1);
	v[2] = cadd(e2, o2); v[6] = csub(e2, o2);
	v[3] = cadd(e3, o3); v[7] = csub(e3, o3);
}

template <>
__device__ __forceinline__ void dft<16>(double2 (&v)[16])
{
	constexpr double c1 = 0.92387953251128675613, s1 = 0.38268343236508977173;   // cos, sin (pi/8)
	constexpr double h = 0.70710678118654752440;
	// n = 4 n1 + n2: 4-point transforms over n1 for each n2; A[n2][k1] lands in v[n2 + 4 k1]
#pragma unroll
	for (int n2 = 0; n2 < 4; ++n2) dft4(v[n2], v[n2 + 4], v[n2 + 8], v[n2 + 12]);
	// twiddle A[n2][k1] *= W16^(n2 k1), W16^m = cos(pi m/8) - i sin(pi m/8)
	const double2 w1 = make_double2(c1, -s1), w2 = make_double2(h, -h), w3 = make_double2(s1, -c1);
	const double2 w6 = make_double2(-h, -h), w9 = make_double2(-c1, s1);
	v[1 + 4] = cmul(v[1 + 4], w1);            // n2=1,k1=1
	v[1 + 8] = cmul(v[1 + 8], w2);            // n2=1,k1=2
	v[1 + 12] = cmul(v[1 + 12], w3);          // n2=1,k1=3
	v[2 + 4] = cmul(v[2 + 4], w2);            // n2=2,k1=1
	v[2 + 8] = mul_neg_i(v[2 + 8]);           // n2=2,k1=2: W16^4 = -i
	v[2 + 12] = cmul(v[2 + 12], w6);          // n2=2,k1=3
	v[3 + 4] = cmul(v[3 + 4], w3);            // n2=3,k1=1
	v[3 + 8] = cmul(v[3 + 8], w6);            // n2=3,k1=2
	v[3 + 12] = cmul(v[3 + 12], w9);          // n2=3,k1=3
	// 4-point transforms over n2 for each k1: X[k1 + 4 k2] lands in v[4 k1 + k2]
#pragma unroll
	for (int k1 = 0; k1 < 4; ++k1) dft4(v[4 * k1], v[4 * k1 + 1], v[4 * k1 + 2], v[4 * k1 + 3]);
	// natural order: X[r] = v[4 (r % 4) + r / 4]  (a 4x4 transpose)
#pragma unroll
	for (int a = 0; a < 4; ++a)
#pragma unroll
		for (int b = a + 1; b < 4; ++b) {
			const double2 t = v[4 * a + b];
			v[4 * a + b] = v[4 * b + a];
			v[4 * b + a] = t;
		}
}

// Who synchronises between the phases of a pass: the whole CTA (default), or -- in warp-specialised kernels
// where only some warps transform -- a named barrier over the transforming threads.
struct CtaSync {
	static __device__ __forceinline__ void sync() { __syncthreads(); }
};
template <int ID, int COUNT>
struct NamedSync {
	static __device__ __forceinline__ void sync() { asm volatile("bar.sync %0, %1;" ::"n"(ID), "n"(COUNT) : "memory"); }
};

// One radix-R Stockham pass over s[0..N) (sub-transform length so far: Ns); ptw = this pass's table.
template <int N, int R, class Sync = CtaSync>
__device__ __forceinline__ void fft_pass(double2 *s, const double2 *__restrict__ ptw, int t, int Ns)
{
	constexpr int T = FftCfg<N>::T;
	constexpr int BF = N / R;       // butterflies in this pass
	constexpr int PT = BF / T;      // butterflies per thread (16 / R)
	double2 v[PT][R];
#pragma unroll
	for (int b = 0; b < PT; ++b) {
		const int j = t + b * T;
#pragma unroll
		for (int r = 0; r < R; ++r) v[b][r] = s[spad(j + r * BF)];
	}
	Sync::sync();
#pragma unroll
	for (int b = 0; b < PT; ++b) {
		const int j = t + b * T;
		const int k = j & (Ns - 1);
		if (Ns > 1) {
			// w = W_{Ns R}^k; powers 1, 2, 4, 8 from the table, the rest by products
			double2 w[R];
			w[1] = __ldg(&ptw[k]);
			if (R > 2) w[2] = __ldg(&ptw[Ns + k]);
			if (R > 4) w[4] = __ldg(&ptw[2 * Ns + k]);
			if (R > 8) w[8] = __ldg(&ptw[3 * Ns + k]);
#pragma unroll
			for (int r = 3; r < R; ++r)
				if (r & (r - 1)) w[r] = cmul(w[r & (r - 1)], w[r & -r]);   // r = (r without lowest bit) + lowest bit
#pragma unroll
			for (int r = 1; r < R; ++r) v[b][r] = cmul(v[b][r], w[r]);
		}
		dft<R>(v[b]);
		const int j0 = (j - k) * R + k;
#pragma unroll
		for (int r = 0; r < R; ++r) s[spad(j0 + r * Ns)] = v[b][r];
	}
	Sync::sync();
}

template <int N, int PASS, int NS, int OFF, class Sync>
struct FftPasses {
	static __device__ __forceinline__ void run(double2 *s, const double2 *__restrict__ ptw, int t)
	{
		constexpr int R = fft_radix(N, PASS);
		if constexpr (R != 0) {
			fft_pass<N, R, Sync>(s, ptw + OFF, t, NS);
			FftPasses<N, PASS + 1, NS * R, OFF + ((NS > 1) ? 4 * NS : 0), Sync>::run(s, ptw, t);
		}
	}
};

// Forward complex DFT (exp(-2 pi i nk/N)) of s[0..N), natural order in and out.
// All threads that Sync synchronises (default: the FftCfg<N>::THREADS threads of the CTA) must call it;
// t = thread index within the FFT.
template <int N, class Sync = CtaSync>
__device__ __forceinline__ void fft_forward_smem(double2 *s, const double2 *__restrict__ ptw, int t)
{
	FftPasses<N, 0, 1, 0, Sync>::run(s, ptw, t);
}

}  // namespace dspb200
