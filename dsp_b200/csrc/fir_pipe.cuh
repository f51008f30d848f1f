// fir_pipe.cuh -- K2's block kernel as a persistent, warp-specialised pipeline (single-level plans).
//
// One CTA per SM walks the channels s = blockIdx.x, blockIdx.x + gridDim.x, ...; for the block j of channel s
//     S = X_j H_0 + sum_{1 <= p < pf} X_{j-p} H_p + V_j        (reference: fft_part_group_compute, fir_p.c:64-103)
//     y = IRFFT(S)[0:B) + carry,  carry' = IRFFT(S)[B:2B)
// with three kinds of warps that only meet at mbarriers:
//   producer (1 warp, one lane)  streams the FDL rows X_{j-p}, the filter rows H_p and V_j of the CTA's channels
//                                from HBM into a ring of shared-memory stages with cp.async.bulk (TMA bulk copies,
//                                completion counted in bytes on the stage's "full" mbarrier) -- it runs ahead of
//                                everybody else, across channel boundaries, limited only by the ring
//   MAC warps (4)                multiply-accumulate the stages into registers (512 bins per stage, 4 per lane) and
//                                drop the finished sums into the spectrum buffer `sbuf`
//   transform team (N/16 threads) reads the caller's interleaved block, forward FFT in shared memory, spectrum to
//                                the FDL, S = sbuf + X_j H_0 for the bin pairs it owns, inverse FFT in place,
//                                overlap-add, writes the caller's interleaved result
// The HBM stream (everything but 8 + 16 + 24 of the bytes per sample) therefore never waits for a transform and
// the transforms never wait for a load they did not issue a whole phase earlier.  sbuf is handed back and forth
// by two mbarriers (s_full: MAC -> team, s_empty: team -> MAC); the team synchronises itself with a named barrier.
#pragma once
#include "common.cuh"
#include "fft.cuh"
#include <cstdint>

namespace dspb200 {

template <int N>
struct PipeCfg {
	static constexpr int TF = N / 16;                 // transform team (one FFT)
	static constexpr int MAC_WARPS = 4;
	static constexpr int TM = MAC_WARPS * 32;
	static constexpr int THREADS = TF + TM + 32;      // + producer warp
	static constexpr int CHUNK = 512;                 // bins per stage
	static constexpr int NCHUNK = N / CHUNK;
	static constexpr int PER = CHUNK / TM;            // bins per MAC lane and stage
	static constexpr int NS = (N >= 4096) ? 5 : 8;    // ring stages
	static constexpr size_t STAGE_BYTES = 2 * (size_t) CHUNK * sizeof(double2);   // X chunk | H chunk
	static constexpr size_t FBUF = (size_t) FftCfg<N>::STRIDE * sizeof(double2);
	static constexpr size_t SBUF = (size_t) N * sizeof(double2);
	static constexpr size_t SMEM = FBUF + SBUF + NS * STAGE_BYTES + (2 * NS + 2) * sizeof(uint64_t);
	static_assert(TF % 32 == 0 && N % CHUNK == 0 && CHUNK % TM == 0, "shape");
};

struct PipeArgs {
	const double *xin;       // caller's interleaved block
	long xin_stride;
	const int *xin_map;      // channel of selected channel s (NULL: s)
	double *yout;
	long yout_stride;
	const int *yout_map;
	double2 *fdl;            // [s][fdl_rows][N]
	long fdl_ch_stride;
	int fdl_rows, slot;      // slot = row of block j
	const double2 *H;        // [s or 0][P][N]
	long h_ch_stride;        // 0: shared filter
	int pf;                  // partitions summed here: 0 .. pf-1
	const double2 *V;        // [s][N] spectrum of the older partitions for block j (NULL: none)
	double *carry;           // [s][N]
	const double2 *tw, *ptw;
	int n_ch;
	int evict_first;         // streaming operands are read once per block: keep them from displacing the rest of L2
	int fake_io;             // MEASUREMENT ONLY (DSP_B200_FIR_PIPE_FAKEIO): contiguous block I/O (wrong results) to time the kernel without the strided accesses
};

// ---- mbarrier / bulk-copy primitives (PTX; SASS: SYNCS.*, UBLKCP) -------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, unsigned bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, unsigned parity)
{
	uint32_t ok;
	asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
	             : "=r"(ok)
	             : "r"(smem_u32(bar)), "r"(parity)
	             : "memory");
	return ok != 0;
}

// try_wait suspends the warp in hardware for a bounded time; a wait that has not come true after 2^20 of them (>= 50 ms) is a
// protocol bug: trap (the launch fails with an error) rather than hang the device
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity)
{
	unsigned spins = 0;
	while (!mbar_try_wait(bar, parity))
		if (++spins > (1u << 20)) __trap();
}

__device__ __forceinline__ uint64_t l2_policy_evict_first()
{
	uint64_t pol;
	asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
	return pol;
}

// global -> shared bulk copy, completion signalled on `bar` in bytes
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, uint64_t *bar, bool hint, uint64_t pol)
{
	if (hint)
		asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(dst)),
		             "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
		             : "memory");
	else
		asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
		             "r"(bytes), "r"(smem_u32(bar))
		             : "memory");
}

__device__ __forceinline__ double2 pipe_cmac(double2 acc, double2 x, double2 h)
{
	return make_double2(fma(x.x, h.x, fma(-x.y, h.y, acc.x)), fma(x.x, h.y, fma(x.y, h.x, acc.y)));
}

__device__ __forceinline__ void pipe_prefetch_rows(const void *p, long bytes, int t, int T)
{
	const char *c = static_cast<const char *>(p);
	for (long off = (long) t * 128; off < bytes; off += (long) T * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(c + off));
}

// Register cap: 416 threads x 104 registers leave room for one 256-thread CTA of the batched MAC kernel on the same
// SM, so that the HBM stream of the look-ahead MAC fills the time this kernel's transform team spends on latency.
#ifndef FIR_PIPE_MAXNREG
#define FIR_PIPE_MAXNREG 104
#endif
template <int N>
__global__ void __maxnreg__(FIR_PIPE_MAXNREG) k_fir_pipe(PipeArgs a)
{
	using Cfg = PipeCfg<N>;
	constexpr int T = Cfg::TF, CHUNK = Cfg::CHUNK, NS = Cfg::NS, PER = Cfg::PER, TM = Cfg::TM;
	using TeamSync = NamedSync<1, T>;
	extern __shared__ __align__(128) unsigned char smem_raw[];
	double2 *fbuf = reinterpret_cast<double2 *>(smem_raw);
	double2 *sbuf = reinterpret_cast<double2 *>(smem_raw + Cfg::FBUF);
	double2 *ring = reinterpret_cast<double2 *>(smem_raw + Cfg::FBUF + Cfg::SBUF);
	uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + Cfg::FBUF + Cfg::SBUF + NS * Cfg::STAGE_BYTES);
	uint64_t *full = bars, *empty = bars + NS, *s_full = bars + 2 * NS, *s_empty = bars + 2 * NS + 1;

	if (threadIdx.x == 0) {
		for (int i = 0; i < NS; ++i) {
			mbar_init(&full[i], 1);                  // the producer's arrive.expect_tx; the copies complete the bytes
			mbar_init(&empty[i], Cfg::MAC_WARPS);    // one arrival per MAC warp
		}
		mbar_init(s_full, Cfg::MAC_WARPS);
		mbar_init(s_empty, T / 32);                  // one arrival per team warp
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();   // the only CTA-wide barrier: from here on the roles meet at mbarriers

	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

	if (warp < T / 32) {
		// ------------------------------------------------------------------------------------------
		// transform team
		// ------------------------------------------------------------------------------------------
		const int t = threadIdx.x;
		double2 *buf = fbuf;
		int it = 0;
		for (int s = blockIdx.x; s < a.n_ch; s += gridDim.x, ++it) {
			double2 *fdl = a.fdl + (long) s * a.fdl_ch_stride;
			const double2 *H0 = a.H + (long) s * a.h_ch_stride;
			double2 *X = fdl + (long) a.slot * N;
			double2 *carry = reinterpret_cast<double2 *>(a.carry + (long) s * N);
			{
				// frames 2n, 2n+1 of this channel (the CTAs next door read the rest of each sector at about the same time)
				const double *xc = a.fake_io ? a.xin + (long) s * N : a.xin + (a.xin_map ? a.xin_map[s] : s);
				const long xs = a.fake_io ? 1 : a.xin_stride;
				double2 v[8];
#pragma unroll
				for (int i = 0; i < 8; ++i) {
					const long n2 = 2L * (t + i * T);
					v[i] = make_double2(xc[n2 * xs], xc[(n2 + 1) * xs]);   // plain loads: the call may be in place
				}
				// what the later phases read with plain loads: on its way to L2 while the first transform runs
				pipe_prefetch_rows(H0, (long) N * sizeof(double2), t, T);
				pipe_prefetch_rows(carry, (long) N * sizeof(double), t, T);
#pragma unroll
				for (int i = 0; i < 8; ++i) {
					buf[spad(t + i * T)] = v[i];
					buf[spad(t + i * T + N / 2)] = make_double2(0.0, 0.0);
				}
			}
			TeamSync::sync();
			fft_forward_smem<N, TeamSync>(buf, a.ptw, t);
			double2 w[8];
#pragma unroll
			for (int i = 0; i < 8; ++i) w[i] = __ldg(&a.tw[t + i * T]);
			// (1) real split: X[k], X[N-k] to the FDL and, in place of Z, to shared memory.
			//     Thread t owns the pairs k = t + i T; thread 0 also owns k = 0 (packed DC/Nyquist) and k = N/2.
#pragma unroll
			for (int i = 0; i < 8; ++i) {
				const int k = t + i * T;
				if (k == 0) {
					const double2 z0 = buf[0], zh = buf[spad(N / 2)];
					const double2 x0 = make_double2(z0.x + z0.y, z0.x - z0.y), xh = cconj(zh);
					X[0] = x0; X[N / 2] = xh;
					buf[0] = x0; buf[spad(N / 2)] = xh;
				}
				else {
					const double2 zk = buf[spad(k)], zn = buf[spad(N - k)];
					const double2 e0 = make_double2(0.5 * (zk.x + zn.x), 0.5 * (zk.y - zn.y));
					const double2 o0 = make_double2(0.5 * (zk.y + zn.y), -0.5 * (zk.x - zn.x));
					const double2 wo = cmul(w[i], o0);
					const double2 xk = cadd(e0, wo), xn = cconj(csub(e0, wo));
					X[k] = xk; X[N - k] = xn;
					buf[spad(k)] = xk; buf[spad(N - k)] = xn;
				}
			}
			// (2) S = X_j H_0 + (what the MAC warps summed for this channel), inverse merge into shared memory
			mbar_wait(s_full, (unsigned) (it & 1));
#pragma unroll
			for (int i = 0; i < 8; ++i) {
				const int k = t + i * T;
				const int n = (k == 0) ? N / 2 : N - k;
				const double2 xk = buf[spad(k)], xn = buf[spad(n)];
				const double2 hk = __ldg(&H0[k]), hn = __ldg(&H0[n]);
				double2 Sk = sbuf[k], Sn = sbuf[n];
				if (k == 0) { Sk.x = fma(xk.x, hk.x, Sk.x); Sk.y = fma(xk.y, hk.y, Sk.y); }   // packed bin: two real products
				else Sk = pipe_cmac(Sk, xk, hk);
				Sn = pipe_cmac(Sn, xn, hn);
				if (k == 0) {
					buf[0] = make_double2(0.5 * (Sk.x + Sk.y), -0.5 * (Sk.x - Sk.y));
					buf[spad(N / 2)] = Sn;   // conj(Z[N/2]) = S[N/2]
				}
				else {
					const double2 e = make_double2(0.5 * (Sk.x + Sn.x), 0.5 * (Sk.y - Sn.y));
					const double2 d = make_double2(0.5 * (Sk.x - Sn.x), 0.5 * (Sk.y + Sn.y));
					const double2 o = cmul(cconj(w[i]), d);
					buf[spad(k)] = make_double2(e.x - o.y, -(e.y + o.x));
					buf[spad(n)] = make_double2(e.x + o.y, -(o.x - e.y));
				}
			}
			__syncwarp();
			if (lane == 0) mbar_arrive(s_empty);   // sbuf may be refilled for the next channel
			TeamSync::sync();
			fft_forward_smem<N, TeamSync>(buf, a.ptw, t);
			// (3) overlap-add, result into the caller's block
			{
				const double scale = 1.0 / N;
				double *yc = a.fake_io ? a.yout + (long) s * N : a.yout + (a.yout_map ? a.yout_map[s] : s);
				const long ys = a.fake_io ? 1 : a.yout_stride;
				double2 c[8];
#pragma unroll
				for (int i = 0; i < 8; ++i) c[i] = carry[t + i * T];
#pragma unroll
				for (int i = 0; i < 8; ++i) {
					const int n = t + i * T;
					const double2 lo = buf[spad(n)], hi = buf[spad(n + N / 2)];
					yc[2L * n * ys] = fma(lo.x, scale, c[i].x);
					yc[(2L * n + 1) * ys] = fma(-lo.y, scale, c[i].y);
					carry[n] = make_double2(hi.x * scale, -hi.y * scale);
				}
			}
			// entry n and n + N/2 of buf are only touched by this thread between the last pass and the next
			// channel's first barrier: no barrier needed here
		}
	}
	else if (warp < T / 32 + Cfg::MAC_WARPS) {
		// ------------------------------------------------------------------------------------------
		// MAC warps
		// ------------------------------------------------------------------------------------------
		const int tm = threadIdx.x - T;
		int stage = 0;
		unsigned phase = 0;
		int it = 0;
		for (int s = blockIdx.x; s < a.n_ch; s += gridDim.x, ++it) {
			mbar_wait(s_empty, (unsigned) ((it & 1) ^ 1));   // the team has taken the previous channel's sums
			for (int c = 0; c < Cfg::NCHUNK; ++c) {
				double2 acc[PER];
#pragma unroll
				for (int i = 0; i < PER; ++i) acc[i] = make_double2(0.0, 0.0);
				const bool dc = (c == 0 && tm == 0);
				for (int p = 1; p < a.pf; ++p) {
					mbar_wait(&full[stage], phase);
					const double2 *Xs = ring + (size_t) stage * 2 * CHUNK, *Hs = Xs + CHUNK;
					double2 x[PER], h[PER];
#pragma unroll
					for (int i = 0; i < PER; ++i) { x[i] = Xs[tm + i * TM]; h[i] = Hs[tm + i * TM]; }
					if (dc) {
						acc[0].x = fma(x[0].x, h[0].x, acc[0].x);
						acc[0].y = fma(x[0].y, h[0].y, acc[0].y);
					}
					else acc[0] = pipe_cmac(acc[0], x[0], h[0]);
#pragma unroll
					for (int i = 1; i < PER; ++i) acc[i] = pipe_cmac(acc[i], x[i], h[i]);
					__syncwarp();
					if (lane == 0) mbar_arrive(&empty[stage]);
					if (++stage == NS) { stage = 0; phase ^= 1; }
				}
				if (a.V) {
					mbar_wait(&full[stage], phase);
					const double2 *Vs = ring + (size_t) stage * 2 * CHUNK;
#pragma unroll
					for (int i = 0; i < PER; ++i) {
						const double2 v = Vs[tm + i * TM];
						acc[i].x += v.x; acc[i].y += v.y;
					}
					__syncwarp();
					if (lane == 0) mbar_arrive(&empty[stage]);
					if (++stage == NS) { stage = 0; phase ^= 1; }
				}
#pragma unroll
				for (int i = 0; i < PER; ++i) sbuf[c * CHUNK + tm + i * TM] = acc[i];
			}
			__syncwarp();
			if (lane == 0) mbar_arrive(s_full);
		}
	}
	else if (lane == 0) {
		// ------------------------------------------------------------------------------------------
		// producer
		// ------------------------------------------------------------------------------------------
		const uint64_t pol = l2_policy_evict_first();
		const bool hint_x = a.evict_first != 0, hint_h = a.evict_first != 0 && a.h_ch_stride != 0;
		int stage = 0;
		unsigned phase = 0;
		constexpr unsigned ROW_BYTES = CHUNK * sizeof(double2);
		for (int s = blockIdx.x; s < a.n_ch; s += gridDim.x) {
			const double2 *fdl = a.fdl + (long) s * a.fdl_ch_stride;
			const double2 *Hc = a.H + (long) s * a.h_ch_stride;
			const double2 *Vc = a.V ? a.V + (long) s * N : nullptr;
			for (int c = 0; c < Cfg::NCHUNK; ++c) {
				for (int p = 1; p < a.pf; ++p) {
					int sl = a.slot - p;
					if (sl < 0) sl += a.fdl_rows;
					mbar_wait(&empty[stage], phase ^ 1);
					mbar_arrive_expect_tx(&full[stage], 2 * ROW_BYTES);
					double2 *dst = ring + (size_t) stage * 2 * CHUNK;
					bulk_g2s(dst, fdl + (long) sl * N + c * CHUNK, ROW_BYTES, &full[stage], hint_x, pol);
					bulk_g2s(dst + CHUNK, Hc + (long) p * N + c * CHUNK, ROW_BYTES, &full[stage], hint_h, pol);
					if (++stage == NS) { stage = 0; phase ^= 1; }
				}
				if (Vc) {
					mbar_wait(&empty[stage], phase ^ 1);
					mbar_arrive_expect_tx(&full[stage], ROW_BYTES);
					bulk_g2s(ring + (size_t) stage * 2 * CHUNK, Vc + c * CHUNK, ROW_BYTES, &full[stage], hint_x, pol);
					if (++stage == NS) { stage = 0; phase ^= 1; }
				}
			}
		}
	}
}

}  // namespace dspb200
