// biquad.cu -- K1: fused, time-parallel biquad cascade (one kernel, one read + one write of the block).
//
// Reference behaviour reproduced (/root/reference): biquad() biquad.h:76-92 (transposed direct form II),
// applied in place per channel by biquad_effect_run[_all] biquad.c:296-315 -- one pass over the block PER
// STAGE there; here a cascade of S stages is ONE operator.
//
// Per stage (biquad.h:79-81):   r = c0 s + m0;  m0' = m1 + c1 s - c3 r;  m1' = c2 s - c4 r
// i.e. the state z = (m0, m1) follows z' = A z + B s with A = [[-c3, 1], [-c4, 0]] (zero-input step).
// The per-sample recurrence is parallelised along TIME as an associative scan over these affine maps:
//   * a CTA owns two adjacent channels and a tile of 4096 frames; a lane owns one chunk of L = 16 frames of one
//     channel (a warp = 16 consecutive chunks x 2 channels, channel fastest), the chunk's samples live in
//     registers through all S stages;
//   * per stage: (1) every lane needs the end state e its chunk would reach from a ZERO state; that is the
//     dot product e = sum_i A^(L-1-i) B s_i of the chunk with a per-stage table (2 L independent FMAs instead
//     of a serial recurrence; table built on the host in long double); (2) warp-level inclusive scan of the
//     states by shuffles -- all chunks share the matrix A^L, so round r is  s += A^(L 2^r) shfl_up(s, 2 * 2^r)
//     with the powers taken on the host in long double; (3) warp totals are chained through shared memory
//     with A^(16 L), every lane turns the carry-in of its warp into its own start state with the binary
//     expansion of its chunk index; (4) the lane re-runs the reference's own recurrence from that TRUE start
//     state -- so every output sample is produced by exactly the reference's arithmetic, started from a
//     state that differs from the sequential one only by rounding in the scan (|eig A| < 1);
//   * longer calls loop over tiles, carrying the per-stage state through shared memory.
// A stage that does not act on a channel carries the identity section {1,0,0,0,0}.
// Global accesses: one load instruction of a warp touches 16 rows (frames) and uses 16 contiguous bytes of
// each -- half of every 32-byte sector, the neighbouring CTA uses the other half at about the same time, so
// HBM sees every sector once (L2 absorbs the rest).  (One channel per CTA = 32 sectors per request, measured
// 16.4 us for a 1-stage pass over a config-2 block against 10.3 us for pairs; four channels per CTA halves
// the grid and loses more than it gains at 256 channels.)  Algorithmic bytes: 16 per sample.
#include "common.cuh"
#include "ops.h"

namespace dspb200 {

constexpr int BQ_L = 16;           // frames per lane
constexpr int BQ_MAX_STAGES = 16;  // stages fused into one operator
constexpr int BQ_NPOW = 6;         // A^(L 2^r), r = 0..5
// per channel and stage: 5 coefficients (+1 pad), BQ_NPOW 2x2 matrices, and G[i] = A^(L-1-i) B (2 x L), contiguous;
// every piece starts on a 16-byte boundary so that the kernel reads the table with 128-bit shared loads
constexpr int BQ_OFF_P = 6, BQ_OFF_G = BQ_OFF_P + 4 * BQ_NPOW;
constexpr int BQ_TBL = BQ_OFF_G + 2 * BQ_L;
static_assert(BQ_TBL % 2 == 0 && BQ_OFF_G % 2 == 0, "table pieces must stay 16-byte aligned");

struct M2 { double a, b, c, d; };  // [[a, b], [c, d]]

__device__ __forceinline__ double2 m2_apply(const M2 &m, double2 v)
{
	return make_double2(fma(m.a, v.x, m.b * v.y), fma(m.c, v.x, m.d * v.y));
}

// CH adjacent channels share a CTA; a warp's lanes are (32 / CH consecutive chunks) x (CH channels), channel
// fastest, so that one load instruction touches 32 / CH rows and uses CH x 8 contiguous bytes of each.
// SPLIT: two CTAs per channel pair, each on one half (2048 frames) of a 4096-frame tile, 8 warps each -- two of
// them fit an SM, so that one CTA's barriers, table loads and block I/O hide behind the other's arithmetic.  The
// second half needs, per stage, the state the first half ends in: the first half publishes it (state, fence, flag)
// as soon as its stage is done and the second half picks it up right before that stage's carry step, i.e. it runs
// one stage behind (decoupled look-back along time; CTAs of the first half have the lower block indices, so they
// are always scheduled before the CTAs that wait for them).
template <int CH, bool SPLIT = false>
struct BqCfg {
	static_assert(CH == 1 || CH == 2 || CH == 4, "1, 2 or 4 channels per CTA");
	static_assert(!SPLIT || CH == 2, "the split form is built for channel pairs");
	static constexpr int CPW = 32 / CH;                         // chunks per warp
	static constexpr int LOGW = (CH == 1) ? 5 : (CH == 2) ? 4 : 3;
	static constexpr int WARPS = SPLIT ? 8 : (CH == 1) ? 8 : 16;
	static constexpr int THREADS = 32 * WARPS;
	static constexpr long TILE = (long) BQ_L * CPW * WARPS;     // frames per tile
};

struct BqSplitArgs {
	double *xstate;   // [C][S][2]: state at the end of the first half, per stage
	int *flag;        // [C][S]: == epoch once xstate[c][st] is valid
	int epoch;
};

// tbl: [C][S][BQ_TBL] (see BQ_TBL), zstate: [C][S][2]
template <int CH, bool SPLIT = false>
__global__ void __launch_bounds__(BqCfg<CH, SPLIT>::THREADS, SPLIT ? 2 : 1) k_bq_cascade(const double *in, double *out, const double *__restrict__ tbl,
                                                                  double *zstate, int C, int S, long frames, BqSplitArgs sp)
{
	using Cfg = BqCfg<CH, SPLIT>;
	const int half = SPLIT ? (int) blockIdx.y : 0;
	if (SPLIT && half == 1 && frames <= Cfg::TILE) return;   // nothing in the second half (the whole CTA leaves before any barrier)
	const bool second_exists = SPLIT && frames > Cfg::TILE;
	constexpr int LOGW = Cfg::LOGW;
	__shared__ double2 tot[Cfg::WARPS][CH];
	__shared__ double2 carry[BQ_MAX_STAGES][CH];
	extern __shared__ __align__(16) double stbl[];   // [CH][S][BQ_TBL]: these channels' tables
	const int c0 = blockIdx.x * CH;
	const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	const int ch = lane & (CH - 1), jl = lane / CH;     // channel within the CTA, chunk within the warp
	const int c = c0 + ch;
	const bool have = c < C;
	const int nch = (C - c0 < CH) ? C - c0 : CH;
	for (int i = threadIdx.x; i < nch * S * BQ_TBL; i += blockDim.x) stbl[i] = tbl[(long) c0 * S * BQ_TBL + i];
	if (threadIdx.x < S * CH) {
		const int st = threadIdx.x / CH, k = threadIdx.x % CH;
		// (the second half's entries are filled stage by stage from the first half's published states)
		carry[st][k] = (k < nch && half == 0) ? make_double2(zstate[((long) (c0 + k) * S + st) * 2], zstate[((long) (c0 + k) * S + st) * 2 + 1])
		                                      : make_double2(0.0, 0.0);
	}
	__syncthreads();
	const double *mytbl = stbl + (have ? ch : 0) * S * BQ_TBL;

	for (long base = SPLIT ? half * Cfg::TILE : 0; base < frames; base += SPLIT ? frames : Cfg::TILE) {
		const long f0 = base + ((long) w * Cfg::CPW + jl) * BQ_L;
		const long rem = frames - f0;
		const int nv = (rem <= 0 || !have) ? 0 : (rem < BQ_L ? (int) rem : BQ_L);
		// the last valid chunk of this tile hands its end state to the next tile / the next call
		const bool last_chunk = nv > 0 && (rem <= BQ_L || f0 + BQ_L >= base + Cfg::TILE);
		double y[BQ_L];
#pragma unroll
		for (int i = 0; i < BQ_L; ++i) y[i] = (i < nv) ? in[(f0 + i) * C + c] : 0.0;

		for (int st = 0; st < S; ++st) {
			const double2 *t = reinterpret_cast<const double2 *>(mytbl + st * BQ_TBL);
			const double2 t0 = t[0], t1 = t[1], t2 = t[2];
			const double c0_ = t0.x, c1 = t0.y, c2 = t1.x, c3 = t1.y, c4 = t2.x;
			M2 P[LOGW + 1];
#pragma unroll
			for (int r = 0; r <= LOGW; ++r) {
				const double2 lo = t[BQ_OFF_P / 2 + 2 * r], hi = t[BQ_OFF_P / 2 + 2 * r + 1];
				P[r] = M2{ lo.x, lo.y, hi.x, hi.y };
			}
			// (1) zero-state end state of a FULL chunk as a dot product (a partial last chunk is nobody's predecessor:
			//     its value is never used; missing samples are zeros)
			const double2 *G = t + BQ_OFF_G / 2;
			double ex0 = 0.0, ex1 = 0.0, ey0 = 0.0, ey1 = 0.0;
#pragma unroll
			for (int i = 0; i < BQ_L; i += 2) {
				const double2 g0 = G[i], g1 = G[i + 1];
				ex0 = fma(g0.x, y[i], ex0);
				ey0 = fma(g0.y, y[i], ey0);
				ex1 = fma(g1.x, y[i + 1], ex1);
				ey1 = fma(g1.y, y[i + 1], ey1);
			}
			// (2) inclusive scan over the chunks of this warp (uniform matrix A^L per chunk); lanes CH apart
			double2 sc = make_double2(ex0 + ex1, ey0 + ey1);
#pragma unroll
			for (int r = 0; r < LOGW; ++r) {
				const double vx = __shfl_up_sync(0xffffffffu, sc.x, CH << r), vy = __shfl_up_sync(0xffffffffu, sc.y, CH << r);
				if (jl >= (1 << r)) {
					const double2 u = m2_apply(P[r], make_double2(vx, vy));
					sc.x += u.x;
					sc.y += u.y;
				}
			}
			// exclusive: state at the start of this lane's chunk if the warp started from zero
			double2 z = make_double2(__shfl_up_sync(0xffffffffu, sc.x, CH), __shfl_up_sync(0xffffffffu, sc.y, CH));
			if (jl == 0) z = make_double2(0.0, 0.0);
			if (jl == Cfg::CPW - 1) tot[w][ch] = sc;
			if (SPLIT && half == 1 && w == 0 && jl == 0 && have) {
				// the state the first half left this stage in: wait for its flag, then read it past L1
				const long o = (long) c * S + st;
				volatile int *fl = sp.flag + o;
				while (*fl != sp.epoch) { }
				__threadfence();
				carry[st][ch] = make_double2(__ldcg(&sp.xstate[o * 2]), __ldcg(&sp.xstate[o * 2 + 1]));
			}
			__syncthreads();
			// (3) carry into this warp, then into this lane
			double2 cin = carry[st][ch];
			for (int k = 0; k < w; ++k) {
				const double2 u = m2_apply(P[LOGW], cin);
				cin = make_double2(u.x + tot[k][ch].x, u.y + tot[k][ch].y);
			}
#pragma unroll
			for (int r = 0; r < LOGW; ++r)
				if ((jl >> r) & 1) cin = m2_apply(P[r], cin);
			double m0 = z.x + cin.x, m1 = z.y + cin.y;
			// (4) the real run from the true start state
#pragma unroll
			for (int i = 0; i < BQ_L; ++i) {
				if (i < nv) {
					const double s = y[i];
					const double r = c0_ * s + m0;
					m0 = m1 + c1 * s - c3 * r;
					m1 = c2 * s - c4 * r;
					y[i] = r;
				}
			}
			__syncthreads();   // everyone has read tot[] and carry[st]
			if (last_chunk) {
				carry[st][ch] = make_double2(m0, m1);
				if (SPLIT && half == 0 && second_exists) {
					const long o = (long) c * S + st;
					sp.xstate[o * 2] = m0;
					sp.xstate[o * 2 + 1] = m1;
					__threadfence();
					*reinterpret_cast<volatile int *>(sp.flag + o) = sp.epoch;
				}
			}
		}
#pragma unroll
		for (int i = 0; i < BQ_L; ++i)
			if (i < nv) out[(f0 + i) * C + c] = y[i];
		__syncthreads();   // carry[] complete before the next tile reads it
	}
	if (threadIdx.x < S * CH && (!SPLIT || half == 1 || !second_exists)) {   // the CTA that holds the end of the block
		const int st = threadIdx.x / CH, k = threadIdx.x % CH;
		if (k < nch) {
			zstate[((long) (c0 + k) * S + st) * 2] = carry[st][k].x;
			zstate[((long) (c0 + k) * S + st) * 2 + 1] = carry[st][k].y;
		}
	}
}

template <int CH>
static int bq_launch(const double *in, double *out, const double *tbl, double *zstate, int C, int S, long frames, cudaStream_t st)
{
	LAUNCH((k_bq_cascade<CH>), (C + CH - 1) / CH, BqCfg<CH>::THREADS, (size_t) CH * S * BQ_TBL * sizeof(double), st, in, out, tbl, zstate, C, S, frames, BqSplitArgs{});
	return 0;
}

// split form: one launch per 4096 frames, two CTAs per channel pair
static int bq_launch_split(const double *in, double *out, const double *tbl, double *zstate, int C, int S, long frames, cudaStream_t st, BqSplitArgs sp, int *epoch)
{
	using Cfg = BqCfg<2, true>;
	const long span = 2 * Cfg::TILE;
	for (long base = 0; base < frames; base += span) {
		const long n = (frames - base < span) ? frames - base : span;
		sp.epoch = ++*epoch;
		LAUNCH((k_bq_cascade<2, true>), dim3((C + 1) / 2, 2), Cfg::THREADS, (size_t) 2 * S * BQ_TBL * sizeof(double), st, in + base * C, out + base * C, tbl,
		       zstate, C, S, n, sp);
	}
	return 0;
}

struct BiquadOp : Op {
	int S = 0;
	std::vector<double> h_coefs;   // [S][C][5] as handed in (kept so neighbouring cascades can be fused)
	double *d_tbl = nullptr, *d_zstate = nullptr;
	double *d_xstate = nullptr;   // split form: per-stage state between the two halves of a tile
	int *d_flag = nullptr;
	int epoch = 0;

	const char *name() const override { return "biquad"; }
	std::string describe() const override
	{
		char buf[96];
		snprintf(buf, sizeof(buf), "{\"op\":\"biquad\",\"stages\":%d}", S);
		return buf;
	}
	~BiquadOp() override { dev_free(d_tbl); dev_free(d_zstate); dev_free(d_xstate); dev_free(d_flag); }

	void reset(cudaStream_t st) override
	{
		cudaMemsetAsync(d_zstate, 0, (size_t) channels * S * 2 * sizeof(double), st);
	}

	long run(long frames, const double *in, double *out, cudaStream_t st) override
	{
		if (frames <= 0) return 0;
		ProfScope prof("biquad", st);
		// channels per CTA: wider rows per load as long as the grid still covers the SMs
		static const int force = getenv("DSP_B200_BQ_CH") ? atoi(getenv("DSP_B200_BQ_CH")) : 0;
		// Always channel pairs: the scan's association order depends on the chunks per warp, and a channel's
		// result must not depend on how the chain was sharded (tests: sharded == unsharded, bit for bit).
		const int chp = force ? force : 2;
		// Split form (two 8-warp CTAs per channel pair with a per-stage look-back of the first half's state): measured
		// SLOWER than one 16-warp CTA on full tiles (35.3 vs 26.7 us on config 2: the second half trails by a stage and
		// the two CTAs of a pair contend for the same SM's FP64 pipe instead of hiding each other) -- but calls of at
		// most 2048 frames have no second half, and 8 warps on such a block beat 16 half-idle ones (16.4 vs 20.6 us at 512
		// frames).  DSP_B200_BQ_SPLIT=1 forces it everywhere, =0 nowhere.
		static const int split_env = getenv("DSP_B200_BQ_SPLIT") ? atoi(getenv("DSP_B200_BQ_SPLIT")) : -1;
		const bool split = (split_env >= 0) ? split_env != 0 : frames <= BqCfg<2, true>::TILE;
		int rc;
		if (chp == 2 && split) {
			BqSplitArgs sp = { d_xstate, d_flag, 0 };
			rc = bq_launch_split(in, out, d_tbl, d_zstate, channels, S, frames, st, sp, &epoch);
		}
		else if (chp == 4) rc = bq_launch<4>(in, out, d_tbl, d_zstate, channels, S, frames, st);
		else if (chp == 2) rc = bq_launch<2>(in, out, d_tbl, d_zstate, channels, S, frames, st);
		else rc = bq_launch<1>(in, out, d_tbl, d_zstate, channels, S, frames, st);
		return rc ? -1 : frames;
	}
};

Op *make_biquad_op(int slab_channels, int fs, int n_stages, const double *coefs)
{
	if (n_stages < 1 || n_stages > BQ_MAX_STAGES) {
		set_error("biquad: %d stages per operator unsupported (1..%d)", n_stages, BQ_MAX_STAGES);
		return nullptr;
	}
	std::unique_ptr<BiquadOp> op(new BiquadOp());
	const int C = slab_channels, S = n_stages;
	op->channels = C; op->fs_in = op->fs_out = fs; op->S = S;
	op->h_coefs.assign(coefs, coefs + (size_t) S * C * 5);

	// coefs arrive as [stage][channel][5]; the kernel wants one contiguous table per channel (a CTA reads one channel)
	std::vector<double> tbl((size_t) C * S * BQ_TBL);
	for (int c = 0; c < C; ++c) {
		for (int st = 0; st < S; ++st) {
			const double *cf = &coefs[((size_t) st * C + c) * 5];
			double *t = &tbl[((size_t) c * S + st) * BQ_TBL];
			for (int k = 0; k < 5; ++k) t[k] = cf[k];
			// A = [[-c3, 1], [-c4, 0]], B = (c1 - c3 c0, c2 - c4 c0): z' = A z + B s   (biquad.h:79-81 with r = c0 s + m0)
			const long double A[4] = { -(long double) cf[3], 1.0L, -(long double) cf[4], 0.0L };
			const long double Bv[2] = { (long double) cf[1] - (long double) cf[3] * cf[0], (long double) cf[2] - (long double) cf[4] * cf[0] };
			// G[i] = A^(L-1-i) B, i = L-1 .. 0, and A^L on the way
			long double g[2] = { Bv[0], Bv[1] };
			for (int i = BQ_L - 1; i >= 0; --i) {
				t[BQ_OFF_G + 2 * i] = (double) g[0];
				t[BQ_OFF_G + 2 * i + 1] = (double) g[1];
				const long double n0 = A[0] * g[0] + A[1] * g[1], n1 = A[2] * g[0] + A[3] * g[1];
				g[0] = n0; g[1] = n1;
			}
			long double a[4] = { 1.0L, 0.0L, 0.0L, 1.0L };
			for (int i = 0; i < BQ_L; ++i) {
				const long double u[4] = { A[0] * a[0] + A[1] * a[2], A[0] * a[1] + A[1] * a[3], A[2] * a[0] + A[3] * a[2], A[2] * a[1] + A[3] * a[3] };
				for (int k = 0; k < 4; ++k) a[k] = u[k];
			}
			// A^(L 2^r) by repeated squaring
			for (int r = 0; r < BQ_NPOW; ++r) {
				for (int k = 0; k < 4; ++k) t[BQ_OFF_P + 4 * r + k] = (double) a[k];
				const long double u[4] = { a[0] * a[0] + a[1] * a[2], a[0] * a[1] + a[1] * a[3], a[2] * a[0] + a[3] * a[2], a[2] * a[1] + a[3] * a[3] };
				for (int k = 0; k < 4; ++k) a[k] = u[k];
			}
		}
	}
	op->d_tbl = dev_alloc<double>(tbl.size(), false);
	op->d_zstate = dev_alloc<double>((size_t) C * S * 2, true);
	op->d_xstate = dev_alloc<double>((size_t) C * S * 2, true);
	op->d_flag = dev_alloc<int>((size_t) C * S, true);
	if (!op->d_tbl || !op->d_zstate || !op->d_xstate || !op->d_flag) return nullptr;
	CUDA_TRY(cudaMemcpy(op->d_tbl, tbl.data(), tbl.size() * sizeof(double), cudaMemcpyHostToDevice), return nullptr);
	return op.release();
}

// Two adjacent cascades over the same slab become one operator (one pass over the block).
// Only legal before the first run()/after reset.
Op *fuse_biquad_ops(Op *a, Op *b)
{
	BiquadOp *x = dynamic_cast<BiquadOp *>(a), *y = dynamic_cast<BiquadOp *>(b);
	if (!x || !y || x->channels != y->channels || x->fs_in != y->fs_in || x->S + y->S > BQ_MAX_STAGES) return nullptr;
	std::vector<double> all(x->h_coefs);
	all.insert(all.end(), y->h_coefs.begin(), y->h_coefs.end());
	return make_biquad_op(x->channels, x->fs_in, x->S + y->S, all.data());
}

}  // namespace dspb200
