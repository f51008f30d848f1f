// biquad.cu -- K1: fused, time-parallel biquad cascade.
//
// Reference behaviour reproduced (/root/reference): biquad() biquad.h:76-92 (transposed direct
// form II), applied in place per channel by biquad_effect_run[_all] biquad.c:296-315, one pass
// over the block PER STAGE.  Here a cascade of S stages is ONE operator and one read + one write
// of the block.
//
// Per stage (biquad.h:79-81):   r = c0 s + m0;  m0' = m1 + c1 s - c3 r;  m1' = c2 s - c4 r
// The cascade is a linear system with state z = (m0,m1 of every stage) in R^D, D = 2S.  Time is
// cut into chunks of L frames; thread (channel c, chunk j):
//   k_bq_local  runs the cascade over its chunk from ZERO state, keeps only the end state b_j
//   k_bq_scan   one warp per channel: z_{j+1} = M_c z_j + b_j, z_0 = carried state, M_c = A_c^L
//               (A_c = zero-input transition of the cascade, powers taken on the host)
//   k_bq_apply  re-runs the cascade over the chunk from its TRUE start state z_j and stores
// so every output sample is produced by the reference's own recurrence, started from a state
// that differs from the sequential one only by rounding in the scan (|eig A| < 1).
// Lanes run along channels: every global access is a coalesced row of the interleaved block.
#include "common.cuh"
#include "ops.h"

namespace dspb200 {

constexpr int BQ_L = 32;          // frames per chunk
constexpr int BQ_MAX_STAGES = 16; // D <= 32: one lane per state component in the scan

// coef layout: [stage][5][C]; state layouts: b, zin: [c][chunk][D]; zstate: [c][D]
template <bool APPLY>
__global__ void __launch_bounds__(128) k_bq_chunks(const double *in, double *out, const double *__restrict__ coef,
                                                   const double *zin, double *bout, double *zstate,
                                                   int C, int S, long frames, int n_chunks)
{
	const long idx = (long) blockIdx.x * blockDim.x + threadIdx.x;
	const int c = (int) (idx % C);
	const long j = idx / C;
	if (j >= n_chunks) return;
	const int D = 2 * S;
	const long f0 = j * BQ_L;
	const int nv = (int) ((frames - f0 < BQ_L) ? frames - f0 : BQ_L);

	double y[BQ_L];
#pragma unroll
	for (int i = 0; i < BQ_L; ++i) y[i] = (i < nv) ? in[(f0 + i) * C + c] : 0.0;

	const double *z0 = APPLY ? zin + ((long) c * n_chunks + j) * D : nullptr;
	double *zend = APPLY ? zstate + (long) c * D : bout + ((long) c * n_chunks + j) * D;
	const bool store_end = APPLY ? (j == n_chunks - 1) : true;

	for (int st = 0; st < S; ++st) {
		const double *cf = coef + (long) st * 5 * C + c;
		const double c0 = cf[0], c1 = cf[C], c2 = cf[2 * C], c3 = cf[3 * C], c4 = cf[4 * C];
		double m0 = 0.0, m1 = 0.0;
		if (APPLY) { m0 = z0[2 * st]; m1 = z0[2 * st + 1]; }
#pragma unroll
		for (int i = 0; i < BQ_L; ++i) {
			if (i < nv) {
				const double s = y[i];
				const double r = c0 * s + m0;
				m0 = m1 + c1 * s - c3 * r;
				m1 = c2 * s - c4 * r;
				y[i] = r;
			}
		}
		if (store_end) { zend[2 * st] = m0; zend[2 * st + 1] = m1; }
	}
	if (APPLY) {
#pragma unroll
		for (int i = 0; i < BQ_L; ++i)
			if (i < nv) out[(f0 + i) * C + c] = y[i];
	}
}

// one warp per channel; lane d owns state component d and row d of M_c
__global__ void __launch_bounds__(128) k_bq_scan(const double *__restrict__ M, const double *__restrict__ b, double *__restrict__ zin,
                                                 const double *__restrict__ zstate, int C, int D, int n_chunks)
{
	const int warp = (blockIdx.x * blockDim.x + threadIdx.x) / 32;
	const int lane = threadIdx.x & 31;
	if (warp >= C) return;
	const int c = warp;
	double row[2 * BQ_MAX_STAGES];
#pragma unroll
	for (int e = 0; e < 2 * BQ_MAX_STAGES; ++e)
		row[e] = (lane < D && e < D) ? M[((long) c * D + lane) * D + e] : 0.0;
	double z = (lane < D) ? zstate[(long) c * D + lane] : 0.0;
	const double *bc = b + (long) c * n_chunks * D;
	double *zc = zin + (long) c * n_chunks * D;
	double bnext = (lane < D) ? bc[lane] : 0.0;
	for (int j = 0; j < n_chunks; ++j) {
		if (lane < D) zc[(long) j * D + lane] = z;
		if (j + 1 == n_chunks) break;
		const double bj = bnext;
		if (lane < D && j + 2 < n_chunks + 1) bnext = bc[(long) (j + 1) * D + lane];
		double acc = bj;
#pragma unroll
		for (int e = 0; e < 2 * BQ_MAX_STAGES; ++e) {
			const double ze = __shfl_sync(0xffffffffu, z, e);
			if (e < D) acc = fma(row[e], ze, acc);
		}
		z = acc;
	}
}

struct BiquadOp : Op {
	int S = 0, D = 0;
	std::vector<double> h_coefs;   // [S][C][5] as handed in (kept so neighbouring cascades can be fused)
	double *d_coef = nullptr, *d_M = nullptr, *d_zstate = nullptr, *d_b = nullptr, *d_zin = nullptr;
	long chunk_cap = 0;

	const char *name() const override { return "biquad"; }
	~BiquadOp() override { dev_free(d_coef); dev_free(d_M); dev_free(d_zstate); dev_free(d_b); dev_free(d_zin); }

	void reset(cudaStream_t st) override
	{
		cudaMemsetAsync(d_zstate, 0, (size_t) channels * D * sizeof(double), st);
	}

	long run(long frames, const double *in, double *out, cudaStream_t st) override
	{
		if (frames <= 0) return 0;
		const int C = channels;
		const long n_chunks = (frames + BQ_L - 1) / BQ_L;
		const long threads = n_chunks * C;
		ProfScope prof("biquad", st);
		if (n_chunks == 1) {
			// the chunk starts from the carried state itself: zin == zstate (layout [c][1][D])
			LAUNCH(k_bq_chunks<true>, ceil_div(threads, 128), 128, 0, st, in, out, d_coef, d_zstate, nullptr, d_zstate, C, S, frames, 1);
			return frames;
		}
		if (n_chunks > chunk_cap) {
			dev_free(d_b); dev_free(d_zin);
			d_b = dev_alloc<double>((size_t) n_chunks * C * D, false);
			d_zin = dev_alloc<double>((size_t) n_chunks * C * D, false);
			if (!d_b || !d_zin) return -1;
			chunk_cap = n_chunks;
		}
		LAUNCH(k_bq_chunks<false>, ceil_div(threads, 128), 128, 0, st, in, nullptr, d_coef, nullptr, d_b, nullptr, C, S, frames, (int) n_chunks);
		LAUNCH(k_bq_scan, ceil_div((long) C * 32, 128), 128, 0, st, d_M, d_b, d_zin, d_zstate, C, D, (int) n_chunks);
		LAUNCH(k_bq_chunks<true>, ceil_div(threads, 128), 128, 0, st, in, out, d_coef, d_zin, nullptr, d_zstate, C, S, frames, (int) n_chunks);
		return frames;
	}
};

// zero-input step of the cascade (same arithmetic as biquad.h:79-81 with s = 0 at stage 0)
static void cascade_zero_input_step(int S, const double *cf /*[S][5]*/, long double *z /*[2S]*/)
{
	long double s = 0.0L;
	for (int st = 0; st < S; ++st) {
		const long double c0 = cf[st * 5 + 0], c1 = cf[st * 5 + 1], c2 = cf[st * 5 + 2], c3 = cf[st * 5 + 3], c4 = cf[st * 5 + 4];
		const long double r = c0 * s + z[2 * st];
		z[2 * st] = z[2 * st + 1] + c1 * s - c3 * r;
		z[2 * st + 1] = c2 * s - c4 * r;
		s = r;
	}
}

Op *make_biquad_op(int slab_channels, int fs, int n_stages, const double *coefs)
{
	if (n_stages < 1 || n_stages > BQ_MAX_STAGES) {
		set_error("biquad: %d stages per operator unsupported (1..%d)", n_stages, BQ_MAX_STAGES);
		return nullptr;
	}
	std::unique_ptr<BiquadOp> op(new BiquadOp());
	const int C = slab_channels, S = n_stages, D = 2 * n_stages;
	op->channels = C; op->fs_in = op->fs_out = fs; op->S = S; op->D = D;
	op->h_coefs.assign(coefs, coefs + (size_t) S * C * 5);

	// coefs arrive as [stage][channel][5]; device wants [stage][5][channel]
	std::vector<double> dev_coef((size_t) S * 5 * C), M((size_t) C * D * D);
	std::vector<double> cf((size_t) S * 5);
	std::vector<long double> z(D);
	for (int c = 0; c < C; ++c) {
		for (int st = 0; st < S; ++st)
			for (int k = 0; k < 5; ++k) {
				const double v = coefs[((size_t) st * C + c) * 5 + k];
				dev_coef[((size_t) st * 5 + k) * C + c] = v;
				cf[st * 5 + k] = v;
			}
		// column e of M_c = A_c^L e_e
		for (int e = 0; e < D; ++e) {
			for (int d = 0; d < D; ++d) z[d] = (d == e) ? 1.0L : 0.0L;
			for (int i = 0; i < BQ_L; ++i) cascade_zero_input_step(S, cf.data(), z.data());
			for (int d = 0; d < D; ++d) M[((size_t) c * D + d) * D + e] = (double) z[d];
		}
	}
	op->d_coef = dev_alloc<double>(dev_coef.size(), false);
	op->d_M = dev_alloc<double>(M.size(), false);
	op->d_zstate = dev_alloc<double>((size_t) C * D, true);
	if (!op->d_coef || !op->d_M || !op->d_zstate) return nullptr;
	CUDA_TRY(cudaMemcpy(op->d_coef, dev_coef.data(), dev_coef.size() * sizeof(double), cudaMemcpyHostToDevice), return nullptr);
	CUDA_TRY(cudaMemcpy(op->d_M, M.data(), M.size() * sizeof(double), cudaMemcpyHostToDevice), return nullptr);
	return op.release();
}

// Two adjacent cascades over the same slab become one operator (one pass over the block) when
// the state still fits one lane per component.  Only legal before the first run()/after reset.
Op *fuse_biquad_ops(Op *a, Op *b)
{
	BiquadOp *x = dynamic_cast<BiquadOp *>(a), *y = dynamic_cast<BiquadOp *>(b);
	if (!x || !y || x->channels != y->channels || x->fs_in != y->fs_in || x->S + y->S > BQ_MAX_STAGES) return nullptr;
	std::vector<double> all(x->h_coefs);
	all.insert(all.end(), y->h_coefs.begin(), y->h_coefs.end());
	return make_biquad_op(x->channels, x->fs_in, x->S + y->S, all.data());
}

}  // namespace dspb200
