// design.cu -- init-time host arithmetic that must agree with the reference's libm results:
// the biquad coefficient design (biquad.c:111-294, RBJ Audio-EQ-Cookbook forms plus the first
// order and Linkwitz-transform sections).  Pure host code; lives in the library so that the C
// shim (shim/*.c) and the Python mirror (dsp_b200/effects.py) share one implementation.
#include "common.cuh"
#include "../../include/dsp_b200.h"
#include <cmath>

namespace {

struct Sos { double b0, b1, b2, a0, a1, a2; };

// Second-order cookbook sections share the denominator 1 + alpha, -2 cos w0, 1 - alpha.
inline Sos rbj(double b0, double b1, double b2, double alpha, double cos_w0)
{
	return Sos{ b0, b1, b2, 1.0 + alpha, -2.0 * cos_w0, 1.0 - alpha };
}

}  // namespace

extern "C" int dspb200_biquad_design(int type, double fs, double arg0, double arg1, double arg2, double arg3,
                                     int width_type, double c5[5])
{
	Sos q = { 1.0, 0.0, 0.0, 1.0, 0.0, 0.0 };
	if (type == DSPB200_BQ_LOWPASS_TRANSFORM || type == DSPB200_BQ_HIGHPASS_TRANSFORM) {
		// biquad.c:114-130: zeros at (fz, qz) replaced by poles at (fp, qp)
		const bool lp = (type == DSPB200_BQ_LOWPASS_TRANSFORM);
		const double w0z = 2 * M_PI * arg0 / fs, w0p = 2 * M_PI * arg2 / fs;
		const double cz = cos(w0z), cp = cos(w0p);
		const double az = sin(w0z) / (2.0 * arg1), ap = sin(w0p) / (2.0 * arg3);
		const double kz = lp ? 2.0 / (1.0 - cz) : 2.0 / (1.0 + cz);
		const double kp = lp ? 2.0 / (1.0 - cp) : 2.0 / (1.0 + cp);
		q = Sos{ (1.0 + az) * kz, (-2.0 * cz) * kz, (1.0 - az) * kz, (1.0 + ap) * kp, (-2.0 * cp) * kp, (1.0 - ap) * kp };
	}
	else {
		double f0 = arg0, width = arg1;
		const double gain = arg2;
		if (width_type == DSPB200_BQ_WIDTH_SLOPE_DB) {
			// biquad.c:136-143: dB/octave slope, corner moved to the -3 dB-ish point
			width_type = DSPB200_BQ_WIDTH_SLOPE;
			width /= 12.0;
			if (type == DSPB200_BQ_LOWSHELF) f0 *= pow(10.0, fabs(gain) / 80.0 / width);
			else if (type == DSPB200_BQ_HIGHSHELF) f0 /= pow(10.0, fabs(gain) / 80.0 / width);
		}
		const double a = pow(10.0, gain / 40.0);
		const double w0 = 2 * M_PI * f0 / fs;
		const double sn = sin(w0), cs = cos(w0);
		double alpha;
		switch (width_type) {
		case DSPB200_BQ_WIDTH_SLOPE: alpha = sn / 2.0 * sqrt((a + 1.0 / a) * (1.0 / width - 1.0) + 2.0); break;
		case DSPB200_BQ_WIDTH_BW_OCT: alpha = sn * sinh(M_LN2 / 2 * width * w0 / sn); break;
		case DSPB200_BQ_WIDTH_BW_HZ: alpha = sn / (2.0 * f0 / width); break;
		default: alpha = sn / (2.0 * width); break;
		}
		const double c1 = 1.0 + cs;
		switch (type) {
		case DSPB200_BQ_LOWPASS_1:
			q = Sos{ sn, sn, 0.0, sn + c1, sn - c1, 0.0 };
			break;
		case DSPB200_BQ_HIGHPASS_1:
			q = Sos{ c1, -c1, 0.0, sn + c1, sn - c1, 0.0 };
			break;
		case DSPB200_BQ_ALLPASS_1:
			q = Sos{ sn - c1, sn + c1, 0.0, sn + c1, sn - c1, 0.0 };
			break;
		case DSPB200_BQ_LOWSHELF_1:
			q = Sos{ a * sn + c1, a * sn - c1, 0.0, sn / a + c1, sn / a - c1, 0.0 };
			break;
		case DSPB200_BQ_HIGHSHELF_1:
			q = Sos{ sn + c1 * a, sn - c1 * a, 0.0, sn + c1 / a, sn - c1 / a, 0.0 };
			break;
		case DSPB200_BQ_LOWPASS_1P: {
			const double c = 1.0 - cs;
			const double b0 = -c + sqrt(c * c + 2.0 * c);
			q = Sos{ b0, 0.0, 0.0, 1.0, -1.0 + b0, 0.0 };
			break;
		}
		case DSPB200_BQ_LOWPASS: {
			const double h = (1.0 - cs) / 2.0;
			q = rbj(h, 1.0 - cs, h, alpha, cs);
			break;
		}
		case DSPB200_BQ_HIGHPASS: {
			const double h = (1.0 + cs) / 2.0;
			q = rbj(h, -(1.0 + cs), h, alpha, cs);
			break;
		}
		case DSPB200_BQ_BANDPASS_SKIRT: q = rbj(sn / 2.0, 0.0, -(sn / 2.0), alpha, cs); break;
		case DSPB200_BQ_BANDPASS_PEAK: q = rbj(alpha, 0.0, -alpha, alpha, cs); break;
		case DSPB200_BQ_NOTCH: q = rbj(1.0, -2.0 * cs, 1.0, alpha, cs); break;
		case DSPB200_BQ_ALLPASS: q = Sos{ 1.0 - alpha, -2.0 * cs, 1.0 + alpha, 1.0 + alpha, -2.0 * cs, 1.0 - alpha }; break;
		case DSPB200_BQ_PEAK:
			q = Sos{ 1.0 + alpha * a, -2.0 * cs, 1.0 - alpha * a, 1.0 + alpha / a, -2.0 * cs, 1.0 - alpha / a };
			break;
		case DSPB200_BQ_LOWSHELF: {
			const double c = 2.0 * sqrt(a) * alpha, ap1 = a + 1.0, am1 = a - 1.0;
			q = Sos{ a * (ap1 - am1 * cs + c), 2.0 * a * (am1 - ap1 * cs), a * (ap1 - am1 * cs - c),
			         ap1 + am1 * cs + c, -2.0 * (am1 + ap1 * cs), ap1 + am1 * cs - c };
			break;
		}
		case DSPB200_BQ_HIGHSHELF: {
			const double c = 2.0 * sqrt(a) * alpha, ap1 = a + 1.0, am1 = a - 1.0;
			q = Sos{ a * (ap1 + am1 * cs + c), -2.0 * a * (am1 + ap1 * cs), a * (ap1 + am1 * cs - c),
			         ap1 - am1 * cs + c, 2.0 * (am1 - ap1 * cs), ap1 - am1 * cs - c };
			break;
		}
		default:
			dspb200::set_error("biquad_design: unknown type %d", type);
			return -1;
		}
	}
	// biquad.c:91-98: normalise by a0
	c5[0] = q.b0 / q.a0; c5[1] = q.b1 / q.a0; c5[2] = q.b2 / q.a0; c5[3] = q.a1 / q.a0; c5[4] = q.a2 / q.a0;
	return 0;
}
