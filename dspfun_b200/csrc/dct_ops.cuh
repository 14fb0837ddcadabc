// dct_ops.cuh -- pointwise stages fused into the first (load side) or last (store side) pass of a plan.
//
// OpNone is the zero-cost default.  OpAny is one POD that carries any of the tool-level stages and dispatches on
// a CTA-uniform `kind`; kernels are instantiated for <OpNone, OpNone> and <OpAny, OpAny> only.
//
// Reference formulas (all /root/reference): spec/spec.c:66-139 (OP_SPEC), spec/ispec.c:84-163 (OP_ISPEC),
// scan/scan.c:296-298 (OP_SCALE), scan/scan.c:429-459 (OP_SCAN_MASK / OP_SCAN_ACCUM).
// The reference rounds to `coeff` after every loop; the stages below round to T at the same points and do the
// transcendental steps in double ("intermediate" = D, the reference default for COEFF_PRECISION=F).
#pragma once
#include "dct_core.cuh"
#include <math.h>

#if DSP_GPU
#define DSP_OPANY_ATTR __device__ __noinline__
#else
#define DSP_OPANY_ATTR inline
#endif

namespace dsp {

enum {
	OP_NONE = 0,
	OP_SCALE,        // v * p[0]
	OP_ACCUM_DC,     // store side of spec's row pass: sum the k = 0 outputs per channel into aux (double[4])
	OP_SPEC,         // store side of spec's column pass
	OP_ISPEC,        // load side of ispec's first pass
	OP_SCAN_MASK,    // load side: keep coefficient iff lo <= index[y][x] < hi, zero DC   (scan/scan.c:429-445)
	OP_SCAN_ACCUM,   // store side: sum[i] += v ; out = sum[i]                             (scan/scan.c:451-456)
	OP_MOTION_COEFF, // load side of motion's inverse: zero-pad/crop, normalise, band-pass damp/boost, threshold,
	                 // preserve-dc, quantise, de-normalise                               (motion/motion.c:617,644-751)
	OP_MOTION_STORE, // store side of motion's inverse: scale, clamp, round to the 8-bit range  (motion/motion.c:757-776)
};

struct OpAny {
	enum { kNeedsCoord = 1 };
	int kind;
	int scaletype, signtype, rangetype;
	int d, w, h, flag;
	int lo, hi;
	double p[4];             // OP_SCALE: p[0] = factor.  OP_SPEC/ISPEC: p[0] = gain, p[1] = norm (2wh); ISPEC: p[2] = 1/gain, p[3] = 255/254
	double q[4];             // OP_ISPEC: log1p(max[z]) or max[z] per channel; OP_SPEC: q[0] = 1/norm, q[1] = 254/255
	double dc[4];            // OP_ISPEC preserve_dc values
	int a3[3], b3[3], e3[3]; // OP_MOTION_COEFF: active box, band-pass begin / end (d, h, w order)
	double m[8];             // OP_MOTION_*: damp, boost, threshold min, max, quantizer, grey offset, output scale, -
	int flag2;               // OP_MOTION_STORE: float pixels
	const void *aux_c;       // OP_SPEC: double[8] device scalars {scale_z[4], -, -, -, -};  OP_ISPEC: u8 signmap;  OP_SCAN_MASK: int32 index map
	void *aux;               // OP_ACCUM_DC: double[4] accumulators;  OP_SPEC: double[4] DC out;  OP_SCAN_ACCUM: T sum buffer

	int skipn, skipd;        // OP_MOTION_COEFF: skip the normalisation (motion --ispec input: motion.c:639) / the de-normalisation (--spec output: :746)
	int fast;                // OP_MOTION_COEFF: no threshold / preserve-dc / quantiser -> the OpMotionCoeff functor serves it
	double nf[4], rnf[4];    // OP_MOTION_COEFF: 2 sqrt2 / sqrt2^k and its reciprocal, k = number of zero coordinates
	// Not inlined on the GPU: the unrolled passes call it per element, and dozens of inlined copies of this switch
	// overflow the instruction cache (measured: "no instruction" became the top stall of the fused kernels).
	template <class T> DSP_OPANY_ATTR T operator()(T v, const Coord &c) const {
		typedef double I;
		const I SQRT2 = 1.41421356237309504880168872420969808;
		switch (kind) {
		default:
		case OP_NONE: return v;
		case OP_SCALE: return (T)(v * (T)p[0]);
		case OP_ACCUM_DC: {
			if (c.i2 == 0) {
#if DSP_GPU
				atomicAdd((double *)aux + c.ch, (double)v);
#else
				((double *)aux)[c.ch] += (double)v;
#endif
			}
			return v;
		}
		case OP_SPEC: {
			const int y = c.i1, x = c.i2, ch = c.ch;
			T f = v;
			if (y == 0 && x == 0) ((double *)aux)[ch] = (double)f / ((double)w * (double)h * 4.0);   // spec.c:66-68
			if (y == 0) f = (T)((I)f / SQRT2);                                                       // :70-71
			if (x == 0) f = (T)((I)f / SQRT2);                                                       // :72-74
			// The double divisions of the reference's `intermediate` chain are multiplications by reciprocals prepared
			// once (q[0] = 1/(2wh), q[1] = 254/255, the resolved 1/log1p(max[z])): the quotients differ from true division
			// by at most an ulp of double before they are rounded to `coeff`, and the stage is FP64-issue bound.
			f = (T)((I)f * q[0]);                                                                    // :76-78
			f = (T)((I)f * p[0]);                                                                    // :89-90
			if (scaletype == 0) {
				const I rsc = DSP_LDG((const double *)aux_c + 8 + ch);                               // 1 / log1p(max[z])
				f = (T)(copysign(log1p(fabs((I)f)), (I)f) * rsc);                                    // :110-117
			} else {
				const I sc = DSP_LDG((const double *)aux_c + ch);                                    // max[z], resolved on device
				f = (T)(f / (T)sc);                                                                  // :118-121
			}
			if (signtype == 0) f = (T)fabs((I)f);                                                    // :126-128
			else if (signtype == 1) f = (T)(((I)f * 0.5 + 0.5) * q[1]);                              // :130-132
			else if (signtype == 2) { if (y != 0 || x != 0) f = signbit((I)f) ? (T)0 : (T)1; }        // :134-136
			return f;
		}
		case OP_ISPEC: {
			const int y = c.i1, x = c.i2, ch = c.ch;
			const bool pix0 = (y == 0 && x == 0);
			T f = v;
			if (signtype == 0) {                                                                     // ispec.c:87-98
				if (aux_c && !pix0) {
					const int sm = (int)DSP_LDG((const unsigned char *)aux_c + ((size_t)y * w + x) * d + ch) - 128;
					f = (T)copysign((I)f, (I)sm);
				}
			} else if (signtype == 1) f = (T)(((I)f * p[3] - 0.5) * 2);                              // :100-103 (p[3] = 255/254)
			else if (signtype == 2) { if (!pix0) f = f * 2 - 1; }                                     // :104-107
			if (scaletype == 0) {                                                                    // :136-143
				const T prod = f * (T)q[ch];
				f = (T)copysign(expm1(fabs((I)prod)), (I)f);
			} else f = f * (T)q[ch];                                                                 // :144-147
			f = (T)((I)f * p[2]);                                                                    // :150-151 (p[2] = 1/gain)
			if (y == 0) f = (T)((I)f * SQRT2);                                                       // :153-154
			if (x == 0) f = (T)((I)f * SQRT2);                                                       // :155-157
			f = f / 2;                                                                               // :158-159
			if (flag && pix0) f = (T)dc[ch];                                                         // :161-163
			return f;
		}
		case OP_SCAN_MASK: {
			const int y = c.i1, x = c.i2;
			if (y == 0 && x == 0) return (T)0;                                                       // scan.c:445
			const int idx = DSP_LDG((const int *)aux_c + (size_t)y * w + x);
			return (idx >= lo && idx < hi) ? v : (T)0;                                               // scan.c:429-432
		}
		case OP_MOTION_COEFF: {
			int z = c.i0, y = c.i1, x = c.i2;
			if (w > 0) {                                  // temporal pass of a slab-sharded volume: [D][slice of flattened h*w]
				const int hw = lo + c.ch;
				z = c.i2; y = hw / w; x = hw - y * w;
			}
			if (z >= a3[0] || y >= a3[1] || x >= a3[2]) return (T)0;                                 // outside the active box: motion.c:617
			const I nf = (2 * SQRT2) / ((x ? 1.0 : SQRT2) * (y ? 1.0 : SQRT2) * (z ? 1.0 : SQRT2));
			T f = skipn ? v : (T)((I)v * nf);                                                        // :644-647 (not with --ispec)
			const T dc0 = f;
			const bool inside = z >= b3[0] && z < e3[0] && y >= b3[1] && y < e3[1] && x >= b3[2] && x < e3[2];
			if (!inside) { if (m[0] != 1.0) f = f * (T)m[0]; }                                       // :683-713
			else if (m[1] != 1.0) f = f * (T)m[1];                                                   // :714-719
			if (m[3] != 0.0) {                                                                       // :721-728
				const T af = f < 0 ? -f : f;
				if (af < (T)m[2] || af > (T)m[3]) f = 0;
			}
			if (flag && !(z | y | x)) {                                                              // :730-738
				const bool dcstop = (b3[0] | b3[1] | b3[2]) != 0;
				if (dcstop || m[1] != 1.0 || m[3] != 0.0) {
					if (flag == 1) f = dc0;
					else f = (T)((I)f + (1.0 - (dcstop ? m[0] : m[1])) * m[5]);
				}
			}
			if (m[4] != 0.0) {                                                                       // :740-744
				f = (T)(round((I)f / m[4]) * m[4]);
				if (f != 0 && aux) {
#if DSP_GPU
					atomicAdd((unsigned long long *)aux, 1ull);
#else
					*(unsigned long long *)aux += 1ull;
#endif
				}
			}
			return skipd ? f : (T)((I)f / nf);                                                       // :748-751 (not with --spec)
		}
		case OP_MOTION_STORE: {
			const I pel = (I)v * m[6];                                                               // :757,767
			if (flag2) return (T)(pel / 255.0);                                                      // :773
			return (T)(pel > 255.0 ? 255.0 : pel < 0.0 ? 0.0 : (I)lround(pel));                      // :776
		}
		case OP_SCAN_ACCUM: {
			T *sum = (T *)aux + (((size_t)c.i1 * w + c.i2) * d + c.ch);
			const T s = *sum + v;                                                                    // scan.c:454
			*sum = s;
			return s;
		}
		}
	}
};

// ---- motion's two hot stages as small functors of their own: kernels instantiated for them carry a few dozen
// instructions per element instead of OpAny's whole switch (ncu r02 on the 256-frame volume: the OpAny instantiation of
// the temporal pass ran 5x the instructions of the plain one and stalled on instruction fetch).
// OP_MOTION_COEFF without threshold / preserve-dc / quantiser (OpAny::fast): normalise, damp | boost, de-normalise.
struct OpMotionCoeff {
	enum { kNeedsCoord = 1 };
	int a3[3], b3[3], e3[3];
	int w, lo;                   // > 0: flat coordinates (see OpAny)
	float m0, m1;                // damp (outside the band-pass box), boost (inside)
	double nf0, nf1, nf2, nf3, rnf0, rnf1, rnf2, rnf3;
	template <class T> DSP_DEVM T operator()(T v, const Coord &c) const {
		typedef double I;
		int z = c.i0, y = c.i1, x = c.i2;
		if (w > 0) {
			const int hw = lo + c.ch;
			z = c.i2; y = hw / w; x = hw - y * w;
		}
		if (z >= a3[0] || y >= a3[1] || x >= a3[2]) return (T)0;                                     // motion.c:617
		const int k = (x == 0) + (y == 0) + (z == 0);
		const I nf = k == 0 ? nf0 : (k == 1 ? nf1 : (k == 2 ? nf2 : nf3)), rnf = k == 0 ? rnf0 : (k == 1 ? rnf1 : (k == 2 ? rnf2 : rnf3));
		T f = (T)((I)v * nf);                                                                        // :644-647
		const bool inside = z >= b3[0] && z < e3[0] && y >= b3[1] && y < e3[1] && x >= b3[2] && x < e3[2];
		f = f * (T)(inside ? m1 : m0);                                                               // :683-719
		return (T)((I)f * rnf);                                                                      // :748-751
	}
	static OpMotionCoeff from(const OpAny &o) {
		OpMotionCoeff r;
		for (int i = 0; i < 3; i++) { r.a3[i] = o.a3[i]; r.b3[i] = o.b3[i]; r.e3[i] = o.e3[i]; }
		r.w = o.w; r.lo = o.lo; r.m0 = (float)o.m[0]; r.m1 = (float)o.m[1];
		r.nf0 = o.skipn ? 1.0 : o.nf[0]; r.nf1 = o.skipn ? 1.0 : o.nf[1]; r.nf2 = o.skipn ? 1.0 : o.nf[2]; r.nf3 = o.skipn ? 1.0 : o.nf[3];   // (x 1.0 is exact)
		r.rnf0 = o.rnf[0]; r.rnf1 = o.rnf[1]; r.rnf2 = o.rnf[2]; r.rnf3 = o.rnf[3];
		return r;
	}
};
// OP_MOTION_STORE: scale, clamp, round to the 8-bit range (or float pels / 255)
struct OpMotionStore {
	enum { kNeedsCoord = 0 };
	double scale;
	int float_pixels;
	template <class T> DSP_DEVM T operator()(T v, const Coord &) const {
		const double pel = (double)v * scale;                                                        // motion.c:757,767
		if (float_pixels) return (T)(pel / 255.0);                                                   // :773
		return (T)(pel > 255.0 ? 255.0 : pel < 0.0 ? 0.0 : (double)lround(pel));                     // :776
	}
	static OpMotionStore from(const OpAny &o) { OpMotionStore r; r.scale = o.m[6]; r.float_pixels = o.flag2; return r; }
};

// ---- spec / ispec stages as functors of their own (same arithmetic, same rounding points as the OpAny cases; the
// kernels instantiated for them inline a few dozen instructions instead of calling the whole switch per element)
struct OpAccumDc {               // store side of spec's row pass: sum the k = 0 outputs per channel (spec.c:92-117 resolved later)
	enum { kNeedsCoord = 1 };
	double *acc;
	template <class T> DSP_DEVM T operator()(T v, const Coord &c) const {
		if (c.i2 == 0) {
#if DSP_GPU
			atomicAdd(acc + c.ch, (double)v);
#else
			acc[c.ch] += (double)v;
#endif
		}
		return v;
	}
	static OpAccumDc from(const OpAny &o) { OpAccumDc r; r.acc = (double *)o.aux; return r; }
};
struct OpSpecStore {             // store side of spec's column pass (spec.c:66-139)
	enum { kNeedsCoord = 1 };
	int scaletype, signtype, w, h;
	double gain, rnorm, c254;
	const double *scale_z;       // device: max[z] | ... | 1 / log1p(max[z]) at +8
	double *dc_out;
	template <class T> DSP_DEVM T operator()(T v, const Coord &c) const {
		typedef double I;
		const I SQRT2 = 1.41421356237309504880168872420969808;
		const int y = c.i1, x = c.i2, ch = c.ch;
		T f = v;
		if (y == 0 && x == 0) dc_out[ch] = (double)f / ((double)w * (double)h * 4.0);            // spec.c:66-68
		if (y == 0) f = (T)((I)f / SQRT2);                                                        // :70-71
		if (x == 0) f = (T)((I)f / SQRT2);                                                        // :72-74
		f = (T)((I)f * rnorm);                                                                    // :76-78
		f = (T)((I)f * gain);                                                                     // :89-90
		if (scaletype == 0) f = (T)(copysign(log1p(fabs((I)f)), (I)f) * DSP_LDG(scale_z + 8 + ch));   // :110-117
		else f = (T)(f / (T)DSP_LDG(scale_z + ch));                                               // :118-121
		if (signtype == 0) f = (T)fabs((I)f);                                                     // :126-128
		else if (signtype == 1) f = (T)(((I)f * 0.5 + 0.5) * c254);                               // :130-132
		else if (signtype == 2) { if (y != 0 || x != 0) f = signbit((I)f) ? (T)0 : (T)1; }         // :134-136
		return f;
	}
	static OpSpecStore from(const OpAny &o) {
		OpSpecStore r;
		r.scaletype = o.scaletype; r.signtype = o.signtype; r.w = o.w; r.h = o.h;
		r.gain = o.p[0]; r.rnorm = o.q[0]; r.c254 = o.q[1];
		r.scale_z = (const double *)o.aux_c; r.dc_out = (double *)o.aux;
		return r;
	}
};
struct OpIspecLoad {             // load side of ispec's first pass (ispec.c:84-163)
	enum { kNeedsCoord = 1 };
	int scaletype, signtype, w, d, preserve;
	double rgain, c255;
	double q[4], dc[4];
	const unsigned char *signmap;
	template <class T> DSP_DEVM T operator()(T v, const Coord &c) const {
		typedef double I;
		const I SQRT2 = 1.41421356237309504880168872420969808;
		const int y = c.i1, x = c.i2, ch = c.ch;
		const bool pix0 = (y == 0 && x == 0);
		const T qc = (T)(ch == 0 ? q[0] : (ch == 1 ? q[1] : (ch == 2 ? q[2] : q[3])));
		T f = v;
		if (signtype == 0) {                                                                      // ispec.c:87-98
			if (signmap && !pix0) {
				const int sm = (int)DSP_LDG(signmap + ((size_t)y * w + x) * d + ch) - 128;
				f = (T)copysign((I)f, (I)sm);
			}
		} else if (signtype == 1) f = (T)(((I)f * c255 - 0.5) * 2);                               // :100-103
		else if (signtype == 2) { if (!pix0) f = f * 2 - 1; }                                      // :104-107
		if (scaletype == 0) {                                                                     // :136-143
			const T prod = f * qc;
			f = (T)copysign(expm1(fabs((I)prod)), (I)f);
		} else f = f * qc;                                                                        // :144-147
		f = (T)((I)f * rgain);                                                                    // :150-151
		if (y == 0) f = (T)((I)f * SQRT2);                                                        // :153-154
		if (x == 0) f = (T)((I)f * SQRT2);                                                        // :155-157
		f = f / 2;                                                                                // :158-159
		if (preserve && pix0) f = (T)(ch == 0 ? dc[0] : (ch == 1 ? dc[1] : (ch == 2 ? dc[2] : dc[3])));   // :161-163
		return f;
	}
	static OpIspecLoad from(const OpAny &o) {
		OpIspecLoad r;
		r.scaletype = o.scaletype; r.signtype = o.signtype; r.w = o.w; r.d = o.d; r.preserve = o.flag;
		r.rgain = o.p[2]; r.c255 = o.p[3];
		for (int z = 0; z < 4; z++) { r.q[z] = o.q[z]; r.dc[z] = o.dc[z]; }
		r.signmap = (const unsigned char *)o.aux_c;
		return r;
	}
};

// Resolves spec's data-dependent range (spec/spec.c:92-117) once the row pass has accumulated the per-channel
// sums S[z] of its k = 0 outputs: Y[0,0,z] = 2 S[z].  Writes scale_z[ch] = log1p(max[ch]) or max[ch].
template <class T>
DSP_DEV void spec_resolve_range(const OpAny &op, const double *acc, double *scale_z) {
	typedef double I;
	const I SQRT2 = 1.41421356237309504880168872420969808;
	T mx[4];
	for (int z = 0; z < op.d && z < 4; z++) {
		T f = (T)(2.0 * acc[z]);
		f = (T)((I)f / SQRT2); f = (T)((I)f / SQRT2);
		f = (T)((I)f / op.p[1]);
		f = (T)((I)f * op.p[0]);
		mx[z] = f;
	}
	if (op.rangetype == 0) for (int z = 0; z < op.d && z < 4; z++) mx[z] = (T)op.p[0];
	else if (op.rangetype == 1) {
		T m = mx[0];
		for (int z = 1; z < op.d && z < 4; z++) if (mx[z] > m) m = mx[z];
		for (int z = 0; z < op.d && z < 4; z++) mx[z] = m;
	}
	for (int z = 0; z < op.d && z < 4; z++) {
		T m = mx[z];
		if (op.scaletype == 0) m = (T)log1p((I)m);     // mc(log1p): rounded to coeff
		scale_z[z] = (double)m;
		scale_z[8 + z] = 1.0 / (double)m;              // OP_SPEC multiplies (d_scalars[12..15])
	}
}

}  // namespace dsp
