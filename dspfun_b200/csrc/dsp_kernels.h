// dsp_kernels.h -- launch entry points of the pass kernels.  Each (element type, row/column, generic/fast)
// combination lives in its own translation unit (kern_*.cu, all generated from kern_inst.cuh) so the library
// builds in parallel.
#pragma once
#include "dct_core.cuh"
#include "dct_ops.cuh"
#include "dct_fast.cuh"
#include "dsp_rt.h"
#include <string>

namespace dsp {

static const size_t kMaxSmem = 227 * 1024;
static const int kThreads = 256;

// fused == false launches the lean <OpMul, OpMul> instantiation (lop / sop may only be OP_NONE or OP_SCALE)
#define DSP_DECL_LAUNCH(NAME, ARGS)                                                                          \
	bool NAME(const ARGS &a, const FastDesc &f, bool fused, const OpAny &lop, const OpAny &sop, int grid,    \
	          int block, size_t smem, rt_stream st, std::string &err);
DSP_DECL_LAUNCH(launch_row_generic_f32, RowArgs)
DSP_DECL_LAUNCH(launch_row_generic_f64, RowArgs)
DSP_DECL_LAUNCH(launch_row_fast_f32, RowArgs)
DSP_DECL_LAUNCH(launch_row_fast_f64, RowArgs)
DSP_DECL_LAUNCH(launch_row_fast_f32p, RowArgs)     // planar specialisation (fixed lengths 256..8192, lean ops)
DSP_DECL_LAUNCH(launch_row_fast_f32i3, RowArgs)    // the same for 3 interleaved channels (the image tools' RGB layout)
DSP_DECL_LAUNCH(launch_col_generic_f32, ColArgs)
DSP_DECL_LAUNCH(launch_col_generic_f64, ColArgs)
DSP_DECL_LAUNCH(launch_col_fast_f32, ColArgs)
DSP_DECL_LAUNCH(launch_col_fast_f64, ColArgs)
#undef DSP_DECL_LAUNCH

struct RingArgs;
bool ring_supports(int n);
bool launch_row_ring_f32(const RingArgs &a, int n, bool fwd, rt_stream st, std::string &err);   // dct_ring.cuh

struct ColRingArgs;
bool colring_supports(int n);
int colring_max_panels();
int colring_scratch_panels();
bool colring_encode(ColRingArgs &a, int n, bool inverse, const float *in, long long ax_is, long long plane_is, float *out, long long ax_os, long long plane_os,
                    int nplanes, int ncols, float *scratch, int P, std::string &err);
bool launch_col_ring_f32(const ColRingArgs &a, int n, bool inverse, rt_stream st, std::string &err);   // dct_colring.cuh

bool launch_l2_prefetch(const void *base, long long pitch_bytes, int nrows, int row_bytes, rt_stream st, std::string &err);
bool launch_spec_resolve(char prec, const OpAny &op, const double *acc, double *scale_z, rt_stream st, std::string &err);
bool launch_block_quant(char prec, void *coeffs, long long n, int H, int W, int bd, int bh, int bw, double quantizer,
                        unsigned long long *count, rt_stream st, std::string &err);
bool block_dquant_supports(int bd);
bool launch_block_dquant(float *coeffs, int D, int H, int W, int bd, int bh, int bw, double quantizer, unsigned long long *count, rt_stream st,
                         std::string &err);
bool launch_block_store_u8(char prec, const void *coeffs, unsigned char *pels, long long n, double scale, rt_stream st, std::string &err);
bool launch_block_load_u8(char prec, const unsigned char *pels, void *coeffs, long long n, rt_stream st, std::string &err);

// 2-D block DCT of whole planes as tensor-core GEMMs (kern_block_mm.cu): tcgen05 / TMEM, 3 x TF32 split for float accuracy
bool block_mm_supports(int B);
void block_mm_cleanup();
bool launch_block_mm_f32(const float *in, float *out, long long nplanes, int H, int W, int B, int kind, double scale, rt_stream st,
                         std::string &err, float *dbg);

// dense float GEMM on the tensor cores, 3 x TF32 (kern_gemm_tc.cu): zoom's general synthesis path
bool gemm_tc_available();
bool launch_tf32_residual(const float *x, float *lo, long long n, rt_stream st, std::string &err);
bool launch_planarize3(const float *src, int rows, int src_cols, int cols, int ch, float *dst, float *dst_lo, int ld, rt_stream st, std::string &err);
bool launch_gemm_tf32x3(int M, int N, int K, const float *A, const float *A_lo, long long lda, const float *B, const float *B_lo, long long ldb,
                        float *D, float *D_lo, long long dr, long long dc, double alpha, rt_stream st, std::string &err);

// pointwise spectrogram stages of motion (kern_misc.cu)
struct MotionSpecArgs {
	int md, mh, mw;              // padded box (minbuf)
	int bd, bh, bw;              // box the kernel covers: the block (ispec) or the scaled block (spec)
	int type;                    // DSP_MOTION_SPEC_*
	int float_pixels;
	double norm, sf;             // motion.c:567, :566
	double c;                    // shift: 127.5 / log1p(sw sh sd norm 255 8)  (:568-569)
	OpAny coeff;                 // spec: the coefficient stages (OP_MOTION_COEFF with skipd), run here instead of in an inverse pass
};

bool launch_motion_coeff(char prec, const OpAny &op, void *coeffs, int D, int H, int W, rt_stream st, std::string &err);
bool launch_motion_ispec(char prec, const MotionSpecArgs &a, const void *pels, void *coeffs, rt_stream st, std::string &err);
bool launch_motion_spec(char prec, const MotionSpecArgs &a, const void *coeffs, void *pels, rt_stream st, std::string &err);

struct SplitArgs;
bool launch_split_fft_f32(const SplitArgs &a, const FastDesc &fM, bool fused, const OpAny &lop, const OpAny &sop, int grid, size_t smem,
                          rt_stream st, std::string &err);
bool launch_split_outer_f32(const SplitArgs &a, const FastDesc &fN, bool fused, const OpAny &lop, const OpAny &sop, int nwarps,
                            rt_stream st, std::string &err);

bool launch_split_inv_fft_f32(const SplitArgs &a, const FastDesc &fM, const FastDesc &fN, const OpAny &lop, int grid, size_t smem,
                              rt_stream st, std::string &err);
bool launch_split_inv_outer_f32(const SplitArgs &a, const FastDesc &fN, const OpAny &sop, int nwarps, rt_stream st, std::string &err);
// double: the forward split and the DIF-style inverse (sub-pass B first) -- the generic, element-type agnostic moves
bool launch_split_fft_f64(const SplitArgs &a, const FastDesc &fM, bool fused, const OpAny &lop, const OpAny &sop, int grid, size_t smem,
                          rt_stream st, std::string &err);
bool launch_split_outer_f64(const SplitArgs &a, const FastDesc &fN, bool fused, const OpAny &lop, const OpAny &sop, int nwarps,
                            rt_stream st, std::string &err);

bool launch_zoom_basis(char prec, void *basis, int nvec, int ncomp, int type, double num, double den, double offset, int len,
                       rt_stream st, std::string &err);
bool launch_zoom_shift_build(char prec, const void *coef, void *planes, int W, int Nh, int Nw, int ch, int cw, double dx, double dy,
                             rt_stream st, std::string &err);
bool launch_zoom_shift_combine(char prec, const void *planes, void *out, int Nh, int Nw, int vh, int vw, double alpha, rt_stream st,
                               std::string &err);
bool launch_zoom_gemm(char prec, int M, int N, int K, const void *A, long long ar, long long ac, const void *B, long long br,
                      long long bc, void *Cm, long long cr, long long cc, double alpha, rt_stream st, std::string &err);

}  // namespace dsp
