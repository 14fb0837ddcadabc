// kern_gemm_tc.cu -- dense float GEMM on the 5th-generation tensor cores with float accuracy (3 x TF32):
//     D[m][n] = alpha * sum_k A[m][k] B[n][k]          A: [M][lda], B: [N][ldb], both K-contiguous ("TN")
// zoom's general path (arbitrary rational scale / centered basis / offset view, zoom/zoom.c:361-375) is two such products
// per channel -- out = Yb (Xb C^T)^T / (W H) -- and was an FP64-accumulating SIMT GEMM in round 1 (kern_zoom.cu, kept for
// double and as the fallback).
//
// Operands are split x = hi + lo, hi = the top 19 bits of the fp32 word (what the MMA reads; the raw array serves), lo =
// x - hi kept in a companion array written once per operand (k_tf32_residual, or by the producing GEMM's epilogue).
// Per 32-wide k block a CTA (128 x 128 output tile) receives four 128 x 32 boxes by TMA (A, A_lo, B, B_lo; 128-byte
// swizzle; B_lo lands right behind B so that [B | B_lo] is ONE K-major operand with N = 256) and issues
//     D[:, 0:256]   += A    [B | B_lo]^T        (hi*hi | hi*lo)
//     D[:, 128:256] += A_lo  B^T                 (lo*hi joins the small half)
// with fp32 accumulators in TMEM; the epilogue adds the halves.  Warp-specialised, no SIMT work in the main loop:
// warp 0 issues TMA into a 3-stage ring, warp 1 issues the MMAs (tcgen05.commit frees each stage), warps 2-5 drain TMEM.
#include "dsp_kernels.h"
#include "dct_tma.cuh"
#include <math.h>

namespace dsp {

constexpr int kGtTile = 128, kGtKb = 32, kGtStages = 3;
constexpr int kGtBox = kGtTile * kGtKb * 4;              // one operand box: 16 KB
constexpr int kGtStage = 4 * kGtBox;                     // A | A_lo | B | B_lo
constexpr int kGtThreads = 192;

struct GemmTcArgs {
	TmaDesc a, alo, b, blo;
	float *d, *dlo;
	long long dr, dc;
	int M, N, K;
	float alpha;
};

#if DSP_GPU
DSP_DEV uint64_t gt_desc(uint32_t saddr) {               // K-major, 128-byte swizzle, 8-row groups 1024 bytes apart
	return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | ((uint64_t)2 << 61);
}
DSP_DEV uint32_t gt_idesc(int N) { return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(kGtTile >> 4) << 24); }
DSP_DEV void gt_mma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
	asm volatile(
	    "{\n"
	    ".reg .pred p;\n"
	    "setp.ne.b32 p, %4, 0;\n"
	    "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
	    "}\n" ::"r"(d_tmem),
	    "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
	    : "memory");
}
DSP_DEV void gt_commit(uint64_t *bar) {
	asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
DSP_DEV void gt_wait(uint64_t *bar, uint32_t parity) {   // bounded: a lost completion traps instead of hanging the device
	uint32_t ok = 0;
	for (uint32_t spin = 0; spin < (1u << 26); spin++) {
		asm volatile(
		    "{\n"
		    ".reg .pred p;\n"
		    "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
		    "selp.u32 %0, 1, 0, p;\n"
		    "}\n"
		    : "=r"(ok)
		    : "r"(smem_u32(bar)), "r"(parity)
		    : "memory");
		if (ok) return;
	}
	asm volatile("trap;");
}
DSP_DEV void gt_ld16(uint32_t taddr, uint32_t (&v)[16]) {
	asm volatile(
	    "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
	    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
	    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
	      "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
	    : "r"(taddr)
	    : "memory");
}
DSP_DEV float gt_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

__global__ void __launch_bounds__(kGtThreads, 1) k_gemm_tf32x3(const __grid_constant__ GemmTcArgs a) {
	extern __shared__ unsigned char gt_smem_raw[];
	const uint32_t raw = smem_u32(gt_smem_raw);
	const uint32_t base = (raw + 1023u) & ~1023u;
	unsigned char *sm = gt_smem_raw + (base - raw);
	uint64_t *bars = (uint64_t *)(sm + kGtStages * kGtStage);       // full[3] | empty[3] | accumulator complete
	uint32_t *tslot = (uint32_t *)(bars + 2 * kGtStages + 1);
	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	const int m0 = blockIdx.y * kGtTile, n0 = blockIdx.x * kGtTile;
	const int nkb = (a.K + kGtKb - 1) / kGtKb;

	if (tid == 0) {
		for (int i = 0; i < 2 * kGtStages + 1; i++) mbar_init(&bars[i], 1);
		mbar_fence_init();
	}
	if (warp == 1) {
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tslot)), "r"(256u) : "memory");
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
	}
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
	const uint32_t tmem = *(volatile uint32_t *)tslot;

	if (warp == 0) {
		if (lane == 0) {                                        // producer: four boxes per k block
			for (int kb = 0; kb < nkb; kb++) {
				const int s = kb % kGtStages;
				if (kb >= kGtStages) gt_wait(&bars[kGtStages + s], (uint32_t)((kb / kGtStages) - 1) & 1u);
				unsigned char *st = sm + s * kGtStage;
				mbar_expect_tx(&bars[s], kGtStage);
				tma_load2(st, &a.a, kb * kGtKb, m0, &bars[s]);
				tma_load2(st + kGtBox, &a.alo, kb * kGtKb, m0, &bars[s]);
				tma_load2(st + 2 * kGtBox, &a.b, kb * kGtKb, n0, &bars[s]);
				tma_load2(st + 3 * kGtBox, &a.blo, kb * kGtKb, n0, &bars[s]);
			}
		}
	} else if (warp == 1) {
		if (lane == 0) {                                        // MMA issuer
			const uint32_t id256 = gt_idesc(256), id128 = gt_idesc(128);
			for (int kb = 0; kb < nkb; kb++) {
				const int s = kb % kGtStages;
				gt_wait(&bars[s], (uint32_t)(kb / kGtStages) & 1u);
				asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
				const uint32_t st = base + s * kGtStage;
				const uint64_t da = gt_desc(st), dl = gt_desc(st + kGtBox), db = gt_desc(st + 2 * kGtBox);
#pragma unroll
				for (int ks = 0; ks < kGtKb / 8; ks++) {
					gt_mma(tmem, da + 2 * ks, db + 2 * ks, id256, (uint32_t)((kb | ks) != 0));      // A [B | B_lo]
					gt_mma(tmem + kGtTile, dl + 2 * ks, db + 2 * ks, id128, 1u);                     // A_lo B
				}
				gt_commit(&bars[kGtStages + s]);
			}
			gt_commit(&bars[2 * kGtStages]);
		}
	} else {
		// epilogue: this warp's quarter of the rows; D = (hi*hi) + (hi*lo + lo*hi), times alpha
		gt_wait(&bars[2 * kGtStages], 0);
		asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
		const int q = warp & 3;
		const int m = m0 + q * 32 + lane;
		const uint32_t lane_base = (uint32_t)(q * 32) << 16;
#pragma unroll 1
		for (int c = 0; c < kGtTile / 16; c++) {
			uint32_t v1[16], v2[16];
			gt_ld16(tmem + lane_base + 16 * c, v1);
			gt_ld16(tmem + lane_base + kGtTile + 16 * c, v2);
			asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
			if (m < a.M) {
				float *dp = a.d + (long long)m * a.dr + (long long)(n0 + 16 * c) * a.dc;
				float *lp = a.dlo ? a.dlo + (long long)m * a.dr + (long long)(n0 + 16 * c) * a.dc : nullptr;
#pragma unroll
				for (int i = 0; i < 16; i++) {
					if (n0 + 16 * c + i < a.N) {
						const float x = (__uint_as_float(v1[i]) + __uint_as_float(v2[i])) * a.alpha;
						dp[(long long)i * a.dc] = x;
						if (lp) lp[(long long)i * a.dc] = x - gt_hi(x);
					}
				}
			}
		}
	}
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
}

// lo[i] = x[i] - (x[i] with the low 13 mantissa bits cleared): exact in fp32
__global__ void k_tf32_residual(const float *x, float *lo, long long n) {
	for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
		const float v = x[i];
		lo[i] = v - gt_hi(v);
	}
}
// channel ch of the first `rows` x `cols` corner of an interleaved [..][src_cols][3] array -> planar [rows][ld] (+ its
// residual), columns beyond `cols` zero
__global__ void k_planarize3(const float *src, int rows, int src_cols, int cols, int ch, float *dst, float *dst_lo, int ld) {
	const long long total = (long long)rows * ld;
	for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
		const int r = (int)(i / ld), c = (int)(i - (long long)r * ld);
		const float v = c < cols ? src[((long long)r * src_cols + c) * 3 + ch] : 0.0f;
		dst[i] = v;
		dst_lo[i] = v - gt_hi(v);
	}
}
#endif  // DSP_GPU

bool gemm_tc_available() {
#if DSP_GPU
	static const bool off = getenv("DSP_ZOOM_NO_TC") != nullptr;
	return !off;
#else
	return false;
#endif
}

bool launch_tf32_residual(const float *x, float *lo, long long n, rt_stream st, std::string &err) {
#if DSP_GPU
	k_tf32_residual<<<148 * 8, 256, 0, st>>>(x, lo, n);
	return rt_ok(cudaGetLastError(), err, "tf32 residual launch");
#else
	(void)x; (void)lo; (void)n; (void)st;
	err = "tensor-core GEMM: GPU only";
	return false;
#endif
}

bool launch_planarize3(const float *src, int rows, int src_cols, int cols, int ch, float *dst, float *dst_lo, int ld, rt_stream st, std::string &err) {
#if DSP_GPU
	k_planarize3<<<148 * 8, 256, 0, st>>>(src, rows, src_cols, cols, ch, dst, dst_lo, ld);
	return rt_ok(cudaGetLastError(), err, "planarize launch");
#else
	(void)src; (void)rows; (void)src_cols; (void)cols; (void)ch; (void)dst; (void)dst_lo; (void)ld; (void)st;
	err = "tensor-core GEMM: GPU only";
	return false;
#endif
}

// A, A_lo: [M][lda]; B, B_lo: [N][ldb]; lda, ldb multiples of 4 floats, bases 16-byte aligned.  D (and D_lo, optional: the
// residual of D for use as an operand of the next product) are written at d[m * dr + n * dc].
bool launch_gemm_tf32x3(int M, int N, int K, const float *A, const float *A_lo, long long lda, const float *B, const float *B_lo, long long ldb,
                        float *D, float *D_lo, long long dr, long long dc, double alpha, rt_stream st, std::string &err) {
#if DSP_GPU
	if (M < 1 || N < 1 || K < 1 || (lda & 3) || (ldb & 3) || ((uintptr_t)A & 15) || ((uintptr_t)A_lo & 15) || ((uintptr_t)B & 15) || ((uintptr_t)B_lo & 15)) {
		err = "tensor-core GEMM: operands must be 16-byte aligned with leading dimensions that are multiples of 4";
		return false;
	}
	GemmTcArgs a;
	memset(&a, 0, sizeof(a));
	const float *ptr[4] = {A, A_lo, B, B_lo};
	TmaDesc *map[4] = {&a.a, &a.alo, &a.b, &a.blo};
	for (int i = 0; i < 4; i++) {
		TmaView v;
		memset(&v, 0, sizeof(v));
		v.base = (void *)ptr[i]; v.rank = 2;
		v.dims[0] = (unsigned long long)K; v.dims[1] = (unsigned long long)(i < 2 ? M : N);
		v.strides[1] = (unsigned long long)(i < 2 ? lda : ldb) * 4;
		v.box[0] = kGtKb; v.box[1] = kGtTile;
		if (!tma_encode(map[i], v, err, true)) return false;
	}
	a.d = D; a.dlo = D_lo; a.dr = dr; a.dc = dc; a.M = M; a.N = N; a.K = K; a.alpha = (float)alpha;
	const size_t smem = 1024 + (size_t)kGtStages * kGtStage + 128;
	static unsigned long long attr_dev = 0;
	const int dev = rt_device() & 63;
	if (!((attr_dev >> dev) & 1ull)) {
		if (!rt_ok(cudaFuncSetAttribute(k_gemm_tf32x3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), err, "smem attribute")) return false;
		attr_dev |= 1ull << dev;
	}
	dim3 grid((unsigned)((N + kGtTile - 1) / kGtTile), (unsigned)((M + kGtTile - 1) / kGtTile));
	k_gemm_tf32x3<<<grid, kGtThreads, smem, st>>>(a);
	return rt_ok(cudaGetLastError(), err, "tensor-core GEMM launch");
#else
	(void)M; (void)N; (void)K; (void)A; (void)A_lo; (void)lda; (void)B; (void)B_lo; (void)ldb; (void)D; (void)D_lo; (void)dr; (void)dc; (void)alpha; (void)st;
	err = "tensor-core GEMM: GPU only";
	return false;
#endif
}

}  // namespace dsp
