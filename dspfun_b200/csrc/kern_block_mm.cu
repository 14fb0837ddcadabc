// kern_block_mm.cu -- 2-D block DCT as two small GEMMs per tile on the 5th-generation tensor cores (tcgen05 + TMEM).
//
// motion -b BxBx1 / -b BxBxD and applybasis' DCT bases transform every B x B block of a plane on its own
// (/root/reference/motion/motion.c:613-641 with the README's 8x8x8 example motion/README.md:75-77;
// /root/reference/applybasis/applybasis.c:410-448 is the same contraction written as an O(N^4) loop).  For B <= 64 the
// FFT passes of the plan path move every sample through HBM once per axis; here a 128 x 128 tile of a plane makes ONE
// round trip and both axes are contracted on chip:
//
//     stage 1 (along w):  D1[r][gB+n] = sum_k X[r][gB+k] M[n][k]      A = X tile, K-major, as the TMA left it (128-byte
//                                                                     swizzle), B = the block matrix, D1 in TMEM
//     stage 2 (along h):  D2[c][gB+n] = sum_k D1[gB+k][c] M[n][k]     A = D1 read back from TMEM and stored transposed into
//                                                                     shared memory, K-major in the same swizzled form;
//                                                                     D2 in TMEM
//     epilogue:           out[gB+n][c] = D2[c][gB+n]                  lane = image column: 128-byte coalesced stores
//
// M is FFTW's unnormalised REDFT10 or REDFT01 matrix of size B (B = 8 uses a block-diagonal pair so that N = 16).
// The tensor cores multiply in TF32; to keep float accuracy every operand is split x = hi + lo (hi = the top 19 bits,
// which is what the MMA reads of an fp32 word, lo = x - hi, exact) and each product is three TF32 products with fp32
// accumulation in TMEM, issued as two MMA sequences: A_hi [M_hi | M_lo] (one operand of width 2B) and A_lo M_hi onto the
// small half; whoever reads the accumulator adds the halves.  The error is ~2^-21 relative per product instead of 2^-11.
//
// One persistent CTA per SM, 384 threads: eight front warps (lo operand, MMA issue by thread 0, the TMEM -> shared
// hand-over between the stages) and four store warps that drain the finished accumulator of the previous tile while the
// front works on the next one.  Thread 0 also issues the TMA loads two tiles ahead; completion comes back through
// mbarriers (cp.async.bulk.tensor complete_tx, tcgen05.commit).  DESIGN.md 4b has the measurements.
#include "dsp_kernels.h"
#include "dct_tma.cuh"
#include <math.h>
#include <map>
#include <mutex>
#include <vector>

namespace dsp {

constexpr int kMmTile = 128;             // tile edge: M of both MMAs
constexpr int kMmFront = 256;           // warps 0-7: operand preparation, MMA issue, TMEM hand-over between the stages
constexpr int kMmEpi = 128;             // warps 8-11: accumulator -> global stores of the previous tile
constexpr int kMmThreads = kMmFront + kMmEpi;
constexpr int kMmBuf = kMmTile * kMmTile * 4;   // one operand buffer: 64 KB
constexpr int kMmTmemCols = 512;         // D1 | D2, each 128 outputs x two partial sums

struct BlockMmArgs {
	TmaDesc in_map;          // [planes][H][W] floats, box {32, 128, 1}, 128-byte swizzle
	const float *consts;     // hi | lo halves of the block matrix, each Be*Be floats in operand layout
	float *out;
	float *dbg;              // bring-up: D1 and D2 of tile 0 (2 x 128 x 128 floats) or null
	int H, W, nplanes, Be;
	int tiles_x, tiles_y;
	long long ntiles;
	float scale;
};

static size_t block_mm_smem(int Be) { return 1024 + 3 * (size_t)kMmBuf + 8 * (size_t)Be * Be + 128; }

// element (n, k) of an N x K operand stored K-major without swizzle: 8 x 16-byte core matrices, K chunks 128 bytes apart,
// groups of 8 rows 32*K bytes apart (float index)
static inline int const_slot(int n, int k, int K) { return (n >> 3) * (8 * K) + (k >> 2) * 32 + (n & 7) * 4 + (k & 3); }

#if DSP_GPU
// ------------------------------------------------------------------------------------------------ tcgen05 wrappers
// shared-memory operand descriptor (cute/arch/mma_sm100_desc.hpp SmemDescriptor): start >> 4 [0,14), leading byte offset
// >> 4 [16,30), stride byte offset >> 4 [32,46), version 1 [46,48), layout [61,64) (0 none, 2 = 128-byte swizzle)
DSP_DEV uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
	return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) |
	       (1ull << 46) | ((uint64_t)layout << 61);
}
// instruction descriptor, kind::tf32 with fp32 accumulation: c_format F32 [4,6), a/b format TF32 [7,10) [10,13),
// a_major [15], b_major [16] (0 = K-major, 1 = MN-major), N >> 3 [17,23), M >> 4 [24,29)
DSP_DEV uint32_t umma_idesc(int M, int N, int a_mn, int b_mn) {
	return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
DSP_DEV void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
	asm volatile(
	    "{\n"
	    ".reg .pred p;\n"
	    "setp.ne.b32 p, %4, 0;\n"
	    "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
	    "}\n" ::"r"(d_tmem),
	    "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
	    : "memory");
}
DSP_DEV void umma_commit(uint64_t *bar) {
	asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
DSP_DEV void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
DSP_DEV void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
DSP_DEV void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// this warp's 32 lanes x 32 consecutive columns: v[i] = TMEM[lane][col + i]
DSP_DEV void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
	asm volatile(
	    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
	    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
	    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
	    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
	      "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
	      "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
	      "=r"(v[31])
	    : "r"(taddr)
	    : "memory");
}
DSP_DEV void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
	asm volatile(
	    "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
	    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
	    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
	      "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
	    : "r"(taddr)
	    : "memory");
}
// 16 consecutive outputs m0 .. m0 + 15 of an accumulator kept as two partial sums per group of BE outputs:
// columns [2 BE g, 2 BE g + BE) hold A_hi B_hi, the next BE columns A_hi B_lo + A_lo B_hi
template <int BE>
DSP_DEV void tmem_ld_sum16(uint32_t tbase, int m0, float (&o)[16]) {
	const uint32_t c = (uint32_t)(2 * BE * (m0 / BE) + (m0 % BE));
	uint32_t v1[16], v2[16];
	tmem_ld16(tbase + c, v1);
	tmem_ld16(tbase + c + BE, v2);
	tmem_wait_ld();
#pragma unroll
	for (int i = 0; i < 16; i++) o[i] = __uint_as_float(v1[i]) + __uint_as_float(v2[i]);
}
// mbarrier wait that gives up (trap: the launch fails instead of hanging the device) if the phase never completes
DSP_DEV void mbar_wait_bounded(uint64_t *bar, uint32_t parity) {
	uint32_t ok = 0;
	for (uint32_t spin = 0; spin < (1u << 24); spin++) {
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
DSP_DEV float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

DSP_DEV void mm_front_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kMmFront) : "memory"); }
DSP_DEV void mbar_arrive(uint64_t *bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }

// barriers: [0] [1] tile landed in buffer 0 / 1 (TMA bytes); [2] stage 1 done; [3] stage 2 done (operand buffers free);
// [4] accumulator D2 complete (tcgen05.commit); [6] D2 drained by the four store warps
template <int BE>
__global__ void __launch_bounds__(kMmThreads, 1) k_block_mm(const __grid_constant__ BlockMmArgs a) {
	extern __shared__ unsigned char mm_smem_raw[];
	const uint32_t raw = smem_u32(mm_smem_raw);
	const uint32_t base = (raw + 1023u) & ~1023u;              // the swizzled boxes need 1024-byte alignment
	unsigned char *sm = mm_smem_raw + (base - raw);
	constexpr int Be = BE, nb = kMmTile / BE;
	float *Chi = (float *)(sm + 3 * kMmBuf);
	uint64_t *bars = (uint64_t *)(sm + 3 * kMmBuf + 8 * Be * Be);
	uint32_t *tslot = (uint32_t *)(bars + 8);
	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

	if (tid == 0) {
		for (int i = 0; i < 6; i++) mbar_init(&bars[i], 1);
		mbar_init(&bars[6], kMmEpi / 32);
		mbar_init(&bars[7], kMmEpi / 32);
		mbar_fence_init();
	}
	for (int i = tid; i < 2 * Be * Be / 4; i += kMmThreads) ((float4 *)Chi)[i] = __ldg((const float4 *)a.consts + i);
	if (warp == 0) {
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tslot)), "r"((uint32_t)kMmTmemCols) : "memory");
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
	}
	fence_proxy_async();                                      // the constants are read by the tensor core (async proxy)
	tc_fence_before();
	__syncthreads();
	tc_fence_after();
	const uint32_t tmem = *(volatile uint32_t *)tslot;
	const long long per_plane = (long long)a.tiles_x * a.tiles_y;
	const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;          // a warp reaches the TMEM lanes of its quarter

	if (warp < kMmFront / 32) {
		// ---------------------------------------------------------------- front: split, both MMA stages, the transposing hand-over
		const uint32_t c_hi = base + 3 * kMmBuf, c_lo = c_hi + 4 * Be * Be, s_lo = base + 2 * kMmBuf;
		const uint32_t sbo_c = 32u * Be;
		const uint32_t id1n = umma_idesc(kMmTile, Be, 0, 0), id2n = umma_idesc(kMmTile, 2 * Be, 0, 0);
		auto issue_load = [&](long long tile, int b) {
			const int plane = (int)(tile / per_plane);
			const int rem = (int)(tile - plane * per_plane);
			const int ty = rem / a.tiles_x, tx = rem - ty * a.tiles_x;
			mbar_expect_tx(&bars[b], kMmBuf);
			for (int j = 0; j < 4; j++) tma_load3(sm + b * kMmBuf + j * (kMmBuf / 4), &a.in_map, tx * kMmTile + 32 * j, ty * kMmTile, plane, &bars[b]);
		};
		// One stage = for every group g of Be columns of K, over Be / 8 k-steps each:
		//     D[:, 2 Be g .. +2 Be) = A_hi [B_hi | B_lo]   (one MMA with N = 2 Be: the lo half of the block matrix follows the
		//                                                   hi half in shared memory, so the stacked operand is one descriptor)
		//     D[:, 2 Be g + Be .. +Be) += A_lo B_hi        (the two small terms share an accumulator)
		// and whoever reads the accumulator adds the two halves.  A_hi is streamed from shared memory once instead of twice:
		// the stage is bound by operand reads (N is small), so 32 MMAs instead of 48 is a third less time.
		// Both stages read their A operand in the same form (K-major, 128-byte swizzle, four boxes of 32 k), so the issue
		// loop is shared; it is fully unrolled with every operand offset a compile-time constant.
		const uint64_t bd_hi = umma_desc(c_hi, 128, sbo_c, 0);
		auto issue_stage = [&](uint32_t a_hi_addr, uint32_t d_tmem) {
			const uint64_t ad_hi = umma_desc(a_hi_addr, 16, 1024, 2), ad_lo = umma_desc(s_lo, 16, 1024, 2);
#pragma unroll
			for (int g = 0; g < nb; g++) {
#pragma unroll
				for (int term = 0; term < 2; term++) {
#pragma unroll
					for (int kk = 0; kk < Be / 8; kk++) {
						const int c0 = g * Be + 8 * kk;
						const uint64_t aoff = (uint64_t)(((c0 >> 5) * (kMmBuf / 4) + (c0 & 31) * 4) >> 4), boff = (uint64_t)((kk * 256) >> 4);
						if (term == 0) umma_tf32(d_tmem + 2 * Be * g, ad_hi + aoff, bd_hi + boff, id2n, kk != 0);      // A_hi [B_hi | B_lo]
						else umma_tf32(d_tmem + 2 * Be * g + Be, ad_lo + aoff, bd_hi + boff, id1n, 1);              // A_lo  B_hi, onto A_hi B_lo
					}
				}
			}
		};

		long long t = blockIdx.x;
		uint32_t ph = 0;
		if (tid == 0) {
			if (t < a.ntiles) issue_load(t, 0);
			if (t + gridDim.x < a.ntiles) issue_load(t + gridDim.x, 1);
		}
		for (int it = 0; t < a.ntiles; t += gridDim.x, it++) {
			const int b = it & 1;
			const uint32_t s_hi = base + b * kMmBuf;
			mbar_wait_bounded(&bars[b], (it >> 1) & 1);

			// split the tile: the MMA reads the top 19 bits of each fp32 word and ignores the rest (measured: masking the words
			// in place changes nothing), so the tile as the TMA left it IS the hi operand; lo = x - hi goes to the same offset of
			// the second buffer
			{
				float4 *X4 = (float4 *)(sm + b * kMmBuf), *S4 = (float4 *)(sm + 2 * kMmBuf);
#pragma unroll 4
				for (int i = 0; i < kMmBuf / 16 / kMmFront; i++) {
					const int f = tid + kMmFront * i;
					const float4 v = X4[f];
					const float4 h = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
					S4[f] = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
				}
			}
			fence_proxy_async();
			mm_front_sync();

			if (tid == 0) {                                                    // stage 1: contract along w
				tc_fence_after();
				issue_stage(s_hi, tmem);
				umma_commit(&bars[2]);
			}
			mbar_wait_bounded(&bars[2], ph);
			tc_fence_after();

			// D1 (lane = row) -> registers -> hi | lo -> the A operand of stage 2, A2[m = column][k = row], K-major in the
			// same swizzled form the TMA gives stage 1: four boxes of 32 k, rows of 128 bytes, 16-byte chunk ^= m % 8.
			// A warp (32 consecutive k, one m) stores one whole 128-byte row: no bank conflicts.
			{
				const int k = (warp & 3) * 32 + lane;
				unsigned char *Thi = sm + b * kMmBuf, *Tlo = sm + 2 * kMmBuf;
				const uint32_t kbox = (uint32_t)(warp & 3) * (kMmBuf / 4), kch = (uint32_t)lane >> 2, kin = ((uint32_t)lane & 3) * 4;
#pragma unroll 1
				for (int u = 0; u < 4; u++) {
					const int m0 = (warp >> 2) * 64 + u * 16;
					float v[16];
					tmem_ld_sum16<BE>(tmem + lane_base, m0, v);
					if (a.dbg && blockIdx.x == 0 && it == 0)
						for (int i = 0; i < 16; i++) a.dbg[k * kMmTile + m0 + i] = v[i];
#pragma unroll
					for (int i = 0; i < 16; i++) {
						const float x = v[i], h = tf32_hi(x);
						const uint32_t m = (uint32_t)(m0 + i);
						const uint32_t off = kbox + m * 128 + ((kch ^ (m & 7)) << 4) + kin;
						*(float *)(Thi + off) = x;
						*(float *)(Tlo + off) = x - h;
					}
				}
			}
			tc_fence_before();
			fence_proxy_async();
			mm_front_sync();

			if (tid == 0) {                                                    // stage 2: contract along h
				if (it >= 1) mbar_wait_bounded(&bars[6], (it - 1) & 1);                     // D2 drained by the store warps (previous tile)
				tc_fence_after();
				issue_stage(s_hi, tmem + 2 * kMmTile);
				umma_commit(&bars[4]);
				umma_commit(&bars[3]);
			}
			mbar_wait_bounded(&bars[3], ph);                                       // operand buffers free again
			ph ^= 1;
			if (tid == 0 && t + 2 * (long long)gridDim.x < a.ntiles) issue_load(t + 2 * (long long)gridDim.x, b);
		}
	} else {
		// ---------------------------------------------------------------- store warps: D2 (lane = image column, TMEM column =
		// image row) -> global while the front works on the next tile; a warp stores 128 contiguous bytes per row
		long long t = blockIdx.x;
		for (int it = 0; t < a.ntiles; t += gridDim.x, it++) {
			const int plane = (int)(t / per_plane);
			const int rem = (int)(t - plane * per_plane);
			const int ty = rem / a.tiles_x, tx = rem - ty * a.tiles_x;
			const int m = (warp & 3) * 32 + lane;
			const int col = tx * kMmTile + m;
			const int rows = a.H - ty * kMmTile;                                   // rows of this tile inside the plane (>= 128: all)
			float *op = a.out + ((long long)plane * a.H + (long long)ty * kMmTile) * a.W + col;
			mbar_wait_bounded(&bars[4], it & 1);
			tc_fence_after();
#pragma unroll 1
			for (int u = 0; u < 8; u++) {
				float v[16];
				tmem_ld_sum16<BE>(tmem + lane_base + 2 * kMmTile, 16 * u, v);
				if (a.dbg && blockIdx.x == 0 && it == 0)
					for (int i = 0; i < 16; i++) a.dbg[kMmTile * kMmTile + (16 * u + i) * kMmTile + m] = v[i];
				if (col < a.W) {
					float *p = op;
					if (16 * u + 16 <= rows) {
#pragma unroll
						for (int i = 0; i < 16; i++, p += a.W) *p = v[i] * a.scale;
					} else {
#pragma unroll
						for (int i = 0; i < 16; i++, p += a.W)
							if (16 * u + i < rows) *p = v[i] * a.scale;
					}
				}
				op += 16 * (long long)a.W;
			}
			tc_fence_before();
			__syncwarp();
			if (lane == 0) mbar_arrive(&bars[6]);
		}
	}
	tc_fence_before();
	__syncthreads();
	if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)kMmTmemCols) : "memory");
}
#endif  // DSP_GPU

// ------------------------------------------------------------------------------------------------ host side
// FFTW's unnormalised matrices: REDFT10 Y[n] = 2 sum_k x[k] cos(pi (k + 1/2) n / B); REDFT01 Y[n] = x[0] + 2 sum_{k>=1} x[k] cos(pi (n + 1/2) k / B)
static double block_matrix(int B, int kind, int n, int k) {
	const double pi = 3.14159265358979323846264338327950288;
	if (kind == DSP_KIND_REDFT10) return 2.0 * cos(pi * (k + 0.5) * n / B);
	return k == 0 ? 1.0 : 2.0 * cos(pi * (n + 0.5) * k / B);
}

bool block_mm_supports(int B) { return B == 8 || B == 16 || B == 32 || B == 64; }

// releases the cached block matrices (dsp_dct_cleanup)
void block_mm_cleanup();

#if DSP_GPU
static float host_tf32_hi(float x) {
	uint32_t u;
	memcpy(&u, &x, 4);
	u &= 0xFFFFE000u;
	memcpy(&x, &u, 4);
	return x;
}
static std::mutex g_mm_mu;
static std::map<long long, float *> g_mm_consts;      // (device, B, kind) -> hi | lo operand images on the device

static const float *block_mm_consts(int B, int kind, int Be, std::string &err) {
	const long long key = ((long long)rt_device() << 16) | (B << 4) | kind;
	std::lock_guard<std::mutex> lock(g_mm_mu);
	auto it = g_mm_consts.find(key);
	if (it != g_mm_consts.end()) return it->second;
	std::vector<float> h(2 * (size_t)Be * Be, 0.0f);
	for (int n = 0; n < Be; n++)
		for (int k = 0; k < Be; k++) {
			if (n / B != k / B) continue;                                  // B = 8: two blocks on the diagonal of a 16 x 16 matrix
			const double c = block_matrix(B, kind, n % B, k % B);
			const float hi = host_tf32_hi((float)c);
			const float lo = host_tf32_hi((float)(c - (double)hi));
			h[const_slot(n, k, Be)] = hi;
			h[(size_t)Be * Be + const_slot(n, k, Be)] = lo;
		}
	float *d = nullptr;
	if (!rt_ok(cudaMalloc(&d, h.size() * 4), err, "block matrix allocation")) return nullptr;
	if (!rt_ok(cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice), err, "block matrix upload")) { cudaFree(d); return nullptr; }
	g_mm_consts[key] = d;
	return d;
}
void block_mm_cleanup() {
	std::lock_guard<std::mutex> lock(g_mm_mu);
	int cur = 0;
	cudaGetDevice(&cur);
	for (auto &kv : g_mm_consts) {
		cudaSetDevice((int)(kv.first >> 16));
		cudaDeviceSynchronize();                     // a launch that still reads the matrix finishes first
		cudaFree(kv.second);
	}
	cudaSetDevice(cur);
	g_mm_consts.clear();
}
#else
void block_mm_cleanup() {}
#endif

// every B x B block of `nplanes` planes [H][W] (floats, W-contiguous): out = M X M^T per block, times `scale`.
// in == out is allowed (each tile is read whole before it is written, tiles are disjoint).
bool launch_block_mm_f32(const float *in, float *out, long long nplanes, int H, int W, int B, int kind, double scale, rt_stream st,
                         std::string &err, float *dbg) {
	if (!block_mm_supports(B)) { err = "block DCT by GEMM: block size must be 8, 16, 32 or 64"; return false; }
	if (H % B || W % B) { err = "block DCT by GEMM: the plane must be a whole number of blocks"; return false; }
	if (kind != DSP_KIND_REDFT10 && kind != DSP_KIND_REDFT01) { err = "block DCT by GEMM: REDFT10 or REDFT01"; return false; }
	if (nplanes < 1 || nplanes > 0x7fffffffLL) { err = "block DCT by GEMM: bad plane count"; return false; }
#if DSP_GPU
	if (((uintptr_t)in & 15) || ((uintptr_t)out & 3)) { err = "block DCT by GEMM: the input must be 16-byte aligned"; return false; }
	const int Be = B < 16 ? 16 : B;
	BlockMmArgs a;
	memset(&a, 0, sizeof(a));
	a.consts = block_mm_consts(B, kind, Be, err);
	if (!a.consts) return false;
	TmaView v;
	memset(&v, 0, sizeof(v));
	v.base = (void *)in; v.rank = 3;
	v.dims[0] = (unsigned long long)W; v.dims[1] = (unsigned long long)H; v.dims[2] = (unsigned long long)nplanes;
	v.strides[1] = (unsigned long long)W * 4; v.strides[2] = (unsigned long long)W * 4 * (unsigned long long)H;
	v.box[0] = 32; v.box[1] = kMmTile; v.box[2] = 1;
	if (!tma_encode(&a.in_map, v, err, true)) return false;
	a.out = out; a.dbg = dbg; a.H = H; a.W = W; a.nplanes = (int)nplanes; a.Be = Be;
	a.tiles_x = (W + kMmTile - 1) / kMmTile; a.tiles_y = (H + kMmTile - 1) / kMmTile;
	a.ntiles = (long long)a.tiles_x * a.tiles_y * nplanes;
	a.scale = (float)scale;
	static int sm_count[64] = {0};
	static unsigned long long attr_dev = 0;
	const int dev = rt_device() & 63;
	if (!sm_count[dev]) {
		int s = 0;
		if (cudaDeviceGetAttribute(&s, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || s < 1) s = 148;
		sm_count[dev] = s;
	}
	if (!((attr_dev >> dev) & 1ull)) {
		if (!rt_ok(cudaFuncSetAttribute(k_block_mm<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)block_mm_smem(16)), err, "smem attribute") ||
		    !rt_ok(cudaFuncSetAttribute(k_block_mm<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)block_mm_smem(32)), err, "smem attribute") ||
		    !rt_ok(cudaFuncSetAttribute(k_block_mm<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)block_mm_smem(64)), err, "smem attribute"))
			return false;
		attr_dev |= 1ull << dev;
	}
	const int grid = (int)(a.ntiles < sm_count[dev] ? a.ntiles : sm_count[dev]);
	if (Be == 16) k_block_mm<16><<<grid, kMmThreads, block_mm_smem(16), st>>>(a);
	else if (Be == 32) k_block_mm<32><<<grid, kMmThreads, block_mm_smem(32), st>>>(a);
	else k_block_mm<64><<<grid, kMmThreads, block_mm_smem(64), st>>>(a);
	return rt_ok(cudaGetLastError(), err, "block DCT GEMM launch");
#else
	// emulation (tests/emu): the same contraction as plain loops in double; the tensor-core path itself only exists on the GPU
	(void)st; (void)dbg;
	std::vector<double> M((size_t)B * B), tmp((size_t)B * B);
	for (int n = 0; n < B; n++)
		for (int k = 0; k < B; k++) M[(size_t)n * B + k] = block_matrix(B, kind, n, k);
	std::vector<float> blk((size_t)B * B);
	for (long long p = 0; p < nplanes; p++)
		for (int by = 0; by < H; by += B)
			for (int bx = 0; bx < W; bx += B) {
				const float *src = in + ((size_t)p * H + by) * W + bx;
				float *dst = out + ((size_t)p * H + by) * W + bx;
				for (int r = 0; r < B; r++)
					for (int c = 0; c < B; c++) blk[(size_t)r * B + c] = src[(size_t)r * W + c];
				for (int r = 0; r < B; r++)
					for (int n = 0; n < B; n++) {
						double s = 0;
						for (int k = 0; k < B; k++) s += (double)blk[(size_t)r * B + k] * M[(size_t)n * B + k];
						tmp[(size_t)r * B + n] = (double)(float)s;                 // D1 is held in fp32
					}
				for (int n = 0; n < B; n++)
					for (int c = 0; c < B; c++) {
						double s = 0;
						for (int k = 0; k < B; k++) s += tmp[(size_t)k * B + c] * M[(size_t)n * B + k];
						dst[(size_t)n * W + c] = (float)(s * (double)(float)scale);
					}
			}
	return true;
#endif
}

}  // namespace dsp
