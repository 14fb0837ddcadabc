/*
 * dsp_dct.h -- C ABI of libdspdct: B200 (sm_100a) DCT-II / DCT-III plans behind dspfun's FFTW call sites.
 *
 * This is the drop-in boundary.  dspfun's tools reach their transform through FFTW's r2r planner interface,
 * spelled fftw(call) -> fftwf_call / fftw_call by /root/reference/include/precision.h:115.  Each entry point
 * below cites the reference call site(s) it replaces.  shim/fftw3.h maps the FFTW names the tools use onto
 * these functions so spec.c / ispec.c / zoom.c / scan.c / motion.c / draw.c compile unchanged.
 *
 * Conventions
 *   prec      'f' (COEFF_PRECISION=F, float buffers) or 'd' (COEFF_PRECISION=D, double buffers).
 *             'l' (long double) has no GPU equivalent and is rejected.
 *   kinds     DSP_DCT_REDFT10 (unnormalised DCT-II,  Y_k = 2 sum_j X_j cos(pi (j+1/2) k / n)) and
 *             DSP_DCT_REDFT01 (unnormalised DCT-III, Y_k = X_0 + 2 sum_{j>=1} X_j cos(pi j (k+1/2) / n)),
 *             the only two kinds the reference ever plans.  Same numeric values as FFTW's fftw_r2r_kind.
 *   pointers  `in`/`out` given at plan time may be host or device pointers (detected).  Host buffers are
 *             staged through plan-owned device memory inside dsp_dct_execute; device buffers are transformed
 *             where they are.  dsp_dct_execute_dev is the device-resident, stream-ordered entry.
 *   errors    a failing call returns NULL / non-zero and leaves a message in dsp_dct_last_error().  There is
 *             no CPU fallback: without a usable CUDA device every plan call fails.
 */
#ifndef DSP_DCT_H
#define DSP_DCT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DSP_DCT_REDFT01 4
#define DSP_DCT_REDFT10 5

/* planner-effort flags are accepted and ignored (values match fftw3.h) */
#define DSP_DCT_MEASURE    (0U)
#define DSP_DCT_EXHAUSTIVE (1U << 3)
#define DSP_DCT_PATIENT    (1U << 5)
#define DSP_DCT_ESTIMATE   (1U << 6)

typedef struct dsp_dct_plan_s *dsp_dct_plan;

/* Replaces fftw(plan_many_r2r): spec/spec.c:63, spec/ispec.c:165, zoom/zoom.c:263, scan/scan.c:292,359,
 * motion/motion.c:535-538,549-552.  Same argument meaning as FFTW's advanced interface.  Supported layouts:
 * planar (istride == ostride == 1, any howmany/dist/embed) and channel-interleaved (stride == howmany,
 * dist == 1, the layout of every image tool).  rank 1..3.  Never touches the user buffers at plan time. */
dsp_dct_plan dsp_dct_plan_many(char prec, int rank, const int *n, int howmany,
                               void *in, const int *inembed, int istride, int idist,
                               void *out, const int *onembed, int ostride, int odist,
                               const int *kind, unsigned flags);

/* Same, with one more outermost batch level (nbatch independent copies of the whole plan_many problem,
 * ibdist / obdist elements apart): the batched-images configuration, one launch per pass for all images. */
dsp_dct_plan dsp_dct_plan_many_batched(char prec, int rank, const int *n, int howmany,
                                       void *in, const int *inembed, int istride, int idist,
                                       void *out, const int *onembed, int ostride, int odist,
                                       const int *kind, unsigned flags,
                                       int nbatch, ptrdiff_t ibdist, ptrdiff_t obdist);

/* Replaces fftw(plan_r2r_2d): applybasis/draw.c:74. */
dsp_dct_plan dsp_dct_plan_2d(char prec, int n0, int n1, void *in, void *out, int kind0, int kind1, unsigned flags);

/* Replaces fftw(execute): spec/spec.c:64, spec/ispec.c:166, zoom/zoom.c:264, scan/scan.c:293,407,447,
 * motion/motion.c:641,753, applybasis/draw.c:75.  Transforms the buffers given at plan time; returns when
 * the result is visible to the host (host buffers) or the work is complete (device buffers). */
void dsp_dct_execute(dsp_dct_plan p);

/* New-array execute on host buffers with the plan's layout (fftw_execute_r2r analogue).  0 on success. */
int dsp_dct_execute_host(dsp_dct_plan p, void *in, void *out);

/* Device-resident execute: d_in / d_out are device pointers with the plan's layout; kernels are enqueued on
 * `stream` (a cudaStream_t, NULL = default stream) and the call returns without synchronising.  0 on success. */
int dsp_dct_execute_dev(dsp_dct_plan p, void *d_in, void *d_out, void *stream);

/* Replaces fftw(destroy_plan): spec/spec.c:65, spec/ispec.c:167, zoom/zoom.c:265, scan/scan.c:294,
 * motion/motion.c:834, applybasis/draw.c:76. */
void dsp_dct_destroy(dsp_dct_plan p);

/* Replace fftw(alloc_real) / fftw(free) (spec/spec.c:59,143, motion/motion.c:500,826): pinned host memory, so
 * dsp_dct_execute's staging copies run at full PCIe rate. */
/* Replaces fftw(plan_with_nthreads)(n): motion/motion.c:485-486, scan/scan.c:289-290 (`--fftw-threads n`).  The knob the
 * reference uses for CPU threads selects the number of GPUs of this process here: rank-3 float plans over one contiguous
 * [D][H][W] volume (motion -b 0x0x0), executed with HOST buffers (dsp_dct_execute / dsp_dct_execute_host), created after the
 * call run slab-sharded over min(n, visible GPUs) devices: frame slabs in, per-frame 2-D passes whose last pass stores
 * straight into the other GPUs' column buffers over NVLink (the all-to-all of SURVEY 8e fused into the transform), temporal
 * pass, strided copy-out -- one exchange per transform, no NCCL, no second process.  Everything else (rank 1 / 2, double,
 * embedded sub-boxes, fused stages, D or H not divisible by the GPU count, no peer access) runs on one GPU as before.
 * dsp_dct_plan_ngpus reports how many GPUs the plan's last host execution used. */
void dsp_dct_plan_with_ngpus(int n);
int dsp_dct_plan_ngpus(dsp_dct_plan p);

void *dsp_dct_alloc(size_t bytes);
void dsp_dct_free(void *p);

/* Replace fftw(cleanup) (zoom/zoom.c:266, motion/motion.c:836): releases cached twiddle tables. */
void dsp_dct_cleanup(void);

/* Thread-local message of the last failing call on this thread ("" if none). */
const char *dsp_dct_last_error(void);

/* Number of this library's kernels launched so far by this process (bench.py's gpu_launches). */
unsigned long long dsp_dct_launch_count(void);

/* 0 for the product (CUDA) build.  1 for the sequential host emulation of the same kernel sources that the CPU test
 * suite compiles (tests/emu); the package's loader refuses such a library. */
int dsp_dct_is_emulation(void);

/* Per-pass device timing (CUDA events on the launching stream around every pass of every execute while enabled).
 * dsp_dct_pass_stat blocks until the recorded executes finished, returns the statistics accumulated since the
 * last call for pass `i` and resets them.  Returns non-zero when `i` is out of range. */
typedef struct {
	int is_row;          /* 1: contiguous-axis (row) kernel, 0: strided-axis (column) kernel */
	int axis, n;         /* transformed axis and its length */
	int grid, block;     /* launch geometry */
	size_t smem_bytes;   /* dynamic shared memory per CTA */
	int launches;        /* launches accumulated */
	double ms_total;     /* their summed device time */
	double samples;      /* scalar samples one launch reads and writes */
	int split_panels;    /* > 0: the pass runs as two sub-kernels per column panel (this many panels per plane);
	                        `launches` then counts passes, not sub-kernel launches */
	int kernel_launches; /* timed launches behind `launches`: a pass of the chunked (L2-resident) schedule is one launch
	                        per chunk of batch elements / frames, `launches` still counts whole passes */
} dsp_dct_pass_stat;
int dsp_dct_profile(dsp_dct_plan p, int enable);
int dsp_dct_num_passes(dsp_dct_plan p);
int dsp_dct_pass_stat_get(dsp_dct_plan p, int i, dsp_dct_pass_stat *out);

/* ---- fused pointwise stages (run inside the first / last pass; no extra trip through HBM) ---------------- */

/* Multiply every element by load_scale as it is read and by store_scale as it is written.
 * scan/scan.c:296-298 (coeffs /= w*h*4) is store_scale = 1/(4wh) on the forward plan. */
int dsp_dct_fuse_scale(dsp_dct_plan p, double load_scale, double store_scale);

/* Segmented output of the plan's LAST pass (no reference counterpart: it replaces the pack + all-to-all of the
 * slab-sharded motion volume).  The last axis transformed (length n = nseg * seg_rows) is cut into nseg
 * runs of seg_rows positions; run g is stored at bases[g] (device pointers, typically peer-mapped buffers of the
 * other GPUs) with the plan's output row stride (row_stride elements when non-zero), plus the plan's outer offset --
 * whose stride becomes outer_stride elements when that is non-zero.  Float plans whose last pass is a one-kernel strided-axis pass over full,
 * 16-byte aligned column tiles; returns non-zero (dsp_dct_last_error) otherwise.  nseg = 0 turns it off. */
int dsp_dct_set_output_segments(dsp_dct_plan p, int nseg, int seg_rows, void *const *bases, long long outer_stride,
                                long long row_stride);

enum { DSP_SPEC_SCALE_LOG = 0, DSP_SPEC_SCALE_LINEAR = 1 };
enum { DSP_SPEC_SIGN_ABS = 0, DSP_SPEC_SIGN_SHIFT = 1, DSP_SPEC_SIGN_SATURATE = 2, DSP_SPEC_SIGN_RETAIN = 3 };
enum { DSP_SPEC_RANGE_ONE = 0, DSP_SPEC_RANGE_DC = 1, DSP_SPEC_RANGE_DCS = 2 };

/* spec's post-transform loops (spec/spec.c:66-139) fused into the last pass of a rank-2 REDFT10 plan:
 * DC capture, row0/col0 /sqrt2, /(2wh), *gain, range (one|dc|dcs), scale (log|linear), sign (abs|shift|
 * saturate|retain).  `gain` is the already-resolved multiplier (spec/spec.c:81-87).  After execute the DC
 * property values (spec/spec.c:66-68) are available from dsp_dct_spec_dc(). */
typedef struct {
	int scaletype, signtype, rangetype;
	double gain;
} dsp_spec_params;
int dsp_dct_fuse_spec(dsp_dct_plan p, const dsp_spec_params *sp);
/* copies the d DC values of the last execute (blocks until that execute finished).  0 on success. */
int dsp_dct_spec_dc(dsp_dct_plan p, double *dc, int d);

/* ispec's pre-transform loops (spec/ispec.c:84-163) fused into the first pass of a rank-2 REDFT01 plan.
 * max[z] is the already-resolved range (spec/ispec.c:119-134, before the log1p of :138); dc[] is only read
 * when preserve_dc is set.  signmap (device or host pointer to u8 [h][w][d], or NULL) is spec/ispec.c:87-98. */
typedef struct {
	int scaletype, signtype;
	double gain;
	double max[4];
	int preserve_dc;
	double dc[4];
	const unsigned char *signmap;
} dsp_ispec_params;
int dsp_dct_fuse_ispec(dsp_dct_plan p, const dsp_ispec_params *ip);

/* ---- scan: progressive reconstruction (scan/scan.c:292-298, 352-459) ------------------------------------------
 * A scan session keeps the coefficient plane, the scan-order index map and the running sum on the GPU:
 *   create : REDFT10 x REDFT10 of the [h][w][d] interleaved pixels, fused 1/(4wh) (scan.c:292-298);
 *            sum[y][x][z] = DC[z] (scan.c:381-383).  index_map[y][x] is the scan step that delivers
 *            coefficient (y, x) -- the "index" serialisation of scan/scan_precomputed.c:133-153.
 *   frame  : one output frame: keep coefficients with lo <= index < hi, clear DC (scan.c:429-445), REDFT01 x
 *            REDFT01 (scan.c:447), sum += image, frame = sum (scan.c:451-456).  The mask rides in the inverse's
 *            first pass and the accumulation in its last pass; `frame` (host, T[h][w][d]) may be NULL.
 *   coeffs / sum : copy the normalised coefficient plane / current sum to host buffers. */
typedef struct dsp_scan_s *dsp_scan;
dsp_scan dsp_scan_create(char prec, int h, int w, int d, const void *pixels, const int32_t *index_map);
int dsp_scan_frame(dsp_scan s, int lo, int hi, void *frame);
int dsp_scan_coeffs(dsp_scan s, void *coeffs);
int dsp_scan_sum(dsp_scan s, void *sum);
void dsp_scan_destroy(dsp_scan s);

/* ---- motion: 3-D DCT -> coefficient-space filters -> 3-D inverse DCT of one plane block ----------------------
 * (motion/motion.c:525-573 plans and constants, :617-788 per-block body; options -p/--bandpass, -D/--damp, -B/--boost,
 * --threshold, -q/--quant, --preserve-dc, -s/--size.)  One block = the staging buffer the reference fills per
 * (plane, spatial block, temporal block): [minbuf.d][minbuf.h][minbuf.w] pels, 8-bit or float32, with
 * minbuf = max(block, scaled).  The forward REDFT10^3 runs over `block`, the inverse REDFT01^3 over `scaled` in the
 * same padded box, so scaled != block resamples by spectral zero-pad / crop.  Fused: 8-bit load in the first pass
 * of the forward transform; zero-pad/crop + normalise + band-pass + threshold + preserve-dc + quantise +
 * de-normalise in the first pass of the inverse; scale + clamp + lround + 8-bit store in its last pass.
 * Spectrogram output / input (--spec / --ispec) run as one pointwise kernel in place of the inverse / forward transform.
 * Not covered, rejected by the host mirror (sequential or host-side in the reference): --coeff-limit (repeated qsort),
 * --eval (FFmpeg expression VM), --dither (Floyd-Steinberg error diffusion), linear-light transfer curves (libavutil).
 * All dims are (d, h, w). */
typedef struct {
	int block[3], scaled[3];
	int float_pixels;             /* 0: 8-bit pels; 1: float32 pels in [0,1] (motion.c:621-624, 773) */
	double damp, boost;           /* 1 = off */
	int bp_begin[3], bp_end[3];   /* band-pass box, 0 <= begin <= end <= min(block, scaled) */
	double threshold_min, threshold_max;   /* as given on the command line; max == 0 = off */
	double quant;                 /* 0 = off */
	int preserve_dc;              /* 0 none, 1 dc, 2 grey */
	int spec;                     /* --spec:  write the (filtered) coefficients out as a spectrogram instead of inverting them
	                                 (motion.c:755-776): 0 none, 1 abs, 2 shift, 3 flat, 4 copy */
	int ispec;                    /* --ispec: the pels are a spectrogram, no forward transform (motion.c:627-637):
	                                 0 none, 2 shift, 3 flat, 4 copy */
} dsp_motion_params;
enum { DSP_MOTION_SPEC_NONE = 0, DSP_MOTION_SPEC_ABS = 1, DSP_MOTION_SPEC_SHIFT = 2, DSP_MOTION_SPEC_FLAT = 3, DSP_MOTION_SPEC_COPY = 4 };
typedef struct dsp_motion_s *dsp_motion;
dsp_motion dsp_motion_create(char prec, const dsp_motion_params *mp);

/* The same three stages on caller-owned plans (what dsp_motion_create does to its own): the slab-sharded full-volume
 * transform (motion -b 0x0x0 over several GPUs) puts them on the plans either side of its exchange.
 *   dsp_dct_fuse_pel_load   first pass (along the contiguous axis) reads 8-bit pels, or float pels * 255 (motion.c:618-624)
 *   dsp_dct_fuse_motion_coeff  first pass of an inverse plan applies motion.c:617,644-751 as it loads (on a forward
 *                           plan the same map rides in the last pass's store instead).  flat_w > 0: the
 *                           plan is the temporal pass over a [D][slice of flattened h*w] array; the axis index is z and
 *                           (y, x) = divmod(flat_base + column, flat_w).  d_counter (device, may be NULL) counts the
 *                           non-zero coefficients after --quant.
 *   dsp_dct_fuse_pel_store  last pass (along the contiguous axis) scales, clamps, rounds and stores 8-bit pels, or
 *                           float pels / 255 (motion.c:757-776); an 8-bit store gives the plan a work buffer. */
int dsp_dct_fuse_pel_load(dsp_dct_plan p, int float_pixels);
int dsp_dct_fuse_motion_coeff(dsp_dct_plan p, const dsp_motion_params *mp, unsigned long long *d_counter, int flat_w,
                              long long flat_base);
/* The same coefficient stage as ONE sweep over a contiguous device volume of coefficients [minbuf.d][minbuf.h][minbuf.w]
 * (block == scaled: the whole volume), in place, for callers that run plain plans around it.  Measured on the
 * 256 x 1080 x 1920 volume the separate sweep is cheaper than carrying the stage in the temporal pass. */
int dsp_motion_coeff_stage(char prec, const dsp_motion_params *mp, void *d_coeffs, unsigned long long *d_counter, void *stream);
/* The same over the temporal layout of a slab-sharded volume: a [D][ncols] array whose columns are the flattened (y, x)
 * positions flat_base .. flat_base + ncols of frames of width flat_w (the coordinates dsp_dct_fuse_motion_coeff's flat mode uses). */
int dsp_motion_coeff_stage_flat(char prec, const dsp_motion_params *mp, void *d_coeffs, int D, long long ncols, int flat_w, long long flat_base,
                                unsigned long long *d_counter, void *stream);
int dsp_dct_fuse_pel_store(dsp_dct_plan p, const dsp_motion_params *mp);
/* host staging buffers (pels_out may equal pels_in); *coeffs_coded += non-zero coefficients after --quant */
int dsp_motion_block(dsp_motion m, const void *pels_in, void *pels_out, unsigned long long *coeffs_coded);
/* device-resident: same layout, device pointers, enqueued on `stream` */
int dsp_motion_block_dev(dsp_motion m, const void *d_pels_in, void *d_pels_out, void *stream);
void dsp_motion_destroy(dsp_motion m);

/* ---- motion over a block-tiled volume (`motion -b BWxBHxBD [--quant q]` with block == scaled): the block DCTs are
 * three per-axis dsp_dct plans over the whole [D][H][W] volume (dspfun_b200/motion.py: MotionTiled builds them);
 * these two entry points are the stages between and after them, one pass each over device memory.
 *   dsp_block_quant    : in place: c *= nf (motion.c:644-647), c = round(c / quantizer) * quantizer and count the
 *                        non-zero results into *d_count (:740-744; quantizer = 0 skips this step), c /= nf (:748-751),
 *                        with nf from the in-block indices of each element.
 *   dsp_block_store_u8 : pel = c * scale, clamp to [0, 255], lround, 8-bit store (:757-776). */
int dsp_block_quant(char prec, void *d_coeffs, int D, int H, int W, int bd, int bh, int bw, double quantizer,
                    unsigned long long *d_count, void *stream);
int dsp_block_store_u8(char prec, const void *d_coeffs, unsigned char *d_pels, long long n, double scale, void *stream);
/*   dsp_block_dquant   : float volumes, block depth bd = 2, 4, 8 or 16, to be called between the spatial transforms: in ONE
 *                        pass per pixel and group of bd frames: REDFT10 along d, the stage of dsp_block_quant on the now
 *                        complete 3-D coefficients, REDFT01 along d (replaces d-plan + dsp_block_quant + inverse d-plan). */
int dsp_block_dquant(void *d_coeffs, int D, int H, int W, int bd, int bh, int bw, double quantizer, unsigned long long *d_count, void *stream);

/* ---- the same block-tiled volume as ONE session behind the C ABI (what dspfun_b200/motion.py: MotionTiled does from
 * Python): `motion -b BWxBHxBD [--quant q]` with block == scaled over a [D][H][W] plane of 8-bit pels in device memory
 * (D, H, W whole numbers of blocks).  create builds the plans (square spatial blocks of 8 / 16 / 32 / 64 take
 * dsp_block_dct2d for both spatial axes and keep a plan only for the d axis; other shapes use three per-axis plans) and a
 * float work volume; process_dev enqueues pels -> block DCT-II -> normalise / quantise / de-normalise -> block DCT-III ->
 * clamp, round, 8-bit pels on `stream` (d_pels_out may equal d_pels_in).  If coeffs_coded is not NULL the call
 * synchronises the stream and adds the number of non-zero quantised coefficients (motion.c:740-744). */
typedef struct dsp_motion_tiled_s *dsp_motion_tiled;
dsp_motion_tiled dsp_motion_tiled_create(int D, int H, int W, int bd, int bh, int bw, double quant);
int dsp_motion_tiled_process_dev(dsp_motion_tiled t, const unsigned char *d_pels_in, unsigned char *d_pels_out,
                                 unsigned long long *coeffs_coded, void *stream);
void dsp_motion_tiled_destroy(dsp_motion_tiled t);

/* ---- 2-D block DCT on the tensor cores: every B x B block (B = 8, 16, 32 or 64) of `nplanes` float planes [H][W]
 * (device memory, W contiguous, H and W whole numbers of blocks, d_in 16-byte aligned; d_out may equal d_in) is
 * replaced by its unnormalised FFTW transform along both axes, times `scale`: kind = DSP_DCT_REDFT10 or
 * DSP_DCT_REDFT01, i.e. what fftw_plan_many_r2r(rank 2, {B, B}) computes per block in motion's block loop
 * (motion/motion.c:559-564, :613-641, :753 with -b BxBx1, or the two spatial axes of -b BxBxD) and what applybasis'
 * DCT contraction (applybasis/applybasis.c:410-448) evaluates with O(N^4) loops.  One pass over the planes: a
 * 128 x 128 tile is contracted along w and then along h by tcgen05 MMAs (TF32 operands split hi + lo, three MMAs
 * per product, fp32 accumulation in tensor memory), so the result carries float accuracy (~1e-6 relative).
 * prec must be 'f'.  Other block sizes / double: use the per-axis plans (dspfun_b200/motion.py MotionTiled). */
int dsp_block_dct2d(char prec, const void *d_in, void *d_out, long long nplanes, int H, int W, int B, int kind, double scale, void *stream);
/* bring-up / test hook: as above for one launch, also copying the first tile's two accumulators ([2][128][128] floats:
 * after the w contraction, after the h contraction transposed) to d_debug */
int dsp_block_dct2d_debug(const void *d_in, void *d_out, long long nplanes, int H, int W, int B, int kind, double scale, void *stream, float *d_debug);

/* ---- zoom: DCT-domain resampling of an RGB image (zoom/zoom.c:263-266 forward plan, :36-68 scaled basis,
 * :361-375 synthesis).  create: REDFT10 x REDFT10 of the [h][w][3] interleaved pixels, kept on the GPU.
 * frame: one output view.  Scale is num/den per axis; (vx, vy) the view offset in output samples; (vw, vh) the
 * view size (0 = the whole scaled image, zoom.c:286-289).  basis: 0 interpolated (default), 1 centered, 2 native.
 * A native basis with integer scaled size and zero offset is a spectral zero-pad / crop and runs as one inverse
 * DCT; the interpolated basis (the default) with an integer scaled size runs as four phase-shifted inverse DCTs;
 * every other case runs the reference's separable cosine synthesis as dense contractions.
 * out: host buffer [vh][vw][3] of the coefficient type, the values zoom hands to ffapi_setpelf (zoom.c:393-400). */
typedef struct {
	int basis;
	double xscale_num, xscale_den, yscale_num, yscale_den;
	double vx, vy;
	int vw, vh;
} dsp_zoom_params;
typedef struct dsp_zoom_s *dsp_zoom;
dsp_zoom dsp_zoom_create(char prec, int h, int w, const void *pixels);
int dsp_zoom_view_size(dsp_zoom z, const dsp_zoom_params *zp, int *vw, int *vh);
int dsp_zoom_frame(dsp_zoom z, const dsp_zoom_params *zp, void *out);
/* which path the last frame took: 1 = inverse-DCT fast path (native basis), 2 = four phase-shifted inverse DCTs
 * (interpolated basis with an integer scaled size, any offset), 0 = dense synthesis (SIMT GEMM, FP64 accumulation),
 * 3 = dense synthesis on the tensor cores (float sessions: two 3 x TF32 GEMMs per channel) */
int dsp_zoom_last_path(dsp_zoom z);
void dsp_zoom_destroy(dsp_zoom z);

#ifdef __cplusplus
}
#endif
#endif
