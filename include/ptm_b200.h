/*
 * ptm_b200.h -- C ABI of the B200 (sm_100a) PTM fit / quantise / relight library.
 *
 * This is the boundary a C host (the ptmlib.h drop-in in rti_b200/csrc/ptmlib_host.c,
 * the command-line mains) or any FFI (ctypes in rti_b200/_lib.py) binds to.  Plain
 * pointers and sizes only.  Every entry names the reference interface it stands
 * in for; paths are relative to the reference tree (cceh/rti).
 *
 * Conventions
 *   - All functions return 0 on success and a non-zero code on failure;
 *     ptmb_last_error() then returns a message for the calling thread.  There
 *     is no CPU fallback: without a usable CUDA device every call fails.
 *   - `*_dev` pointers are device memory on the context's GPU.  Work is
 *     enqueued on the context's stream and is asynchronous unless noted.
 *   - A "stack" is the decoded image stack of the reference's encoder
 *     (rti-builder/ptm-encoder.c:245-264): samples [light][pixel][3], pixels in
 *     stored (already flipped) row order.  `image_stride` is the distance between two
 *     lights' images: in BYTES for the stack entry points, in SAMPLES for ptmb_fit_channel
 *     (which mirrors the reference's sample-indexed signature).
 *   - Coefficient order everywhere: u^2, v^2, uv, u, v, 1 (rti-builder/ptmlib.h:37-55).
 *   - "keys": the 12 running extrema a fit leaves behind, as order-preserving
 *     int32 images of floats (keys[k] = key(-min_k), keys[6+k] = key(max_k)), so
 *     that ONE integer max-reduction (atomicMax on a GPU, an int32 MAX allreduce
 *     across GPUs) merges partial results exactly.
 */
#ifndef PTM_B200_H
#define PTM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PTMB_COEFFICIENTS 6
#define PTMB_KEYS 12

/* sample types of a stack */
#define PTMB_U8  0
#define PTMB_U16 1
#define PTMB_U32 2

/* synthetic stack content (rti_b200/csrc/ptm_synth.cu) */
#define PTMB_SYNTH_YCBCR 0
#define PTMB_SYNTH_RGB   1

typedef struct ptmb_ctx ptmb_ctx_t;

/* What ptm_scale_coefficients writes into the PTM header
 * (rti-builder/ptmlib.h:126-127, rti-builder/ptmlib.c:858-863). */
typedef struct {
    float scale[PTMB_COEFFICIENTS];
    int32_t bias[PTMB_COEFFICIENTS];
} ptmb_scale_bias_t;

const char *ptmb_last_error (void);
int ptmb_device_count (int *count);

/* One context per (process, GPU).  `stream` is a cudaStream_t owned by the caller (e.g. the
 * framework's current stream); NULL is the CUDA default stream. */
int ptmb_create (int device, void *stream, ptmb_ctx_t **ctx);
void ptmb_destroy (ptmb_ctx_t *ctx);
int ptmb_set_stream (ptmb_ctx_t *ctx, void *stream);
int ptmb_synchronize (ptmb_ctx_t *ctx);
int ptmb_sm_count (const ptmb_ctx_t *ctx);
/* the context's cudaStream_t and device ordinal (so that a host can record its own events on the stream) */
void *ptmb_stream (const ptmb_ctx_t *ctx);
int ptmb_device (const ptmb_ctx_t *ctx);

/* Device memory helpers for hosts that have no allocator of their own. */
int ptmb_malloc (ptmb_ctx_t *ctx, size_t bytes, void **dev);
int ptmb_free (ptmb_ctx_t *ctx, void *dev);
int ptmb_upload (ptmb_ctx_t *ctx, void *dev, const void *host, size_t bytes);
int ptmb_download (ptmb_ctx_t *ctx, void *host, const void *dev, size_t bytes);
int ptmb_host_register (const void *host, size_t bytes);
int ptmb_host_unregister (const void *host);
/* Pinned host memory usable from every GPU of the process (cudaHostAllocPortable).  write_combined: faster for the
 * GPU to read over PCIe on some hosts, but very slow for the CPU to READ -- only for buffers the CPU just fills. */
int ptmb_host_alloc (size_t bytes, int write_combined, void **host);
int ptmb_host_free (void *host);

/* The 6 x n_lights pseudo-inverse from ptm_svd (rti-builder/ptmlib.c:630-689), row
 * major, HOST pointer.  Copied to the device; valid for all later fits. */
int ptmb_set_matrix (ptmb_ctx_t *ctx, const float *M, unsigned n_lights);

/* Which kernels serve the fits (both produce the reference's results within the same bars):
 *   PTMB_ENGINE_FFMA    FP32 FMA streaming kernels (ptm_fit_tma*.cu)
 *   PTMB_ENGINE_MMA     byte-wise integer GEMM on the tensor cores (ptm_fit_mma.cu): the
 *                       pseudo-inverse as 32-bit fixed point, exact integer dot products, one rounding
 *   PTMB_ENGINE_DEFAULT the library's choice per format (environment PTM_FIT_ENGINE=ffma|mma overrides)
 * Layouts an engine cannot take (alignment, light count) fall through to the other GPU kernels. */
#define PTMB_ENGINE_DEFAULT 0
#define PTMB_ENGINE_FFMA    1
#define PTMB_ENGINE_MMA     2
int ptmb_set_fit_engine (ptmb_ctx_t *ctx, int engine);
/* launches of the tensor-core kernel by this process so far (tests use it to prove which engine ran) */
unsigned long long ptmb_mma_launches (void);
/* Host only (no GPU needed): the B operand of the tensor-core engine for a 6 x n_lights matrix --
 * balanced base-256 digits of round(M * 2^F_k), as the shared-memory image
 * [n_lights/16 chunks][4 column groups][8 columns][16 lights]; column 4k+j = digit j of coefficient k,
 * column 24 = ones.  scale[k] = 2^-F_k.  image_bytes >= 32 * kpad. */
int ptmb_mma_pack_digits (const float *M, unsigned n_lights, int8_t *image, size_t image_bytes,
                          float *scale, unsigned *kpad);

/* Reset keys to the reference's initial extrema (min = FLT_MAX, max = -FLT_MAX,
 * rti-builder/ptmlib.c:826-827). */
int ptmb_keys_init (ptmb_ctx_t *ctx, int32_t *keys_dev);

/* K1, LRGB.  Fuses, for every pixel of an interleaved 3-channel stack:
 *   ptm_fit_poly_jsample on channel 0      rti-builder/ptmlib.c:692-725
 *   ptm_cbcr_avg on all three channels     rti-builder/ptmlib.c:762-802
 *   the LRGB colour + normalise loop       rti-builder/ptm-encoder.c:327-353
 *   phase A of ptm_scale_coefficients      rti-builder/ptmlib.c:826-847 (into keys)
 * stack_dev: samples [n_lights][pixels][3] of `sample_type`, image_stride in BYTES.
 * coeffs_dev: float [pixels][6] (normalised), rgb_dev: uint8 [pixels][3].
 * keys_dev is merged into (call ptmb_keys_init first). */
int ptmb_fit_lrgb (ptmb_ctx_t *ctx, const void *stack_dev, int sample_type, size_t image_stride,
                   size_t pixels, float *coeffs_dev, uint8_t *rgb_dev, int32_t *keys_dev);

/* Variants of the LRGB-shaped entries (ptmb_fit_lrgb, ptmb_encode_lrgb_dev, ptmb_encode_host* with
 * format_planes = 1); flags, 0 = plain PTM_FORMAT_LRGB.  uint8 stacks only.
 *   PTMB_FIT_LUM        PTM_FORMAT_LUM (rti-builder/ptm-encoder.c:286-305): fit on Y, the colour block is the
 *                       averaged YCbCr triplet as ptm_cbcr_avg leaves it, no colour conversion, no normalisation
 *   PTMB_FIT_RGB_INPUT  the stack holds RGB triplets (not the YCbCr the reference's decoder delivers,
 *                       rti-builder/ptm-encoder.c:188): every sample is converted to YCbCr first with libjpeg's
 *                       integer JFIF formula, fused into K1 */
#define PTMB_FIT_LUM       1
#define PTMB_FIT_RGB_INPUT 2
int ptmb_set_fit_mode (ptmb_ctx_t *ctx, int flags);

/* K1, RGB formats: three fits (rti-builder/ptm-encoder.c:272-284) in one pass.
 * coeffs_dev: float [3][pixels][6] with `plane_stride` floats between planes.  Only plane 0
 * feeds the extrema, as in the reference (rti-builder/ptmlib.c:831). */
int ptmb_fit_rgb (ptmb_ctx_t *ctx, const void *stack_dev, int sample_type, size_t image_stride,
                  size_t pixels, float *coeffs_dev, size_t plane_stride, int32_t *keys_dev);

/* ptm_fit_poly_jsample / ptm_fit_poly_uint in their general form
 * (rti-builder/ptmlib.c:692-760): one channel, arbitrary pixel stride (in samples),
 * image_stride in SAMPLES (= height * width * pixel_stride in the reference). */
int ptmb_fit_channel (ptmb_ctx_t *ctx, const void *samples_dev, int sample_type, size_t pixel_stride,
                      size_t image_stride, size_t pixels, float *coeffs_dev);

/* ptm_cbcr_avg alone (rti-builder/ptmlib.c:762-802): uint8 stack [n][pixels][3]. */
int ptmb_chroma_avg (ptmb_ctx_t *ctx, const uint8_t *stack_dev, size_t image_stride,
                     size_t pixels, unsigned n_lights, uint8_t *avg_dev);

/* The LRGB colour + normalise loop alone (rti-builder/ptm-encoder.c:327-353), in place. */
int ptmb_lrgb_colour_normalise (ptmb_ctx_t *ctx, uint8_t *ycc_to_rgb_dev, float *coeffs_dev, size_t pixels);

/* Phase A of ptm_scale_coefficients alone (rti-builder/ptmlib.c:826-847). */
int ptmb_minmax (ptmb_ctx_t *ctx, const float *coeffs_dev, size_t pixels, int32_t *keys_dev);

/* K2: phases B and C of ptm_scale_coefficients (rti-builder/ptmlib.c:858-880).
 * Derives scale / bias from keys_dev on the device (written to sb_dev if non-NULL) and
 * quantises one coefficient plane: bytes_dev uint8 [pixels][6]. */
int ptmb_quantise (ptmb_ctx_t *ctx, const float *coeffs_dev, size_t pixels, const int32_t *keys_dev,
                   uint8_t *bytes_dev, ptmb_scale_bias_t *sb_dev);

/* K2 for a float plane that is scratch (the fused encode: nobody reads the floats afterwards).
 * Same bytes as ptmb_quantise, but the plane is walked from its end -- the part K1 wrote last,
 * still dirty in L2 -- and each line is discarded from L2 after use, so it is never written
 * back to HBM.  The contents of coeffs_dev are UNDEFINED afterwards. */
int ptmb_quantise_consume (ptmb_ctx_t *ctx, float *coeffs_dev, size_t pixels, const int32_t *keys_dev,
                           uint8_t *bytes_dev, ptmb_scale_bias_t *sb_dev);

/* K3: the pixel loop of ptm_write_jpeg (rti-builder/ptmlib.c:553-562,587-621) for `n_dirs`
 * light directions at once.  uv: HOST array of n_dirs (u, v) pairs.  out_dev: uint8
 * [n_dirs][height][width][3], top row first (stored row height-1-y, :587-589).
 * LRGB: coef_dev [pixels][6] + rgb_dev [pixels][3].  RGB: three coefficient planes. */
/* Arithmetic of K3:
 *   PTMB_RELIGHT_EXACT  every intermediate rounding of the reference's loop is reproduced: rasters are
 *                       bit-identical to the reference decoder's (issue bound, ~60 % of the HBM roofline)
 *   PTMB_RELIGHT_FAST   fused multiply-adds, "/ 255" folded into the coefficients, integer CLIP: every byte is
 *                       within +-1 of the reference's (the tolerance BASELINE.json states for relit output)
 *   PTMB_RELIGHT_AUTO   (default) exact below 4 directions per call, fast for sweeps; the environment
 *                       variable PTM_RELIGHT=exact|fast overrides AUTO. */
#define PTMB_RELIGHT_AUTO  0
#define PTMB_RELIGHT_EXACT 1
#define PTMB_RELIGHT_FAST  2
int ptmb_set_relight_mode (ptmb_ctx_t *ctx, int mode);
int ptmb_relight_lrgb (ptmb_ctx_t *ctx, const uint8_t *coef_dev, const uint8_t *rgb_dev,
                       size_t width, size_t height, const float *scale, const int32_t *bias,
                       const float *uv, unsigned n_dirs, uint8_t *out_dev);
int ptmb_relight_rgb (ptmb_ctx_t *ctx, const uint8_t *r_dev, const uint8_t *g_dev, const uint8_t *b_dev,
                      size_t width, size_t height, const float *scale, const int32_t *bias,
                      const float *uv, unsigned n_dirs, uint8_t *out_dev);

/* PTM_FORMAT_LUM relighting (rti-builder/ptmlib.c:601-609): out triplets are (CLIP(POLY), Cb, Cr) -- a YCbCr
 * raster -- with (Cr, Cb) read from crcb_dev as 2-byte records at the pixel's index, which is how the
 * reference's loop addresses block 1.  Always the exact arithmetic. */
int ptmb_relight_lum (ptmb_ctx_t *ctx, const uint8_t *coef_dev, const uint8_t *crcb_dev, size_t width,
                      size_t height, const float *scale, const int32_t *bias, const float *uv, unsigned n_dirs,
                      uint8_t *out_dev);

/* Surface normals, ptm_normal (rti-builder/ptmlib.c:808-814, [Malzbender2001] eq. 16-17) for every pixel:
 * normals_dev float [pixels][3] = (nu, nv, nw), the reference's expressions including the sign it has inside
 * the square root.  From a float coefficient plane [pixels][6], or from a quantised block unscaled with the
 * header's scale / bias (rti-builder/ptmlib.c:54). */
int ptmb_normal_map (ptmb_ctx_t *ctx, const float *coeffs_dev, size_t pixels, float *normals_dev);
int ptmb_normal_map_blocks (ptmb_ctx_t *ctx, const uint8_t *coef_dev, const float *scale, const int32_t *bias,
                            size_t pixels, float *normals_dev);

/* The reference encoder's disabled LRGB variant (rti-builder/ptm-encoder.c:356-400): luminance L = R + G + B of
 * an RGB uint8 stack fitted like ptm_fit_poly_uint (coefficients NOT normalised), colour block
 * (JSAMPLE) (256 * chroma_c / L_0) with L_0 the first light's sum, chroma_c = truncated mean over the lights
 * (what the reference computes, PTMB_CHROMA_MEAN) or the median its FIXME asks for (:383, PTMB_CHROMA_MEDIAN:
 * (lower + upper middle) / 2).  L_0 == 0 (a division by zero in the reference) stores 0. */
#define PTMB_CHROMA_MEAN   0
#define PTMB_CHROMA_MEDIAN 1
int ptmb_fit_lsum (ptmb_ctx_t *ctx, const uint8_t *stack_dev, size_t image_stride, size_t pixels, int chroma,
                   float *coeffs_dev, uint8_t *colour_dev, int32_t *keys_dev);

/* Self test of K3's arithmetic shortcuts: compares the 3-instruction x / 255.0f and the CLIP
 * replacement with the IEEE division / the reference's CLIP (rti-builder/ptmlib.h:156-158) for
 * the float bit patterns [first_bits, first_bits + count); *mismatches must come back 0. */
int ptmb_selftest_div255 (ptmb_ctx_t *ctx, uint64_t first_bits, uint64_t count, uint64_t *mismatches);

/* Synthetic stack for pixels [pixel0, pixel0 + pixels) of every light, generated on the
 * device (integer only; the host statement is oracle/synth.c).  lights_q14: HOST array of
 * n_lights (x, y, z) in Q14. */
int ptmb_synth (ptmb_ctx_t *ctx, uint32_t seed, int mode, int sample_type, uint32_t pixel0, size_t pixels,
                unsigned n_lights, const int32_t *lights_q14, void *stack_dev, size_t image_stride);

/* Whole encode with HOST buffers -- the call the ptmlib.h drop-in and the new encoder main
 * make (it stands in for rti-builder/ptm-encoder.c:270-353,406 between "stack decoded" and
 * "blocks ready").  The stack is streamed to the GPU in row slabs (two staging buffers, copy
 * stream overlapped with K1; give it pinned memory for full PCIe rate), K1 runs per slab, then
 * K2, and the blocks come back.  format_planes = 1 (LRGB: blocks[0] coefficients, blocks[1]
 * RGB) or 3 (RGB formats: three coefficient blocks).  coeffs_host may be NULL.
 *
 * The two-call form exposes the one cross-GPU step: _fit leaves the slab's extrema keys on
 * the device (copied to keys_dev when non-NULL) so a multi-GPU caller can MAX-all-reduce
 * them before _finish (keys_dev NULL = the context's own keys).  _finish is synchronous.
 * Scratch is kept in the context between calls; _release frees it. */
/* Optional, LRGB only: where the RGB block of the NEXT ptmb_encode_host_fit goes on the HOST.  The block is
 * final after K1, so each slab's part is copied back on its own stream while later slabs are still coming
 * up (PCIe carries both directions at once); ptmb_encode_host_finish then skips blocks_host[1] if it is the
 * same pointer.  NULL (the default after every finish) = everything is copied in finish. */
int ptmb_encode_host_early_rgb (ptmb_ctx_t *ctx, uint8_t *rgb_host);
int ptmb_encode_host_fit (ptmb_ctx_t *ctx, const void *stack_host, int sample_type, size_t width, size_t height,
                          unsigned n_lights, const float *M, int format_planes, int32_t *keys_dev);
int ptmb_encode_host_finish (ptmb_ctx_t *ctx, const int32_t *keys_dev, uint8_t *const *blocks_host,
                             float *coeffs_host, ptmb_scale_bias_t *sb_host);
int ptmb_encode_host (ptmb_ctx_t *ctx, const void *stack_host, int sample_type, size_t width, size_t height,
                      unsigned n_lights, const float *M, int format_planes, uint8_t *const *blocks_host,
                      float *coeffs_host, ptmb_scale_bias_t *sb_host);
int ptmb_encode_host_release (ptmb_ctx_t *ctx);

/* Cross-GPU merge of the extrema keys over NVLink peer memory, one process per GPU -- the one
 * exchange step of the path, 48 bytes, latency bound.  Each rank creates a mailbox
 * and exports a 64-byte IPC handle; the hosts swap handles out of band (any transport), every
 * rank connects, and ptmb_xchg_allreduce_max then replaces keys_dev by the element-wise maximum
 * over all ranks with two one-block kernels around plain peer stores: stream ordered, no host
 * synchronisation, no collective library.  Integer max on order-preserving keys is exact, so the
 * result is bit-identical on every rank.  An int32 MAX all-reduce of the same 12 values through
 * NCCL / MPI is an equivalent alternative (rti_b200/encoder.py falls back to it). */
#define PTMB_XCHG_HANDLE_BYTES 64
typedef struct ptmb_xchg ptmb_xchg_t;
int ptmb_xchg_create (ptmb_ctx_t *ctx, int rank, int world, ptmb_xchg_t **xchg, void *handle_out);
int ptmb_xchg_connect (ptmb_xchg_t *xchg, const void *handles /* world x 64 bytes, rank order */);
int ptmb_xchg_allreduce_max (ptmb_ctx_t *ctx, ptmb_xchg_t *xchg, int32_t *keys_dev);
int ptmb_xchg_status (ptmb_ctx_t *ctx, ptmb_xchg_t *xchg, int *status);
void ptmb_xchg_destroy (ptmb_xchg_t *xchg);

/* Whole device-resident LRGB encode of one row slab in TWO launches (what the benchmark times):
 * K1's last block hands the merged keys straight to the mailbox of every rank (its own when
 * xchg is NULL) and K2 picks them up in its prologue -- no keys-init, no exchange kernels, no
 * host synchronisation.  Falls back to the separate-kernel sequence for layouts the streaming
 * kernel does not take.  Call sequence replaced: rti-builder/ptm-encoder.c:313-353,406. */
int ptmb_encode_lrgb_dev (ptmb_ctx_t *ctx, ptmb_xchg_t *xchg, const void *stack_dev, int sample_type,
                          size_t image_stride, size_t pixels, float *coeffs_dev, uint8_t *rgb_dev,
                          uint8_t *coef_bytes_dev, ptmb_scale_bias_t *sb_dev);

/* ---------------------------------------------------------------------------------------------
 * One process, several GPUs.  The reference spreads ONE encode over the host's cores with an OpenMP
 * loop over image rows (rti-builder/ptmlib.c:704,739,774,829,869); this is the same split over GPUs:
 * GPU i of n owns the contiguous row slab [floor(i*H/n), floor((i+1)*H/n)), fits it, the 12 extrema
 * keys are merged through in-process peer mailboxes over NVLink (cudaDeviceEnablePeerAccess -- no IPC
 * handles, no collective library), and every GPU quantises its slab.  Results are byte-identical to
 * the single-GPU call (integer max of order-preserving keys is exact).
 *
 *   ptmb_multi_create          devices = NULL / n_dev <= 0: all visible GPUs.  One context, one stream
 *                              and one issuing host thread per GPU.
 *   ptmb_multi_ctx             the context of GPU i (memory helpers, ptmb_synth, ptmb_stream, ...)
 *   ptmb_multi_slab            the rows GPU i owns for an image of `height` rows
 *   ptmb_multi_encode_host     ptmb_encode_host across the GPUs: ONE host stack / ONE set of host blocks
 *                              in the caller's layout; each GPU streams its rows (synchronous).  The rows are
 *                              split in proportion to the host bandwidth each GPU gets when all copy at once
 *                              (measured once, by the first call with a stack of 256 MB or more, and corrected
 *                              from the copy times of later encodes; equal slabs with PTM_MULTI_BALANCE=0) --
 *                              the result does not depend on the split
 *   ptmb_multi_host_slab       the rows GPU i takes in ptmb_multi_encode_host for an image of `height` rows
 *   ptmb_multi_set_host_weights  explicit shares (n_dev positive numbers; NULL = equal)
 *   ptmb_multi_calibrate_h2d   measure the shares now: every GPU copies pinned host memory for `seconds`
 *                              (<= 0: 12 ms), all at once; gbs_out (may be NULL) receives GB/s per GPU
 *   ptmb_multi_encode_lrgb_dev ptmb_encode_lrgb_dev on every GPU's resident slab (arrays of n_dev
 *                              entries); asynchronous, ptmb_multi_synchronize waits
 * This is what ptm_encode_stack / the ptm-encoder command line use when PTM_B200_DEVICES=0,1,... is set. */
typedef struct ptmb_multi ptmb_multi_t;
int ptmb_multi_create (const int *devices, int n_dev, ptmb_multi_t **multi);
void ptmb_multi_destroy (ptmb_multi_t *multi);
int ptmb_multi_device_count (const ptmb_multi_t *multi);
ptmb_ctx_t *ptmb_multi_ctx (ptmb_multi_t *multi, int i);
int ptmb_multi_slab (const ptmb_multi_t *multi, size_t height, int i, size_t *row0, size_t *row1);
int ptmb_multi_host_slab (const ptmb_multi_t *multi, size_t height, int i, size_t *row0, size_t *row1);
int ptmb_multi_set_host_weights (ptmb_multi_t *multi, const double *weights);
int ptmb_multi_calibrate_h2d (ptmb_multi_t *multi, double seconds, double *gbs_out);
int ptmb_multi_set_matrix (ptmb_multi_t *multi, const float *M, unsigned n_lights);
int ptmb_multi_set_fit_mode (ptmb_multi_t *multi, int flags);
int ptmb_multi_synchronize (ptmb_multi_t *multi);
int ptmb_multi_encode_host (ptmb_multi_t *multi, const void *stack_host, int sample_type, size_t width,
                            size_t height, unsigned n_lights, const float *M, int format_planes,
                            uint8_t *const *blocks_host, float *coeffs_host, ptmb_scale_bias_t *sb_host);
int ptmb_multi_encode_lrgb_dev (ptmb_multi_t *multi, const void *const *stack_dev, int sample_type,
                                const size_t *image_stride, const size_t *pixels, float *const *coeffs_dev,
                                uint8_t *const *rgb_dev, uint8_t *const *coef_bytes_dev,
                                ptmb_scale_bias_t *const *sb_dev);

/* The same for the RGB formats: coeffs_dev float [3][pixels][6] and coef_bytes_dev uint8 [3][pixels][6], planes
 * contiguous; the tensor-core fit kernel hands the keys over, ONE K2 launch quantises the three planes.  Falls back
 * to separate launches when the tensor-core engine does not take the layout.
 * Call sequence replaced: rti-builder/ptm-encoder.c:272-284,406. */
int ptmb_encode_rgb_dev (ptmb_ctx_t *ctx, ptmb_xchg_t *xchg, const void *stack_dev, int sample_type,
                         size_t image_stride, size_t pixels, float *coeffs_dev, uint8_t *coef_bytes_dev,
                         ptmb_scale_bias_t *sb_dev);

/* A promise by the caller about the stacks it passes to ptmb_encode_lrgb_dev: nothing queued earlier on the
 * context's stream writes them (they were complete before the PREVIOUS call on the stream was enqueued -- e.g. a
 * resident stack encoded repeatedly, or slabs uploaded on another stream and joined by an event earlier on).
 * K1 is launched programmatically dependent on the previous kernel; with the promise its copy engine warp
 * starts streaming the stack while that kernel still drains (a few microseconds per encode, which matters when
 * an 8-GPU slab takes 0.12 ms).  Default 0: K1 waits for the previous kernel before it reads anything. */
int ptmb_set_stack_stable (ptmb_ctx_t *ctx, int stable);

/* The float coefficients are an INTERMEDIATE of a whole encode (the reference's encoder allocates them, hands them
 * to ptm_scale_coefficients and frees them, rti-builder/ptm-encoder.c:313,406,409).  scratch != 0 declares the
 * coeffs_dev plane of ptmb_encode_lrgb_dev / ptmb_encode_rgb_dev scratch: its contents are undefined after the call.
 * K2 then quantises the end of the plane first -- the part K1 wrote last, still dirty in L2 -- and discards every L2
 * line it has read, so those lines are neither fetched from nor written back to HBM (24 MP: 80 MB less traffic and
 * 1 % per encode; a 3 MP slab, whose plane fits L2: 4 %).  The bytes, the colour block and scale / bias are the
 * same either way.  Default 0: the float plane is an output. */
int ptmb_set_coeffs_scratch (ptmb_ctx_t *ctx, int scratch);

/* Timing probes for benchmarks: the next `pairs` ptmb_encode_lrgb_dev calls bracket their fit
 * stage with CUDA events on the context's stream; _read synchronises and returns milliseconds. */
int ptmb_probe_arm (ptmb_ctx_t *ctx, unsigned pairs);
int ptmb_probe_read (ptmb_ctx_t *ctx, float *ms_out, unsigned pairs, unsigned *recorded);

/* JPEG in and out through nvJPEG (the codec the reference takes from libjpeg:
 * rti-builder/ptm-encoder.c:181-190,248-264, rti-builder/ptmlib.c:371-412,473-520,565-627).
 *   ptmb_jpeg_info    dimensions and component count (1 = grayscale, else 3) of a JPEG in memory.
 *   ptmb_jpeg_decode  -> interleaved [height][width][components].  want_rgb = 0 delivers YCbCr the
 *                     way libjpeg does with out_color_space = JCS_YCbCr (no colour conversion, chroma
 *                     upsampled with libjpeg's triangle filter); 1 delivers RGB.  `out` is host memory,
 *                     or device memory when out_is_device (e.g. one light's image inside the stack, so
 *                     that only the compressed bytes cross PCIe); flip stores the bottom row first like
 *                     rti-builder/ptm-encoder.c:252-256.
 *   ptmb_jpeg_encode  interleaved host image (1 = gray, 3 = RGB or YCbCr) -> malloc'd JPEG, baseline,
 *                     4:2:0 for colour, given quality.
 * The inverse DCT is nvJPEG's, so decoded samples may differ from libjpeg's by an LSB. */
int ptmb_jpeg_info (ptmb_ctx_t *ctx, const uint8_t *data, size_t len, unsigned *width, unsigned *height,
                    unsigned *components);
int ptmb_jpeg_decode (ptmb_ctx_t *ctx, const uint8_t *data, size_t len, int want_rgb, int flip, void *out,
                      int out_is_device);
int ptmb_jpeg_encode (ptmb_ctx_t *ctx, const uint8_t *pixels_host, unsigned width, unsigned height,
                      unsigned components, int input_is_ycbcr, int quality, uint8_t **jpeg_out, size_t *jpeg_len);

/* Host-pointer forms of the individual steps, for callers that keep the reference's own
 * call sequence through ptmlib.h (each streams its operands through device slabs):
 *   ptmb_fit_channel_host            ptm_fit_poly_jsample / _uint   rti-builder/ptmlib.c:692-760
 *   ptmb_chroma_avg_host             ptm_cbcr_avg                   rti-builder/ptmlib.c:762-802
 *   ptmb_lrgb_colour_normalise_host  the loop in the encoder main   rti-builder/ptm-encoder.c:327-353
 *   ptmb_scale_host                  ptm_scale_coefficients         rti-builder/ptmlib.c:816-881
 *   ptmb_relight_host                the loop of ptm_write_jpeg     rti-builder/ptmlib.c:587-621
 * samples_host points at the first sample of the wanted channel; sample (light l, pixel p) is
 * samples_host[(l * width * height + p) * pixel_stride], as in the reference. */
int ptmb_fit_channel_host (ptmb_ctx_t *ctx, const void *samples_host, int sample_type, size_t pixel_stride,
                           size_t width, size_t height, unsigned n_lights, const float *M, float *coeffs_host);
int ptmb_chroma_avg_host (ptmb_ctx_t *ctx, const uint8_t *stack_host, size_t pixels, unsigned n_lights,
                          uint8_t *avg_host);
int ptmb_lrgb_colour_normalise_host (ptmb_ctx_t *ctx, uint8_t *ycc_to_rgb_host, float *coeffs_host, size_t pixels);
int ptmb_scale_host (ptmb_ctx_t *ctx, const float *coeffs_host, size_t pixels, int planes,
                     uint8_t *const *blocks_host, ptmb_scale_bias_t *sb_host);
int ptmb_relight_host (ptmb_ctx_t *ctx, const uint8_t *const *blocks_host, int planes, size_t width, size_t height,
                       const float *scale, const int32_t *bias, const float *uv, unsigned n_dirs,
                       uint8_t *out_host);   /* planes: 1 = LRGB, 3 = RGB formats, 2 = LUM (blocks[1] = (cr, cb) records) */
/* ptm_normal (rti-builder/ptmlib.c:808-814) for every pixel: exactly one of coeffs_host (float [pixels][6])
 * and coef_bytes_host (uint8 [pixels][6] + scale / bias) is given; normals_host float [pixels][3]. */
int ptmb_normal_map_host (ptmb_ctx_t *ctx, const float *coeffs_host, const uint8_t *coef_bytes_host,
                          const float *scale, const int32_t *bias, size_t pixels, float *normals_host);

#ifdef __cplusplus
}
#endif

#endif
