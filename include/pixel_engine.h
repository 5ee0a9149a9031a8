/* pixel_engine.h -- C ABI of the B200 per-frame pixel engine (libpe_b200.so).
 *
 * Drop-in boundary for the LiVES hot path
 *     palette conversion -> resize / letterbox -> effect blend / composite -> gamma
 * Plain C: opaque handles, raw pointers, ints and doubles only.  Every entry point cites the
 * reference interface it replaces (file:line in the LiVES tree).  INTEGRATION.md shows the
 * reference-side bindings (the weed_layer_t drop-ins of include/pe_weed_layer.h = libpe_weed_layer.so, and the dlopen'd
 * effect plugin libpe_weed_plugin.so).
 *
 * Conventions (same as the reference):
 *   - "boolean" results are int PE_TRUE / PE_FALSE; a frame is MUTATED IN PLACE by the layer ops
 *     (palette, size, rowstrides, plane pointers, gamma / clamping fields are rewritten) and is
 *     left untouched when PE_FALSE is returned (colourspace.c:13906-13927).
 *   - widths are in PIXELS everywhere in this header (the weed shim converts macropixels).
 *   - rowstride rule: ALIGN_CEIL(width * bytes_per_pixel, 32); chroma planes of 4:2:0 / 4:2:2
 *     use rowstride[0] >> 1 (colourspace.c:11299-11357).
 *   - all work is enqueued on the engine's CUDA stream; device-frame calls return without
 *     synchronising, host-frame calls (pe_host_*) synchronise before returning.
 *   - the library FAILS LOUDLY (PE_ERR_CUDA / PE_FALSE + pe_last_error) when no CUDA device is
 *     usable; there is no CPU fallback.
 */
#ifndef PIXEL_ENGINE_H
#define PIXEL_ENGINE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PE_ABI_VERSION 2
#define PE_TRUE 1
#define PE_FALSE 0
#define PE_MAXPLANES 4 /* WEED_MAXPPLANES */

/* error codes of the int-returning calls (0 = ok) */
enum {
  PE_OK = 0,
  PE_ERR_CUDA = 1,      /* CUDA runtime / driver failure (see pe_last_error) */
  PE_ERR_ARG = 2,       /* NULL / out-of-range argument */
  PE_ERR_PALETTE = 3,   /* palette (pair) not handled by this build */
  PE_ERR_MEMORY = 4,    /* device allocation failed -> WEED_ERROR_MEMORY_ALLOCATION */
  PE_ERR_SIZE = 5       /* frame sizes do not match */
};

/* palettes, clamping, subspace, gamma: numeric values of libweed/weed-palettes.h:43-183 */
enum {
  PE_PALETTE_NONE = 0,
  PE_PALETTE_RGB24 = 1, PE_PALETTE_BGR24 = 2, PE_PALETTE_RGBA32 = 3, PE_PALETTE_BGRA32 = 4, PE_PALETTE_ARGB32 = 5,
  PE_PALETTE_YUV420P = 512, PE_PALETTE_YVU420P = 513, PE_PALETTE_YUV422P = 522, PE_PALETTE_YUV444P = 544,
  PE_PALETTE_YUVA4444P = 545, PE_PALETTE_UYVY = 564, PE_PALETTE_YUYV = 565, PE_PALETTE_YUV888 = 588,
  PE_PALETTE_YUVA8888 = 589,
  PE_PALETTE_YUV411 = 595 /* IYU1: {u2, y0, y1, v2, y2, y3} per 4 pixels; as a conversion SOURCE only (what the dv decoder delivers) */
};
enum { PE_YUV_CLAMPING_CLAMPED = 0, PE_YUV_CLAMPING_UNCLAMPED = 1 };
enum { PE_YUV_SAMPLING_DEFAULT = 0, PE_YUV_SAMPLING_JPEG = 0, PE_YUV_SAMPLING_MPEG = 1 };
enum { PE_YUV_SUBSPACE_YUV = 0, PE_YUV_SUBSPACE_YCBCR = 1, PE_YUV_SUBSPACE_BT709 = 2 };
enum { PE_GAMMA_UNKNOWN = 0, PE_GAMMA_LINEAR = -1, PE_GAMMA_SRGB = 1, PE_GAMMA_BT709 = 2,
       PE_GAMMA_MONITOR = 1024, PE_GAMMA_FILE = 1025, PE_GAMMA_VARIANT = 2048 }; /* colourspace.h:28-30 */
enum { PE_QUALITY_LOW = 1, PE_QUALITY_MED = 2, PE_QUALITY_HIGH = 3 };             /* preferences.h:100-104 */
enum { PE_INTERP_FAST = 0, PE_INTERP_NORMAL = 1, PE_INTERP_BEST = 2 };             /* LiVESInterpType */
enum { PE_DIRECTION_REVERSE = -1, PE_DIRECTION_FORWARD = 1 };                      /* defs.h:427-442 */
#define PE_LAYER_ALPHA_PREMULT 1                                                    /* colourspace.h:26 */

typedef struct pe_engine pe_engine_t; /* one per (process, GPU): stream, tables and LUT cache in HBM */
typedef struct pe_frame pe_frame_t;   /* a device-resident frame (the weed_layer_t of src/layers.c) */

/* the `prefs` fields that change pixel results */
typedef struct pe_config {
  int device;          /* CUDA ordinal */
  int pb_quality;      /* prefs->pb_quality; render/transcode force HIGH (colourspace.c:2103-2114) */
  double screen_gamma; /* prefs->screen_gamma (colourspace.c:677) */
  int apply_gamma;     /* prefs->apply_gamma (colourspace.c:12311) */
  int alpha_post;      /* prefs->alpha_post (colourspace.c:12290) */
  int ref_quirks;      /* 1: reproduce the deterministic slips of colourspace.c:3461,3544,3600 (parity default) */
  void *stream;        /* optional cudaStream_t to run on (e.g. the caller's); NULL -> engine-owned stream */
} pe_config_t;

/* frame descriptor: the leaves of a layer that the path reads (libweed/weed-effects.h:269-277,350-375) */
typedef struct pe_frame_desc {
  int palette;
  int width;  /* pixels */
  int height;
  int nplanes;
  int rowstrides[PE_MAXPLANES];
  void *planes[PE_MAXPLANES]; /* device pointers (pe_frame_t) or host pointers (pe_host_*) */
  int yuv_clamping, yuv_sampling, yuv_subspace;
  int gamma_type;
  int flags; /* PE_LAYER_ALPHA_PREMULT */
} pe_frame_desc_t;

/* ---- engine ------------------------------------------------------------------------------ */

void pe_config_default(pe_config_t *cfg);
/* init_colour_engine (colourspace.c:1973): builds every table on the host and uploads it */
int pe_engine_create(const pe_config_t *cfg, pe_engine_t **out);
void pe_engine_destroy(pe_engine_t *e);
int pe_engine_sync(pe_engine_t *e);
void *pe_engine_stream(pe_engine_t *e); /* cudaStream_t */
const char *pe_last_error(void);
/* kernels launched by this engine since creation (bench.py "gpu_launches") */
long pe_engine_launch_count(pe_engine_t *e);
/* CUDA-event timing on the engine stream (torch.cuda.Event only sees torch's stream) */
int pe_timer_start(pe_engine_t *e);
int pe_timer_stop_ms(pe_engine_t *e, float *ms); /* synchronises on the stop event */
int pe_sm_count(pe_engine_t *e);
/* The persistent kernels (one CTA per SM: the fused chain, the marching converters) size their grids by the SM count.  n > 0 caps that
 * at n SMs, 0 lifts the cap: a host that runs a collective (the operand broadcast of a multitrack crossfade) beside the
 * engine leaves the collective's CTAs somewhere to run, instead of having them time-sliced against 148 resident CTAs. */
int pe_engine_set_sm_limit(pe_engine_t *e, int n);
/* resize coefficients.  The reference delegates resizing to libswscale (src/colourspace.c:15059-15228; version unpinned,
 * configure.ac:562) with one flag per LiVESInterpType (:14991-14997).  recipe 1 (default) = libswscale's own coefficient recipes:
 * PE_INTERP_NORMAL SWS_BILINEAR, PE_INTERP_BEST SWS_LANCZOS when the frame grows / SWS_BICUBIC when it shrinks, PE_INTERP_FAST
 * SWS_FAST_BILINEAR (DESIGN.md section 5 states the measured distance to a real libswscale).  recipe 0 = the round-1 triangle
 * contract for every interpolation type.
 * pe_resize_filter_host returns the bank a recipe produces (host arithmetic, no GPU): kind 0 triangle, 1 bilinear, 2 bicubic,
 * 3 Lanczos, 4 / 5 fast bilinear vertical / horizontal; tap count, or -1 */
int pe_engine_set_resize_recipe(pe_engine_t *e, int recipe);
int pe_resize_filter_host(int recipe, int src_n, int dst_n, int shift_bits, int32_t *first, int16_t *coefs, int max_taps);
/* avg_chroma (colourspace.c:2079, the tables of init_average :190-217) in the closed form the chroma up-sampling kernels compute:
 * table[x][y] = clamp((((x + y) * A + B) * M) >> 32, lo, hi); out = {A, B, M, lo, hi}.  Returns 1 when the form equals the table the
 * library builds for this clamping, entry by entry (host arithmetic, no GPU). */
int pe_avg_closed_form(int clamped, uint32_t out[5]);
/* Pageable host planes (what lives_calloc_safety hands out) of 1 MB and more do not go to cudaMemcpyAsync as they are -- the driver
 * would stage them with one thread, ~10 GB/s -- but through page-locked ring buffers filled / drained by a small pool of copy threads
 * (PE_HOST_COPY_THREADS, default min(8, cores / 2); PE_HOST_NO_STAGING=1 switches it off).  Page-locked planes (pe_host_alloc,
 * pe_host_register) are never staged.  This is the pool's copy on its own: rows of row_bytes bytes at the given strides, `threads`
 * workers (host arithmetic, no GPU). */
int pe_host_parallel_copy2d(void *dst, size_t dst_stride, const void *src, size_t src_stride, size_t row_bytes, size_t rows, int threads);

/* one engine per process and GPU, shared by the weed_layer_t drop-ins (libpe_weed_layer.so) and the effect plugin
 * (libpe_weed_plugin.so): created on first use with the configuration given to pe_engine_shared_configure (default:
 * pe_config_default, device = $PE_DEVICE or 0); NULL + pe_last_error when no CUDA device is usable */
int pe_engine_shared_configure(const pe_config_t *cfg);
pe_engine_t *pe_engine_shared(void);
/* the host's prefs as they change at run time (prefs->pb_quality, screen_gamma, apply_gamma, alpha_post; src/preferences.h) */
int pe_engine_set_prefs(pe_engine_t *e, int pb_quality, double screen_gamma, int apply_gamma, int alpha_post);
int pe_engine_get_config(pe_engine_t *e, pe_config_t *out);

/* ---- frames (create_empty_pixel_data colourspace.c:11434, weed_layer_* src/layers.c) ------- */

/* plane geometry for a palette: fills nplanes / rowstrides / plane heights; returns total bytes when the
 * planes are laid out contiguously (what pe_frame_create allocates) */
size_t pe_frame_layout(int palette, int width, int height, int *nplanes, int rowstrides[PE_MAXPLANES],
                       int plane_heights[PE_MAXPLANES]);
int pe_frame_create(pe_engine_t *e, int palette, int width, int height, int yuv_clamping, int yuv_sampling,
                    int yuv_subspace, int gamma_type, int black_fill, pe_frame_t **out);
/* wrap caller-owned DEVICE memory (e.g. a torch tensor); never freed by the engine */
int pe_frame_wrap(pe_engine_t *e, const pe_frame_desc_t *desc, pe_frame_t **out);
void pe_frame_destroy(pe_frame_t *f);
int pe_frame_get_desc(const pe_frame_t *f, pe_frame_desc_t *out);
int pe_frame_set_gamma(pe_frame_t *f, int gamma_type);
int pe_frame_set_flags(pe_frame_t *f, int flags);
/* host <-> device (pinned host memory recommended); rowstrides may differ from the device ones */
int pe_frame_upload(pe_engine_t *e, pe_frame_t *f, const void *const host_planes[PE_MAXPLANES],
                    const int host_rowstrides[PE_MAXPLANES]);
int pe_frame_download(pe_engine_t *e, const pe_frame_t *f, void *const host_planes[PE_MAXPLANES],
                      const int host_rowstrides[PE_MAXPLANES]);
/* weed_layer_copy (src/layers.c:840) deep copy on the device */
int pe_frame_copy(pe_engine_t *e, const pe_frame_t *src, pe_frame_t **out);
/* pinned host buffers for callers without their own allocator */
void *pe_host_alloc(size_t bytes);
void pe_host_free(void *p);
/* page-lock caller-owned host memory in place (LiVES recycles its big pixel buffers, src/memory.c:37-47: register once) */
int pe_host_register(void *p, size_t bytes);
int pe_host_unregister(void *p);

/* ---- boundary B2: frame ops (same argument meaning as the reference functions) -------------- */

/* boolean convert_layer_palette_full(weed_layer_t *, int outpl, int oclamping, int osampling,
 *                                    int osubspace, int tgt_gamma)            colourspace.h:395, .c:12190 */
int pe_convert_layer_palette_full(pe_engine_t *e, pe_frame_t *layer, int outpl, int oclamping, int osampling,
                                  int osubspace, int tgt_gamma);
/* boolean convert_layer_palette(weed_layer_t *, int outpl, int op_clamping)   colourspace.h:393, .c:13931 */
int pe_convert_layer_palette(pe_engine_t *e, pe_frame_t *layer, int outpl, int op_clamping);
/* boolean resize_layer_full(weed_layer_t *, int width, int height, LiVESInterpType interp, int opal_hint,
 *                           int oclamp_hint, int osamp_hint, int osubs_hint, int tgt_gamma)
 *                                                                             colourspace.h:409, .c:14759 */
int pe_resize_layer_full(pe_engine_t *e, pe_frame_t *layer, int width, int height, int interp, int opal_hint,
                         int oclamp_hint, int osamp_hint, int osubs_hint, int tgt_gamma);
/* boolean resize_layer(weed_layer_t *, int w, int h, LiVESInterpType, int opal, int oclamp)
 *                                                                             colourspace.h:413, .c:15331 */
int pe_resize_layer(pe_engine_t *e, pe_frame_t *layer, int width, int height, int interp, int opal_hint,
                    int oclamp_hint);
/* batches of independent layers (what the render-to-disk loop, src/events.c:4239-4253, issues one by one): one lock, the
 * launches go out back to back.  Return the number of layers for which the per-layer call returned TRUE. */
int pe_resize_layer_batch(pe_engine_t *e, int n, pe_frame_t *const *layers, int width, int height, int interp, int opal_hint,
                          int oclamp_hint);
int pe_convert_layer_palette_batch(pe_engine_t *e, int n, pe_frame_t *const *layers, int outpl, int op_clamping);
/* boolean letterbox_layer(weed_layer_t *, int nwidth, int nheight, int width, int height, LiVESInterpType,
 *                         int tpal, int tclamp)                               colourspace.h:415, .c:15343 */
int pe_letterbox_layer(pe_engine_t *e, pe_frame_t *layer, int nwidth, int nheight, int width, int height,
                       int interp, int tpal, int tclamp);
/* boolean gamma_convert_layer(int gamma_type, weed_layer_t *)                 colourspace.h:389, .c:14146 */
int pe_gamma_convert_layer(pe_engine_t *e, int gamma_type, pe_frame_t *layer);
/* boolean gamma_convert_sub_layer(int gamma_type, double fileg, weed_layer_t *, int x, int y, int width,
 *                                 int height, boolean may_thread)             colourspace.h:391, .c:14069 */
int pe_gamma_convert_sub_layer(pe_engine_t *e, int gamma_type, double fileg, pe_frame_t *layer, int x, int y,
                               int width, int height, int may_thread);
/* void alpha_premult(weed_layer_t *, int direction)                           colourspace.h:387, .c:11968 */
void pe_alpha_premult(pe_engine_t *e, pe_frame_t *layer, int direction);
/* uint8_t *create_gamma_lut8(double fileg, int gamma_from, int gamma_to)      colourspace.c:655 (host copy) */
int pe_gamma_lut8(pe_engine_t *e, double fileg, int gamma_from, int gamma_to, uint8_t out[256]);

/* The reference's float ("experimental") YUV -> RGB arithmetic: float tables colourspace.c:101-172 (filled :1040-1104, BT.709 only),
 * clamp0255f :592, yuv2rgb_float :2367 (the reference compiles it but routes yuv2rgb to yuv2rgb_int, :2374-2375).  layer: YUV888 /
 * YUVA8888 with the BT.709 subspace -> outpl (an RGB palette), converted in place like convert_layer_palette.
 *   mode 0: yuv2rgb_float exactly as written (`int yy = RGB_Y[y]`, the 16.16 integer table, plus the float chroma tables);
 *   mode 1: the form of the commented-out variant at :2398-2400 (RGBf_Y[y] + Rf_Cr[v], ...).
 * sums_host (optional): width * height * 3 floats, the sums before clamp0255f -- equal to the compiled reference's bit for bit
 * (0 ULP; every addition is an IEEE single-precision add in the source's order).
 * pe_float_yuv_table: the float tables on the host (which 0 RGBf_Y 1 Rf_Cr 2 Gf_Cb 3 Gf_Cr 4 Bf_Cb). */
int pe_convert_yuv888_to_rgb_float(pe_engine_t *e, pe_frame_t *layer, int outpl, int mode, float *sums_host);
int pe_float_yuv_table(int clamping, int which, float out[256]);

/* ---- boundary B1 arithmetic: effect process functions on device frames ---------------------- */

/* simple_blend.c common_process :58.  type 0 "chroma blend", 1 "luma overlay", 2 "luma underlay",
 * 3 "negative luma overlay", 4 "averaged luma overlay" (its averaging branch is unreachable in the reference, :153-169: == type 1).
 * out may be in1 (in-place, effects-weed.c:2304-2314). */
int pe_fx_simple_blend(pe_engine_t *e, int type, const pe_frame_t *in1, const pe_frame_t *in2, pe_frame_t *out,
                       int blend_factor);
/* convert_layer_palette(clip, outpl) (colourspace.c:13521-13685) + the 'chroma blend' of simple_blend.c:58 with in1 = the
 * converted clip, in2 = operand, result in the clip: the multitrack crossfade (src/multitrack.h:84) as one kernel.
 * clip: YUV420P / YVU420P / YUV422P, outpl = operand palette = RGB24 or BGR24. */
int pe_fx_convert_crossfade(pe_engine_t *e, pe_frame_t *clip, const pe_frame_t *operand, int outpl, int op_clamping,
                            int blend_factor);
/* the same for the n clips of a multitrack stack that fade against ONE shared operand (BASELINE config 5: the operand every
 * rank receives by broadcast): same-shaped clips leave as one kernel launch per 32.  Returns the number of clips converted
 * (clips that fail are left untouched, as convert_layer_palette leaves its layer, colourspace.c:13906-13927). */
int pe_fx_convert_crossfade_batch(pe_engine_t *e, int n, pe_frame_t *const *clips, const pe_frame_t *operand, int outpl,
                                  int op_clamping, int blend_factor);
/* ... and with one operand per clip (operands[i] for clips[i]): successive frames of one clip against successive frames of the
 * other track, e.g. a group of operand frames received in one broadcast */
int pe_fx_convert_crossfade_batchv(pe_engine_t *e, int n, pe_frame_t *const *clips, const pe_frame_t *const *operands, int outpl,
                                   int op_clamping, int blend_factor);
/* multi_blends.c common_process :26.  type 0 multiply .. 6 burn; RGB24 / BGR24 only */
int pe_fx_multi_blend(pe_engine_t *e, int type, const pe_frame_t *in1, const pe_frame_t *in2, pe_frame_t *out,
                      int blend_factor);
/* slide_over.c sover_process :55 (the "slide over" transition): out = in1 ("upper") on one side of a dividing line, in2 ("lower") on
 * the other; the line moves with transval 0 .. 255.  direction = the plugin's "plugin_direction" 1 .. 4 as sover_init :38-52 derives
 * it from the radio parameters (1 / 2: the line runs along x, 3 / 4: along y; the caller resolves "random"); mvlower / mvupper = the
 * "Slide lower / upper clip" switches.  Every packed palette (ALL_PACKED_PALETTES_PLUS :158); not in-place.
 * pe_fx_slide_over_bound: the dividing line in rows / macropixels exactly as the reference's -ffast-math build computes it
 * (host arithmetic, no GPU needed). */
int pe_fx_slide_over(pe_engine_t *e, const pe_frame_t *in1, const pe_frame_t *in2, pe_frame_t *out, int transval, int direction,
                     int mvlower, int mvupper);
int pe_fx_slide_over_bound(int direction, int transval, int width, int height);
/* softlight.c softlight_process :62: an edge-magnitude "soft light" on the luma plane of a planar YUV frame (YUV444P, YUVA4444P,
 * YUV422P, YUV420P, YVU420P :169); chroma (and alpha) planes are copied.  in's yuv_clamping picks the luma range (:100-106). */
int pe_fx_softlight(pe_engine_t *e, const pe_frame_t *in, pe_frame_t *out);
/* layout_blends.c common_process :19, the "triple split" filter (RGB24 / BGR24): in1 inside a band, in2 outside, a border of
 * borderw in bordercol (RGB order) between them.  Parameters as the plugin's templates :137-144: start, sym ("make symmetrical"),
 * end, vert ("split horizontally"), borderw, bordercol.  out may be in1 (CAN_DO_INPLACE :135). */
int pe_fx_triple_split(pe_engine_t *e, const pe_frame_t *in1, const pe_frame_t *in2, pe_frame_t *out, double start, int sym, double end,
                       int vert, double borderw, const int bordercol[3]);
/* the column / row tests of layout_blends.c:92-99 as the reference evaluates them (bit 0 "outside", bit 1 "inside"; host arithmetic) */
void pe_fx_triple_split_classes(int width, int height, double start, int sym, double end, int vert, double borderw, uint8_t *colclass,
                                uint8_t *rowclass);
/* multi_transitions.c common_process :85.  type 0 "iris rectangle", 1 "iris circle", 2 "4 way split", 3 "dissolve"; amount = the
 * transition parameter 0 .. 1; every packed palette (ALL_PACKED_PALETTES :232).  out may be in1 except for type 2 (:268).  The float
 * geometry is evaluated in the form the plugins' -ffast-math build (lives-plugins/weed-plugins/Makefile.am:49) compiles to.
 * "dissolve" needs the per-instance mask dissolve_init :42 draws from the host's random seed (WEED_LEAF_RANDOM_SEED); type 4 ("rand
 * replace") is a whole-frame choice made by the plugin's host-side random stream: no pixel arithmetic, handled by the caller. */
typedef struct pe_dissolve_mask pe_dissolve_mask_t;
int pe_fx_dissolve_mask_create(pe_engine_t *e, int width, int height, int64_t random_seed, pe_dissolve_mask_t **out);
void pe_fx_dissolve_mask_destroy(pe_dissolve_mask_t *m);
int pe_fx_multi_transition(pe_engine_t *e, int type, const pe_frame_t *in1, const pe_frame_t *in2, pe_frame_t *out, double amount,
                           const pe_dissolve_mask_t *mask);
/* gdk/compositor.c compositor_process :127 at scale 1 / offset 0: out = bgcol, then paint_pixel(:120) of every
 * layer, last first (revz == WEED_FALSE, :189-197); alpha[i] is the scalar per-layer alpha */
int pe_fx_compositor(pe_engine_t *e, pe_frame_t *out, const pe_frame_t *const *layers, const double *alpha,
                     int nlayers, const int bgcol[3]);
/* compositor_process followed by gamma_convert_layer(gamma_to, out) (colourspace.c:14146; the APPLY_INST + gamma substeps of
 * src/nodemodel.c:1119-1333) with the LUT folded into the last paint: same bytes, one pass less over the frame */
int pe_fx_compositor_gamma(pe_engine_t *e, pe_frame_t *out, const pe_frame_t *const *layers, const double *alpha,
                           int nlayers, const int bgcol[3], int gamma_to);
/* the same for n independent output frames (render-to-disk, src/events.c:4239-4253): layers[i * nlayers + z] is layer z of frame
 * i, alpha[z] the per-layer alpha shared by all frames.  Integer paints (alpha = k / 256) of same-shaped frames leave as one
 * launch per 32 frames and paint pass.  Returns the number of frames composited. */
int pe_fx_compositor_gamma_batch(pe_engine_t *e, int n, pe_frame_t *const *outs, const pe_frame_t *const *layers,
                                 const double *alpha, int nlayers, const int bgcol[3], int gamma_to);
/* batches of independent frames (render-to-disk / multitrack): one launch, frames spread over the SMs */
int pe_fx_simple_blend_batch(pe_engine_t *e, int type, int n, const pe_frame_t *const *in1,
                             const pe_frame_t *const *in2, pe_frame_t *const *out, int blend_factor);

/* ---- fused chain ---------------------------------------------------------------------------- */

/* One kernel for: convert_layer_palette(fg -> RGBA32) ; letterbox_layer(fg, outer = bg size, inner = inner_w x
 * inner_h, bilinear) ; compositor over bg with scalar alpha (paint_pixel) ; gamma_convert_layer(out, gamma_to).
 * fg: YUV420P / YUV422P, bg and out: RGBA32 of the same size.  Bit-identical to calling the four ops above. */
int pe_fused_convert_letterbox_over_gamma(pe_engine_t *e, const pe_frame_t *fg, const pe_frame_t *bg,
                                          pe_frame_t *out, int inner_w, int inner_h, double alpha,
                                          int gamma_from, int gamma_to);
int pe_fused_convert_letterbox_over_gamma_batch(pe_engine_t *e, int n, const pe_frame_t *const *fg,
                                                const pe_frame_t *const *bg, pe_frame_t *const *out, int inner_w,
                                                int inner_h, double alpha, int gamma_from, int gamma_to);

/* ---- the node model's CONVERT step as one descriptor ----------------------------------------------------- */

/* What get_op_order (src/nodemodel.c:161, called :1981) decides per layer -- which of resize / palette conversion / gamma / letterbox
 * are needed and in which order (equal numbers = one operation does both) -- with their arguments: the binding hands the whole
 * chain over in one call instead of issuing res -> pconv -> lbox -> gamma substeps (:1065-1282). */
enum { PE_OP_RESIZE = 0, PE_OP_PCONV = 1, PE_OP_GAMMA = 2, PE_OP_LETTERBOX = 3, PE_N_OP_TYPES = 6 }; /* src/nodemodel.h:717-722 */
typedef struct pe_convert_plan {
  int op_order[PE_N_OP_TYPES]; /* get_op_order's output: 0 not needed, 1 .. the substep an op runs in */
  int width, height;           /* OP_RESIZE target = the inner rectangle when letterboxing (pixels) */
  int lb_width, lb_height;     /* OP_LETTERBOX outer size */
  int interp;                  /* LiVESInterpType */
  int out_palette, out_clamping, out_sampling, out_subspace; /* OP_PCONV */
  int out_gamma;               /* OP_GAMMA */
  int no_fuse;                 /* 1: never take the fused kernel (tests) */
} pe_convert_plan_t;
/* the substeps, one reference op at a time, in the plan's order; boolean */
int pe_run_convert_plan(pe_engine_t *e, pe_frame_t *layer, const pe_convert_plan_t *plan);
/* CONVERT + the compositor APPLY_INST over one background layer + gamma (:1119-1333): ONE fused launch when the plan is planar YUV ->
 * RGBA32 + letterbox (bilinear), the op-by-op sequence otherwise; same bytes.  pe_last_plan_path(): 1 fused, 0 op by op */
int pe_run_convert_plan_over(pe_engine_t *e, const pe_frame_t *fg, const pe_convert_plan_t *plan, const pe_frame_t *bg, pe_frame_t *out,
                             double alpha, int gamma_to);
int pe_last_plan_path(void);

/* ---- frame ingest / egress in device memory -------------------------------------------------------------- */

/* boolean get_frame(const lives_clip_data_t *, int64_t frame, int *rowstrides, int height, void **pixel_data)
 *                                                  src/plugins.h:442, lives-plugins/plugins/decoders/decplugin.h:280
 * with pixel_data in DEVICE memory: a device decoder (NVDEC / nvJPEG) writes the planes of frame `frame` on `cuda_stream` and returns
 * TRUE; nothing crosses PCIe.  rowstrides / height are the device frame's (rowstride rule of colourspace.c:11299-11357). */
typedef int (*pe_device_get_frame_f)(void *clip_data, int64_t frame, const int *rowstrides, int height, void *const *pixel_data_dev,
                                     void *cuda_stream);
typedef struct pe_clip_source {
  void *clip_data;                 /* lives_clip_data_t of the decoder */
  pe_device_get_frame_f get_frame;
  int palette, width, height;      /* cdata->current_palette, width, height (pixels) */
  int yuv_clamping, yuv_sampling, yuv_subspace, gamma_type;
} pe_clip_source_t;
/* pull_frame (src/frameloader.c:1494,1841) on the device: a new layer filled by the source */
int pe_ingest_frame(pe_engine_t *e, const pe_clip_source_t *src, int64_t frame, pe_frame_t **out);

/* the built-in source: a clip that already sits in HBM (decoded once, or written there by a device decoder) */
typedef struct pe_clip_cache pe_clip_cache_t;
int pe_clip_cache_create(pe_engine_t *e, int palette, int width, int height, int nframes, int yuv_clamping, int yuv_sampling,
                         int yuv_subspace, int gamma_type, pe_clip_cache_t **out);
void pe_clip_cache_destroy(pe_clip_cache_t *c);
int pe_clip_cache_load(pe_clip_cache_t *c, int64_t frame, const void *const host_planes[PE_MAXPLANES],
                       const int host_rowstrides[PE_MAXPLANES]);                 /* the one-time fill from host memory */
int pe_clip_cache_frame_desc(pe_clip_cache_t *c, int64_t frame, pe_frame_desc_t *out); /* device pointers of a cached frame */
int pe_clip_cache_source(pe_clip_cache_t *c, pe_clip_source_t *out);            /* get_frame = device-to-device plane copies */
int pe_clip_cache_borrow(pe_clip_cache_t *c, int64_t frame, pe_frame_t **out);  /* zero copy: the cached frame as a read-only layer */

/* the render-to-disk tail (src/events.c:4247-4263: convert to the clip's palette, layer_to_pixbuf): convert_layer_palette(layer,
 * out_palette) when needed, then ONLY that packed frame is copied to (pinned) host memory, on its own stream.  _begin returns at once;
 * slot 0 .. 3 names the copy for _wait; the layer stays alive and unwritten until then. */
int pe_render_out_begin(pe_engine_t *e, pe_frame_t *layer, int out_palette, void *host_dst, int host_rowstride, int slot);
int pe_render_out_wait(pe_engine_t *e, int slot);
int pe_render_out(pe_engine_t *e, pe_frame_t *layer, int out_palette, void *host_dst, int host_rowstride);

/* ---- the one exchange step of the path (BASELINE config 5) ---------------------------------------------------------
 * The shared transition operand of a multitrack crossfade (src/multitrack.h:84: every track's clip fades against the same frame)
 * lives on ONE rank; with one clip per GPU every rank needs it for every output frame.  pe_mc_publish is the owner's side as ONE
 * kernel: `bytes` bytes (a multiple of 16; 16-byte aligned addresses) are read from `src` in the owner's memory and stored through
 * `mc_dst`, an NVSwitch multicast address that maps a buffer on every GPU of the group (cuMulticast* / symmetric memory: the host's
 * plumbing, lives_b200/shard.py), so the operand leaves the owner's NVLink once whatever the number of receivers.  Asynchronous on
 * `cuda_stream` (NULL: the engine's stream); max_ctas <= 0: one CTA per SM (the kernel is small enough to share the SMs with the
 * conversion kernels).  Completion is ordered like any kernel's: an event on the stream, then the group's barrier. */
int pe_mc_publish(pe_engine_t *e, void *mc_dst, const void *src, size_t bytes, void *cuda_stream, int max_ctas);

/* ---- per-frame diagnostics (is_all_black_ish colourspace.c:2554, hash_cmp_layer :16044) ------ */

typedef struct pe_frame_stats {
  uint8_t min[4], max[4]; /* per byte position of a packed pixel (plane 0 for planar) */
  uint32_t hist[256];     /* histogram of the colour bytes (the alpha byte excluded) of plane 0 */
  uint64_t sum;           /* sum of all payload bytes of plane 0 */
  int all_black_ish;      /* is_all_black_ish(..., exact = FALSE): the reference's bit expression on bytes 0 .. 2 of every pixel
                             (colourspace.c:2583-2587); -1 when the palette is not packed RGB */
  int all_black;          /* is_all_black_ish(..., exact = TRUE): bytes 0 .. 2 of every pixel are 0 (:2589); -1 as above */
} pe_frame_stats_t;
int pe_frame_stats(pe_engine_t *e, const pe_frame_t *f, pe_frame_stats_t *out);
/* hash_cmp_layer (colourspace.c:16044-16075): minimd5 (src/maths.c:575, the reference's own MD5 variant) of the first nbytes bytes of
 * every row of plane 0 (nbytes <= 0: the reference's `width` bytes, width in macropixels) + their XOR (the "parity", :16068) */
int pe_frame_row_hashes(pe_engine_t *e, const pe_frame_t *f, int nbytes, uint64_t *hashes, uint64_t *parity);

/* ---- host-frame drop-ins: H2D -> device op -> D2H around the calls above --------------------- */

/* allocator the caller wants new pixel buffers to come from (LiVES: create_empty_pixel_data / bigblocks) */
typedef void *(*pe_host_alloc_f)(size_t bytes, void *user);
typedef void (*pe_host_free_f)(void *ptr, void *user);
typedef struct pe_host_allocator { pe_host_alloc_f alloc; pe_host_free_f free; void *user; } pe_host_allocator_t;

/* layer->planes are HOST pointers; on success they are replaced (old ones released with the allocator) when the
 * byte size changes, exactly where the reference swaps pixel_data (colourspace.c:13859-13900) */
int pe_host_convert_layer_palette_full(pe_engine_t *e, pe_frame_desc_t *layer, int outpl, int oclamping,
                                       int osampling, int osubspace, int tgt_gamma,
                                       const pe_host_allocator_t *alloc);
int pe_host_resize_layer(pe_engine_t *e, pe_frame_desc_t *layer, int width, int height, int interp, int opal_hint,
                         int oclamp_hint, const pe_host_allocator_t *alloc);
int pe_host_letterbox_layer(pe_engine_t *e, pe_frame_desc_t *layer, int nwidth, int nheight, int width, int height,
                            int interp, int tpal, int tclamp, const pe_host_allocator_t *alloc);
int pe_host_gamma_convert_layer(pe_engine_t *e, int gamma_type, pe_frame_desc_t *layer);
/* the weed process_func body: host channels in, host channel out (what libpe_weed_plugin.so calls) */
int pe_host_simple_blend(pe_engine_t *e, int type, const pe_frame_desc_t *in1, const pe_frame_desc_t *in2,
                         pe_frame_desc_t *out, int blend_factor);
int pe_host_multi_blend(pe_engine_t *e, int type, const pe_frame_desc_t *in1, const pe_frame_desc_t *in2,
                        pe_frame_desc_t *out, int blend_factor);
/* slide_over.c:55 on host frames (H2D of both clips, k_slide_over, D2H of the result) */
int pe_host_slide_over(pe_engine_t *e, const pe_frame_desc_t *in1, const pe_frame_desc_t *in2, pe_frame_desc_t *out, int transval,
                       int direction, int mvlower, int mvupper);
/* softlight.c:62 / layout_blends.c:19 / multi_transitions.c:85 on host channels */
int pe_host_softlight(pe_engine_t *e, const pe_frame_desc_t *in, pe_frame_desc_t *out);
int pe_host_triple_split(pe_engine_t *e, const pe_frame_desc_t *in1, const pe_frame_desc_t *in2, pe_frame_desc_t *out, double start, int sym,
                         double end, int vert, double borderw, const int bordercol[3]);
int pe_host_multi_transition(pe_engine_t *e, int type, const pe_frame_desc_t *in1, const pe_frame_desc_t *in2, pe_frame_desc_t *out,
                             double amount, const pe_dissolve_mask_t *mask);
/* gdk/compositor.c compositor_process :127 on host channels (layers[z] NULL = a channel the host disabled, :195-199) */
int pe_host_compositor(pe_engine_t *e, pe_frame_desc_t *out, const pe_frame_desc_t *const *layers, const double *alpha, int nlayers,
                       const int bgcol[3]);
int pe_host_fused_convert_letterbox_over_gamma(pe_engine_t *e, const pe_frame_desc_t *fg, const pe_frame_desc_t *bg,
                                               pe_frame_desc_t *out, int inner_w, int inner_h, double alpha,
                                               int gamma_from, int gamma_to);

/* a batch of independent host frames (render-to-disk) through the fused chain: uploads, kernels and downloads of
 * consecutive frames overlap on three streams; host buffers should be pinned (pe_host_alloc) */
int pe_host_fused_convert_letterbox_over_gamma_batch(pe_engine_t *e, int n, const pe_frame_desc_t *const *fg,
                                                     const pe_frame_desc_t *const *bg, pe_frame_desc_t *const *out,
                                                     int inner_w, int inner_h, double alpha, int gamma_from, int gamma_to);

#ifdef __cplusplus
}
#endif
#endif /* PIXEL_ENGINE_H */
