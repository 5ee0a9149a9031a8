/* pe_weed_plugin.c -- libpe_weed_plugin.so: a libweed effect plugin (boundary B1) whose process functions run on the B200.
 *
 * Loaded exactly like the reference's plugins: dlopen + dlsym("weed_setup") + setup(weed_bootstrap)
 * (src/effects-weed.c:4468-4568).  It registers, under the SAME filter names, the filters of
 *     lives-plugins/weed-plugins/simple_blend.c   "chroma blend", "luma overlay", "luma underlay", "negative luma overlay", "averaged luma overlay"
 *     lives-plugins/weed-plugins/multi_blends.c   "blend_multiply" ... "blend_burn"
 *     lives-plugins/weed-plugins/slide_over.c     "slide over"
 *     lives-plugins/weed-plugins/gdk/compositor.c "compositor" (layers at scale 1 / offset 0: BASELINE config 3 through weed_apply_instance)
 *     lives-plugins/weed-plugins/softlight.c      "softlight"
 *     lives-plugins/weed-plugins/layout_blends.c  "triple split"
 *     lives-plugins/weed-plugins/multi_transitions.c "iris rectangle", "iris circle", "4 way split", "dissolve", "rand replace"
 * with the same channel / parameter templates, so weed_apply_instance() (src/effects-weed.c:1850) drives it unchanged.
 * Differences from the originals, on purpose:
 *   - WEED_FILTER_HINT_MAY_THREAD is NOT set: the host must call process_func once per frame, not once per row band
 *     (the CUDA grid is the row-band fan-out);  STATEFUL is not needed either (no per-instance blend table);
 *   - every pixel is computed by libpe_b200.so (pe_host_simple_blend / pe_host_multi_blend: H2D, kernel, D2H); when no
 *     CUDA device is usable process_func returns WEED_ERROR_PLUGIN_INVALID -- there is no CPU fallback in here.
 * The host owns all pixel buffers; this file only reads the leaves the reference plugins read
 * (libweed/weed-plugin-utils.c:140-157).
 */
#include <stdio.h>
#include <string.h>

#include "pe_weed_abi.h"
#include "pixel_engine.h"

static pe_weed_leaf_get_f w_leaf_get;
static pe_weed_leaf_set_f w_leaf_set;
static pe_weed_plant_new_f w_plant_new;
static pe_weed_leaf_num_elements_f w_num_elements;
static pe_weed_malloc_f w_malloc;
static pe_weed_free_f w_free;

/* the process-wide engine of libpe_b200.so, shared with the weed_layer_t drop-ins (libpe_weed_layer.so): the host's prefs reach it
 * through pe_engine_set_prefs(pe_engine_shared(), prefs->pb_quality, ...), the GPU is picked by pe_engine_shared_configure / $PE_DEVICE */
static pe_engine_t *engine(void) {
  pe_engine_t *e = pe_engine_shared();
  if (!e) fprintf(stderr, "pe_weed_plugin: %s\n", pe_last_error());
  return e;
}

/* ---- leaf helpers ---------------------------------------------------------------------------------------------- */

static int get_int(pe_weed_plant_t *p, const char *key) { int32_t v = 0; w_leaf_get(p, key, 0, &v); return v; }
static void *get_ptr(pe_weed_plant_t *p, const char *key) { void *v = NULL; w_leaf_get(p, key, 0, &v); return v; }
static pe_weed_plant_t *get_plant(pe_weed_plant_t *p, const char *key, int idx) {
  pe_weed_plant_t *v = NULL;
  w_leaf_get(p, key, (pe_weed_size_t)idx, &v);
  return v;
}
static void set_int(pe_weed_plant_t *p, const char *key, int32_t v) { w_leaf_set(p, key, PE_WEED_SEED_INT, 1, &v); }
static void set_str(pe_weed_plant_t *p, const char *key, const char *s) { w_leaf_set(p, key, PE_WEED_SEED_STRING, 1, &s); }

/* the frame a channel plant describes (width leaf is in macropixels == pixels for the RGB palettes) */
static void channel_desc(pe_weed_plant_t *ch, pe_frame_desc_t *d) {
  memset(d, 0, sizeof(*d));
  d->palette = get_int(ch, PE_LEAF_CURRENT_PALETTE);
  d->width = get_int(ch, PE_LEAF_WIDTH);
  if (d->palette == PE_PALETTE_UYVY || d->palette == PE_PALETTE_YUYV) d->width *= 2; /* weed counts macropixels, pixel_engine.h pixels */
  d->height = get_int(ch, PE_LEAF_HEIGHT);
  d->nplanes = 1;
  d->rowstrides[0] = get_int(ch, PE_LEAF_ROWSTRIDES);
  d->planes[0] = get_ptr(ch, PE_LEAF_PIXEL_DATA);
}

/* ---- process functions --------------------------------------------------------------------------------------------- */

static pe_weed_error_t run_blend(int family, int type, pe_weed_plant_t *inst) {
  pe_frame_desc_t in1, in2, out;
  pe_engine_t *e = engine();
  int bf, rc;
  if (!e) return PE_WEED_ERROR_PLUGIN_INVALID;
  channel_desc(get_plant(inst, PE_LEAF_IN_CHANNELS, 0), &in1);
  channel_desc(get_plant(inst, PE_LEAF_IN_CHANNELS, 1), &in2);
  channel_desc(get_plant(inst, PE_LEAF_OUT_CHANNELS, 0), &out);
  bf = get_int(get_plant(inst, PE_LEAF_IN_PARAMETERS, 0), PE_LEAF_VALUE);
  rc = family == 0 ? pe_host_simple_blend(e, type, &in1, &in2, &out, bf) : pe_host_multi_blend(e, type, &in1, &in2, &out, bf);
  if (rc == PE_OK) return PE_WEED_SUCCESS;
  fprintf(stderr, "pe_weed_plugin: %s\n", pe_last_error());
  return rc == PE_ERR_MEMORY ? PE_WEED_ERROR_MEMORY_ALLOCATION : PE_WEED_ERROR_PLUGIN_INVALID;
}

#define PE_PROCESS(name, family, type) \
  static pe_weed_error_t name(pe_weed_plant_t *inst, pe_weed_timecode_t tc) { (void)tc; return run_blend(family, type, inst); }
PE_PROCESS(chroma_process, 0, 0)
PE_PROCESS(lumo_process, 0, 1)
PE_PROCESS(lumu_process, 0, 2)
PE_PROCESS(nlumo_process, 0, 3)
PE_PROCESS(avlumo_process, 0, 4)
PE_PROCESS(mpy_process, 1, 0)
PE_PROCESS(screen_process, 1, 1)
PE_PROCESS(darken_process, 1, 2)
PE_PROCESS(lighten_process, 1, 3)
PE_PROCESS(overlay_process, 1, 4)
PE_PROCESS(dodge_process, 1, 5)
PE_PROCESS(burn_process, 1, 6)

static pe_weed_error_t common_init(pe_weed_plant_t *inst) {
  (void)inst;
  return engine() ? PE_WEED_SUCCESS : PE_WEED_ERROR_PLUGIN_INVALID;
}

/* slide_over.c sover_init :38-52: the radio parameters pick "plugin_direction" (0 = random, resolved at the first process call) */
static pe_weed_error_t sover_init(pe_weed_plant_t *inst) {
  int dirpref = 4, k;
  if (!engine()) return PE_WEED_ERROR_PLUGIN_INVALID;
  for (k = 1; k <= 4; k++)
    if (get_int(get_plant(inst, PE_LEAF_IN_PARAMETERS, k), PE_LEAF_VALUE) == 1) { dirpref = k - 1; break; }
  set_int(inst, "plugin_direction", dirpref);
  return PE_WEED_SUCCESS;
}

/* slide_over.c sover_process :55-145 */
static pe_weed_error_t sover_process(pe_weed_plant_t *inst, pe_weed_timecode_t tc) {
  static unsigned int seed = 0x2545F491u;
  pe_frame_desc_t in1, in2, out;
  pe_engine_t *e = engine();
  int transval, dirn, mvlower, mvupper, rc;
  if (!e) return PE_WEED_ERROR_PLUGIN_INVALID;
  channel_desc(get_plant(inst, PE_LEAF_IN_CHANNELS, 0), &in1);
  channel_desc(get_plant(inst, PE_LEAF_IN_CHANNELS, 1), &in2);
  channel_desc(get_plant(inst, PE_LEAF_OUT_CHANNELS, 0), &out);
  transval = get_int(get_plant(inst, PE_LEAF_IN_PARAMETERS, 0), PE_LEAF_VALUE);
  mvlower = get_int(get_plant(inst, PE_LEAF_IN_PARAMETERS, 6), PE_LEAF_VALUE);
  mvupper = get_int(get_plant(inst, PE_LEAF_IN_PARAMETERS, 7), PE_LEAF_VALUE);
  dirn = get_int(inst, "plugin_direction");
  if (dirn == 0) { /* random, kept for the life of the instance (:79-82) */
    seed = seed * 1664525u + 1013904223u + (unsigned int)tc;
    dirn = (int)((seed >> 24) & 3u) + 1;
    set_int(inst, "plugin_direction", dirn);
  }
  rc = pe_host_slide_over(e, &in1, &in2, &out, transval, dirn, mvlower, mvupper);
  if (rc == PE_OK) return PE_WEED_SUCCESS;
  fprintf(stderr, "pe_weed_plugin: %s\n", pe_last_error());
  return rc == PE_ERR_MEMORY ? PE_WEED_ERROR_MEMORY_ALLOCATION : PE_WEED_ERROR_PLUGIN_INVALID;
}

/* gdk/compositor.c compositor_process :127-297: N repeating in channels painted onto the background colour, last first (or first
 * to last with "revz").  Offsets / scales other than 0 / 1 need gdk_pixbuf_scale_simple (absent from this image, nothing to pin
 * against): such an instance fails loudly.  A disabled channel (WEED_LEAF_DISABLED, :195) is skipped. */
#define PE_MAX_COMP_LAYERS 64
static double get_dbl(pe_weed_plant_t *p, const char *key, int idx, double dflt) {
  double v = dflt;
  if ((int)w_num_elements(p, key) > idx) w_leaf_get(p, key, (pe_weed_size_t)idx, &v);
  return v;
}
static pe_weed_error_t compositor_process(pe_weed_plant_t *inst, pe_weed_timecode_t tc) {
  pe_frame_desc_t out, in[PE_MAX_COMP_LAYERS];
  const pe_frame_desc_t *layers[PE_MAX_COMP_LAYERS];
  double alpha[PE_MAX_COMP_LAYERS];
  pe_weed_plant_t *par[7], *ch;
  pe_engine_t *e = engine();
  int n, z, k, bgcol[3] = {0, 0, 0}, revz = 0, rc;
  (void)tc;
  if (!e) return PE_WEED_ERROR_PLUGIN_INVALID;
  n = (int)w_num_elements(inst, PE_LEAF_IN_CHANNELS);
  if (n > PE_MAX_COMP_LAYERS) n = PE_MAX_COMP_LAYERS;
  for (k = 0; k < 7; k++) par[k] = get_plant(inst, PE_LEAF_IN_PARAMETERS, k);
  channel_desc(get_plant(inst, PE_LEAF_OUT_CHANNELS, 0), &out);
  for (k = 0; k < 3; k++) { int32_t c = 0; if ((int)w_num_elements(par[5], PE_LEAF_VALUE) > k) w_leaf_get(par[5], PE_LEAF_VALUE, (pe_weed_size_t)k, &c); bgcol[k] = c; }
  { int32_t b = 0; w_leaf_get(par[6], PE_LEAF_VALUE, 0, &b); revz = b; }
  for (z = 0; z < n; z++) {
    /* pe_fx_compositor paints the LAST layer first (revz == WEED_FALSE, :189-192): with revz the order of the list is reversed */
    const int src = revz ? n - 1 - z : z;
    int32_t disabled = 0;
    ch = get_plant(inst, PE_LEAF_IN_CHANNELS, src);
    layers[z] = NULL;
    alpha[z] = get_dbl(par[4], PE_LEAF_VALUE, src, 1.);
    if (!ch) continue;
    if (w_num_elements(ch, "disabled") > 0) w_leaf_get(ch, "disabled", 0, &disabled);
    if (disabled) continue;
    channel_desc(ch, &in[z]);
    if (!in[z].planes[0]) continue;
    if (get_dbl(par[0], PE_LEAF_VALUE, src, 0.) != 0. || get_dbl(par[1], PE_LEAF_VALUE, src, 0.) != 0. ||
        get_dbl(par[2], PE_LEAF_VALUE, src, 1.) != 1. || get_dbl(par[3], PE_LEAF_VALUE, src, 1.) != 1.) {
      fprintf(stderr, "pe_weed_plugin: compositor layer %d: offsets / scales other than 0 / 1 are not handled by this build\n", src);
      return PE_WEED_ERROR_FILTER_INVALID;
    }
    layers[z] = &in[z];
  }
  rc = pe_host_compositor(e, &out, layers, alpha, n, bgcol);
  if (rc == PE_OK) return PE_WEED_SUCCESS;
  fprintf(stderr, "pe_weed_plugin: %s\n", pe_last_error());
  return rc == PE_ERR_MEMORY ? PE_WEED_ERROR_MEMORY_ALLOCATION : PE_WEED_ERROR_PLUGIN_INVALID;
}

/* a planar channel: pixel_data / rowstrides are arrays (softlight.c:66-73) */
static void channel_desc_planar(pe_weed_plant_t *ch, pe_frame_desc_t *d) {
  int np, k;
  channel_desc(ch, d);
  np = (int)w_num_elements(ch, PE_LEAF_PIXEL_DATA);
  if (np > PE_MAXPLANES) np = PE_MAXPLANES;
  d->nplanes = np;
  for (k = 0; k < np; k++) {
    int32_t rs = 0;
    void *pp = NULL;
    w_leaf_get(ch, PE_LEAF_PIXEL_DATA, (pe_weed_size_t)k, &pp);
    w_leaf_get(ch, PE_LEAF_ROWSTRIDES, (pe_weed_size_t)k, &rs);
    d->planes[k] = pp;
    d->rowstrides[k] = rs;
  }
  d->yuv_clamping = PE_YUV_CLAMPING_CLAMPED;
  if (w_num_elements(ch, "YUV_clamping") > 0) d->yuv_clamping = get_int(ch, "YUV_clamping");
}

static pe_weed_error_t rc_to_weed(int rc) {
  if (rc == PE_OK) return PE_WEED_SUCCESS;
  fprintf(stderr, "pe_weed_plugin: %s\n", pe_last_error());
  return rc == PE_ERR_MEMORY ? PE_WEED_ERROR_MEMORY_ALLOCATION : PE_WEED_ERROR_PLUGIN_INVALID;
}

/* softlight.c softlight_process :62 */
static pe_weed_error_t softlight_process(pe_weed_plant_t *inst, pe_weed_timecode_t tc) {
  pe_frame_desc_t in, out;
  pe_engine_t *e = engine();
  (void)tc;
  if (!e) return PE_WEED_ERROR_PLUGIN_INVALID;
  channel_desc_planar(get_plant(inst, PE_LEAF_IN_CHANNELS, 0), &in);
  channel_desc_planar(get_plant(inst, PE_LEAF_OUT_CHANNELS, 0), &out);
  return rc_to_weed(pe_host_softlight(e, &in, &out));
}

/* layout_blends.c tsplit_process :122 */
static pe_weed_error_t tsplit_process(pe_weed_plant_t *inst, pe_weed_timecode_t tc) {
  pe_frame_desc_t in1, in2, out;
  pe_engine_t *e = engine();
  pe_weed_plant_t *par[7];
  double xstart = 0., xend = 0., bw = 0.;
  int32_t sym = 0, vert = 0, bc[3] = {0, 0, 0};
  int k, col[3];
  (void)tc;
  if (!e) return PE_WEED_ERROR_PLUGIN_INVALID;
  channel_desc(get_plant(inst, PE_LEAF_IN_CHANNELS, 0), &in1);
  channel_desc(get_plant(inst, PE_LEAF_IN_CHANNELS, 1), &in2);
  channel_desc(get_plant(inst, PE_LEAF_OUT_CHANNELS, 0), &out);
  for (k = 0; k < 7; k++) par[k] = get_plant(inst, PE_LEAF_IN_PARAMETERS, k);
  w_leaf_get(par[0], PE_LEAF_VALUE, 0, &xstart);
  w_leaf_get(par[1], PE_LEAF_VALUE, 0, &sym);
  w_leaf_get(par[3], PE_LEAF_VALUE, 0, &xend);
  w_leaf_get(par[4], PE_LEAF_VALUE, 0, &vert);
  w_leaf_get(par[5], PE_LEAF_VALUE, 0, &bw);
  for (k = 0; k < 3 && k < (int)w_num_elements(par[6], PE_LEAF_VALUE); k++) w_leaf_get(par[6], PE_LEAF_VALUE, (pe_weed_size_t)k, &bc[k]);
  for (k = 0; k < 3; k++) col[k] = bc[k];
  return rc_to_weed(pe_host_triple_split(e, &in1, &in2, &out, xstart, sym, xend, vert, bw, col));
}

/* multi_transitions.c: dissolve_init :42 (the mask drawn from the host's random seed), common_deinit :75, common_process :85 */
static pe_weed_error_t dissolve_init(pe_weed_plant_t *inst) {
  pe_engine_t *e = engine();
  pe_weed_plant_t *ch = get_plant(inst, PE_LEAF_IN_CHANNELS, 0);
  pe_dissolve_mask_t *m = NULL;
  int64_t seed = 0;
  void *vp;
  if (!e || !ch) return PE_WEED_ERROR_PLUGIN_INVALID;
  if (w_num_elements(inst, "random_seed") > 0) w_leaf_get(inst, "random_seed", 0, &seed);
  if (pe_fx_dissolve_mask_create(e, get_int(ch, PE_LEAF_WIDTH), get_int(ch, PE_LEAF_HEIGHT), seed, &m) != PE_OK) {
    fprintf(stderr, "pe_weed_plugin: %s\n", pe_last_error());
    return PE_WEED_ERROR_MEMORY_ALLOCATION;
  }
  vp = m;
  w_leaf_set(inst, "plugin_internal", PE_WEED_SEED_VOIDPTR, 1, &vp);
  return PE_WEED_SUCCESS;
}

static pe_weed_error_t dissolve_deinit(pe_weed_plant_t *inst) {
  void *vp = get_ptr(inst, "plugin_internal");
  if (vp) {
    pe_fx_dissolve_mask_destroy((pe_dissolve_mask_t *)vp);
    vp = NULL;
    w_leaf_set(inst, "plugin_internal", PE_WEED_SEED_VOIDPTR, 1, &vp);
  }
  return PE_WEED_SUCCESS;
}

static pe_weed_error_t run_transition(int type, pe_weed_plant_t *inst, pe_weed_timecode_t tc) {
  static uint64_t rnd = 0;
  pe_frame_desc_t in1, in2, out;
  pe_engine_t *e = engine();
  const pe_dissolve_mask_t *mask = NULL;
  double bfd = 0.;
  if (!e) return PE_WEED_ERROR_PLUGIN_INVALID;
  channel_desc(get_plant(inst, PE_LEAF_IN_CHANNELS, 0), &in1);
  channel_desc(get_plant(inst, PE_LEAF_IN_CHANNELS, 1), &in2);
  channel_desc(get_plant(inst, PE_LEAF_OUT_CHANNELS, 0), &out);
  w_leaf_get(get_plant(inst, PE_LEAF_IN_PARAMETERS, 0), PE_LEAF_VALUE, 0, &bfd);
  if (type == 3) mask = (const pe_dissolve_mask_t *)get_ptr(inst, "plugin_internal");
  if (type == 4) {
    /* "rand replace" :102-119: the whole frame is in2 with probability `amount`, else in1 -- host logic (the plugin's own random
     * stream); the copy itself runs on the device as the degenerate iris rectangle (amount 1: all in2, amount 0: all in1) */
    if (!rnd) rnd = 0x9E3779B97F4A7C15ull ^ (uint64_t)tc ^ (uint64_t)(uintptr_t)inst;
    rnd ^= rnd << 13; rnd ^= rnd >> 7; rnd ^= rnd << 17;
    bfd = ((double)rnd * (1. / 18446744073709551616.) >= bfd) ? 0. : 1.;
    if (bfd == 0. && out.planes[0] == in1.planes[0]) return PE_WEED_SUCCESS;
    type = 0;
  }
  return rc_to_weed(pe_host_multi_transition(e, type, &in1, &in2, &out, bfd, mask));
}
#define PE_TRANSITION(name, type) \
  static pe_weed_error_t name(pe_weed_plant_t *inst, pe_weed_timecode_t tc) { return run_transition(type, inst, tc); }
PE_TRANSITION(irisr_process, 0)
PE_TRANSITION(irisc_process, 1)
PE_TRANSITION(fourw_process, 2)
PE_TRANSITION(dissolve_process, 3)
PE_TRANSITION(rreplace_process, 4)

/* ---- plant construction (what weed_channel_template_init / weed_integer_init / weed_filter_class_init of
 *      libweed/weed-plugin-utils.c:247-336 produce) ----------------------------------------------------------------- */

static pe_weed_plant_t *chantmpl(const char *name, int flags) {
  pe_weed_plant_t *t = w_plant_new(PE_WEED_PLANT_CHANNEL_TEMPLATE);
  if (!t) return NULL;
  set_str(t, PE_LEAF_NAME, name);
  set_int(t, PE_LEAF_FLAGS, flags);
  /* rows 32-byte aligned lets every kernel use 128-bit accesses (honoured at src/effects-weed.c:2319-2324) */
  set_int(t, PE_LEAF_ALIGNMENT_HINT, 32);
  return t;
}

static pe_weed_plant_t *int_param(const char *name, const char *label, int def, int min, int max) {
  pe_weed_plant_t *p = w_plant_new(PE_WEED_PLANT_PARAMETER_TEMPLATE), *gui;
  int32_t one = 1;
  if (!p) return NULL;
  set_str(p, PE_LEAF_NAME, name);
  set_int(p, PE_LEAF_PARAM_TYPE, PE_WEED_PARAM_INTEGER);
  set_int(p, PE_LEAF_DEFAULT, def);
  set_int(p, PE_LEAF_MIN, min);
  set_int(p, PE_LEAF_MAX, max);
  gui = w_plant_new(PE_WEED_PLANT_GUI);
  if (gui) {
    w_leaf_set(p, PE_LEAF_GUI, PE_WEED_SEED_PLANTPTR, 1, &gui);
    set_str(gui, PE_LEAF_LABEL, label);
    w_leaf_set(gui, PE_LEAF_USE_MNEMONIC, PE_WEED_SEED_BOOLEAN, 1, &one);
  }
  w_leaf_set(p, PE_LEAF_IS_TRANSITION, PE_WEED_SEED_BOOLEAN, 1, &one); /* weed_paramtmpl_declare_transition */
  return p;
}

/* weed_switch_init / weed_radio_init (weed-plugin-utils.c:350-367); group < 0: a plain switch */
static pe_weed_plant_t *switch_param(const char *name, const char *label, int def, int group, int flags) {
  pe_weed_plant_t *p = w_plant_new(PE_WEED_PLANT_PARAMETER_TEMPLATE), *gui;
  int32_t one = 1, d = def;
  if (!p) return NULL;
  set_str(p, PE_LEAF_NAME, name);
  set_int(p, PE_LEAF_PARAM_TYPE, PE_WEED_PARAM_SWITCH);
  w_leaf_set(p, PE_LEAF_DEFAULT, PE_WEED_SEED_BOOLEAN, 1, &d);
  gui = w_plant_new(PE_WEED_PLANT_GUI);
  if (gui) {
    w_leaf_set(p, PE_LEAF_GUI, PE_WEED_SEED_PLANTPTR, 1, &gui);
    set_str(gui, PE_LEAF_LABEL, label);
    w_leaf_set(gui, PE_LEAF_USE_MNEMONIC, PE_WEED_SEED_BOOLEAN, 1, &one);
  }
  if (group >= 0) set_int(p, PE_LEAF_GROUP, group);
  if (flags) set_int(p, PE_LEAF_FLAGS, flags);
  return p;
}

static int register_filter(pe_weed_plant_t *plugin_info, pe_weed_plant_t *fc);

/* weed_float_init / weed_colRGBi_init (weed-plugin-utils.c:369-382, :397-413) */
static pe_weed_plant_t *param_gui(pe_weed_plant_t *p, const char *label) {
  pe_weed_plant_t *gui = w_plant_new(PE_WEED_PLANT_GUI);
  int32_t one = 1;
  if (gui) {
    w_leaf_set(p, PE_LEAF_GUI, PE_WEED_SEED_PLANTPTR, 1, &gui);
    set_str(gui, PE_LEAF_LABEL, label);
    w_leaf_set(gui, PE_LEAF_USE_MNEMONIC, PE_WEED_SEED_BOOLEAN, 1, &one);
  }
  return gui;
}
static pe_weed_plant_t *float_param(const char *name, const char *label, double def, double min, double max) {
  pe_weed_plant_t *p = w_plant_new(PE_WEED_PLANT_PARAMETER_TEMPLATE);
  if (!p) return NULL;
  set_str(p, PE_LEAF_NAME, name);
  set_int(p, PE_LEAF_PARAM_TYPE, PE_WEED_PARAM_FLOAT);
  w_leaf_set(p, PE_LEAF_DEFAULT, PE_WEED_SEED_DOUBLE, 1, &def);
  w_leaf_set(p, PE_LEAF_MIN, PE_WEED_SEED_DOUBLE, 1, &min);
  w_leaf_set(p, PE_LEAF_MAX, PE_WEED_SEED_DOUBLE, 1, &max);
  param_gui(p, label);
  return p;
}
static pe_weed_plant_t *rgb_param(const char *name, const char *label, int r, int g, int b) {
  pe_weed_plant_t *p = w_plant_new(PE_WEED_PLANT_PARAMETER_TEMPLATE);
  int32_t def[3] = {r, g, b};
  if (!p) return NULL;
  set_str(p, PE_LEAF_NAME, name);
  set_int(p, PE_LEAF_PARAM_TYPE, PE_WEED_PARAM_COLOR);
  set_int(p, PE_LEAF_COLORSPACE, PE_WEED_COLORSPACE_RGB);
  w_leaf_set(p, PE_LEAF_DEFAULT, PE_WEED_SEED_INT, 3, def);
  set_int(p, PE_LEAF_MIN, 0);
  set_int(p, PE_LEAF_MAX, 255);
  param_gui(p, label);
  return p;
}

/* gdk/compositor.c:300-340: ONE repeating in channel template (max_repeats 0 = any number), one out channel, seven parameters, channel
 * sizes may vary.  (The RFX layout strings of the original only arrange its parameter window; they are not reproduced.) */
static int add_compositor(pe_weed_plant_t *plugin_info) {
  int palettes[] = {PE_PALETTE_RGB24, PE_PALETTE_BGR24, PE_PALETTE_RGBA32, PE_PALETTE_BGRA32};
  pe_weed_plant_t *fc = w_plant_new(PE_WEED_PLANT_FILTER_CLASS);
  pe_weed_plant_t *in_ct[1], *out_ct[1], *in_pt[7];
  pe_weed_process_f process_fn = compositor_process;
  pe_weed_init_f init_fn = common_init;
  const char *author = "lives_b200";
  int k;
  if (!fc) return -1;
  in_ct[0] = chantmpl("in channel 0", 0);
  out_ct[0] = chantmpl("out channel 0", 0);
  in_pt[0] = float_param("xoffs", "_X offset", 0., 0., 1.);
  in_pt[1] = float_param("yoffs", "_Y offset", 0., 0., 1.);
  in_pt[2] = float_param("scalex", "Scale _width", 1., 0., 1.);
  in_pt[3] = float_param("scaley", "Scale _height", 1., 0., 1.);
  in_pt[4] = float_param("alpha", "_Alpha", 1., 0., 1.);
  in_pt[5] = rgb_param("bgcol", "_Background color", 0, 0, 0);
  in_pt[6] = switch_param("revz", "Invert _Z Index", 0, -1, 0);
  if (!in_ct[0] || !out_ct[0]) return -1;
  for (k = 0; k < 7; k++) if (!in_pt[k]) return -1;
  set_int(in_ct[0], PE_LEAF_MAX_REPEATS, 0);
  set_str(fc, PE_LEAF_NAME, "compositor");
  w_leaf_set(fc, PE_LEAF_AUTHOR, PE_WEED_SEED_STRING, 1, &author);
  set_int(fc, PE_LEAF_VERSION, 1);
  set_int(fc, PE_LEAF_FLAGS, PE_WEED_FILTER_CHANNEL_SIZES_MAY_VARY);
  w_leaf_set(fc, PE_LEAF_INIT_FUNC, PE_WEED_SEED_FUNCPTR, 1, &init_fn);
  w_leaf_set(fc, PE_LEAF_PROCESS_FUNC, PE_WEED_SEED_FUNCPTR, 1, &process_fn);
  w_leaf_set(fc, PE_LEAF_IN_CHANNEL_TEMPLATES, PE_WEED_SEED_PLANTPTR, 1, in_ct);
  w_leaf_set(fc, PE_LEAF_OUT_CHANNEL_TEMPLATES, PE_WEED_SEED_PLANTPTR, 1, out_ct);
  w_leaf_set(fc, PE_LEAF_IN_PARAMETER_TEMPLATES, PE_WEED_SEED_PLANTPTR, 7, in_pt);
  w_leaf_set(fc, PE_LEAF_OUT_PARAMETER_TEMPLATES, PE_WEED_SEED_PLANTPTR, 0, NULL);
  w_leaf_set(fc, PE_LEAF_PALETTE_LIST, PE_WEED_SEED_INT, 4, palettes);
  return register_filter(plugin_info, fc);
}

/* weed_plugin_info_add_filter_class (weed-plugin-utils.c:308-320) */
static int register_filter(pe_weed_plant_t *plugin_info, pe_weed_plant_t *fc) {
  pe_weed_size_t n = w_num_elements(plugin_info, PE_LEAF_FILTERS), i;
  pe_weed_plant_t **filters = (pe_weed_plant_t **)w_malloc((n + 1) * sizeof(pe_weed_plant_t *));
  if (!filters) return -1;
  for (i = 0; i < n; i++) w_leaf_get(plugin_info, PE_LEAF_FILTERS, i, &filters[i]);
  filters[n] = fc;
  w_leaf_set(plugin_info, PE_LEAF_FILTERS, PE_WEED_SEED_PLANTPTR, n + 1, filters);
  w_leaf_set(fc, PE_LEAF_PLUGIN_INFO, PE_WEED_SEED_PLANTPTR, 1, &plugin_info);
  w_free(filters);
  return 0;
}

/* slide_over.c:157-196: two in channels, one out channel (not in-place), the transition value, five direction radios (a change
 * re-inits the instance), two switches */
static int add_slide_over(pe_weed_plant_t *plugin_info) {
  int palettes[] = {PE_PALETTE_RGB24, PE_PALETTE_BGR24, PE_PALETTE_RGBA32, PE_PALETTE_BGRA32, PE_PALETTE_ARGB32, PE_PALETTE_YUV888,
                    PE_PALETTE_YUVA8888, PE_PALETTE_UYVY, PE_PALETTE_YUYV}; /* ALL_PACKED_PALETTES_PLUS */
  pe_weed_plant_t *fc = w_plant_new(PE_WEED_PLANT_FILTER_CLASS);
  pe_weed_plant_t *in_ct[2], *out_ct[1], *in_pt[8];
  pe_weed_init_f init_fn = sover_init;
  pe_weed_process_f process_fn = sover_process;
  const char *author = "lives_b200";
  const int re = PE_WEED_PARAMETER_REINIT_ON_VALUE_CHANGE;
  int k;
  if (!fc) return -1;
  in_ct[0] = chantmpl("in channel 0", 0);
  in_ct[1] = chantmpl("in channel 1", 0);
  out_ct[0] = chantmpl("out channel 0", 0);
  in_pt[0] = int_param("amount", "Transition _value", 0, 0, 255);
  in_pt[1] = switch_param("dir_rand", "_Random", 1, 1, re);
  in_pt[2] = switch_param("dir_r2l", "_Right to left", 0, 1, re);
  in_pt[3] = switch_param("dir_l2r", "_Left to right", 0, 1, re);
  in_pt[4] = switch_param("dir_b2t", "_Bottom to top", 0, 1, re);
  in_pt[5] = switch_param("dir_t2b", "_Top to bottom", 0, 1, re);
  in_pt[6] = switch_param("mlower", "_Slide lower clip", 1, -1, 0);
  in_pt[7] = switch_param("mupper", "_Slide upper clip", 0, -1, 0);
  if (!in_ct[0] || !in_ct[1] || !out_ct[0]) return -1;
  for (k = 0; k < 8; k++) if (!in_pt[k]) return -1;
  set_str(fc, PE_LEAF_NAME, "slide over");
  w_leaf_set(fc, PE_LEAF_AUTHOR, PE_WEED_SEED_STRING, 1, &author);
  set_int(fc, PE_LEAF_VERSION, 1);
  set_int(fc, PE_LEAF_FLAGS, 0);
  w_leaf_set(fc, PE_LEAF_INIT_FUNC, PE_WEED_SEED_FUNCPTR, 1, &init_fn);
  w_leaf_set(fc, PE_LEAF_PROCESS_FUNC, PE_WEED_SEED_FUNCPTR, 1, &process_fn);
  w_leaf_set(fc, PE_LEAF_IN_CHANNEL_TEMPLATES, PE_WEED_SEED_PLANTPTR, 2, in_ct);
  w_leaf_set(fc, PE_LEAF_OUT_CHANNEL_TEMPLATES, PE_WEED_SEED_PLANTPTR, 1, out_ct);
  w_leaf_set(fc, PE_LEAF_IN_PARAMETER_TEMPLATES, PE_WEED_SEED_PLANTPTR, 8, in_pt);
  w_leaf_set(fc, PE_LEAF_OUT_PARAMETER_TEMPLATES, PE_WEED_SEED_PLANTPTR, 0, NULL);
  w_leaf_set(fc, PE_LEAF_PALETTE_LIST, PE_WEED_SEED_INT, 9, palettes);
  return register_filter(plugin_info, fc);
}

static int add_filter(pe_weed_plant_t *plugin_info, const char *name, int flags, int *palettes, int npal, pe_weed_init_f init_fn,
                      pe_weed_process_f process_fn, const char *pname, const char *plabel, int pdef) {
  pe_weed_plant_t *fc = w_plant_new(PE_WEED_PLANT_FILTER_CLASS);
  pe_weed_plant_t *in_ct[2], *out_ct[1], *in_pt[1];
  const char *author = "lives_b200";
  int32_t version = 1;
  if (!fc) return -1;
  in_ct[0] = chantmpl("in channel 0", 0);
  in_ct[1] = chantmpl("in channel 1", 0);
  out_ct[0] = chantmpl("out channel 0", PE_WEED_CHANNEL_CAN_DO_INPLACE);
  in_pt[0] = int_param(pname, plabel, pdef, 0, 255);
  if (!in_ct[0] || !in_ct[1] || !out_ct[0] || !in_pt[0]) return -1;
  set_str(fc, PE_LEAF_NAME, name);
  w_leaf_set(fc, PE_LEAF_AUTHOR, PE_WEED_SEED_STRING, 1, &author);
  set_int(fc, PE_LEAF_VERSION, version);
  set_int(fc, PE_LEAF_FLAGS, flags);
  if (init_fn) w_leaf_set(fc, PE_LEAF_INIT_FUNC, PE_WEED_SEED_FUNCPTR, 1, &init_fn);
  w_leaf_set(fc, PE_LEAF_PROCESS_FUNC, PE_WEED_SEED_FUNCPTR, 1, &process_fn);
  w_leaf_set(fc, PE_LEAF_IN_CHANNEL_TEMPLATES, PE_WEED_SEED_PLANTPTR, 2, in_ct);
  w_leaf_set(fc, PE_LEAF_OUT_CHANNEL_TEMPLATES, PE_WEED_SEED_PLANTPTR, 1, out_ct);
  w_leaf_set(fc, PE_LEAF_IN_PARAMETER_TEMPLATES, PE_WEED_SEED_PLANTPTR, 1, in_pt);
  w_leaf_set(fc, PE_LEAF_OUT_PARAMETER_TEMPLATES, PE_WEED_SEED_PLANTPTR, 0, NULL);
  w_leaf_set(fc, PE_LEAF_PALETTE_LIST, PE_WEED_SEED_INT, (pe_weed_size_t)npal, palettes);
  return register_filter(plugin_info, fc);
}

/* a filter class from ready-made templates (weed_filter_class_init, weed-plugin-utils.c:258-305) */
static int add_class(pe_weed_plant_t *plugin_info, const char *name, int flags, const int *palettes, int npal, pe_weed_init_f init_fn,
                     pe_weed_process_f process_fn, pe_weed_init_f deinit_fn, pe_weed_plant_t **in_ct, int nin, pe_weed_plant_t **out_ct,
                     pe_weed_plant_t **in_pt, int npar) {
  pe_weed_plant_t *fc = w_plant_new(PE_WEED_PLANT_FILTER_CLASS);
  const char *author = "lives_b200";
  int k;
  if (!fc) return -1;
  for (k = 0; k < nin; k++) if (!in_ct[k]) return -1;
  for (k = 0; k < npar; k++) if (!in_pt[k]) return -1;
  if (!out_ct[0]) return -1;
  set_str(fc, PE_LEAF_NAME, name);
  w_leaf_set(fc, PE_LEAF_AUTHOR, PE_WEED_SEED_STRING, 1, &author);
  set_int(fc, PE_LEAF_VERSION, 1);
  set_int(fc, PE_LEAF_FLAGS, flags);
  if (init_fn) w_leaf_set(fc, PE_LEAF_INIT_FUNC, PE_WEED_SEED_FUNCPTR, 1, &init_fn);
  if (deinit_fn) w_leaf_set(fc, PE_LEAF_DEINIT_FUNC, PE_WEED_SEED_FUNCPTR, 1, &deinit_fn);
  w_leaf_set(fc, PE_LEAF_PROCESS_FUNC, PE_WEED_SEED_FUNCPTR, 1, &process_fn);
  w_leaf_set(fc, PE_LEAF_IN_CHANNEL_TEMPLATES, PE_WEED_SEED_PLANTPTR, (pe_weed_size_t)nin, in_ct);
  w_leaf_set(fc, PE_LEAF_OUT_CHANNEL_TEMPLATES, PE_WEED_SEED_PLANTPTR, 1, out_ct);
  w_leaf_set(fc, PE_LEAF_IN_PARAMETER_TEMPLATES, PE_WEED_SEED_PLANTPTR, (pe_weed_size_t)npar, npar ? in_pt : NULL);
  w_leaf_set(fc, PE_LEAF_OUT_PARAMETER_TEMPLATES, PE_WEED_SEED_PLANTPTR, 0, NULL);
  w_leaf_set(fc, PE_LEAF_PALETTE_LIST, PE_WEED_SEED_INT, (pe_weed_size_t)npal, (void *)palettes);
  return register_filter(plugin_info, fc);
}

/* softlight.c:168-183: one in, one out channel, no parameters; the in channel template prefers unclamped YUV */
static int add_softlight(pe_weed_plant_t *plugin_info) {
  const int palettes[] = {PE_PALETTE_YUV444P, PE_PALETTE_YUVA4444P, PE_PALETTE_YUV422P, PE_PALETTE_YUV420P, PE_PALETTE_YVU420P};
  pe_weed_plant_t *in_ct[1], *out_ct[1];
  in_ct[0] = chantmpl("in channel 0", 0);
  out_ct[0] = chantmpl("out channel 0", 0);
  if (in_ct[0]) set_int(in_ct[0], "YUV_clamping", PE_YUV_CLAMPING_UNCLAMPED);
  return add_class(plugin_info, "softlight", 0, palettes, 5, common_init, softlight_process, NULL, in_ct, 1, out_ct, NULL, 0);
}

/* layout_blends.c:127-158 */
static int add_triple_split(pe_weed_plant_t *plugin_info) {
  const int palettes[] = {PE_PALETTE_RGB24, PE_PALETTE_BGR24};
  pe_weed_plant_t *in_ct[2], *out_ct[1], *in_pt[7];
  in_ct[0] = chantmpl("in channel 0", 0);
  in_ct[1] = chantmpl("in channel 1", 0);
  out_ct[0] = chantmpl("out channel 0", PE_WEED_CHANNEL_CAN_DO_INPLACE);
  in_pt[0] = float_param("start", "_Start", 0.666667, 0., 1.);
  in_pt[1] = switch_param("sym", "Make s_ymmetrical", 1, 1, 0);
  in_pt[2] = switch_param("usend", "Use _end value", 0, 1, 0);
  in_pt[3] = float_param("end", "_End", 0.333333, 0., 1.);
  in_pt[4] = switch_param("vert", "Split _horizontally", 0, -1, 0);
  in_pt[5] = float_param("borderw", "Border _width", 0., 0., 0.5);
  in_pt[6] = rgb_param("borderc", "Border _colour", 0, 0, 0);
  return add_class(plugin_info, "triple split", 0, palettes, 2, common_init, tsplit_process, NULL, in_ct, 2, out_ct, in_pt, 7);
}

/* multi_transitions.c:229-330: five filters over every packed palette, one transition parameter each */
static int add_transition(pe_weed_plant_t *plugin_info, const char *name, int flags, int out_flags, pe_weed_init_f init_fn,
                          pe_weed_process_f process_fn, pe_weed_init_f deinit_fn) {
  const int palettes[] = {PE_PALETTE_RGB24, PE_PALETTE_BGR24, PE_PALETTE_RGBA32, PE_PALETTE_BGRA32, PE_PALETTE_ARGB32, PE_PALETTE_UYVY,
                          PE_PALETTE_YUYV, PE_PALETTE_YUV888, PE_PALETTE_YUVA8888}; /* ALL_PACKED_PALETTES */
  pe_weed_plant_t *in_ct[2], *out_ct[1], *in_pt[1];
  int32_t one = 1;
  in_ct[0] = chantmpl("in channel 0", 0);
  in_ct[1] = chantmpl("in channel 1", 0);
  out_ct[0] = chantmpl("out channel 0", out_flags);
  in_pt[0] = float_param("amount", "_Transition", 0., 0., 1.);
  if (in_pt[0]) w_leaf_set(in_pt[0], PE_LEAF_IS_TRANSITION, PE_WEED_SEED_BOOLEAN, 1, &one);
  return add_class(plugin_info, name, flags, palettes, 9, init_fn, process_fn, deinit_fn, in_ct, 2, out_ct, in_pt, 1);
}

/* ---- entry points --------------------------------------------------------------------------------------------------- */

pe_weed_plant_t *weed_setup(pe_weed_bootstrap_f weed_boot) {
  pe_weed_default_getter_f getp = NULL;
  pe_weed_plant_t *host_info, *plugin_info = NULL;
  int all_rgb[] = {PE_PALETTE_RGB24, PE_PALETTE_BGR24, PE_PALETTE_RGBA32, PE_PALETTE_BGRA32, PE_PALETTE_ARGB32}; /* ALL_RGB_PALETTES */
  int rgb24[] = {PE_PALETTE_RGB24, PE_PALETTE_BGR24};
  int32_t package_version = 1;
  if (!weed_boot) return NULL;
  /* weed_plugin_info_init (weed-plugin-utils.c:164-233): the default getter bootstraps weed_leaf_get, the rest follows */
  host_info = (*weed_boot)(&getp, PE_WEED_API_MIN, PE_WEED_API_MAX, PE_WEED_FILTER_API_MIN, PE_WEED_FILTER_API_MAX);
  if (!host_info || !getp) return NULL;
  if ((*getp)(host_info, PE_LEAF_GET_FUNC, (void *)&w_leaf_get) != PE_WEED_SUCCESS || !w_leaf_get) return NULL;
  if ((*getp)(host_info, PE_LEAF_MALLOC_FUNC, (void *)&w_malloc) != PE_WEED_SUCCESS) return NULL;
  if ((*getp)(host_info, PE_LEAF_FREE_FUNC, (void *)&w_free) != PE_WEED_SUCCESS) return NULL;
  if (w_leaf_get(host_info, PE_LEAF_SET_FUNC, 0, &w_leaf_set) != PE_WEED_SUCCESS) return NULL;
  if (w_leaf_get(host_info, PE_LEAF_PLANT_NEW_FUNC, 0, &w_plant_new) != PE_WEED_SUCCESS) return NULL;
  if (w_leaf_get(host_info, PE_LEAF_NUM_ELEMENTS_FUNC, 0, &w_num_elements) != PE_WEED_SUCCESS) return NULL;
  if (!w_leaf_set || !w_plant_new || !w_num_elements || !w_malloc || !w_free) return NULL;
  /* the host may hand us a PLUGIN_INFO to fill in (weed-plugin-utils.c:220-229) */
  if (w_num_elements(host_info, PE_LEAF_PLUGIN_INFO) > 0) {
    int32_t type = 0;
    if (w_leaf_get(host_info, PE_LEAF_PLUGIN_INFO, 0, &plugin_info) != PE_WEED_SUCCESS) return NULL;
    if (plugin_info) w_leaf_get(plugin_info, PE_LEAF_TYPE, 0, &type);
    if (type != PE_WEED_PLANT_PLUGIN_INFO) plugin_info = NULL;
  }
  if (!plugin_info && !(plugin_info = w_plant_new(PE_WEED_PLANT_PLUGIN_INFO))) return NULL;
  w_leaf_set(plugin_info, PE_LEAF_HOST_INFO, PE_WEED_SEED_PLANTPTR, 1, &host_info);

  /* simple_blend.c:218-292 (filter order kept) */
  if (add_filter(plugin_info, "chroma blend", PE_WEED_FILTER_PREF_LINEAR_GAMMA, all_rgb, 5, common_init, chroma_process, "amount",
                 "Blend _amount", 128) ||
      add_filter(plugin_info, "luma overlay", 0, all_rgb, 5, common_init, lumo_process, "threshold", "luma _threshold", 64) ||
      add_filter(plugin_info, "luma underlay", 0, all_rgb, 5, common_init, lumu_process, "threshold", "luma _threshold", 64) ||
      add_filter(plugin_info, "negative luma overlay", 0, all_rgb, 5, common_init, nlumo_process, "threshold", "luma _threshold", 64))
    return NULL;
  /* multi_blends.c:196-296 */
  if (add_filter(plugin_info, "blend_multiply", PE_WEED_FILTER_PREF_LINEAR_GAMMA, rgb24, 2, common_init, mpy_process, "amount",
                 "Blend _amount", 128) ||
      add_filter(plugin_info, "blend_screen", PE_WEED_FILTER_PREF_LINEAR_GAMMA, rgb24, 2, common_init, screen_process, "amount",
                 "Blend _amount", 128) ||
      add_filter(plugin_info, "blend_darken", PE_WEED_FILTER_PREF_LINEAR_GAMMA, rgb24, 2, common_init, darken_process, "amount",
                 "Blend _amount", 128) ||
      add_filter(plugin_info, "blend_lighten", PE_WEED_FILTER_PREF_LINEAR_GAMMA, rgb24, 2, common_init, lighten_process, "amount",
                 "Blend _amount", 128) ||
      add_filter(plugin_info, "blend_overlay", PE_WEED_FILTER_PREF_LINEAR_GAMMA, rgb24, 2, common_init, overlay_process, "amount",
                 "Blend _amount", 128) ||
      add_filter(plugin_info, "blend_dodge", PE_WEED_FILTER_PREF_LINEAR_GAMMA, rgb24, 2, common_init, dodge_process, "amount",
                 "Blend _amount", 128) ||
      add_filter(plugin_info, "blend_burn", PE_WEED_FILTER_PREF_LINEAR_GAMMA, rgb24, 2, common_init, burn_process, "amount",
                 "Blend _amount", 128))
    return NULL;
  if (add_slide_over(plugin_info)) return NULL;
  if (add_compositor(plugin_info)) return NULL;
  if (add_softlight(plugin_info) || add_triple_split(plugin_info)) return NULL;
  /* multi_transitions.c:242-326 (filter order kept; "4 way split" is not in-place :268, "dissolve" re-inits when the size changes :281) */
  if (add_transition(plugin_info, "iris rectangle", 0, PE_WEED_CHANNEL_CAN_DO_INPLACE, common_init, irisr_process, NULL) ||
      add_transition(plugin_info, "iris circle", 0, PE_WEED_CHANNEL_CAN_DO_INPLACE, common_init, irisc_process, NULL) ||
      add_transition(plugin_info, "4 way split", 0, 0, common_init, fourw_process, NULL) ||
      add_transition(plugin_info, "dissolve", 0, PE_WEED_CHANNEL_CAN_DO_INPLACE | PE_WEED_CHANNEL_REINIT_ON_SIZE_CHANGE, dissolve_init,
                     dissolve_process, dissolve_deinit) ||
      add_transition(plugin_info, "rand replace", 0, PE_WEED_CHANNEL_CAN_DO_INPLACE, common_init, rreplace_process, NULL))
    return NULL;
  /* simple_blend.c:281-287, the file's fifth filter (registered last here so that the filter indices of the earlier rounds stay) */
  if (add_filter(plugin_info, "averaged luma overlay", PE_WEED_FILTER_PREF_LINEAR_GAMMA, all_rgb, 5, common_init, avlumo_process,
                 "threshold", "luma _threshold", 64))
    return NULL;
  set_int(plugin_info, PE_LEAF_VERSION, package_version);
  return plugin_info;
}

void weed_desetup(void) {
  /* the engine is the process-wide one (pe_engine_shared): it outlives this plugin */
}
