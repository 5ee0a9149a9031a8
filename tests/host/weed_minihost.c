/* weed_minihost.c -- TEST INFRASTRUCTURE ONLY.
 *
 * A minimal libweed *host*, linked against the reference's own libweed
 * (oracle/_ref/libweed*.so, compiled unchanged by oracle/build_ref.py).
 * It loads ANY weed effect plugin exactly the way LiVES does
 * (src/effects-weed.c:4468-4568: dlopen RTLD_NOW|RTLD_LOCAL, dlsym "weed_setup",
 * setup_fn(weed_bootstrap)), builds channel / parameter / instance plants and
 * calls init_func / process_func on caller-supplied frames.
 *
 * The parity tests use it twice with identical inputs: once on the reference's
 * simple_blend.so / multi_blends.so and once on our CUDA plugin
 * (lives_b200/csrc/libpe_weed_plugin.so) -- the drop-in check for boundary B1.
 *
 * Exported flat API (ctypes):
 *   int  mh_open(const char *plugin_path);              -> handle >= 0
 *   int  mh_num_filters(int h);
 *   int  mh_filter_name(int h, int idx, char *buf, int len);
 *   int  mh_filter_flags(int h, int idx);
 *   int  mh_run2(int h, int filter_idx, int palette, int width, int height,
 *                void *src1, int rs1, void *src2, int rs2, void *dst, int rsd,
 *                int int_param0, int nframes);
 *        two in channels, one out channel, one integer in-parameter; dst may
 *        equal src1 (in-place, as LiVES does for CAN_DO_INPLACE channels);
 *        calls init once, process nframes times, deinit once.
 *   int  mh_run2v(... as mh_run2 up to rsd ..., int nparams, const int *vals, int nframes);
 *        the same with one in-parameter per template of the filter: vals[k] is stored as an int or a boolean,
 *        whichever seed type the template's default has (radio / switch parameters are booleans).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <dlfcn.h>

#include <weed/weed-host.h>
#include <weed/weed.h>
#include <weed/weed-effects.h>
#include <weed/weed-palettes.h>
#include <weed/weed-utils.h>
#include <weed/weed-host-utils.h>

#define MH_MAX 64

typedef struct {
  void *dl;
  char path[512];
  weed_plant_t *plugin_info;
  weed_plant_t **filters;
  int nfilters;
} mh_plugin;

static mh_plugin mh_tab[MH_MAX];
static int mh_inited = 0;

static void mh_init_once(void) {
  if (mh_inited) return;
  libweed_init(WEED_ABI_VERSION, 0);
  mh_inited = 1;
}

int mh_open(const char *path) {
  int h;
  weed_setup_f setup_fn;
  mh_init_once();
  /* a plugin is set up once per process: opening the same path again returns its handle */
  for (h = 0; h < MH_MAX; h++) if (mh_tab[h].dl && !strcmp(mh_tab[h].path, path)) return h;
  for (h = 0; h < MH_MAX; h++) if (!mh_tab[h].dl) break;
  if (h == MH_MAX) return -1;
  mh_tab[h].dl = dlopen(path, RTLD_NOW | RTLD_LOCAL);
  if (!mh_tab[h].dl) { fprintf(stderr, "minihost: dlopen %s: %s\n", path, dlerror()); return -2; }
  setup_fn = (weed_setup_f)dlsym(mh_tab[h].dl, "weed_setup");
  if (!setup_fn) { dlclose(mh_tab[h].dl); mh_tab[h].dl = NULL; return -3; }
  mh_tab[h].plugin_info = (*setup_fn)(weed_bootstrap);
  if (!mh_tab[h].plugin_info) { dlclose(mh_tab[h].dl); mh_tab[h].dl = NULL; return -4; }
  mh_tab[h].filters = weed_get_plantptr_array_counted(mh_tab[h].plugin_info, WEED_LEAF_FILTERS, &mh_tab[h].nfilters);
  strncpy(mh_tab[h].path, path, sizeof(mh_tab[h].path) - 1);
  return h;
}

int mh_num_filters(int h) { return (h < 0 || h >= MH_MAX || !mh_tab[h].dl) ? -1 : mh_tab[h].nfilters; }

int mh_filter_name(int h, int idx, char *buf, int len) {
  char *name;
  if (mh_num_filters(h) <= idx || idx < 0) return -1;
  name = weed_get_string_value(mh_tab[h].filters[idx], WEED_LEAF_NAME, NULL);
  if (!name) return -2;
  strncpy(buf, name, len - 1); buf[len - 1] = 0;
  free(name);
  return 0;
}

int mh_filter_flags(int h, int idx) {
  if (mh_num_filters(h) <= idx || idx < 0) return -1;
  return weed_get_int_value(mh_tab[h].filters[idx], WEED_LEAF_FLAGS, NULL);
}

static weed_plant_t *mh_channel(weed_plant_t *tmpl, int palette, int width, int height, void *pixels, int rowstride) {
  weed_plant_t *ch = weed_plant_new(WEED_PLANT_CHANNEL);
  weed_set_plantptr_value(ch, WEED_LEAF_TEMPLATE, tmpl);
  weed_set_int_value(ch, WEED_LEAF_WIDTH, width); /* macropixels == pixels for the RGB palettes */
  weed_set_int_value(ch, WEED_LEAF_HEIGHT, height);
  weed_set_int_value(ch, WEED_LEAF_CURRENT_PALETTE, palette);
  weed_set_int_value(ch, WEED_LEAF_ROWSTRIDES, rowstride);
  weed_set_voidptr_value(ch, WEED_LEAF_PIXEL_DATA, pixels);
  return ch;
}

int mh_run2(int h, int fidx, int palette, int width, int height, void *src1, int rs1, void *src2, int rs2,
            void *dst, int rsd, int int_param0, int nframes) {
  weed_plant_t *filter, *inst, *in_ch[2], *out_ch, *param;
  weed_plant_t **ictm, **octm, **iptm;
  weed_init_f init_fn;
  weed_process_f process_fn;
  weed_deinit_f deinit_fn;
  weed_error_t err = WEED_SUCCESS;
  int n;
  if (mh_num_filters(h) <= fidx || fidx < 0) return -1;
  filter = mh_tab[h].filters[fidx];
  ictm = weed_get_plantptr_array_counted(filter, WEED_LEAF_IN_CHANNEL_TEMPLATES, &n);
  if (n < 2) return -2;
  octm = weed_get_plantptr_array_counted(filter, WEED_LEAF_OUT_CHANNEL_TEMPLATES, &n);
  if (n < 1) return -3;
  iptm = weed_get_plantptr_array_counted(filter, WEED_LEAF_IN_PARAMETER_TEMPLATES, &n);
  if (n < 1) return -4;

  in_ch[0] = mh_channel(ictm[0], palette, width, height, src1, rs1);
  in_ch[1] = mh_channel(ictm[1], palette, width, height, src2, rs2);
  out_ch = mh_channel(octm[0], palette, width, height, dst, rsd);
  param = weed_plant_new(WEED_PLANT_PARAMETER);
  weed_set_plantptr_value(param, WEED_LEAF_TEMPLATE, iptm[0]);
  weed_set_int_value(param, WEED_LEAF_VALUE, int_param0);

  inst = weed_plant_new(WEED_PLANT_FILTER_INSTANCE);
  weed_set_plantptr_value(inst, WEED_LEAF_FILTER_CLASS, filter);
  weed_set_plantptr_array(inst, WEED_LEAF_IN_CHANNELS, 2, in_ch);
  weed_set_plantptr_array(inst, WEED_LEAF_OUT_CHANNELS, 1, &out_ch);
  weed_set_plantptr_array(inst, WEED_LEAF_IN_PARAMETERS, 1, &param);

  init_fn = (weed_init_f)weed_get_funcptr_value(filter, WEED_LEAF_INIT_FUNC, NULL);
  process_fn = (weed_process_f)weed_get_funcptr_value(filter, WEED_LEAF_PROCESS_FUNC, NULL);
  deinit_fn = (weed_deinit_f)weed_get_funcptr_value(filter, WEED_LEAF_DEINIT_FUNC, NULL);
  if (!process_fn) return -5;
  if (init_fn) err = (*init_fn)(inst);
  for (int f = 0; f < nframes && err == WEED_SUCCESS; f++) err = (*process_fn)(inst, (weed_timecode_t)f);
  if (deinit_fn) (*deinit_fn)(inst);

  weed_plant_free(inst); weed_plant_free(param);
  weed_plant_free(in_ch[0]); weed_plant_free(in_ch[1]); weed_plant_free(out_ch);
  free(ictm); free(octm); free(iptm);
  return (int)err;
}

int mh_run2v(int h, int fidx, int palette, int width, int height, void *src1, int rs1, void *src2, int rs2,
             void *dst, int rsd, int nparams, const int *vals, int nframes) {
  weed_plant_t *filter, *inst, *in_ch[2], *out_ch, *params[32];
  weed_plant_t **ictm, **octm, **iptm;
  weed_init_f init_fn;
  weed_process_f process_fn;
  weed_deinit_f deinit_fn;
  weed_error_t err = WEED_SUCCESS;
  int n, np;
  if (mh_num_filters(h) <= fidx || fidx < 0) return -1;
  filter = mh_tab[h].filters[fidx];
  ictm = weed_get_plantptr_array_counted(filter, WEED_LEAF_IN_CHANNEL_TEMPLATES, &n);
  if (n < 2) return -2;
  octm = weed_get_plantptr_array_counted(filter, WEED_LEAF_OUT_CHANNEL_TEMPLATES, &n);
  if (n < 1) return -3;
  iptm = weed_get_plantptr_array_counted(filter, WEED_LEAF_IN_PARAMETER_TEMPLATES, &np);
  if (np < 1 || np > 32 || np != nparams) return -4;

  in_ch[0] = mh_channel(ictm[0], palette, width, height, src1, rs1);
  in_ch[1] = mh_channel(ictm[1], palette, width, height, src2, rs2);
  out_ch = mh_channel(octm[0], palette, width, height, dst, rsd);
  for (int k = 0; k < np; k++) {
    params[k] = weed_plant_new(WEED_PLANT_PARAMETER);
    weed_set_plantptr_value(params[k], WEED_LEAF_TEMPLATE, iptm[k]);
    if (weed_leaf_seed_type(iptm[k], WEED_LEAF_DEFAULT) == WEED_SEED_BOOLEAN) weed_set_boolean_value(params[k], WEED_LEAF_VALUE, vals[k]);
    else weed_set_int_value(params[k], WEED_LEAF_VALUE, vals[k]);
  }

  inst = weed_plant_new(WEED_PLANT_FILTER_INSTANCE);
  weed_set_plantptr_value(inst, WEED_LEAF_FILTER_CLASS, filter);
  weed_set_plantptr_array(inst, WEED_LEAF_IN_CHANNELS, 2, in_ch);
  weed_set_plantptr_array(inst, WEED_LEAF_OUT_CHANNELS, 1, &out_ch);
  weed_set_plantptr_array(inst, WEED_LEAF_IN_PARAMETERS, np, params);

  init_fn = (weed_init_f)weed_get_funcptr_value(filter, WEED_LEAF_INIT_FUNC, NULL);
  process_fn = (weed_process_f)weed_get_funcptr_value(filter, WEED_LEAF_PROCESS_FUNC, NULL);
  deinit_fn = (weed_deinit_f)weed_get_funcptr_value(filter, WEED_LEAF_DEINIT_FUNC, NULL);
  if (!process_fn) return -5;
  if (init_fn) err = (*init_fn)(inst);
  for (int f = 0; f < nframes && err == WEED_SUCCESS; f++) err = (*process_fn)(inst, (weed_timecode_t)f);
  if (deinit_fn) (*deinit_fn)(inst);

  weed_plant_free(inst);
  for (int k = 0; k < np; k++) weed_plant_free(params[k]);
  weed_plant_free(in_ch[0]); weed_plant_free(in_ch[1]); weed_plant_free(out_ch);
  free(ictm); free(octm); free(iptm);
  return (int)err;
}

/* ---- layers: a weed_layer_t built the way src/layers.c builds one (WEED_PLANT_LAYER = 128, src/layers.h:14; leaves of
 *      libweed/weed-effects.h:269-277,350-375), for the weed_layer_t drop-ins of lives_b200/libpe_weed_layer.so.  Pixel planes are
 *      COPIED into malloc'ed (64-byte aligned) buffers the layer then owns: the drop-ins release / replace them with free(). */
#define MH_PLANT_LAYER 128

void *mh_layer_new(int palette, int width_macropixels, int height, int nplanes, void **planes, int *rowstrides, int *plane_rows,
                   int clamping, int sampling, int subspace, int gamma_type, int with_yuv_leaves) {
  weed_plant_t *l;
  void *pd[4] = {NULL, NULL, NULL, NULL};
  int i;
  mh_init_once();
  l = weed_plant_new(MH_PLANT_LAYER);
  weed_set_int_value(l, WEED_LEAF_CURRENT_PALETTE, palette);
  weed_set_int_value(l, WEED_LEAF_WIDTH, width_macropixels);
  weed_set_int_value(l, WEED_LEAF_HEIGHT, height);
  if (nplanes > 0 && planes) {
    for (i = 0; i < nplanes; i++) {
      size_t n = (size_t)rowstrides[i] * (size_t)plane_rows[i];
      if (posix_memalign(&pd[i], 64, n + 64)) return NULL;
      memcpy(pd[i], planes[i], n);
      memset((char *)pd[i] + n, 0, 64);
    }
    weed_set_voidptr_array(l, WEED_LEAF_PIXEL_DATA, nplanes, pd);
    weed_set_int_array(l, WEED_LEAF_ROWSTRIDES, nplanes, rowstrides);
  }
  if (with_yuv_leaves) {
    weed_set_int_value(l, WEED_LEAF_YUV_CLAMPING, clamping);
    weed_set_int_value(l, WEED_LEAF_YUV_SAMPLING, sampling);
    weed_set_int_value(l, WEED_LEAF_YUV_SUBSPACE, subspace);
  }
  if (gamma_type) weed_set_int_value(l, WEED_LEAF_GAMMA_TYPE, gamma_type);
  return l;
}

int mh_layer_has(void *layer, const char *key) { return weed_plant_has_leaf((weed_plant_t *)layer, key) ? 1 : 0; }
int mh_layer_int(void *layer, const char *key, int dflt) {
  return weed_plant_has_leaf((weed_plant_t *)layer, key) ? weed_get_int_value((weed_plant_t *)layer, key, NULL) : dflt;
}
void mh_layer_set_int(void *layer, const char *key, int v) { weed_set_int_value((weed_plant_t *)layer, key, v); }
int mh_layer_nplanes(void *layer) { return weed_leaf_num_elements((weed_plant_t *)layer, WEED_LEAF_PIXEL_DATA); }
void *mh_layer_plane(void *layer, int p) {
  int n = 0;
  void **pd = weed_get_voidptr_array_counted((weed_plant_t *)layer, WEED_LEAF_PIXEL_DATA, &n);
  void *r = (pd && p < n) ? pd[p] : NULL;
  if (pd) free(pd);
  return r;
}
int mh_layer_rowstride(void *layer, int p) {
  int n = 0, r;
  int *rs = weed_get_int_array_counted((weed_plant_t *)layer, WEED_LEAF_ROWSTRIDES, &n);
  r = (rs && p < n) ? rs[p] : 0;
  if (rs) free(rs);
  return r;
}
void mh_layer_free(void *layer) {
  int n = 0, i;
  void **pd = weed_get_voidptr_array_counted((weed_plant_t *)layer, WEED_LEAF_PIXEL_DATA, &n);
  for (i = 0; pd && i < n; i++) free(pd[i]);
  if (pd) free(pd);
  weed_plant_free((weed_plant_t *)layer);
}

/* ---- "compositor" (gdk/compositor.c:300-340): nlayers repeats of the one in-channel template, the seven parameters with per-layer
 *      double arrays, as weed_apply_instance hands them over.  srcs[z] == NULL: a disabled channel. */
int mh_run_compositor(int h, int fidx, int palette, int width, int height, int nlayers, void **srcs, int *rss, void *dst, int rsd,
                      const double *offsx, const double *offsy, const double *scalex, const double *scaley, const double *alpha,
                      const int *bgcol, int revz) {
  weed_plant_t *filter, *inst, *in_ch[64], *out_ch, *params[7];
  weed_plant_t **ictm, **octm, **iptm;
  weed_init_f init_fn;
  weed_process_f process_fn;
  weed_deinit_f deinit_fn;
  weed_error_t err = WEED_SUCCESS;
  int n, z;
  if (mh_num_filters(h) <= fidx || fidx < 0 || nlayers > 64) return -1;
  filter = mh_tab[h].filters[fidx];
  ictm = weed_get_plantptr_array_counted(filter, WEED_LEAF_IN_CHANNEL_TEMPLATES, &n);
  if (n < 1) return -2;
  octm = weed_get_plantptr_array_counted(filter, WEED_LEAF_OUT_CHANNEL_TEMPLATES, &n);
  if (n < 1) return -3;
  iptm = weed_get_plantptr_array_counted(filter, WEED_LEAF_IN_PARAMETER_TEMPLATES, &n);
  if (n != 7) return -4;
  if (!weed_plant_has_leaf(ictm[0], WEED_LEAF_MAX_REPEATS)) return -5;
  for (z = 0; z < nlayers; z++) {
    in_ch[z] = mh_channel(ictm[0], palette, width, height, srcs[z], rss[z]);
    if (!srcs[z]) weed_set_boolean_value(in_ch[z], WEED_LEAF_DISABLED, WEED_TRUE);
  }
  out_ch = mh_channel(octm[0], palette, width, height, dst, rsd);
  for (z = 0; z < 7; z++) {
    params[z] = weed_plant_new(WEED_PLANT_PARAMETER);
    weed_set_plantptr_value(params[z], WEED_LEAF_TEMPLATE, iptm[z]);
  }
  weed_set_double_array(params[0], WEED_LEAF_VALUE, nlayers, (double *)offsx);
  weed_set_double_array(params[1], WEED_LEAF_VALUE, nlayers, (double *)offsy);
  weed_set_double_array(params[2], WEED_LEAF_VALUE, nlayers, (double *)scalex);
  weed_set_double_array(params[3], WEED_LEAF_VALUE, nlayers, (double *)scaley);
  weed_set_double_array(params[4], WEED_LEAF_VALUE, nlayers, (double *)alpha);
  weed_set_int_array(params[5], WEED_LEAF_VALUE, 3, (int *)bgcol);
  weed_set_boolean_value(params[6], WEED_LEAF_VALUE, revz ? WEED_TRUE : WEED_FALSE);
  inst = weed_plant_new(WEED_PLANT_FILTER_INSTANCE);
  weed_set_plantptr_value(inst, WEED_LEAF_FILTER_CLASS, filter);
  weed_set_plantptr_array(inst, WEED_LEAF_IN_CHANNELS, nlayers, in_ch);
  weed_set_plantptr_array(inst, WEED_LEAF_OUT_CHANNELS, 1, &out_ch);
  weed_set_plantptr_array(inst, WEED_LEAF_IN_PARAMETERS, 7, params);
  init_fn = (weed_init_f)weed_get_funcptr_value(filter, WEED_LEAF_INIT_FUNC, NULL);
  process_fn = (weed_process_f)weed_get_funcptr_value(filter, WEED_LEAF_PROCESS_FUNC, NULL);
  deinit_fn = (weed_deinit_f)weed_get_funcptr_value(filter, WEED_LEAF_DEINIT_FUNC, NULL);
  if (!process_fn) return -6;
  if (init_fn) err = (*init_fn)(inst);
  if (err == WEED_SUCCESS) err = (*process_fn)(inst, 0);
  if (deinit_fn) (*deinit_fn)(inst);
  for (z = 0; z < nlayers; z++) weed_plant_free(in_ch[z]);
  for (z = 0; z < 7; z++) weed_plant_free(params[z]);
  weed_plant_free(out_ch);
  weed_plant_free(inst);
  free(ictm); free(octm); free(iptm);
  return (int)err;
}

/* ---- generic runner: any number of in channels (planar or packed), one out channel, one in-parameter per template of the filter.
 *      vals holds the parameter values back to back, counts[k] of them for parameter k; each is stored with the seed type of the
 *      template's default (boolean / int / double; a colour is three ints).  `seed` becomes the instance's WEED_LEAF_RANDOM_SEED
 *      (what the host hands "dissolve", multi_transitions.c:55).  init once, process nframes times, deinit once. */
typedef struct {
  int palette, width, height, nplanes, yuv_clamping;
  void *planes[4];
  int rowstrides[4];
} mh_chan_desc;

static weed_plant_t *mh_channel_n(weed_plant_t *tmpl, const mh_chan_desc *d) {
  weed_plant_t *ch = weed_plant_new(WEED_PLANT_CHANNEL);
  weed_set_plantptr_value(ch, WEED_LEAF_TEMPLATE, tmpl);
  weed_set_int_value(ch, WEED_LEAF_WIDTH, d->width);
  weed_set_int_value(ch, WEED_LEAF_HEIGHT, d->height);
  weed_set_int_value(ch, WEED_LEAF_CURRENT_PALETTE, d->palette);
  weed_set_int_array(ch, WEED_LEAF_ROWSTRIDES, d->nplanes, (int *)d->rowstrides);
  weed_set_voidptr_array(ch, WEED_LEAF_PIXEL_DATA, d->nplanes, (void **)d->planes);
  if (d->yuv_clamping >= 0) weed_set_int_value(ch, WEED_LEAF_YUV_CLAMPING, d->yuv_clamping);
  return ch;
}

int mh_run_generic(int h, int fidx, int nin, const mh_chan_desc *in, const mh_chan_desc *out, int nparams, const double *vals,
                   const int *counts, long long seed, int nframes) {
  weed_plant_t *filter, *inst, *in_ch[8], *out_ch, *params[32];
  weed_plant_t **ictm, **octm, **iptm = NULL;
  weed_init_f init_fn;
  weed_process_f process_fn;
  weed_deinit_f deinit_fn;
  weed_error_t err = WEED_SUCCESS;
  int n, np = 0, k, j, pos = 0;
  if (mh_num_filters(h) <= fidx || fidx < 0 || nin < 1 || nin > 8) return -1;
  filter = mh_tab[h].filters[fidx];
  ictm = weed_get_plantptr_array_counted(filter, WEED_LEAF_IN_CHANNEL_TEMPLATES, &n);
  if (n < nin) return -2;
  octm = weed_get_plantptr_array_counted(filter, WEED_LEAF_OUT_CHANNEL_TEMPLATES, &n);
  if (n < 1) return -3;
  if (weed_plant_has_leaf(filter, WEED_LEAF_IN_PARAMETER_TEMPLATES))
    iptm = weed_get_plantptr_array_counted(filter, WEED_LEAF_IN_PARAMETER_TEMPLATES, &np);
  if (np > 32 || np != nparams) return -4;
  for (k = 0; k < nin; k++) in_ch[k] = mh_channel_n(ictm[k], &in[k]);
  out_ch = mh_channel_n(octm[0], out);
  for (k = 0; k < np; k++) {
    const uint32_t st = weed_leaf_seed_type(iptm[k], WEED_LEAF_DEFAULT);
    params[k] = weed_plant_new(WEED_PLANT_PARAMETER);
    weed_set_plantptr_value(params[k], WEED_LEAF_TEMPLATE, iptm[k]);
    if (st == WEED_SEED_DOUBLE) weed_set_double_array(params[k], WEED_LEAF_VALUE, counts[k], (double *)(vals + pos));
    else {
      int iv[8];
      for (j = 0; j < counts[k] && j < 8; j++) iv[j] = (int)vals[pos + j];
      if (st == WEED_SEED_BOOLEAN) weed_set_boolean_array(params[k], WEED_LEAF_VALUE, counts[k], iv);
      else weed_set_int_array(params[k], WEED_LEAF_VALUE, counts[k], iv);
    }
    pos += counts[k];
  }
  inst = weed_plant_new(WEED_PLANT_FILTER_INSTANCE);
  weed_set_plantptr_value(inst, WEED_LEAF_FILTER_CLASS, filter);
  weed_set_plantptr_array(inst, WEED_LEAF_IN_CHANNELS, nin, in_ch);
  weed_set_plantptr_array(inst, WEED_LEAF_OUT_CHANNELS, 1, &out_ch);
  if (np > 0) weed_set_plantptr_array(inst, WEED_LEAF_IN_PARAMETERS, np, params);
  weed_set_int64_value(inst, WEED_LEAF_RANDOM_SEED, (int64_t)seed);
  init_fn = weed_plant_has_leaf(filter, WEED_LEAF_INIT_FUNC) ? (weed_init_f)weed_get_funcptr_value(filter, WEED_LEAF_INIT_FUNC, NULL) : NULL;
  process_fn = (weed_process_f)weed_get_funcptr_value(filter, WEED_LEAF_PROCESS_FUNC, NULL);
  deinit_fn = weed_plant_has_leaf(filter, WEED_LEAF_DEINIT_FUNC) ? (weed_deinit_f)weed_get_funcptr_value(filter, WEED_LEAF_DEINIT_FUNC, NULL) : NULL;
  if (!process_fn) return -5;
  if (init_fn) err = (*init_fn)(inst);
  for (k = 0; k < nframes && err == WEED_SUCCESS; k++) err = (*process_fn)(inst, (weed_timecode_t)k);
  if (deinit_fn) (*deinit_fn)(inst);
  weed_plant_free(inst);
  for (k = 0; k < np; k++) weed_plant_free(params[k]);
  for (k = 0; k < nin; k++) weed_plant_free(in_ch[k]);
  weed_plant_free(out_ch);
  free(ictm); free(octm); if (iptm) free(iptm);
  return (int)err;
}
