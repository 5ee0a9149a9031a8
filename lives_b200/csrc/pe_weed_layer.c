/* pe_weed_layer.c -- libpe_weed_layer.so: the reference's frame ops (src/colourspace.h:387-415) over weed_layer_t *, computed on the
 * B200 by libpe_b200.so.  See include/pe_weed_layer.h for the contract.  Host side of boundary B2: leaf traffic in, device op,
 * leaf traffic out; the layer is untouched until the device op has succeeded. */
#define _GNU_SOURCE
#include "pe_weed_layer.h"

#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define PE_LEAF_YUV_CLAMPING "YUV_clamping" /* libweed/weed-effects.h:272-274 */
#define PE_LEAF_YUV_SAMPLING "YUV_sampling"
#define PE_LEAF_YUV_SUBSPACE "YUV_subspace"
#define PE_LEAF_GAMMA_TYPE "gamma_type"      /* :375 */

static pe_weed_host_funcs_t H;
static int bound;
static pe_host_allocator_t A;
static int pinning;

static void *def_alloc(size_t n, void *u) { void *p = NULL; (void)u; return posix_memalign(&p, 64, n ? n : 64) == 0 ? p : NULL; }
static void def_free(void *p, void *u) { (void)u; free(p); }

int pe_weed_layer_bind(const pe_weed_host_funcs_t *funcs) {
  if (funcs) {
    H = *funcs;
  } else {
    /* libweed exports its API as function-pointer VARIABLES filled by weed_init() (weed.h:340-351) */
    void **g = (void **)dlsym(RTLD_DEFAULT, "weed_leaf_get"), **s = (void **)dlsym(RTLD_DEFAULT, "weed_leaf_set"),
         **n = (void **)dlsym(RTLD_DEFAULT, "weed_leaf_num_elements"), **d = (void **)dlsym(RTLD_DEFAULT, "weed_leaf_delete");
    H.leaf_get = g ? (pe_weed_leaf_get_f)*g : NULL;
    H.leaf_set = s ? (pe_weed_leaf_set_f)*s : NULL;
    H.leaf_num_elements = n ? (pe_weed_leaf_num_elements_f)*n : NULL;
    H.leaf_delete = d ? (pe_weed_leaf_delete_f)*d : NULL;
  }
  bound = H.leaf_get && H.leaf_set && H.leaf_num_elements && H.leaf_delete;
  return bound ? PE_OK : PE_ERR_ARG;
}

void pe_weed_layer_set_allocator(const pe_host_allocator_t *alloc) {
  if (alloc && alloc->alloc && alloc->free) A = *alloc;
  else { A.alloc = def_alloc; A.free = def_free; A.user = NULL; }
}
void pe_weed_layer_set_pinning(int on) { pinning = on != 0; }
pe_engine_t *pe_weed_layer_engine(void) { return pe_engine_shared(); }

static int ready(void) {
  if (!bound && pe_weed_layer_bind(NULL) != PE_OK) {
    fprintf(stderr, "pe_weed_layer: libweed is not bound (weed_init() first, or pe_weed_layer_bind with the host's functions)\n");
    return 0;
  }
  if (!A.alloc) pe_weed_layer_set_allocator(NULL);
  return 1;
}

/* ---- leaves (src/layers.c:292-510) ------------------------------------------------------------------------------------------ */

static int has_leaf(weed_layer_t *l, const char *key) { return H.leaf_num_elements(l, key) > 0; }
static int get_int(weed_layer_t *l, const char *key, int dflt) {
  int32_t v = dflt;
  if (has_leaf(l, key)) H.leaf_get(l, key, 0, &v);
  return v;
}
static void set_int(weed_layer_t *l, const char *key, int32_t v) { H.leaf_set(l, key, PE_WEED_SEED_INT, 1, &v); }

typedef struct {
  pe_frame_desc_t d;      /* width in PIXELS */
  int plane_bytes_known;  /* rowstrides and planes present */
} layer_view;

static int ppmp(int pal) { return (pal == PE_PALETTE_UYVY || pal == PE_PALETTE_YUYV) ? 2 : pal == PE_PALETTE_YUV411 ? 4 : 1; }

static int read_layer(weed_layer_t *l, layer_view *v) {
  int np, nrs, i;
  memset(v, 0, sizeof(*v));
  if (!l || get_int(l, PE_LEAF_TYPE, 0) != PE_WEED_PLANT_LAYER) return 0; /* WEED_IS_LAYER */
  v->d.palette = get_int(l, PE_LEAF_CURRENT_PALETTE, PE_PALETTE_NONE);
  v->d.width = get_int(l, PE_LEAF_WIDTH, 0) * ppmp(v->d.palette);
  v->d.height = get_int(l, PE_LEAF_HEIGHT, 0);
  v->d.yuv_clamping = get_int(l, PE_LEAF_YUV_CLAMPING, PE_YUV_CLAMPING_CLAMPED);
  v->d.yuv_sampling = get_int(l, PE_LEAF_YUV_SAMPLING, PE_YUV_SAMPLING_DEFAULT);
  v->d.yuv_subspace = get_int(l, PE_LEAF_YUV_SUBSPACE, PE_YUV_SUBSPACE_YUV);
  v->d.gamma_type = get_int(l, PE_LEAF_GAMMA_TYPE, PE_GAMMA_UNKNOWN);
  v->d.flags = get_int(l, PE_LEAF_FLAGS, 0);
  np = (int)H.leaf_num_elements(l, PE_LEAF_PIXEL_DATA);
  nrs = (int)H.leaf_num_elements(l, PE_LEAF_ROWSTRIDES);
  if (np > PE_MAXPLANES) np = PE_MAXPLANES;
  for (i = 0; i < np; i++) H.leaf_get(l, PE_LEAF_PIXEL_DATA, (pe_weed_size_t)i, &v->d.planes[i]);
  for (i = 0; i < nrs && i < PE_MAXPLANES; i++) { int32_t rs = 0; H.leaf_get(l, PE_LEAF_ROWSTRIDES, (pe_weed_size_t)i, &rs); v->d.rowstrides[i] = rs; }
  v->d.nplanes = np;
  v->plane_bytes_known = np > 0 && nrs >= np && v->d.planes[0] != NULL;
  return 1;
}

/* upload the layer into a device frame; NULL when it has no pixel data / no device */
static pe_frame_t *to_device(pe_engine_t *e, const layer_view *v) {
  pe_frame_t *f = NULL;
  int np = 0, rs[PE_MAXPLANES], ph[PE_MAXPLANES], p;
  if (!v->plane_bytes_known) return NULL;
  if (!pe_frame_layout(v->d.palette, v->d.width, v->d.height, &np, rs, ph) || np != v->d.nplanes) return NULL;
  if (pe_frame_create(e, v->d.palette, v->d.width, v->d.height, v->d.yuv_clamping, v->d.yuv_sampling, v->d.yuv_subspace, v->d.gamma_type, 0,
                      &f) != PE_OK)
    return NULL;
  pe_frame_set_flags(f, v->d.flags);
  /* page-locking is TRANSIENT: the buffers belong to the host, which frees them without telling us -- a registration left behind
   * would poison whatever malloc() later places at the same address */
  if (pinning)
    for (p = 0; p < np; p++) pe_host_register(v->d.planes[p], (size_t)v->d.rowstrides[p] * (size_t)ph[p]);
  p = pe_frame_upload(e, f, (const void *const *)v->d.planes, v->d.rowstrides);
  if (pinning) {
    int q;
    pe_engine_sync(e);
    for (q = 0; q < np; q++) pe_host_unregister(v->d.planes[q]);
  }
  if (p != PE_OK) { pe_frame_destroy(f); return NULL; }
  return f;
}

/* bring the device frame back into the layer: same plane sizes -> into the caller's buffers; otherwise new buffers from the
 * allocator, the old ones released (weed_layer_pixel_data_free).  Leaves are only written once the pixels are in host memory. */
static int from_device(pe_engine_t *e, pe_frame_t *f, weed_layer_t *l, const layer_view *old) {
  pe_frame_desc_t nd;
  int np = 0, rs[PE_MAXPLANES], ph[PE_MAXPLANES], oph[PE_MAXPLANES], ors[PE_MAXPLANES], onp = 0, p, same;
  void *planes[PE_MAXPLANES] = {0};
  int32_t rs32[PE_MAXPLANES];
  if (pe_frame_get_desc(f, &nd) != PE_OK) return 0;
  if (!pe_frame_layout(nd.palette, nd.width, nd.height, &np, rs, ph)) return 0;
  pe_frame_layout(old->d.palette, old->d.width, old->d.height, &onp, ors, oph);
  /* the caller's buffers are kept (with the caller's rowstrides) when every plane still fits row for row: in-place ops, and the
   * conversions the reference also does in place (pconv_can_inplace, colourspace.c:12148) */
  same = np == old->d.nplanes && np == onp;
  for (p = 0; p < np && same; p++) same = ph[p] == oph[p] && old->d.rowstrides[p] >= rs[p];
  if (same && nd.palette == old->d.palette && nd.width == old->d.width) same = 1;
  if (same) {
    for (p = 0; p < np; p++) { planes[p] = old->d.planes[p]; rs[p] = old->d.rowstrides[p]; }
  } else {
    for (p = 0; p < np; p++) {
      planes[p] = A.alloc((size_t)rs[p] * (size_t)ph[p] + 64, A.user); /* + slack: the reference's converters read a byte past a chroma row */
      if (!planes[p]) { while (p--) A.free(planes[p], A.user); return 0; }
    }
  }
  if (pinning)
    for (p = 0; p < np; p++) pe_host_register(planes[p], (size_t)rs[p] * (size_t)ph[p]);
  p = pe_frame_download(e, f, planes, rs); /* synchronises */
  if (pinning) {
    int q;
    for (q = 0; q < np; q++) pe_host_unregister(planes[q]);
  }
  if (p != PE_OK) {
    if (!same) for (p = 0; p < np; p++) A.free(planes[p], A.user);
    return 0;
  }
  /* ---- the layer changes from here on */
  if (!same) {
    for (p = 0; p < old->d.nplanes; p++)
      if (old->d.planes[p]) A.free(old->d.planes[p], A.user);
    H.leaf_set(l, PE_LEAF_PIXEL_DATA, PE_WEED_SEED_VOIDPTR, (pe_weed_size_t)np, planes);
  }
  for (p = 0; p < np; p++) rs32[p] = rs[p];
  H.leaf_set(l, PE_LEAF_ROWSTRIDES, PE_WEED_SEED_INT, (pe_weed_size_t)np, rs32);
  set_int(l, PE_LEAF_CURRENT_PALETTE, nd.palette);
  set_int(l, PE_LEAF_WIDTH, nd.width / ppmp(nd.palette));
  set_int(l, PE_LEAF_HEIGHT, nd.height);
  if (nd.palette >= 512) { /* YUV */
    set_int(l, PE_LEAF_YUV_CLAMPING, nd.yuv_clamping);
    set_int(l, PE_LEAF_YUV_SAMPLING, nd.yuv_sampling);
    set_int(l, PE_LEAF_YUV_SUBSPACE, nd.yuv_subspace);
  } else { /* conv_done deletes them for RGB (:13878-13881) */
    if (has_leaf(l, PE_LEAF_YUV_CLAMPING)) H.leaf_delete(l, PE_LEAF_YUV_CLAMPING);
    if (has_leaf(l, PE_LEAF_YUV_SAMPLING)) H.leaf_delete(l, PE_LEAF_YUV_SAMPLING);
    if (has_leaf(l, PE_LEAF_YUV_SUBSPACE)) H.leaf_delete(l, PE_LEAF_YUV_SUBSPACE);
  }
  if (nd.gamma_type != old->d.gamma_type || has_leaf(l, PE_LEAF_GAMMA_TYPE)) set_int(l, PE_LEAF_GAMMA_TYPE, nd.gamma_type);
  if (nd.flags != old->d.flags || has_leaf(l, PE_LEAF_FLAGS)) set_int(l, PE_LEAF_FLAGS, nd.flags);
  return 1;
}

/* ---- one driver for every op -------------------------------------------------------------------------------------------------- */

typedef struct { int op, i[8]; double d; } op_args;
enum { OP_CONVERT, OP_RESIZE, OP_LETTERBOX, OP_GAMMA, OP_GAMMA_SUB, OP_PREMULT };

static int run_device_op(pe_engine_t *e, pe_frame_t *f, const op_args *a) {
  switch (a->op) {
  case OP_CONVERT: return pe_convert_layer_palette_full(e, f, a->i[0], a->i[1], a->i[2], a->i[3], a->i[4]);
  case OP_RESIZE: return pe_resize_layer_full(e, f, a->i[0], a->i[1], a->i[2], a->i[3], a->i[4], a->i[5], a->i[6], a->i[7]);
  case OP_LETTERBOX: return pe_letterbox_layer(e, f, a->i[0], a->i[1], a->i[2], a->i[3], a->i[4], a->i[5], a->i[6]);
  case OP_GAMMA: return pe_gamma_convert_layer(e, a->i[0], f);
  case OP_GAMMA_SUB: return pe_gamma_convert_sub_layer(e, a->i[0], a->d, f, a->i[1], a->i[2], a->i[3], a->i[4], a->i[5]);
  case OP_PREMULT: pe_alpha_premult(e, f, a->i[0]); return PE_TRUE;
  }
  return PE_FALSE;
}

static boolean layer_op(weed_layer_t *layer, const op_args *a) {
  layer_view v;
  pe_engine_t *e;
  pe_frame_t *f;
  int ok;
  if (!ready() || !read_layer(layer, &v)) return PE_FALSE;
  if (!v.plane_bytes_known) {
    /* a layer without pixel data: resize_layer_full still records the target size / palette (:14820-14832) and returns FALSE */
    if (a->op == OP_RESIZE) {
      set_int(layer, PE_LEAF_HEIGHT, a->i[1]);
      if (a->i[3] != PE_PALETTE_NONE) set_int(layer, PE_LEAF_CURRENT_PALETTE, a->i[3]);
      set_int(layer, PE_LEAF_WIDTH, a->i[0] / ppmp(a->i[3] != PE_PALETTE_NONE ? a->i[3] : v.d.palette));
      set_int(layer, PE_LEAF_YUV_CLAMPING, a->i[4]);
    }
    return PE_FALSE;
  }
  e = pe_engine_shared();
  if (!e) { fprintf(stderr, "pe_weed_layer: %s\n", pe_last_error()); return PE_FALSE; }
  f = to_device(e, &v);
  if (!f) { fprintf(stderr, "pe_weed_layer: %s\n", pe_last_error()); return PE_FALSE; }
  ok = run_device_op(e, f, a) == PE_TRUE && from_device(e, f, layer, &v);
  pe_frame_destroy(f);
  return ok ? PE_TRUE : PE_FALSE;
}

/* ---- src/colourspace.h:387-415 --------------------------------------------------------------------------------------------------- */

boolean convert_layer_palette_full(weed_layer_t *layer, int outpl, int oclamping, int osampling, int osubspace, int tgt_gamma) {
  op_args a = {OP_CONVERT, {outpl, oclamping, osampling, osubspace, tgt_gamma, 0, 0, 0}, 0.};
  return layer_op(layer, &a);
}

boolean convert_layer_palette(weed_layer_t *layer, int outpl, int op_clamping) { /* :13931 */
  return convert_layer_palette_full(layer, outpl, op_clamping, PE_YUV_SAMPLING_DEFAULT, PE_YUV_SUBSPACE_YUV, PE_GAMMA_UNKNOWN);
}

boolean resize_layer_full(weed_layer_t *layer, int width, int height, LiVESInterpType interp, int opal_hint, int oclamp_hint,
                          int osamp_hint, int osubs_hint, int tgt_gamma) {
  op_args a = {OP_RESIZE, {width, height, interp, opal_hint, oclamp_hint, osamp_hint, osubs_hint, tgt_gamma}, 0.};
  return layer_op(layer, &a);
}

boolean resize_layer(weed_layer_t *layer, int width, int height, LiVESInterpType interp, int opal_hint, int oclamp_hint) { /* :15331 */
  return resize_layer_full(layer, width, height, interp, opal_hint, oclamp_hint, PE_YUV_SAMPLING_DEFAULT, PE_YUV_SUBSPACE_YCBCR,
                           PE_GAMMA_UNKNOWN);
}

boolean letterbox_layer(weed_layer_t *layer, int nwidth, int nheight, int width, int height, LiVESInterpType interp, int tpal,
                        int tclamp) {
  op_args a = {OP_LETTERBOX, {nwidth, nheight, width, height, interp, tpal, tclamp, 0}, 0.};
  return layer_op(layer, &a);
}

boolean gamma_convert_layer(int gamma_type, weed_layer_t *layer) {
  op_args a = {OP_GAMMA, {gamma_type, 0, 0, 0, 0, 0, 0, 0}, 0.};
  return layer_op(layer, &a);
}

boolean gamma_convert_sub_layer(int gamma_type, double fileg, weed_layer_t *layer, int x, int y, int width, int height,
                                boolean may_thread) {
  op_args a = {OP_GAMMA_SUB, {gamma_type, x, y, width, height, may_thread, 0, 0}, fileg};
  return layer_op(layer, &a);
}

void alpha_premult(weed_layer_t *layer, int direction) {
  op_args a = {OP_PREMULT, {direction, 0, 0, 0, 0, 0, 0, 0}, 0.};
  layer_op(layer, &a);
}
