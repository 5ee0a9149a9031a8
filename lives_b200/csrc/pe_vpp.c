/* pe_vpp.c -- libpe_vpp.so: a LiVES video playback plugin whose screen is a ring of frames in B200 HBM (include/pe_vpp.h).
 * Reference side: lives-plugins/plugins/playback/video/videoplugin.h (the ABI), openGL.cpp:2058-2204 (play_frame + return data),
 * src/player.c:1358-1508 (the host converts to the plugin's palette, applies the screen gamma, then calls play_frame). */
#define _GNU_SOURCE
#include "pe_vpp.h"

#include <dlfcn.h>
#include <stdio.h>
#include <string.h>

static pe_weed_leaf_get_f w_get;
static pe_weed_leaf_num_elements_f w_num;

static struct {
  pe_engine_t *e;
  int palette, width, height, inited;
  pe_frame_t *ring[PE_VPP_RING];
  int64_t tc[PE_VPP_RING];
  int newest;
  uint64_t frames, h2d, d2h;
} S = {NULL, PE_PALETTE_RGBA32, 0, 0, 0, {NULL}, {0}, -1, 0, 0, 0};

static int bind_weed(void) {
  /* libweed exports its API as function-pointer VARIABLES filled by weed_init() (weed.h:340-351) */
  void **g, **n;
  if (w_get && w_num) return 1;
  g = (void **)dlsym(RTLD_DEFAULT, "weed_leaf_get");
  n = (void **)dlsym(RTLD_DEFAULT, "weed_leaf_num_elements");
  w_get = g ? (pe_weed_leaf_get_f)*g : NULL;
  w_num = n ? (pe_weed_leaf_num_elements_f)*n : NULL;
  if (!w_get || !w_num) fprintf(stderr, "pe_vpp: libweed is not loaded in this process (weed_init() first)\n");
  return w_get && w_num;
}

static int psize_of(int pal) { return (pal == PE_PALETTE_RGB24 || pal == PE_PALETTE_BGR24) ? 3 : 4; }
static int palette_ok(int pal) { return pal == PE_PALETTE_RGBA32 || pal == PE_PALETTE_BGRA32 || pal == PE_PALETTE_RGB24 || pal == PE_PALETTE_BGR24; }

const char *module_check_init(void) {
  S.e = pe_engine_shared();
  if (!S.e) return "lives_b200 playback plugin: no usable CUDA device (there is no CPU path in this plugin)";
  return NULL;
}

const char *get_description(void) {
  return "The lives_b200 playback plugin keeps the final frame in GPU memory (a ring of device surfaces) for a CUDA presenter.\n";
}

const int *get_palette_list(void) {
  static const int pals[] = {PE_PALETTE_RGBA32, PE_PALETTE_BGRA32, PE_PALETTE_RGB24, PE_PALETTE_BGR24, PE_PALETTE_NONE /* WEED_PALETTE_END 0 */};
  return pals;
}

boolean set_palette(int palette) {
  if (!palette_ok(palette)) return 0;
  S.palette = palette;
  return 1;
}

uint64_t get_capabilities(int palette) {
  (void)palette;
  return PE_VPP_CAN_RESIZE | PE_VPP_CAN_RETURN;
}

static void drop_ring(void) {
  int k;
  for (k = 0; k < PE_VPP_RING; k++) {
    if (S.ring[k]) pe_frame_destroy(S.ring[k]);
    S.ring[k] = NULL;
  }
  S.newest = -1;
}

boolean init_screen(int width, int height, boolean fullscreen, uint64_t window_id, int argc, char **argv) {
  (void)fullscreen; (void)window_id; (void)argc; (void)argv;
  if (!S.e && module_check_init()) return 0;
  if (width <= 0 || height <= 0) return 0;
  drop_ring();
  S.width = width; S.height = height;
  S.frames = S.h2d = S.d2h = 0;
  S.inited = 1;
  return 1;
}

void exit_screen(int16_t mouse_x, int16_t mouse_y) {
  (void)mouse_x; (void)mouse_y;
  if (S.e) pe_engine_sync(S.e);
  drop_ring();
  S.inited = 0;
}

void module_unload(void) { exit_screen(0, 0); }

/* f becomes the newest surface (the ring owns it from here on) */
static boolean present(pe_frame_t *f, int64_t tc) {
  const int slot = (S.newest + 1) % PE_VPP_RING;
  pe_frame_desc_t d;
  if (pe_frame_get_desc(f, &d) != PE_OK) { pe_frame_destroy(f); return 0; }
  if (d.width != S.width || d.height != S.height) { /* VPP_CAN_RESIZE: the screen size, bilinear as the player asks elsewhere */
    if (!pe_resize_layer(S.e, f, S.width, S.height, 1 /* LIVES_INTERP_NORMAL */, d.palette, 0)) {
      fprintf(stderr, "pe_vpp: %s\n", pe_last_error());
      pe_frame_destroy(f);
      return 0;
    }
  }
  if (S.ring[slot]) pe_frame_destroy(S.ring[slot]); /* stream ordered: the block goes back to the engine's pool */
  S.ring[slot] = f;
  S.tc[slot] = tc;
  S.newest = slot;
  S.frames++;
  return 1;
}

static boolean play_host(int width, int height, const void *pixels, int rowstride, int64_t tc, void *ret_pixels, int ret_rowstride) {
  pe_frame_t *f = NULL;
  const void *planes[PE_MAXPLANES] = {pixels, NULL, NULL, NULL};
  void *rplanes[PE_MAXPLANES] = {ret_pixels, NULL, NULL, NULL};
  int rs[PE_MAXPLANES] = {rowstride, 0, 0, 0}, rrs[PE_MAXPLANES] = {ret_rowstride, 0, 0, 0};
  if (!S.inited || !pixels || width <= 0 || height <= 0) return 0;
  if (pe_frame_create(S.e, S.palette, width, height, 0, 0, 0, 0, 0, &f) != PE_OK || pe_frame_upload(S.e, f, planes, rs) != PE_OK) {
    fprintf(stderr, "pe_vpp: %s\n", pe_last_error());
    if (f) pe_frame_destroy(f);
    return 0;
  }
  S.h2d += (uint64_t)width * psize_of(S.palette) * (uint64_t)height;
  if (ret_pixels) { /* VPP_CAN_RETURN: the unresized data, from the device copy that is about to be shown */
    if (pe_frame_download(S.e, f, rplanes, rrs) != PE_OK) { pe_frame_destroy(f); return 0; }
    S.d2h += (uint64_t)width * psize_of(S.palette) * (uint64_t)height;
  }
  if (!ret_pixels && pe_engine_sync(S.e) != PE_OK) { pe_frame_destroy(f); return 0; } /* the host may free the layer's pixels once we return */
  return present(f, tc);
}

boolean play_frame(weed_layer_t *frame, int64_t tc, weed_layer_t *ret) {
  int32_t pal = 0, w = 0, h = 0, rs = 0, rrs = 0;
  void *px = NULL, *rpx = NULL;
  if (!frame || !bind_weed()) return 0;
  w_get(frame, PE_LEAF_CURRENT_PALETTE, 0, &pal);
  w_get(frame, PE_LEAF_WIDTH, 0, &w);
  w_get(frame, PE_LEAF_HEIGHT, 0, &h);
  if (w_num(frame, PE_LEAF_PIXEL_DATA) < 1 || w_num(frame, PE_LEAF_ROWSTRIDES) < 1) return 0;
  w_get(frame, PE_LEAF_PIXEL_DATA, 0, &px);
  w_get(frame, PE_LEAF_ROWSTRIDES, 0, &rs);
  if (pal != S.palette) { /* the host converts to the plugin's palette before the call (src/player.c:1359-1369) */
    fprintf(stderr, "pe_vpp: frame palette %d is not the palette set by set_palette (%d)\n", pal, S.palette);
    return 0;
  }
  if (ret && w_num(ret, PE_LEAF_PIXEL_DATA) > 0 && w_num(ret, PE_LEAF_ROWSTRIDES) > 0) { /* the host created space for it (openGL.cpp:2074) */
    w_get(ret, PE_LEAF_PIXEL_DATA, 0, &rpx);
    w_get(ret, PE_LEAF_ROWSTRIDES, 0, &rrs);
  }
  return play_host(w, h, px, rs, tc, rpx, rrs);
}

boolean render_frame(int hsize, int vsize, int64_t timecode, void **pixel_data, void **return_data, void **play_params) {
  const int rs = hsize * psize_of(S.palette); /* "no extra padding (rowstrides) is allowed", videoplugin.h:134 */
  (void)play_params;
  if (!pixel_data) return 0;
  return play_host(hsize, vsize, pixel_data[0], rs, timecode, return_data ? return_data[0] : NULL, rs);
}

boolean pe_vpp_play_device_frame(const pe_frame_t *frame, int64_t tc) {
  pe_frame_t *f = NULL;
  pe_frame_desc_t d;
  if (!S.inited || !frame || pe_frame_get_desc(frame, &d) != PE_OK) return 0;
  if (d.palette != S.palette) {
    fprintf(stderr, "pe_vpp: frame palette %d is not the palette set by set_palette (%d)\n", d.palette, S.palette);
    return 0;
  }
  if (pe_frame_copy(S.e, frame, &f) != PE_OK) { fprintf(stderr, "pe_vpp: %s\n", pe_last_error()); return 0; }
  return present(f, tc);
}

int pe_vpp_acquire(pe_vpp_surface_t *out) {
  if (!out || !S.inited || S.newest < 0) return PE_ERR_ARG;
  if (pe_engine_sync(S.e) != PE_OK) return PE_ERR_CUDA;
  if (pe_frame_get_desc(S.ring[S.newest], &out->desc) != PE_OK) return PE_ERR_ARG;
  out->timecode = S.tc[S.newest];
  out->serial = S.frames;
  return PE_OK;
}

int pe_vpp_read_surface(void *host, int rowstride) {
  void *planes[PE_MAXPLANES] = {host, NULL, NULL, NULL};
  int rs[PE_MAXPLANES] = {rowstride, 0, 0, 0};
  if (!host || !S.inited || S.newest < 0) return PE_ERR_ARG;
  return pe_frame_download(S.e, S.ring[S.newest], planes, rs);
}

void pe_vpp_counters(uint64_t *frames, uint64_t *h2d_bytes, uint64_t *d2h_bytes) {
  if (frames) *frames = S.frames;
  if (h2d_bytes) *h2d_bytes = S.h2d;
  if (d2h_bytes) *d2h_bytes = S.d2h;
}
