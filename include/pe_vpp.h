/* pe_vpp.h -- SURVEY 8f rank 4, the display hand-off: lives_b200/libpe_vpp.so is a LiVES VIDEO PLAYBACK PLUGIN
 * (lives-plugins/plugins/playback/video/videoplugin.h; the host's view of it: _vid_playback_plugin, src/plugins.h:153-215; loaded
 * with dlopen + dlsym of the names below, src/plugins.c) whose "screen" is a ring of frames in B200 HBM: play_frame() leaves the
 * final frame on the device, where a presenter maps it (CUDA - GL interop in the role of openGL.cpp's texture upload, or an encoder)
 * instead of the host walking it once more.  There is no window system on the GPU boxes: this library stops at the device surface
 * (pe_vpp_acquire) and proves the path by reading it back (pe_vpp_read_surface, the VPP_CAN_RETURN data of play_frame).
 *
 * The plugin reads the layer's leaves through the libweed already loaded in the process (the function-pointer variables
 * weed_leaf_get / weed_leaf_num_elements, libweed/weed.h:340-351), like libpe_weed_layer.so.  No CPU fallback: module_check_init()
 * returns an error string when no CUDA device is usable and the host then refuses the plugin (src/plugins.c). */
#ifndef PE_VPP_H
#define PE_VPP_H

#include <stdint.h>

#include "pe_weed_abi.h"
#include "pixel_engine.h"

#ifdef __cplusplus
extern "C" {
#endif

#ifndef PE_WEED_LAYER_NO_TYPEDEFS
#ifndef PE_WEED_LAYER_H
typedef pe_weed_plant_t weed_layer_t; /* src/main.h:95 */
typedef int boolean;
#endif
#endif

/* capabilities, videoplugin.h:115-124 */
#define PE_VPP_CAN_RESIZE (1 << 0)
#define PE_VPP_CAN_RETURN (1 << 1)

/* ---- the playback plugin ABI (videoplugin.h:62-150) ------------------------------------------------------------------------ */
const char *module_check_init(void);                 /* NULL = usable */
const char *get_description(void);
const int *get_palette_list(void);                   /* RGBA32, BGRA32, RGB24, BGR24, WEED_PALETTE_END */
boolean set_palette(int palette);
uint64_t get_capabilities(int palette);              /* VPP_CAN_RESIZE | VPP_CAN_RETURN */
boolean init_screen(int width, int height, boolean fullscreen, uint64_t window_id, int argc, char **argv);
/* display one frame: the layer's pixels go up once (H2D), are resized on the device when they are not screen sized (VPP_CAN_RESIZE;
 * resize_layer's own kernels), and become the newest surface of the ring.  ret (optional): receives the unresized pixels back
 * (VPP_CAN_RETURN, videoplugin.h:131-134). */
boolean play_frame(weed_layer_t *frame, int64_t tc, weed_layer_t *ret);
/* the older entry point (videoplugin.h:136-137): packed pixel_data without row padding */
boolean render_frame(int hsize, int vsize, int64_t timecode, void **pixel_data, void **return_data, void **play_params);
void exit_screen(int16_t mouse_x, int16_t mouse_y);
void module_unload(void);

/* ---- the device side of the hand-off ----------------------------------------------------------------------------------------- */
/* a frame that is ALREADY on the device (the output of pe_fused_* / pe_fx_* / pe_run_convert_plan_over) becomes the newest surface
 * without crossing PCIe: one device-to-device copy (the caller keeps its frame) */
boolean pe_vpp_play_device_frame(const pe_frame_t *frame, int64_t tc);
typedef struct pe_vpp_surface {
  pe_frame_desc_t desc;   /* planes[] are DEVICE pointers; valid until PE_VPP_RING - 1 newer frames have been played */
  int64_t timecode;
  uint64_t serial;        /* frames played since init_screen */
} pe_vpp_surface_t;
#define PE_VPP_RING 3
/* the newest surface, complete (the engine stream is synchronised); PE_ERR_ARG before the first frame */
int pe_vpp_acquire(pe_vpp_surface_t *out);
/* the newest surface copied to host memory (tests / screenshots) */
int pe_vpp_read_surface(void *host, int rowstride);
/* bytes that crossed PCIe since init_screen */
void pe_vpp_counters(uint64_t *frames, uint64_t *h2d_bytes, uint64_t *d2h_bytes);

#ifdef __cplusplus
}
#endif
#endif /* PE_VPP_H */
