/* pe_weed_layer.h -- boundary B2 for real: the frame ops of the LiVES hot path with the EXACT signatures of
 * src/colourspace.h:387-415, over weed_layer_t * (= weed_plant_t *, src/main.h:95), implemented by lives_b200/libpe_weed_layer.so.
 *
 * A layer is read and rewritten through the host's own libweed (weed_leaf_get / weed_leaf_set / weed_leaf_num_elements /
 * weed_leaf_delete: the function-pointer variables libweed exports, libweed/weed.h:340-351), leaf by leaf as src/layers.c:292-510
 * does: "pixel_data" (voidptr array), "rowstrides" (int array), "width" (MACROpixels), "height", "current_palette", "YUV_clamping",
 * "YUV_sampling", "YUV_subspace", "gamma_type", "flags" (libweed/weed-effects.h:269-277,350-375).  Pixel buffers are host memory the
 * caller owns; an op that changes the byte size of the planes releases the old ones and installs new ones through the allocator
 * given to pe_weed_layer_set_allocator (default malloc / free), where the reference calls weed_layer_pixel_data_free +
 * create_empty_pixel_data (src/colourspace.c:13859-13900).  Every pixel is computed on the GPU by libpe_b200.so (H2D -> kernels ->
 * D2H around the device ops of pixel_engine.h); there is no CPU fallback: without a usable CUDA device every op returns FALSE and
 * leaves the layer untouched, exactly as the reference does on failure (:13906-13927).
 *
 * Integration: LiVES links this library next to libweed, calls pe_weed_layer_bind(NULL) once after weed_init(),
 * pe_engine_set_prefs(pe_engine_shared(), prefs->pb_quality, ...) when the prefs change, and compiles colourspace.c with the bodies
 * of these eight functions left out.
 */
#ifndef PE_WEED_LAYER_H
#define PE_WEED_LAYER_H

#include "pe_weed_abi.h"
#include "pixel_engine.h"

#ifdef __cplusplus
extern "C" {
#endif

#ifndef PE_WEED_LAYER_NO_TYPEDEFS /* define it when the reference's own headers are in scope */
typedef pe_weed_plant_t weed_layer_t; /* src/main.h:95 */
typedef int boolean;                   /* src/main.h: TRUE / FALSE */
typedef int LiVESInterpType;           /* LIVES_INTERP_FAST 0, NORMAL 1, BEST 2 */
#endif

#define PE_WEED_PLANT_LAYER 128 /* src/layers.h:14 */

/* ---- the drop-ins: src/colourspace.h:387-415 ------------------------------------------------------------------------------- */
void alpha_premult(weed_layer_t *, int direction);                                                            /* :387, .c:11968 */
boolean gamma_convert_layer(int gamma_type, weed_layer_t *);                                                  /* :389, .c:14146 */
boolean gamma_convert_sub_layer(int gamma_type, double fileg, weed_layer_t *, int x, int y, int width, int height,
                                boolean may_thread);                                                          /* :391, .c:14069 */
boolean convert_layer_palette(weed_layer_t *, int outpl, int op_clamping);                                    /* :393, .c:13931 */
boolean convert_layer_palette_full(weed_layer_t *, int outpl, int oclamping, int osampling, int osubspace,
                                   int tgt_gamma);                                                            /* :395, .c:12190 */
boolean resize_layer_full(weed_layer_t *, int width, int height, LiVESInterpType interp, int opal_hint, int oclamp_hint,
                          int osamp_hint, int osubs_hint, int tgt_gamma);                                     /* :409, .c:14759 */
boolean resize_layer(weed_layer_t *, int width, int height, LiVESInterpType interp, int opal_hint, int oclamp_hint); /* :413, .c:15331 */
boolean letterbox_layer(weed_layer_t *, int nwidth, int nheight, int width, int height, LiVESInterpType interp, int tpal,
                        int tclamp);                                                                          /* :415, .c:15343 */

/* ---- binding to the host ---------------------------------------------------------------------------------------------------- */

typedef pe_weed_error_t (*pe_weed_leaf_delete_f)(pe_weed_plant_t *, const char *key); /* libweed/weed.h:236 */
typedef struct pe_weed_host_funcs {
  pe_weed_leaf_get_f leaf_get;
  pe_weed_leaf_set_f leaf_set;
  pe_weed_leaf_num_elements_f leaf_num_elements;
  pe_weed_leaf_delete_f leaf_delete;
} pe_weed_host_funcs_t;

/* funcs == NULL: resolve the variables weed_leaf_get, weed_leaf_set, weed_leaf_num_elements, weed_leaf_delete of the libweed already
 * loaded in this process (dlsym RTLD_DEFAULT), after the host's weed_init().  Returns PE_OK, or PE_ERR_ARG when one is missing. */
int pe_weed_layer_bind(const pe_weed_host_funcs_t *funcs);
/* where new pixel buffers come from / old ones go (LiVES: lives_calloc_safety / lives_free_maybe_big); NULL = malloc / free */
void pe_weed_layer_set_allocator(const pe_host_allocator_t *alloc);
/* 1: page-lock the planes around each transfer (pe_host_register ... pe_host_unregister inside the call: the buffers are the host's,
 * it frees them without telling this library, so no registration outlives a call); 0 (default): copy from / to pageable memory as it
 * is.  A host that owns its pixel allocator does better: register its big blocks once (pe_host_register) and leave this off. */
void pe_weed_layer_set_pinning(int on);
/* the engine the drop-ins run on (pe_engine_shared()); NULL + pe_last_error when no CUDA device is usable */
pe_engine_t *pe_weed_layer_engine(void);

#ifdef __cplusplus
}
#endif
#endif /* PE_WEED_LAYER_H */
