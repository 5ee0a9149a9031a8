/* pe_weed_abi.h -- the slice of the libweed effect-plugin ABI (ABI 200..203, filter API 200..202) that a plugin needs in
 * order to be loaded by LiVES: dlopen + dlsym("weed_setup") + setup(weed_bootstrap)  (src/effects-weed.c:4468-4568).
 *
 * This is OUR restatement of the binary interface -- type widths, enum values and leaf-name strings -- so that
 * libpe_weed_plugin.so builds without the reference tree.  Every value cites where the reference defines it; values must
 * stay identical to those for the plugin to interoperate (tests/test_weed_plugin.py loads the plugin through the reference's
 * own weed_bootstrap to prove they are).
 */
#ifndef PE_WEED_ABI_H
#define PE_WEED_ABI_H
#include <stddef.h>
#include <stdint.h>

typedef struct pe_weed_leaf pe_weed_plant_t;            /* opaque: weed_plant_t == weed_leaf_t, libweed/weed.h:205 */
typedef uint32_t pe_weed_size_t;                          /* weed.h:127 */
typedef int32_t pe_weed_error_t;                          /* weed.h:128 */
typedef uint32_t pe_weed_seed_t;                          /* weed.h:131 */
typedef int64_t pe_weed_timecode_t;                       /* weed-effects.h:164 */
typedef void (*pe_weed_funcptr_t)(void);                  /* weed.h:130 */

/* host functions handed over in the HOST_INFO plant (weed.h:215-236) */
typedef void *(*pe_weed_malloc_f)(size_t);
typedef void (*pe_weed_free_f)(void *);
typedef pe_weed_plant_t *(*pe_weed_plant_new_f)(int32_t plant_type);
typedef pe_weed_error_t (*pe_weed_leaf_set_f)(pe_weed_plant_t *, const char *key, pe_weed_seed_t seed_type, pe_weed_size_t num_elems,
                                              void *values);
typedef pe_weed_error_t (*pe_weed_leaf_get_f)(pe_weed_plant_t *, const char *key, pe_weed_size_t idx, void *value);
typedef pe_weed_size_t (*pe_weed_leaf_num_elements_f)(pe_weed_plant_t *, const char *key);

/* bootstrap (weed-effects.h:171-187) */
typedef pe_weed_error_t (*pe_weed_default_getter_f)(pe_weed_plant_t *plant, const char *key, void *value);
typedef pe_weed_plant_t *(*pe_weed_bootstrap_f)(pe_weed_default_getter_f *, int32_t plugin_weed_min_api_version,
                                                int32_t plugin_weed_max_api_version, int32_t plugin_filter_min_api_version,
                                                int32_t plugin_filter_max_api_version);
typedef pe_weed_error_t (*pe_weed_process_f)(pe_weed_plant_t *filter_instance, pe_weed_timecode_t timestamp);
typedef pe_weed_error_t (*pe_weed_init_f)(pe_weed_plant_t *filter_instance);
typedef pe_weed_error_t (*pe_weed_deinit_f)(pe_weed_plant_t *filter_instance);

/* versions we speak (weed.h:65, weed-effects.h:44; the reference plugins ask for 200..200, WEED_SETUP_START(200, 200)) */
#define PE_WEED_API_MIN 200
#define PE_WEED_API_MAX 203
#define PE_WEED_FILTER_API_MIN 200
#define PE_WEED_FILTER_API_MAX 202

/* errors (weed.h:373-381, weed-effects.h:151-155) */
#define PE_WEED_SUCCESS 0
#define PE_WEED_ERROR_MEMORY_ALLOCATION 1
#define PE_WEED_ERROR_PLUGIN_INVALID 64
#define PE_WEED_ERROR_FILTER_INVALID 65

/* seed types (weed.h:386-454) */
#define PE_WEED_SEED_INT 1
#define PE_WEED_SEED_DOUBLE 2
#define PE_WEED_SEED_BOOLEAN 3
#define PE_WEED_SEED_STRING 4
#define PE_WEED_SEED_INT64 5
#define PE_WEED_SEED_FUNCPTR 64
#define PE_WEED_SEED_VOIDPTR 65
#define PE_WEED_SEED_PLANTPTR 66

/* plant types (weed-effects.h:61-69) */
#define PE_WEED_PLANT_PLUGIN_INFO 1
#define PE_WEED_PLANT_FILTER_CLASS 2
#define PE_WEED_PLANT_CHANNEL_TEMPLATE 4
#define PE_WEED_PLANT_PARAMETER_TEMPLATE 5
#define PE_WEED_PLANT_GUI 8

/* flags (weed-effects.h:73,107-125) */
#define PE_WEED_PARAM_INTEGER 1
#define PE_WEED_PARAM_FLOAT 2
#define PE_WEED_PARAM_COLOR 5
#define PE_WEED_COLORSPACE_RGB 1                       /* weed-effects.h:96 */
#define PE_WEED_FILTER_CHANNEL_SIZES_MAY_VARY (1 << 8) /* weed-effects.h:113 */
#define PE_WEED_PARAM_SWITCH 4
#define PE_WEED_PARAMETER_REINIT_ON_VALUE_CHANGE (1 << 0)
#define PE_WEED_FILTER_HINT_STATEFUL (1 << 2)
#define PE_WEED_FILTER_PREF_LINEAR_GAMMA (1 << 3)
#define PE_WEED_FILTER_HINT_MAY_THREAD (1 << 6)
#define PE_WEED_CHANNEL_CAN_DO_INPLACE (1 << 4)
#define PE_WEED_CHANNEL_REINIT_ON_SIZE_CHANGE (1 << 0) /* weed-effects.h:121 */

/* leaf names (weed.h:493-496, weed-effects.h:195-424) */
#define PE_LEAF_TYPE "type"
#define PE_LEAF_WEED_API_VERSION "weed_api_version"
#define PE_LEAF_FILTER_API_VERSION "filter_api_version"
#define PE_LEAF_GET_FUNC "weed_leaf_get_func"
#define PE_LEAF_SET_FUNC "weed_leaf_set_func"
#define PE_LEAF_PLANT_NEW_FUNC "weed_plant_new_func"
#define PE_LEAF_NUM_ELEMENTS_FUNC "weed_leaf_num_elements_func"
#define PE_LEAF_MALLOC_FUNC "weed_malloc_func"
#define PE_LEAF_FREE_FUNC "weed_free_func"
#define PE_LEAF_FILTERS "filters"
#define PE_LEAF_HOST_INFO "host_info"
#define PE_LEAF_PLUGIN_INFO "plugin_info"
#define PE_LEAF_VERSION "version"
#define PE_LEAF_FLAGS "flags"
#define PE_LEAF_NAME "name"
#define PE_LEAF_AUTHOR "author"
#define PE_LEAF_PALETTE_LIST "palette_list"
#define PE_LEAF_INIT_FUNC "init_func"
#define PE_LEAF_DEINIT_FUNC "deinit_func"
#define PE_LEAF_PROCESS_FUNC "process_func"
#define PE_LEAF_IN_PARAMETER_TEMPLATES "in_param_tmpls"
#define PE_LEAF_OUT_PARAMETER_TEMPLATES "out_param_tmpls"
#define PE_LEAF_IN_CHANNEL_TEMPLATES "in_chan_tmpls"
#define PE_LEAF_OUT_CHANNEL_TEMPLATES "out_chan_tmpls"
#define PE_LEAF_GUI "gui"
#define PE_LEAF_LABEL "label"
#define PE_LEAF_USE_MNEMONIC "use_mnemonic"
#define PE_LEAF_WIDTH "width"
#define PE_LEAF_HEIGHT "height"
#define PE_LEAF_PIXEL_DATA "pixel_data"
#define PE_LEAF_CURRENT_PALETTE "current_palette"
#define PE_LEAF_ROWSTRIDES "rowstrides"
#define PE_LEAF_IN_PARAMETERS "in_parameters"
#define PE_LEAF_IN_CHANNELS "in_channels"
#define PE_LEAF_OUT_CHANNELS "out_channels"
#define PE_LEAF_VALUE "value"
#define PE_LEAF_DEFAULT "default"
#define PE_LEAF_MIN "min"
#define PE_LEAF_MAX "max"
#define PE_LEAF_PARAM_TYPE "param_type"
#define PE_LEAF_IS_TRANSITION "is_transition"
#define PE_LEAF_GROUP "group"
#define PE_LEAF_MAX_REPEATS "max_repeats" /* weed-effects.h:339 */
#define PE_LEAF_COLORSPACE "colorspace"   /* weed-effects.h:392 */
#define PE_LEAF_ALIGNMENT_HINT "alignment_hint" /* weed-effects.h:265, honoured at src/effects-weed.c:2319-2324 */

#ifdef __cplusplus
extern "C" {
#endif
/* what the host dlsym()s (weed-effects.h:181-183) */
pe_weed_plant_t *weed_setup(pe_weed_bootstrap_f weed_boot);
void weed_desetup(void);
#ifdef __cplusplus
}
#endif
#endif /* PE_WEED_ABI_H */
