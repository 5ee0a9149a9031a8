/* ref_shim.h -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Minimal environment that lets line-range slices of the reference's
 * src/colourspace.c compile stand-alone (the file cannot be built whole: it
 * pulls GTK through main.h).  oracle/build_ref.py concatenates
 *     this header + slices of colourspace.h + slices of colourspace.c
 *     + oracle/ref_wrappers.c
 * into oracle/_ref/ (git-ignored) and builds libref_oracle.so from it.
 * Nothing in here restates reference code; it only supplies the names the
 * slices expect from the rest of LiVES.
 */
#ifndef PE_REF_SHIM_H
#define PE_REF_SHIM_H

#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <math.h>
#include <pthread.h>

typedef int boolean;
#ifndef TRUE
#define TRUE 1
#define FALSE 0
#endif
typedef int64_t ticks_t;
typedef void *livespointer;

#define LIVES_RESTRICT __restrict__
#define LIVES_HOT __attribute__((hot))
#define LIVES_FLATTEN
#define LIVES_INLINE static inline
#define LIVES_LOCAL_INLINE static inline
#define LIVES_GLOBAL_INLINE
#define LIVES_UNLIKELY(x) __builtin_expect(!!(x), 0)
#define LIVES_LIKELY(x) __builtin_expect(!!(x), 1)
#define LIVES_CONST
#define LIVES_PURE
#define LIVES_ASSERT(x) ((void)0)
#define LIVES_DEBUG(x) ((void)0)
#define LIVES_WARN(x) ((void)0)
#define MIN(a, b) ((a) < (b) ? (a) : (b))
#define MAX(a, b) ((a) > (b) ? (a) : (b))

#define WEED_MAXPPLANES 4
#define WEED_MAXPCHANS 8

/* palettes / clamping / subspace / gamma enums come from the reference's own
 * libweed/weed-palettes.h, added to the include path by build_ref.py */
#include "weed-palettes.h"

/* the three src/maths.h macros the slices use are extracted by build_ref.py
 * into ref_maths_slice.h (maths.h:88,101,104,118) */
#include "ref_maths_slice.h"

#define PB_QUALITY_LOW 1
#define PB_QUALITY_MED 2
#define PB_QUALITY_HIGH 3
#define OBJ_INTENTION_PLAY 0
#define OBJ_INTENTION_RENDER 1
#define OBJ_INTENTION_TRANSCODE 2
#define EFFORT_RANGE_MAX 64

typedef enum {
  LIVES_DIRECTION_BACKWARD = -1,
  LIVES_DIRECTION_NONE = 0,
  LIVES_DIRECTION_FORWARD = 1,
} lives_direction_t;
#define LIVES_DIRECTION_REVERSE LIVES_DIRECTION_BACKWARD

typedef struct {
  int nfx_threads;
  short pb_quality;
  double screen_gamma;
  boolean apply_gamma;
  boolean alpha_post;
} ref_prefs_t;
extern ref_prefs_t *prefs, *future_prefs;

typedef struct { int effort; } ref_mainw_t;
extern ref_mainw_t *mainw;

/* a flat stand-in for the weed layer plant: only what alpha_premult() and the
 * gamma code read */
typedef struct ref_layer {
  int width, height, palette, clamping, gamma, flags, nplanes;
  int rowstrides[WEED_MAXPPLANES];
  void *pixel_data[WEED_MAXPPLANES];
} weed_layer_t;

static inline int weed_layer_get_width(weed_layer_t *l) { return l->width; }
static inline int weed_layer_get_height(weed_layer_t *l) { return l->height; }
static inline int weed_layer_get_rowstride(weed_layer_t *l) { return l->rowstrides[0]; }
static inline int weed_layer_get_palette(weed_layer_t *l) { return l->palette; }
static inline int weed_layer_get_yuv_clamping(weed_layer_t *l) { return l->clamping; }
static inline int weed_layer_get_flags(weed_layer_t *l) { return l->flags; }
static inline void weed_layer_set_flags(weed_layer_t *l, int f) { l->flags = f; }
static inline void *weed_layer_get_pixel_data(weed_layer_t *l) { return l->pixel_data[0]; }
static inline void **weed_layer_get_pixel_data_planar(weed_layer_t *l, int *n) {
  void **r = (void **)malloc(WEED_MAXPPLANES * sizeof(void *));
  memcpy(r, l->pixel_data, WEED_MAXPPLANES * sizeof(void *));
  if (n) *n = l->nplanes;
  return r;
}
static inline int *weed_layer_get_rowstrides(weed_layer_t *l, int *n) {
  int *r = (int *)malloc(WEED_MAXPPLANES * sizeof(int));
  memcpy(r, l->rowstrides, WEED_MAXPPLANES * sizeof(int));
  if (n) *n = l->nplanes;
  return r;
}

/* THREADVAR(conv_arrays): thread-local conversion-table selection */
struct _conv_array;
#define THREADVAR(x) (ref_tls_##x)

/* thread pool -> plain pthreads */
typedef pthread_t lives_thread_t; /* slices declare "lives_thread_t *threads[n]" */
#define LIVES_THRDATTR_PRIORITY 0
typedef void *(*ref_thread_fn)(void *);
static inline int ref_thread_create(lives_thread_t **slot, ref_thread_fn fn, void *arg) {
  pthread_t *t = (pthread_t *)malloc(sizeof(pthread_t));
  *slot = t;
  return pthread_create(t, NULL, fn, arg);
}
static inline void ref_thread_join(lives_thread_t *slot) {
  pthread_join(*slot, NULL);
  free(slot);
}
#define lives_thread_create(pslot, attr, fn, arg) ref_thread_create((pslot), (ref_thread_fn)(fn), (arg))
#define lives_thread_join(slot, ret) ref_thread_join(slot)

#define lives_calloc calloc
#define lives_malloc malloc
#define lives_free free
#define lives_memcpy memcpy
#define lives_memset memset
#define lives_memcmp memcmp

static inline void swab4(void *to, const void *from, size_t gran) {
  (void)gran;
  const uint8_t *s = (const uint8_t *)from;
  uint8_t t[4] = {s[3], s[2], s[1], s[0]};
  memcpy(to, t, 4);
}

#endif
