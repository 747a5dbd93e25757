/* cmos_b200.h — C ABI of the B200-native ORB front-end and bundle-adjustment engine.
 *
 * This is the drop-in boundary for the two hot paths of b51/ceres_mono_orb_slam2.
 * The reference has no FFI layer: its boundary is three C++ classes.  The C++ adapters in
 * include/ORBextractor.h, include/ORBmatcher.h and include/CeresOptimizer.h keep those class
 * signatures and call the functions declared here; INTEGRATION.md shows the binding.
 *
 * Conventions
 *  - plain pointers and sizes only; caller allocates, callee fills; capacity in, count out.
 *  - every function returns 0 on success or a negative cmos_status; cmos_last_error() gives the text.
 *    Nothing throws across this boundary.  There is NO CPU fallback: without a CUDA device every compute
 *    entry point returns CMOS_ERR_CUDA.
 *  - handles are not thread-safe; use one handle per calling thread (the reference calls the extractor
 *    from the Tracking thread only, Frame.cc:175-177).
 *  - `stream` arguments are cudaStream_t passed as void*; NULL means the handle's own stream.
 *  - "_device" variants take device pointers, enqueue on `stream` and do not synchronise.
 */
#ifndef CMOS_B200_H
#define CMOS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  CMOS_OK = 0,
  CMOS_ERR_ARG = -1,        /* bad argument / size over the handle's capacity */
  CMOS_ERR_CUDA = -2,       /* CUDA runtime error, or no device */
  CMOS_ERR_CAPACITY = -3,   /* an internal or caller buffer overflowed; results are not valid */
  CMOS_ERR_STATE = -4       /* call order violated (e.g. download before extract) */
} cmos_status;

const char* cmos_last_error(void);
/* Library version string, e.g. "cmos_b200 0.1 sm_100a". */
const char* cmos_version(void);

/* ------------------------------------------------------------------------------------------------
 * Path 1a: ORBextractor  (replaces ORBextractor::ORBextractor / operator(), reference
 * include/ORBextractor.h:46-61, src/ORBextractor.cc:410-470,1043-1105)
 * ---------------------------------------------------------------------------------------------- */

/* Same memory layout as cv::KeyPoint (28 bytes): what operator() appends to its std::vector. */
typedef struct {
  float x, y;        /* pt, level-0 pixel coordinates (already multiplied by the level scale) */
  float size;        /* (int)(31 * scale[octave])  — ORBextractor.cc:837,846 */
  float angle;       /* degrees in [0,360), cv::fastAtan2 of the intensity centroid — :103 */
  float response;    /* FAST score */
  int32_t octave;
  int32_t class_id;  /* always -1 */
} cmos_keypoint;

typedef struct {
  int32_t nfeatures;       /* ORBextractor.nFeatures  */
  float scale_factor;      /* ORBextractor.scaleFactor */
  int32_t nlevels;         /* ORBextractor.nLevels (1..CMOS_MAX_LEVELS) */
  int32_t ini_th_fast;     /* ORBextractor.iniThFAST */
  int32_t min_th_fast;     /* ORBextractor.minThFAST */
  int32_t max_width;       /* largest image this handle will see */
  int32_t max_height;
  int32_t max_batch;       /* frames per call */
  int32_t device;          /* CUDA device ordinal */
} cmos_orb_params;

#define CMOS_MAX_LEVELS 16

typedef struct cmos_orb* cmos_orb_t;

int cmos_orb_create(const cmos_orb_params* params, cmos_orb_t* out);
int cmos_orb_destroy(cmos_orb_t h);

/* Getters of ORBextractor.h:63-83.  Each fills nlevels entries. */
int cmos_orb_get_levels(cmos_orb_t h, int32_t* nlevels);
int cmos_orb_get_scale_factors(cmos_orb_t h, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2);
int cmos_orb_get_features_per_level(cmos_orb_t h, int32_t* quota);
/* Largest number of keypoints one frame can return: sum over levels of (quota + 3). */
int cmos_orb_keypoint_capacity(cmos_orb_t h, int32_t* cap);

/* operator() over a batch of n_frames same-sized CV_8UC1 images in HOST memory (pinned or pageable).
 *   images       frame f, row y starts at images + f*frame_stride + y*pitch
 *   keypoints    [n_frames][capacity]            (frame f's first counts[f] entries are valid)
 *   descriptors  [n_frames][capacity][32]        rows follow keypoint order (ORBextractor.cc:1088-1090)
 *   counts       [n_frames]
 * Synchronous.  Empty image (width or height 0) -> counts = 0, like ORBextractor.cc:1046-1047. */
int cmos_orb_extract(cmos_orb_t h, const uint8_t* images, int64_t frame_stride, int32_t pitch,
                     int32_t width, int32_t height, int32_t n_frames, cmos_keypoint* keypoints,
                     uint8_t* descriptors, int32_t* counts, int32_t capacity);

/* Same, images already in DEVICE memory; results stay in the handle's device buffers (see
 * cmos_orb_device_results); enqueues on `stream` and returns without synchronising. */
int cmos_orb_extract_device(cmos_orb_t h, const uint8_t* d_images, int64_t frame_stride, int32_t pitch,
                            int32_t width, int32_t height, int32_t n_frames, void* stream);

/* Device views of the last extraction: keypoints [max_batch][cap], descriptors [max_batch][cap][32],
 * counts [max_batch], level_counts [max_batch][CMOS_MAX_LEVELS]. */
int cmos_orb_device_results(cmos_orb_t h, cmos_keypoint** d_keypoints, uint8_t** d_descriptors,
                            int32_t** d_counts, int32_t** d_level_counts, int32_t* capacity);

/* Copies the last extraction to host buffers (synchronises the stream).  Returns CMOS_ERR_CAPACITY if any
 * internal candidate buffer overflowed during the extraction. */
int cmos_orb_download(cmos_orb_t h, int32_t n_frames, cmos_keypoint* keypoints, uint8_t* descriptors,
                      int32_t* counts, int32_t capacity, void* stream);

/* Debug/verification taps (used by the parity tests; not needed by a caller of operator()).
 * Level geometry of the last extraction; image = bordered plane of `level` for frame `frame`
 * copied into a tightly packed (w+38) x (h+38) host buffer; blurred = w x h. */
int cmos_orb_level_size(cmos_orb_t h, int32_t level, int32_t* w, int32_t* ht);
int cmos_orb_debug_level_image(cmos_orb_t h, int32_t frame, int32_t level, uint8_t* out);
int cmos_orb_debug_level_blurred(cmos_orb_t h, int32_t frame, int32_t level, uint8_t* out);
/* FAST candidates of (frame, level) before the quadtree: packed x | y<<12 | score<<24, coordinates
 * relative to minBorder (ORBextractor.cc:820-825), unordered.  Returns the count through *n. */
int cmos_orb_debug_level_candidates(cmos_orb_t h, int32_t frame, int32_t level, uint32_t* out,
                                    int32_t capacity, int32_t* n);
/* Selects the FAST cell (index into the handle's cell table, frame 0) whose shared-memory tile and score map the
 * next extraction records; if `out` is non-NULL first copies the previous record: 2*72*80 bytes + 16 ints. */
int cmos_orb_debug_fast_cell(cmos_orb_t h, int32_t cell, uint8_t* out);
/* Number of kernels launched by the last cmos_orb_extract* call. */
int cmos_orb_last_launch_count(cmos_orb_t h, int32_t* n);
/* Device evaluation of the two float helpers that must round like the CPU (cosf/sinf of angle*pi/180 as
 * computeOrbDescriptor uses them, ORBextractor.cc:112-113; cv::fastAtan2, :103).  Host arrays in and out. */
int cmos_debug_sincos_deg(const float* deg, float* cos_out, float* sin_out, int32_t n);
int cmos_debug_fast_atan2(const float* y, const float* x, float* out, int32_t n);

#ifdef __cplusplus
}
#endif
#endif /* CMOS_B200_H */
