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

/* Host buffers like cmos_orb_extract, but only ENQUEUES on `stream` (NULL = the handle's stream): the image upload,
 * the kernels and the download of keypoints / descriptors / counts.  The caller's buffers must stay valid — and should
 * be page-locked, otherwise the copies serialise — until cmos_orb_finish returns.  cmos_orb_finish waits for the
 * stream and returns CMOS_ERR_CAPACITY if a candidate buffer overflowed (results invalid). */
int cmos_orb_extract_async(cmos_orb_t h, const uint8_t* images, int64_t frame_stride, int32_t pitch,
                           int32_t width, int32_t height, int32_t n_frames, cmos_keypoint* keypoints,
                           uint8_t* descriptors, int32_t* counts, int32_t capacity, void* stream);
int cmos_orb_finish(cmos_orb_t h, void* stream);
/* Only the candidate-overflow check of cmos_orb_finish, for callers that synchronised through an event of their own. */
int cmos_orb_check_overflow(cmos_orb_t h);

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
/* Per-stage device timing (CUDA events recorded on the launching stream around each stage of every
 * cmos_orb_extract* call while enabled).  Stages: 0 pyramid, 1 FAST+NMS, 2 quadtree, 3 blur, 4 orientation +
 * descriptors.  cmos_orb_stage_times returns the milliseconds accumulated since profiling was enabled, summed
 * over `calls` extractions (it waits for the recorded events). */
#define CMOS_ORB_STAGES 5
int cmos_orb_set_profiling(cmos_orb_t h, int32_t enable);
int cmos_orb_stage_times(cmos_orb_t h, double* ms, int32_t capacity, int32_t* n_stages, int64_t* calls);
/* Number of kernels launched by the last cmos_orb_extract* call. */
int cmos_orb_last_launch_count(cmos_orb_t h, int32_t* n);
/* Device evaluation of the two float helpers that must round like the CPU (cosf/sinf of angle*pi/180 as
 * computeOrbDescriptor uses them, ORBextractor.cc:112-113; cv::fastAtan2, :103).  Host arrays in and out. */
int cmos_debug_sincos_deg(const float* deg, float* cos_out, float* sin_out, int32_t n);
int cmos_debug_fast_atan2(const float* y, const float* x, float* out, int32_t n);

/* ------------------------------------------------------------------------------------------------
 * Path 1b: ORBmatcher  (replaces ORBmatcher::SearchByProjection x2 and the Frame grid they read;
 * reference include/ORBmatcher.h:43-53, src/ORBmatcher.cc:42-126,1161-1271,1386-1437,
 * src/Frame.cc:158-173,191-320).  Pointer graphs are flattened: a MapPoint* is an index,
 * "F.map_points_[i] && F.map_points_[i]->Observations() > 0" is the per-keypoint `claimed` byte.
 * Every array argument is batched: [n_frames][stride]...; with on_device != 0 all array pointers are device
 * pointers, the call enqueues on `stream` and does not synchronise; with on_device == 0 they are host
 * pointers and the call is synchronous.
 * ---------------------------------------------------------------------------------------------- */

#define CMOS_GRID_COLS 64   /* FRAME_GRID_COLS, Frame.h:46 */
#define CMOS_GRID_ROWS 48   /* FRAME_GRID_ROWS, Frame.h:45 */
#define CMOS_TH_HIGH 100    /* ORBmatcher::TH_HIGH, ORBmatcher.cc:35 */
#define CMOS_TH_LOW 50      /* ORBmatcher::TH_LOW,  ORBmatcher.cc:36 */
#define CMOS_HISTO_LENGTH 30

/* The Frame statics a search reads (Frame.h:158-189): image bounds, grid pitch, intrinsics, scale table. */
typedef struct {
  float min_x, max_x, min_y, max_y;                         /* Frame::ComputeImageBounds, Frame.cc:357-385 */
  float grid_element_width_inv, grid_element_height_inv;   /* Frame.cc:133-136 */
  float fx, fy, cx, cy;                                     /* float32 like Frame.cc:138-141 */
  int32_t nlevels;
  float scale_factors[CMOS_MAX_LEVELS];
  float log_scale_factor;                                   /* log(scaleFactor) as float, Frame.cc:110 */
} cmos_camera;

/* Fills `cam` for an undistorted width x height image (the k1 == 0 branch of ComputeImageBounds). */
int cmos_camera_init(cmos_camera* cam, int32_t width, int32_t height, float fx, float fy, float cx, float cy,
                     const float* scale_factors, int32_t nlevels, float scale_factor);

/* Same for a camera with lens distortion (k1 != 0): the bounds are the undistorted image corners
 * (Frame::ComputeImageBounds, Frame.cc:357-385); dist_coef = k1, k2, p1, p2[, k3] as in the settings file. */
int cmos_camera_init_distorted(cmos_camera* cam, int32_t width, int32_t height, float fx, float fy, float cx, float cy,
                               const float* dist_coef, int32_t n_dist, const float* scale_factors, int32_t nlevels,
                               float scale_factor, int32_t device);

typedef struct {
  int32_t max_batch;       /* frames per call */
  int32_t max_keypoints;   /* keypoints per frame (stride upper bound), <= 65535 */
  int32_t max_points;      /* map points per frame for search_by_projection_points, <= 65535 */
  int32_t device;
} cmos_match_params;

typedef struct cmos_match* cmos_match_t;

int cmos_match_create(const cmos_match_params* params, cmos_match_t* out);
int cmos_match_destroy(cmos_match_t h);

/* Binds the current frames of the batch and runs Frame::AssignFeaturesToGrid (Frame.cc:158-173) for each:
 *   keypoints [n_frames][stride] (undistort_keypoints_), descriptors [n_frames][stride][32], counts [n_frames].
 * With on_device the handle keeps the pointers (they must stay valid until the searches have run). */
int cmos_match_set_frames(cmos_match_t h, const cmos_camera* cam, const cmos_keypoint* keypoints,
                          const uint8_t* descriptors, const int32_t* counts, int32_t n_frames, int32_t stride,
                          int32_t on_device, void* stream);

/* Frame::UndistortKeyPoints (Frame.cc:329-355) for a batch: keypoints [n_frames][stride] -> undistorted (same layout;
 * pt replaced, every other field copied).  cv::undistortPoints(mat, mat, K, dist, Mat(), K) semantics of OpenCV 4.13
 * (5 iterations, double arithmetic); k1 == 0 copies.  May run in place (undistorted == keypoints) with on_device. */
int cmos_match_undistort_keypoints(cmos_match_t h, const float* K4, const float* dist_coef, int32_t n_dist,
                                   const cmos_keypoint* keypoints, const int32_t* counts, int32_t n_frames,
                                   int32_t stride, cmos_keypoint* undistorted, int32_t on_device, void* stream);

/* Verification tap: CSR grid of one frame, cell = ix*48+iy: grid_start[64*48+1], grid_idx[stride]. Host out. */
int cmos_match_debug_grid(cmos_match_t h, int32_t frame, int32_t* grid_start, int32_t* grid_idx);

/* ORBmatcher::SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, th)  (ORBmatcher.cc:1161-1271)
 *   Tcw          [n_frames][16]  row-major CurrentFrame.Tcw_ (double)
 *   last_*       [n_frames][last_stride]...: keypoints (octave, angle read), counts, flags (bit0: map point
 *                present and not outlier; bit1: that point has Observations() > 0), world position
 *                [..][3] double, MapPoint::GetDescriptor() [..][32]
 *   claimed      [n_frames][stride] in/out, may be NULL (= all zero, not returned)
 *   match        [n_frames][stride] out: last-frame index assigned to each current keypoint, or -1
 *   nmatches     [n_frames] out: the function's return value */
int cmos_match_search_by_projection_frame(cmos_match_t h, const double* Tcw, const cmos_keypoint* last_keypoints,
                                          const int32_t* last_counts, const uint8_t* last_flags,
                                          const double* last_xw, const uint8_t* last_descriptors,
                                          int32_t last_stride, float th, int32_t check_orientation,
                                          uint8_t* claimed, int32_t* match, int32_t* nmatches, int32_t on_device,
                                          void* stream);

/* ORBmatcher::SearchByProjection(Frame& F, const vector<MapPoint*>&, th)  (ORBmatcher.cc:42-119)
 *   per point [n_frames][point_stride]: in_view (is_track_in_view_ && !isBad()), level (track_scale_level_),
 *   view_cos, proj_xy [..][2], descriptors [..][32], has_obs (Observations() > 0); n_points [n_frames]
 *   assign [n_frames][stride] out: point index newly written to F.map_points_[i], or -1 */
int cmos_match_search_by_projection_points(cmos_match_t h, const int32_t* n_points, const uint8_t* in_view,
                                           const int32_t* level, const float* view_cos, const float* proj_xy,
                                           const uint8_t* descriptors, const uint8_t* has_obs,
                                           int32_t point_stride, float th, float nn_ratio, uint8_t* claimed,
                                           int32_t* assign, int32_t* nmatches, int32_t on_device, void* stream);

/* Frame::isInFrustum + MapPoint::PredictScale for a batch of points (Frame.cc:191-241, MapPoint.cc:405-420).
 *   pose15 [n_frames][15]: Rcw row-major (9), tcw (3), Ow (3), double;  per point: world position, normal
 *   (double[3]), min/max distance (float, un-scaled: the 0.8 / 1.2 factors are applied inside).
 *   Outputs feed cmos_match_search_by_projection_points. */
int cmos_match_is_in_frustum(cmos_match_t h, const cmos_camera* cam, const double* pose15, float view_cos_limit,
                             const int32_t* n_points, const double* xw, const double* normal,
                             const float* min_distance, const float* max_distance, int32_t point_stride,
                             int32_t n_frames, uint8_t* in_view, float* proj_xy, int32_t* level, float* view_cos,
                             int32_t on_device, void* stream);

/* Per-kernel device timing, as cmos_orb_set_profiling.  ms[4] / calls[4]: 0 grid, 1 search(frame,last),
 * 2 search(frame,points), 3 isInFrustum. */
int cmos_match_set_profiling(cmos_match_t h, int32_t enable);
int cmos_match_stage_times(cmos_match_t h, double* ms, int64_t* calls);
/* Number of kernels launched by the last cmos_match_* call. */
int cmos_match_last_launch_count(cmos_match_t h, int32_t* n);

/* ------------------------------------------------------------------------------------------------
 * Path 1c: the remaining ORBmatcher searches: relocalisation / loop-closing projections,
 * the two SearchByBoW, SearchForInitialization, SearchForTriangulation, both Fuse and SearchBySim3
 * (reference include/ORBmatcher.h:50-96, src/ORBmatcher.cc:128-1159,1273-1384).  One problem per call, HOST
 * pointers, synchronous.  A Frame / KeyFrame is bound to one of two view slots first (keypoints, descriptors, the
 * grid of Frame::AssignFeaturesToGrid).  For a KeyFrame the searches use KeyFrame::GetFeaturesInArea / IsInImage,
 * whose image bounds are the Frame's truncated to int (KeyFrame.h:179-182, KeyFrame.cc:575-626); the grid itself
 * is the Frame's (KeyFrame.cc:98-104).
 * Pointer graphs are flattened like Path 1b.  Predicates that cannot change during a call are bytes prepared by the
 * caller.  The map-mutating Fuse functions return DECISIONS (best keypoint per map point); the host adapter applies
 * Replace / AddObservation in the reference's order (include/orb_slam2/ORBmatcher.h).
 * A DBoW2::FeatureVector (map<NodeId, vector<unsigned>>) is passed flattened: node ids ascending, start[n_nodes+1],
 * feature indices in insertion order.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t max_keypoints;   /* keypoints per view, <= 16384 */
  int32_t max_points;      /* map points per call */
  int32_t max_nodes;       /* FeatureVector nodes per view */
  int32_t device;
} cmos_kfmatch_params;

typedef struct cmos_kfmatch* cmos_kfmatch_t;

typedef struct {           /* flattened DBoW2::FeatureVector */
  int32_t n_nodes;
  const int32_t* node_ids;   /* [n_nodes] ascending */
  const int32_t* start;      /* [n_nodes + 1] */
  const int32_t* features;   /* [start[n_nodes]] */
} cmos_feature_vector;

int cmos_kfmatch_create(const cmos_kfmatch_params* params, cmos_kfmatch_t* out);
int cmos_kfmatch_destroy(cmos_kfmatch_t h);
/* Binds view `slot` (0 or 1): undistort_keypoints_, descriptors_, N_, and builds the grid.  is_keyframe != 0 selects
 * the KeyFrame flavour of GetFeaturesInArea / IsInImage (int-truncated bounds). */
int cmos_kfmatch_set_view(cmos_kfmatch_t h, int32_t slot, const cmos_camera* cam, int32_t is_keyframe,
                          const cmos_keypoint* keypoints, const uint8_t* descriptors, int32_t n);

/* SearchByProjection(Frame& CurrentFrame, KeyFrame* pKF, const set<MapPoint*>& sAlreadyFound, th, ORBdist)
 * (ORBmatcher.cc:1273-1384).  Slot 0 = CurrentFrame.  Per keyframe keypoint i (n_kf of them): kf_valid = pMP &&
 * !isBad && !sAlreadyFound.count(pMP); world position, raw min/max distance (the 0.8 / 1.2 factors are applied
 * inside), MapPoint::GetDescriptor(), angle of pKF->undistort_keypoints_[i].
 *   cur_has_point [n] in/out: CurrentFrame.map_points_[i2] != NULL;  cur_match [n] out: i or -1 */
int cmos_kfmatch_search_by_projection_reloc(cmos_kfmatch_t h, const double* Tcw16, int32_t n_kf, const uint8_t* kf_valid,
                                            const double* kf_xw, const float* kf_min_distance,
                                            const float* kf_max_distance, const uint8_t* kf_descriptors,
                                            const float* kf_angle, float th, int32_t orb_dist,
                                            int32_t check_orientation, uint8_t* cur_has_point, int32_t* cur_match,
                                            int32_t* nmatches);
/* SearchByProjection(KeyFrame* pKF, Scw, vpPoints, vpMatched, th) (ORBmatcher.cc:258-361).  Slot 0 = pKF.
 *   pt_skip = isBad || spAlreadyFound.count;  matched [n] in/out: vpMatched[idx] != NULL;  assign [n] out: point or -1 */
int cmos_kfmatch_search_by_projection_sim3(cmos_kfmatch_t h, const double* Scw16, int32_t n_points,
                                           const uint8_t* pt_skip, const double* xw, const double* normal,
                                           const float* min_distance, const float* max_distance,
                                           const uint8_t* pt_descriptors, int32_t th, uint8_t* matched, int32_t* assign,
                                           int32_t* nmatches);
/* Fuse(KeyFrame*, vpMapPoints, th) (ORBmatcher.cc:724-842; sim3 == 0, pose = Rcw row-major (9), tcw (3), Ow (3)) and
 * Fuse(KeyFrame*, Scw, vpPoints, th, vpReplacePoint) (:844-954; sim3 == 1, pose = Scw row-major 4x4).  Slot 0 = pKF.
 *   pt_skip = !pMP || isBad || IsInKeyFrame(pKF) (resp. spAlreadyFound.count);  inv_level_sigma2 = pKF's table.
 *   best_idx [n_points] out: keypoint to fuse with or -1;  best_dist [n_points];  n_fused = decisions made */
int cmos_kfmatch_fuse(cmos_kfmatch_t h, int32_t sim3, const double* pose, const float* inv_level_sigma2,
                      int32_t n_points, const uint8_t* pt_skip, const double* xw, const double* normal,
                      const float* min_distance, const float* max_distance, const uint8_t* pt_descriptors, float th,
                      int32_t* best_idx, int32_t* best_dist, int32_t* n_fused);
/* SearchBySim3 (ORBmatcher.cc:956-1159).  Slot 0 = pKF1, slot 1 = pKF2.  pose{1,2} = R (9) + t (3) of the keyframes.
 * Per keypoint of each keyframe: valid = pMP && !isBad, already = vbAlreadyMatched, world position, raw min/max
 * distance, MapPoint descriptor.  match12 [n1] out: idx2 or -1 (the mutual matches). */
int cmos_kfmatch_search_by_sim3(cmos_kfmatch_t h, const double* pose1, const double* pose2, float s12,
                                const double* R12, const double* t12, const uint8_t* valid1, const uint8_t* already1,
                                const double* xw1, const float* min_distance1, const float* max_distance1,
                                const uint8_t* mp_descriptors1, const uint8_t* valid2, const uint8_t* already2,
                                const double* xw2, const float* min_distance2, const float* max_distance2,
                                const uint8_t* mp_descriptors2, float th, int32_t* match12, int32_t* n_found);
/* SearchByBoW(KeyFrame*, Frame&, matches) (ORBmatcher.cc:151-256; mode 0: slot 0 = pKF, slot 1 = F, match [n2] out:
 * keyframe keypoint whose map point F's keypoint receives) and SearchByBoW(KeyFrame*, KeyFrame*, matches12)
 * (:470-580; mode 1: match [n1] out: idx2).  valid = pMP && !isBad per keypoint (valid2 unused in mode 0). */
int cmos_kfmatch_search_by_bow(cmos_kfmatch_t h, int32_t mode, const uint8_t* valid1, const cmos_feature_vector* fv1,
                               const uint8_t* valid2, const cmos_feature_vector* fv2, float nn_ratio,
                               int32_t check_orientation, int32_t* match, int32_t* nmatches);
/* SearchForTriangulation (ORBmatcher.cc:582-722, monocular).  Slot 0 = pKF1, slot 1 = pKF2.  has_point = GetMapPoint(i)
 * != NULL.  F12 row-major 3x3; Cw = pKF1->GetCameraCenter(); R2w, t2w of pKF2; level_sigma2_2 = pKF2's table.
 *   match12 [n1] out: idx2 or -1 (vMatchedPairs = the non-negative entries in index order) */
int cmos_kfmatch_search_for_triangulation(cmos_kfmatch_t h, const uint8_t* has_point1, const cmos_feature_vector* fv1,
                                          const uint8_t* has_point2, const cmos_feature_vector* fv2, const double* F12,
                                          const double* Cw, const double* R2w, const double* t2w,
                                          const float* level_sigma2_2, int32_t check_orientation, int32_t* match12,
                                          int32_t* nmatches);
/* SearchForInitialization (ORBmatcher.cc:363-468).  Slot 0 = F1, slot 1 = F2.  prev_matched [n1][2] in/out. */
int cmos_kfmatch_search_for_initialization(cmos_kfmatch_t h, float* prev_matched, int32_t window_size, float nn_ratio,
                                           int32_t check_orientation, int32_t* matches12, int32_t* nmatches);
/* Kernels launched by the last cmos_kfmatch_* call. */
int cmos_kfmatch_last_launch_count(cmos_kfmatch_t h, int32_t* n);

/* ------------------------------------------------------------------------------------------------
 * MapPoint maintenance, batched over map points: what the reference runs for every touched
 * point after tracking, fusing and bundle adjustment.  Observations are passed as a CSR over points, in the iteration
 * order of each point's std::map<KeyFrame*, size_t> (tie-breaks and fp64 sums depend on it), bad keyframes dropped.
 * HOST pointers, synchronous.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t max_points;
  int32_t max_observations;   /* total over the points of one call */
  int32_t max_keyframes;
  int32_t device;
} cmos_map_params;
typedef struct cmos_map* cmos_map_t;
int cmos_map_create(const cmos_map_params* params, cmos_map_t* out);
int cmos_map_destroy(cmos_map_t h);
/* MapPoint::ComputeDistinctiveDescriptors (MapPoint.cc:256-315): descriptors [obs_start[n_points]][32] are the
 * observed keypoint descriptors;  best_index [n_points] out: position inside the point's list of the descriptor with
 * the least median Hamming distance to the others (-1 for a point without observations: descriptor_ untouched);
 * best_descriptor [n_points][32] out (may be NULL) = the new descriptor_.  At most 1024 observations per point. */
int cmos_map_distinctive_descriptors(cmos_map_t h, int32_t n_points, const int32_t* obs_start, const uint8_t* descriptors,
                                     int32_t* best_index, uint8_t* best_descriptor);
/* MapPoint::UpdateNormalAndDepth (MapPoint.cc:335-378): obs_keyframe = keyframe index of every observation,
 * camera_centers [n_keyframes][3] = GetCameraCenter(), world_pos [n_points][3], per point the reference keyframe and the
 * octave of its keypoint there, scale_factors = the keyframes' table.  normal [n_points][3], min_distance, max_distance
 * in/out: points without observations keep their values. */
int cmos_map_update_normal_and_depth(cmos_map_t h, int32_t n_points, const int32_t* obs_start, const int32_t* obs_keyframe,
                                     int32_t n_keyframes, const double* camera_centers, const double* world_pos,
                                     const int32_t* ref_keyframe, const int32_t* ref_level, const float* scale_factors,
                                     int32_t n_levels, double* normal, float* min_distance, float* max_distance);
int cmos_map_last_launch_count(cmos_map_t h, int32_t* n);

/* ------------------------------------------------------------------------------------------------
 * DBoW2 vocabulary transform: Frame::ComputeBoW / KeyFrame::ComputeBoW call
 * orb_vocabulary_->transform(descriptors, bow_vector_, feature_vector_, 4) (Frame.cc:322-327, KeyFrame.cc:107-117;
 * lib/DBoW2/DBoW2/TemplatedVocabulary.h:1127-1260) with ORBvoc's TF_IDF weighting and L1 scoring.
 * The vocabulary tree is uploaded once, flattened: node 0 = root, the children of node i are
 * children[child_start[i] .. child_start[i+1]) in stored order (m_nodes[i].children), a node without children is a
 * word (node_word_ids[i], node_weights[i] = its idf); inner nodes carry word id -1.
 * ---------------------------------------------------------------------------------------------- */
typedef struct cmos_voc* cmos_voc_t;
int cmos_voc_create(int32_t n_nodes, const int32_t* child_start, const int32_t* children, const uint8_t* node_descriptors,
                    const double* node_weights, const int32_t* node_word_ids, int32_t depth_levels, int32_t device,
                    cmos_voc_t* out);
int cmos_voc_destroy(cmos_voc_t h);
/* descriptors [n][32], n <= 8192.  BowVector out: bow_words / bow_values [n] (ascending word id, L1-normalised), *n_words;
 * FeatureVector out (cmos_feature_vector layout): fv_nodes [n], fv_start [n + 1], fv_features [n], *n_fv_nodes.  Features
 * whose word has weight 0 (stop words) appear in neither. */
int cmos_voc_transform(cmos_voc_t h, const uint8_t* descriptors, int32_t n, int32_t levelsup, int32_t* bow_words,
                       double* bow_values, int32_t* n_words, int32_t* fv_nodes, int32_t* fv_start, int32_t* fv_features,
                       int32_t* n_fv_nodes);
int cmos_voc_last_launch_count(cmos_voc_t h, int32_t* n);

/* ------------------------------------------------------------------------------------------------
 * Path 1a+1b fused for a batch of frames with HOST buffers: what the Tracking thread does per frame —
 * Frame::Frame (src/Frame.cc:98-156: ExtractORB at :116 -> ORBextractor::operator() at :175-177, AssignFeaturesToGrid
 * at :155) followed by TrackWithMotionModel's ORBmatcher::SearchByProjection(current_frame_, last_frame_, th)
 * (src/Tracking.cc:632; src/ORBmatcher.cc:1161-1271).  The batch is cut into chunks of `chunk_frames` that travel
 * down `lanes` CUDA streams, so uploads, kernels and downloads of neighbouring chunks overlap; results are
 * identical to cmos_orb_extract + cmos_match_set_frames + cmos_match_search_by_projection_frame.
 * Pass page-locked host memory for the overlap to happen.  Synchronous: everything is complete on return.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  cmos_orb_params orb;      /* max_batch is ignored (each lane is sized for chunk_frames) */
  int32_t lanes;            /* streams / handle sets, 1..16 */
  int32_t chunk_frames;     /* frames per chunk */
} cmos_track_params;

typedef struct cmos_track* cmos_track_t;

int cmos_track_create(const cmos_track_params* params, const cmos_camera* cam, cmos_track_t* out);
int cmos_track_destroy(cmos_track_t h);
int cmos_track_keypoint_capacity(cmos_track_t h, int32_t* cap);
/* images as cmos_orb_extract; Tcw / last_* / th / check_orientation as cmos_match_search_by_projection_frame with
 * [n_frames][last_stride] host arrays (last_stride <= capacity of the handle); keypoints / descriptors / counts as
 * cmos_orb_extract ([n_frames][capacity]...); match [n_frames][capacity], nmatches [n_frames]. */
int cmos_track_frames(cmos_track_t h, const uint8_t* images, int64_t frame_stride, int32_t pitch, int32_t width,
                      int32_t height, int32_t n_frames, const double* Tcw, const cmos_keypoint* last_keypoints,
                      const int32_t* last_counts, const uint8_t* last_flags, const double* last_xw,
                      const uint8_t* last_descriptors, int32_t last_stride, float th, int32_t check_orientation,
                      cmos_keypoint* keypoints, uint8_t* descriptors, int32_t* counts, int32_t capacity,
                      int32_t* match, int32_t* nmatches);
/* The same work split into an asynchronous pair, so that a caller can keep several batches in flight (upload of batch k + 1
 * under the kernels of batch k and the download of batch k - 1) instead of paying the pipeline fill and drain on every call:
 * cmos_track_submit enqueues everything and returns a ticket at once; cmos_track_wait blocks until that batch's outputs are
 * in the caller's buffers.  All buffers of a submitted batch must stay valid and untouched until its wait returns; at most
 * 4 batches may be in flight (CMOS_ERR_STATE beyond that).  cmos_track_frames == submit + wait.
 * (Tracking.cc:617-646 consumes frame t while the camera thread already holds frame t + 1.) */
int cmos_track_submit(cmos_track_t h, const uint8_t* images, int64_t frame_stride, int32_t pitch, int32_t width,
                      int32_t height, int32_t n_frames, const double* Tcw, const cmos_keypoint* last_keypoints,
                      const int32_t* last_counts, const uint8_t* last_flags, const double* last_xw,
                      const uint8_t* last_descriptors, int32_t last_stride, float th, int32_t check_orientation,
                      cmos_keypoint* keypoints, uint8_t* descriptors, int32_t* counts, int32_t capacity, int32_t* match,
                      int32_t* nmatches, int64_t* ticket);
int cmos_track_wait(cmos_track_t h, int64_t ticket);

/* The same call with the last-frame inputs in COMPACT form: one 64-byte record per last-frame keypoint that carries a usable
 * map point — exactly what the loop of ORBmatcher.cc:1187-1195 visits (`pMP && !LastFrame.is_outliers_[i]`) — instead of
 * 85 bytes for every keypoint slot (cmos_keypoint 28 + flag 1 + position 24 + descriptor 32).  The end-to-end call is bound
 * by the host->device copy, so the bytes matter.  Records of frame f are points[point_start[f] .. point_start[f + 1]), in
 * increasing keypoint index (the reference's visiting order: it decides who wins a contested keypoint); `match` reports
 * the records' `index`, i.e. the same last-frame keypoint indices cmos_track_submit returns.  Results are identical to
 * cmos_track_submit on the arrays the records were taken from. */
typedef struct cmos_last_point {
  uint8_t descriptor[32];  /* MapPoint::GetDescriptor() */
  double xw[3];            /* MapPoint::GetWorldPos() */
  float angle;             /* LastFrame.undistort_keypoints_[index].angle */
  uint16_t index;          /* keypoint index in the last frame */
  int8_t octave;           /* LastFrame.undistort_keypoints_[index].octave */
  uint8_t flags;           /* bit0 set (usable), bit1: the point has Observations() > 0 */
} cmos_last_point;         /* 64 bytes */
int cmos_track_submit_points(cmos_track_t h, const uint8_t* images, int64_t frame_stride, int32_t pitch, int32_t width,
                             int32_t height, int32_t n_frames, const double* Tcw, const cmos_last_point* points,
                             const int32_t* point_start, float th, int32_t check_orientation, cmos_keypoint* keypoints,
                             uint8_t* descriptors, int32_t* counts, int32_t capacity, int32_t* match, int32_t* nmatches,
                             int64_t* ticket);

/* Device-resident map points.  The reference's search reads pMP->GetWorldPos() / GetDescriptor() of every map point the last
 * frame holds (src/ORBmatcher.cc:1183, :1211) out of the Map in host memory; a map point changes when LocalMapping / BA moves it
 * (MapPoint::SetWorldPos, src/MapPoint.cc:112) or MapPoint::ComputeDistinctiveDescriptors picks another descriptor
 * (src/MapPoint.cc:256), i.e. per keyframe, not per frame.
 * The table mirrors those two fields on the device, one slot per map point (the caller's numbering, e.g. an index it keeps beside
 * MapPoint::id_): cmos_track_map_reserve sizes it (growing keeps the contents), cmos_track_map_update writes n slots — the
 * slots listed in `slots`, or first_slot .. first_slot + n - 1 when `slots` is NULL — from host arrays xw [n][3] and
 * descriptors [n][32].  Both calls wait for the batches in flight and return when the table is updated.
 * cmos_track_submit_map is cmos_track_submit_points with 12-byte association records instead of 64-byte ones: per frame the
 * upload names which last-frame keypoint holds which slot (records of frame f: assoc[assoc_start[f] .. assoc_start[f+1]), in
 * increasing keypoint index); position and descriptor come from the table.  Records whose slot lies outside the table are
 * treated as keypoints without a map point.  Results are identical to cmos_track_submit on the arrays the records and the
 * table were taken from. */
typedef struct cmos_track_assoc {
  int32_t slot;            /* map-point slot in the device table */
  float angle;             /* LastFrame.undistort_keypoints_[index].angle */
  uint16_t index;          /* keypoint index in the last frame */
  int8_t octave;           /* LastFrame.undistort_keypoints_[index].octave */
  uint8_t flags;           /* bit0 set (usable), bit1: the point has Observations() > 0 */
} cmos_track_assoc;        /* 12 bytes */
int cmos_track_map_reserve(cmos_track_t h, int32_t n_slots);
int cmos_track_map_update(cmos_track_t h, int32_t n, const int32_t* slots, int32_t first_slot, const double* xw,
                          const uint8_t* descriptors);
int cmos_track_submit_map(cmos_track_t h, const uint8_t* images, int64_t frame_stride, int32_t pitch, int32_t width,
                          int32_t height, int32_t n_frames, const double* Tcw, const cmos_track_assoc* assoc,
                          const int32_t* assoc_start, float th, int32_t check_orientation, cmos_keypoint* keypoints,
                          uint8_t* descriptors, int32_t* counts, int32_t capacity, int32_t* match, int32_t* nmatches,
                          int64_t* ticket);

/* Kernels launched by the last cmos_track_frames / cmos_track_submit call. */
int cmos_track_last_launch_count(cmos_track_t h, int32_t* n);

/* ------------------------------------------------------------------------------------------------
 * Path 2: CeresOptimizer  (replaces CeresOptimizer::PoseOptimization / LocalBundleAdjustment / BundleAdjustment /
 * GlobalBundleAdjustemnt, reference include/CeresOptimizer.h:353-387, src/CeresOptimizer.cc:49-342,344-599).
 * Pointer graphs are flattened: a keyframe pose is the reference's 7-vector [t(3), q(x,y,z,w)]
 * (MatEigenConverter.cc:68-77, camera-from-world), a map point is double[3], an observation is
 * (keyframe index, point index, undistorted keypoint u,v as float, inv_level_sigma2 of its octave as float).
 * Intrinsics are the keyframes' float fx, fy, cx, cy (one camera: monocular).  All arithmetic is fp64.
 * ---------------------------------------------------------------------------------------------- */

typedef struct {
  int32_t max_cams;            /* keyframes per problem (constant ones included) */
  int32_t max_points;
  int32_t max_obs;
  int32_t max_pairs_per_obs;   /* capacity of the co-observation lists: sum over points of n_obs(point)^2 / 2 is
                                  bounded by max_obs * max_pairs_per_obs; 0 = default 8 */
  int32_t max_pose_batch;      /* frames per cmos_ba_pose_optimization call with host buffers */
  int32_t max_pose_corr;       /* correspondences per frame (stride) for the same */
  int32_t device;
} cmos_ba_params;

/* What the callers / tests read out of ceres::Solver::Summary. */
typedef struct {
  int32_t iterations;            /* LM iterations performed: successful + unsuccessful + invalid */
  int32_t successful_steps;
  int32_t termination;           /* 0 max iterations, 1 function tolerance, 2 parameter tolerance, 3 gradient
                                    tolerance, 4 stop flag, 5 failure, 6 minimum trust-region radius */
  int32_t jacobian_evaluations;
  double initial_cost, final_cost;
} cmos_ba_summary;

typedef struct cmos_ba* cmos_ba_t;

int cmos_ba_create(const cmos_ba_params* params, cmos_ba_t* out);
int cmos_ba_destroy(cmos_ba_t h);

/* CeresOptimizer::PoseOptimization(Frame*) for a batch of frames (CeresOptimizer.cc:275-342).
 *   pose7       [n_frames][7] in/out: frame->Tcw_ as [t, q_xyzw]; q is normalised on output (:335)
 *   n_corr      [n_frames]: matched keypoints with a map point; fewer than 3 -> n_inliers 0, pose untouched (:330)
 *   xw          [n_frames][stride][3] map_point->GetWorldPos();  uv [..][2], inv_sigma2 [..] of the keypoints
 *   K4          fx, fy, cx, cy (float, widened like MatEigenConverter::MatToMatrix3d)
 *   max_iterations  the reference passes 100 (:300)
 *   is_outlier  [n_frames][stride] out = frame->is_outliers_ after CheckOutliers (:243-269)
 *   n_inliers   [n_frames] out = the function's return value
 * With on_device != 0 every array is a device pointer, no capacity limit applies, and the call does not
 * synchronise; with host pointers it is synchronous. */
int cmos_ba_pose_optimization(cmos_ba_t h, int32_t n_frames, double* pose7, const int32_t* n_corr, const double* xw,
                              const float* uv, const float* inv_sigma2, int32_t stride, const float* K4,
                              int32_t max_iterations, uint8_t* is_outlier, int32_t* n_inliers,
                              cmos_ba_summary* summaries, int32_t on_device, void* stream);

/* Uploads a keyframe / map-point graph and builds its block structure (host pointers; synchronous).
 *   cam_flags [n_cams]: bit0 = constant (a fixed keyframe, or keyframe id 0 — CeresOptimizer.cc:115-120,476-502),
 *                       bit1 = not a local keyframe: its observations are never scanned for outliers (:545). */
int cmos_ba_set_problem(cmos_ba_t h, int32_t n_cams, const double* cams, const uint8_t* cam_flags, int32_t n_points,
                        const double* points, int32_t n_obs, const int32_t* obs_cam, const int32_t* obs_pt,
                        const float* uv, const float* inv_sigma2, const float* K4);

/* Enqueues LocalBundleAdjustment on the uploaded graph, starting from the uploaded values: Huber pass
 * (iterations_pass0, reference 5), outlier scan, pass with the Huber blocks kept and the inliers added again
 * without loss (iterations_pass1, reference 10 — quirk Q2), outlier scan.  Asynchronous on `stream`.
 * stop_flag: host byte polled from the device once per LM iteration (StopFlagCallback) and at the start of
 * each pass, where a raised flag means "return, nothing written" (:509-512); may be NULL. */
int cmos_ba_run_local(cmos_ba_t h, int32_t iterations_pass0, int32_t iterations_pass1, const uint8_t* stop_flag,
                      void* stream);
/* Enqueues BundleAdjustment / GlobalBundleAdjustemnt: one solve of n_iterations, Huber iff robust. */
int cmos_ba_run_global(cmos_ba_t h, int32_t n_iterations, int32_t robust, const uint8_t* stop_flag, void* stream);
/* Waits for the enqueued solve and copies out: cams [n_cams][7], points [n_points][3], erase [n_obs] (LocalBA:
 * observations the reference erases from the map, :573-581; in the caller's observation order), summaries[2]
 * (pass 0, pass 1; global BA fills [0]).  Any pointer may be NULL. */
int cmos_ba_get_results(cmos_ba_t h, double* cams, double* points, uint8_t* erase, cmos_ba_summary* summaries,
                        void* stream);

/* One-shot forms with the reference's defaults (what the C++ adapters call). */
int cmos_ba_local_bundle_adjustment(cmos_ba_t h, int32_t n_cams, double* cams, const uint8_t* cam_flags, int32_t n_points,
                                    double* points, int32_t n_obs, const int32_t* obs_cam, const int32_t* obs_pt,
                                    const float* uv, const float* inv_sigma2, const float* K4, const uint8_t* stop_flag,
                                    uint8_t* erase, cmos_ba_summary* summaries);
int cmos_ba_bundle_adjustment(cmos_ba_t h, int32_t n_cams, double* cams, const uint8_t* cam_const, int32_t n_points,
                              double* points, int32_t n_obs, const int32_t* obs_cam, const int32_t* obs_pt,
                              const float* uv, const float* inv_sigma2, const float* K4, int32_t n_iterations,
                              int32_t robust, const uint8_t* stop_flag, cmos_ba_summary* summary);

/* CeresOptimizer::OptimizeSim3(keyframe_1, keyframe_2, matches12, S12, th2, bFixScale) (CeresOptimizer.cc:601-735;
 * Sim3ErrorTerm / Sim3Parameterization include/CeresOptimizer.h:168-268, CeresOptimizer.cc:24-47).
 *   S12 in/out as scale, rotation (row-major 3x3), translation — Sophus::Sim3d's scale(), rotationMatrix(), translation()
 *   K1, K2          fx, fy, cx, cy of the two keyframes (float, widened)
 *   per correspondence (the reference's n_correspondences loop, :646-693): obs1 = keypoint of keyframe 1, inv_sigma1 =
 *   its inv_level_sigma2, P3D2c = keyframe 2's map point in camera-2 coordinates (R2cw * Xw + t2cw); obs2, inv_sigma2,
 *   P3D1c likewise
 *   th2             squared Huber width (LoopClosing passes 10); max_iterations: the reference uses 100
 *   is_bad [n] out  is_outlier_12 || is_outlier_21 (:702-726);  lie7 out (may be NULL) = the optimised sim12
 *   n_inliers out   the function's return value (0 when fewer than 10 remain); bFixScale is ignored by the reference */
int cmos_ba_optimize_sim3(cmos_ba_t h, int32_t n, double* s12, double* R12, double* t12, const float* K1, const float* K2,
                          const float* obs1, const float* inv_sigma1, const double* P3D2c, const float* obs2,
                          const float* inv_sigma2, const double* P3D1c, float th2, int32_t max_iterations,
                          uint8_t* is_bad, double* lie7, int32_t* n_inliers, cmos_ba_summary* summary);

/* CeresOptimizer::OptimizeEssentialGraph(map, loop_keyframe, current_keyframe, keyframes_non_corrected_sim3,
 * keyframes_corrected_sim3, loop_connections, is_fixed_scale) (CeresOptimizer.cc:736-957; EssentialGraphErrorTerm
 * include/CeresOptimizer.h:270-328).  The adapter flattens the pointer graph in the reference's insertion order:
 *   Scw [n_kf][13]    scale, rotation (row-major 3x3), translation of every keyframe's initial Scw: its entry of
 *                     keyframes_corrected_sim3 where present (:767-769), else (1, Rcw, tcw) (:770-774)
 *   kf_flags [n_kf]   bit 0: constant block = the loop keyframe (:780-783); bit 1: Snc holds the keyframe's entry of
 *                     keyframes_non_corrected_sim3
 *   Snc [n_kf][13]    non-corrected Siw (read only where bit 1 is set)
 *   edges             one per residual block, parameter block 0 = keyframe edge_j, block 1 = keyframe edge_i;
 *                     edge_kind 0 = loop-connection edge, measured between the initial (corrected) values (:797-809);
 *                     edge_kind 1 = spanning-tree, loop or co-visibility edge, each side's non-corrected Sim3 where it has
 *                     one (:822-895)
 *   max_iterations    the reference uses 100 (:747)
 *   points            Xw [n_points][3] + ref_kf [n_points] (corrected_reference_ or the reference keyframe, :939-946)
 * Out (any may be NULL except Xw_out when n_points > 0): lie_out [n_kf][7] optimised logs (upsilon, omega, sigma);
 * Tiw_out [n_kf][16] row-major [R | t / s; 0 0 0 1] for KeyFrame::SetPose (:918-926); Xw_out [n_points][3] corrected
 * positions (:948-953).  Synchronous; host pointers.  is_fixed_scale is ignored by the reference. */
int cmos_ba_optimize_essential_graph(cmos_ba_t h, int32_t n_kf, const double* Scw, const uint8_t* kf_flags, const double* Snc,
                                     int32_t n_edges, const int32_t* edge_j, const int32_t* edge_i, const uint8_t* edge_kind,
                                     int32_t max_iterations, int32_t n_points, const double* Xw, const int32_t* ref_kf,
                                     double* lie_out, double* Tiw_out, double* Xw_out, cmos_ba_summary* summary);

/* Multi-GPU global bundle adjustment (no counterpart in the reference, which is single process).
 * One process per GPU.  Map points — and with them their observations — are partitioned over the ranks;
 * keyframes are replicated.  Each rank uploads ALL keyframes but only ITS points/observations with
 * cmos_ba_set_problem, then every rank calls cmos_ba_run_global with the same arguments.  Per LM iteration the
 * ranks exchange, with NCCL all-reduces enqueued on the solve's stream: the keyframe blocks H_cc/g_c, the partial
 * reduced camera system S + rhs (the one large message), and three short vectors of scalars (cost, norms, step
 * statistics).  Once the communicator exists cmos_ba_set_problem is collective too: one byte-mask all-reduce makes
 * every rank adopt the union of the ranks' co-visibility block patterns, so the per-block sums line up.
 * The reduced system is then factorised redundantly on every rank, so keyframe poses stay
 * replicated bit for bit and the points never leave their rank.  NCCL is bound with dlopen("libnccl.so.2").
 *   cmos_ba_comm_unique_id: rank 0 creates the 128-byte NCCL id; the caller ships it to the other ranks
 *   cmos_ba_comm_init:      collective over all ranks; after it the handle's solves are sharded */
int cmos_ba_comm_unique_id(uint8_t* id128);
int cmos_ba_comm_init(cmos_ba_t h, const uint8_t* id128, int32_t n_ranks, int32_t rank);

/* Verification taps: per-iteration trace rows [iteration][8] = cost, cost_change, gradient_max_norm, step_norm,
 * relative_decrease, trust-region radius, accepted, valid (row 0 = initial evaluation). */
int cmos_ba_debug_trace(cmos_ba_t h, int32_t pass, double* trace, int32_t rows);

/* Test hook for StopFlagCallback (include/CeresOptimizer.h:332-349, src/CeresOptimizer.cc:509-514): in the following solves
 * the DEVICE raises the caller's stop flag itself once `pass` (0 / 1 of LocalBundleAdjustment, 0 of the others) has finished
 * `iteration` iterations — what another thread setting the flag at that moment would do, but reproducible.  -1 disables.
 * Semantics under test: a flag that is up when a LocalBundleAdjustment pass starts returns without writing anything back
 * (:509-512); a flag seen by the callback in the middle of a solve ends it with USER_SUCCESS (the callback returns
 * SOLVER_TERMINATE_SUCCESSFULLY, CeresOptimizer.h:340), so the function carries on with the iterate reached at that moment —
 * the same result as a solve whose max_num_iterations was that iteration. */
int cmos_ba_debug_stop_at(cmos_ba_t h, int32_t pass, int32_t iteration);
int cmos_ba_debug_pose_trace(cmos_ba_t h, int32_t frame, double* trace, int32_t rows);
/* Test tap of the small dense reduced-camera-system solver (what Ceres' DENSE_SCHUR / SPARSE_SCHUR hand to LAPACK / CHOLMOD,
 * src/CeresOptimizer.cc:178-187, 516-519): solves A x = b for a symmetric positive definite A (n x n row-major, n a multiple
 * of 6, <= 228) with exactly the device routines the BA kernels use: variant 0 = factorisation into the 24-block format +
 * 24-row back substitution (GlobalBA's nested-dissection nodes), variant 1 = the whole solve as LocalBA's reduced camera
 * system takes it.  failed = 1 if a pivot was not positive.  cycles2
 * (optional, 2 entries): SM cycles of the factorisation and of the back substitution. */
int cmos_debug_solve_spd(const double* A, const double* b, int32_t n, int32_t variant, double* x, int32_t* failed, int64_t* cycles2);
/* How the last cmos_ba_optimize_essential_graph call solved its normal equations (what Ceres' SPARSE_NORMAL_CHOLESKY does
 * with CHOLMOD, src/CeresOptimizer.cc:897-901): info6 = {1 if nested dissection over a band + border structure was used
 * (0: blocked Cholesky inside the row envelope), keyframes per node, unknowns per node, nodes, border keyframes, border
 * unknowns (padded)}. */
int cmos_ba_debug_essential_graph_plan(cmos_ba_t h, int32_t* info6);
/* Kernels launched by the last cmos_ba_run_* / cmos_ba_pose_optimization call. */
int cmos_ba_last_launch_count(cmos_ba_t h, int32_t* n);
/* Device time of the solves (CUDA events on the launching stream), as cmos_orb_set_profiling. */
int cmos_ba_set_profiling(cmos_ba_t h, int32_t enable);
int cmos_ba_solve_time(cmos_ba_t h, double* ms, int64_t* calls);

#ifdef __cplusplus
}
#endif
#endif /* CMOS_B200_H */
