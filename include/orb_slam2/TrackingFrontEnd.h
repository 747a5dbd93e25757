// Batched tracking front end: what the reference's Tracking thread does per frame — Frame::Frame (ExtractORB,
// AssignFeaturesToGrid; src/Frame.cc:98-156) followed by ORBmatcher::SearchByProjection(current, last, th)
// (src/Tracking.cc:632, src/ORBmatcher.cc:1161-1271) — for a whole batch of host images in one asynchronous call
// (several cameras, offline sequences).  RAII wrapper over cmos_track_* of libcmos_b200.so; header-only like the
// other adapters.  The reference has no class of this shape: the per-frame drop-ins are ORBextractor / ORBmatcher.
#ifndef ORB_SLAM2_CMOS_TRACKINGFRONTEND_H
#define ORB_SLAM2_CMOS_TRACKINGFRONTEND_H

#include "views.h"

namespace ORB_SLAM2 {

class TrackingFrontEnd {
 public:
  // host buffers of one batch: n_frames images in, [n_frames][capacity] keypoints / descriptors / matches out
  struct Batch {
    const uint8_t* images = nullptr;      // frame f at images + f * frame_stride, rows `pitch` bytes apart
    int64_t frame_stride = 0;
    int32_t pitch = 0, width = 0, height = 0, n_frames = 0;
    const double* Tcw = nullptr;          // [n_frames][16] row-major camera-from-world of the CURRENT frames
    cmos_keypoint* keypoints = nullptr;   // out
    uint8_t* descriptors = nullptr;       // out, 32 bytes per keypoint
    int32_t* counts = nullptr;            // out [n_frames]
    int32_t capacity = 0;                 // >= capacity()
    int32_t* match = nullptr;             // out: match[f][k] = last-frame keypoint index or -1
    int32_t* nmatches = nullptr;          // out [n_frames]
  };

  // lanes x chunk_frames: 8 x 32 with four batches in flight runs at the rate of the kernels on one B200
  TrackingFrontEnd(const cmos_camera& camera, int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST,
                   int max_width = 1241, int max_height = 376, int lanes = 8, int chunk_frames = 32, int device = 0) {
    cmos_track_params p = {{nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, max_width, max_height, chunk_frames, device},
                           lanes, chunk_frames};
    cmos_throw_if(cmos_track_create(&p, &camera, &h_), "TrackingFrontEnd");
    cmos_track_keypoint_capacity(h_, &capacity_);
  }
  ~TrackingFrontEnd() { cmos_track_destroy(h_); }
  TrackingFrontEnd(const TrackingFrontEnd&) = delete;
  TrackingFrontEnd& operator=(const TrackingFrontEnd&) = delete;

  int32_t capacity() const { return capacity_; }        // keypoint slots per frame of the output arrays

  // Last-frame inputs as per-keypoint arrays [n_frames][last_stride] (what LastFrameView holds for one frame).
  // Returns a ticket at once; every buffer must stay valid until Wait(ticket).  Up to four batches may be in flight.
  int64_t Submit(const Batch& b, const cmos_keypoint* last_keypoints, const int32_t* last_counts, const uint8_t* last_flags,
                 const double* last_world_pos, const uint8_t* last_descriptors, int32_t last_stride, float th,
                 bool checkOri = true) {
    int64_t t = -1;
    cmos_throw_if(cmos_track_submit(h_, b.images, b.frame_stride, b.pitch, b.width, b.height, b.n_frames, b.Tcw, last_keypoints,
                                    last_counts, last_flags, last_world_pos, last_descriptors, last_stride, th, checkOri ? 1 : 0,
                                    b.keypoints, b.descriptors, b.counts, b.capacity, b.match, b.nmatches, &t),
                  "TrackingFrontEnd::Submit");
    return t;
  }
  // ... as one 64-byte record per last-frame keypoint that carries a usable map point (records of frame f:
  // points[point_start[f] .. point_start[f + 1]), increasing keypoint index)
  int64_t SubmitPoints(const Batch& b, const cmos_last_point* points, const int32_t* point_start, float th, bool checkOri = true) {
    int64_t t = -1;
    cmos_throw_if(cmos_track_submit_points(h_, b.images, b.frame_stride, b.pitch, b.width, b.height, b.n_frames, b.Tcw, points,
                                           point_start, th, checkOri ? 1 : 0, b.keypoints, b.descriptors, b.counts, b.capacity,
                                           b.match, b.nmatches, &t),
                  "TrackingFrontEnd::SubmitPoints");
    return t;
  }
  // ... as 12-byte association records into the device-resident map-point table below
  int64_t SubmitMap(const Batch& b, const cmos_track_assoc* assoc, const int32_t* assoc_start, float th, bool checkOri = true) {
    int64_t t = -1;
    cmos_throw_if(cmos_track_submit_map(h_, b.images, b.frame_stride, b.pitch, b.width, b.height, b.n_frames, b.Tcw, assoc,
                                        assoc_start, th, checkOri ? 1 : 0, b.keypoints, b.descriptors, b.counts, b.capacity,
                                        b.match, b.nmatches, &t),
                  "TrackingFrontEnd::SubmitMap");
    return t;
  }
  void Wait(int64_t ticket) { cmos_throw_if(cmos_track_wait(h_, ticket), "TrackingFrontEnd::Wait"); }

  // Device-resident mirror of MapPoint::GetWorldPos() / GetDescriptor(), one slot per map point (the caller's numbering).
  // Update when the map changes (MapPoint::SetWorldPos, ComputeDistinctiveDescriptors); both calls wait for the batches in flight.
  void ReserveMapPoints(int32_t n_slots) { cmos_throw_if(cmos_track_map_reserve(h_, n_slots), "TrackingFrontEnd::ReserveMapPoints"); }
  void UpdateMapPoints(int32_t n, const int32_t* slots, const double* world_pos, const uint8_t* descriptors) {
    cmos_throw_if(cmos_track_map_update(h_, n, slots, 0, world_pos, descriptors), "TrackingFrontEnd::UpdateMapPoints");
  }
  void UpdateMapPointRange(int32_t first_slot, int32_t n, const double* world_pos, const uint8_t* descriptors) {
    cmos_throw_if(cmos_track_map_update(h_, n, nullptr, first_slot, world_pos, descriptors), "TrackingFrontEnd::UpdateMapPointRange");
  }

  cmos_track_t handle() { return h_; }

 private:
  cmos_track_t h_ = nullptr;
  int32_t capacity_ = 0;
};

}  // namespace ORB_SLAM2
#endif
