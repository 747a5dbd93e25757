// Tracking front-end: Frame construction + the motion-model search, pipelined over CUDA streams.
//
// What the reference does per frame on the Tracking thread (GrabImageMonocular -> Frame::Frame, src/Frame.cc:98-156:
// ExtractORB :116 -> ORBextractor::operator() :175-177, AssignFeaturesToGrid :155; then TrackWithMotionModel ->
// ORBmatcher::SearchByProjection(current_frame_, last_frame_, th), src/Tracking.cc:632, src/ORBmatcher.cc:1161-1271) is one call here for a
// batch of frames with HOST buffers.  The batch is cut into chunks that are dealt round-robin — across batches — to the lanes
// (one compute stream + one upload stream, one extractor handle, one matcher handle per lane), so the host->device copies of
// one chunk, the kernels of another and the device->host copy of a third overlap; cmos_track_submit / cmos_track_wait keep up
// to four batches in flight, which is what keeps the copy engine busy between batches.  The last-frame inputs come in three
// forms with identical results: per-keypoint arrays, packed 64-byte records, or 12-byte association records into a
// device-resident map-point table (cmos_track_map_reserve / _update).  Only public entry points of the same C ABI are used.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <vector>

#include "cmos_common.h"

using namespace cmos;

namespace {
struct Lane {
  cudaStream_t st = nullptr;
  // The last-frame inputs upload on their own stream BEHIND the chunk's images (the extraction needs only the images; a kernel or
  // a copy queued ahead of the image upload on the lane's stream delays everything): ev_up = inputs of this chunk are on the device,
  // ev_done = the lane's previous chunk no longer reads the input buffers.
  cudaStream_t st_up = nullptr;
  cudaEvent_t ev_up = nullptr, ev_done = nullptr;
  bool used = false;
  cmos_orb_t orb = nullptr;
  cmos_match_t match = nullptr;
  double *d_T = nullptr, *d_xw = nullptr;
  cmos_keypoint* d_last_kps = nullptr;
  int *d_last_counts = nullptr, *d_match = nullptr, *d_nm = nullptr;
  uint8_t *d_flags = nullptr, *d_last_desc = nullptr;
  // compact last-frame input (cmos_track_submit_points): packed records, per-frame offsets, record -> last-frame index
  cmos_last_point* d_points = nullptr;   // cmos_track_submit_map stages its 12-byte records in the same buffer
  int *d_point_start = nullptr, *d_pidx = nullptr;
};

static_assert(sizeof(cmos_last_point) == 64, "cmos_last_point is a 64-byte record");
// One record per usable last-frame map point -> the per-keypoint arrays the search kernels read (frame f of the chunk,
// record j): octave / angle of the keypoint, flags, world position, descriptor; pidx keeps the keypoint's own index.
__global__ void __launch_bounds__(256) k_unpack_last(const cmos_last_point* __restrict__ pts, const int* __restrict__ start, int base,
                                                     int stride, cmos_keypoint* __restrict__ kps, int* __restrict__ counts,
                                                     uint8_t* __restrict__ flags, double* __restrict__ xw, uint8_t* __restrict__ desc,
                                                     int* __restrict__ pidx) {
  const int f = blockIdx.y, j = blockIdx.x * 256 + threadIdx.x;
  const int s0 = start[f] - base, n = min(start[f + 1] - start[f], stride);
  if (j == 0) counts[f] = n;
  if (j >= n) return;
  const cmos_last_point& r = pts[s0 + j];
  const size_t o = (size_t)f * stride + j;
  cmos_keypoint kp;
  kp.x = 0.f; kp.y = 0.f; kp.size = 0.f; kp.angle = r.angle; kp.response = 0.f; kp.octave = r.octave; kp.class_id = -1;
  kps[o] = kp;
  flags[o] = r.flags;
  xw[3 * o] = r.xw[0]; xw[3 * o + 1] = r.xw[1]; xw[3 * o + 2] = r.xw[2];
  const uint4* d4 = (const uint4*)r.descriptor;
  ((uint4*)(desc + 32 * o))[0] = d4[0]; ((uint4*)(desc + 32 * o))[1] = d4[1];
  pidx[o] = r.index;
}
static_assert(sizeof(cmos_track_assoc) == 12, "cmos_track_assoc is a 12-byte record");
// cmos_track_submit_map: the same per-keypoint arrays from association records; world position and descriptor are gathered
// out of the device-resident map-point table.  A record whose slot lies outside the table leaves an unusable entry (flags 0).
__global__ void __launch_bounds__(256) k_unpack_assoc(const cmos_track_assoc* __restrict__ recs, const int* __restrict__ start, int base,
                                                      int stride, const double* __restrict__ map_xw, const uint8_t* __restrict__ map_desc,
                                                      int n_slots, cmos_keypoint* __restrict__ kps, int* __restrict__ counts,
                                                      uint8_t* __restrict__ flags, double* __restrict__ xw, uint8_t* __restrict__ desc,
                                                      int* __restrict__ pidx) {
  const int f = blockIdx.y, j = blockIdx.x * 256 + threadIdx.x;
  const int s0 = start[f] - base, n = min(start[f + 1] - start[f], stride);
  if (j == 0) counts[f] = n;
  if (j >= n) return;
  const cmos_track_assoc r = recs[s0 + j];
  const size_t o = (size_t)f * stride + j;
  const bool ok = (unsigned)r.slot < (unsigned)n_slots;
  cmos_keypoint kp;
  kp.x = 0.f; kp.y = 0.f; kp.size = 0.f; kp.angle = r.angle; kp.response = 0.f; kp.octave = r.octave; kp.class_id = -1;
  kps[o] = kp;
  flags[o] = ok ? r.flags : (uint8_t)0;
  pidx[o] = r.index;
  if (!ok) return;
  const double* X = map_xw + 3 * (size_t)r.slot;
  xw[3 * o] = X[0]; xw[3 * o + 1] = X[1]; xw[3 * o + 2] = X[2];
  const uint4* d4 = (const uint4*)(map_desc + 32 * (size_t)r.slot);
  ((uint4*)(desc + 32 * o))[0] = __ldg(d4); ((uint4*)(desc + 32 * o))[1] = __ldg(d4 + 1);
}
// slots[i] <- (xw[i], desc[i]) of an incremental map update
__global__ void __launch_bounds__(256) k_map_scatter(const int* __restrict__ slots, const double* __restrict__ xw_in,
                                                     const uint8_t* __restrict__ desc_in, int n, int n_slots,
                                                     double* __restrict__ map_xw, uint8_t* __restrict__ map_desc) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const int s = slots[i];
  if ((unsigned)s >= (unsigned)n_slots) return;
  map_xw[3 * (size_t)s] = xw_in[3 * (size_t)i]; map_xw[3 * (size_t)s + 1] = xw_in[3 * (size_t)i + 1];
  map_xw[3 * (size_t)s + 2] = xw_in[3 * (size_t)i + 2];
  const uint4* d4 = (const uint4*)(desc_in + 32 * (size_t)i);
  ((uint4*)(map_desc + 32 * (size_t)s))[0] = d4[0]; ((uint4*)(map_desc + 32 * (size_t)s))[1] = d4[1];
}
// match holds record numbers: back to last-frame keypoint indices
__global__ void __launch_bounds__(256) k_remap_match(int* __restrict__ match, const int* __restrict__ pidx, int stride, int n_total) {
  const int t = blockIdx.x * 256 + threadIdx.x;
  if (t >= n_total) return;
  const int m = match[t];
  if (m >= 0) match[t] = pidx[(size_t)(t / stride) * stride + m];
}


constexpr int kTrackSlots = 4;        // batches that may be in flight between cmos_track_submit and cmos_track_wait
struct Slot {
  bool busy = false;
  int64_t ticket = -1;
  std::vector<cudaEvent_t> done;      // one event per lane, recorded behind the batch's last copy on that lane
  int rc = 0;
  int32_t* match = nullptr;           // for the tail padding done at wait()
  int n_frames = 0, capacity = 0;
};
}  // namespace

struct cmos_track {
  cmos_track_params p{};
  cmos_camera cam{};
  int kp_cap = 0;
  std::vector<Lane> lanes;
  int launches = 0;
  Slot slots[kTrackSlots];
  int64_t next_ticket = 0;
  uint64_t next_chunk = 0;            // chunks are dealt to the lanes round-robin ACROSS batches
  // device-resident map points (cmos_track_map_reserve / _update, read by cmos_track_submit_map)
  double* d_map_xw = nullptr;
  uint8_t* d_map_desc = nullptr;
  int map_slots = 0;
};

extern "C" {

int cmos_track_destroy(cmos_track_t h) {
  if (!h) return CMOS_OK;
  cudaSetDevice(h->p.orb.device);
  for (Slot& sl : h->slots)
    for (cudaEvent_t e : sl.done)
      if (e) cudaEventDestroy(e);
  if (h->d_map_xw) cudaFree(h->d_map_xw);
  if (h->d_map_desc) cudaFree(h->d_map_desc);
  for (Lane& L : h->lanes) {
    if (L.orb) cmos_orb_destroy(L.orb);
    if (L.match) cmos_match_destroy(L.match);
    void* bufs[] = {L.d_T, L.d_xw, L.d_last_kps, L.d_last_counts, L.d_match, L.d_nm, L.d_flags, L.d_last_desc, L.d_points, L.d_point_start,
                    L.d_pidx};
    for (void* b : bufs)
      if (b) cudaFree(b);
    if (L.ev_up) cudaEventDestroy(L.ev_up);
    if (L.ev_done) cudaEventDestroy(L.ev_done);
    if (L.st_up) cudaStreamDestroy(L.st_up);
    if (L.st) cudaStreamDestroy(L.st);
  }
  delete h;
  return CMOS_OK;
}

int cmos_track_create(const cmos_track_params* params, const cmos_camera* cam, cmos_track_t* out) {
  CMOS_REQUIRE(params && cam && out, "null argument");
  CMOS_REQUIRE(params->lanes >= 1 && params->lanes <= 16 && params->chunk_frames >= 1, "lanes must be 1..16, chunk_frames >= 1");
  cmos_track* h = new cmos_track();
  h->p = *params;
  h->cam = *cam;
  h->lanes.resize(params->lanes);
  cmos_orb_params op = params->orb;
  op.max_batch = params->chunk_frames;
  int rc = CMOS_OK;
  for (Lane& L : h->lanes) {
    if ((rc = cmos_orb_create(&op, &L.orb))) break;
    if ((rc = cmos_orb_keypoint_capacity(L.orb, &h->kp_cap))) break;
    cmos_match_params mp{params->chunk_frames, h->kp_cap, 1, op.device};
    if ((rc = cmos_match_create(&mp, &L.match))) break;
    if (cudaStreamCreateWithFlags(&L.st, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&L.st_up, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&L.ev_up, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&L.ev_done, cudaEventDisableTiming) != cudaSuccess) {
      set_error("stream / event creation failed");
      rc = CMOS_ERR_CUDA;
      break;
    }
    cudaError_t err = cudaSuccess;
    const size_t n = (size_t)params->chunk_frames * h->kp_cap;
    L.d_T = dev_alloc<double>((size_t)params->chunk_frames * 16, &err);
    L.d_xw = dev_alloc<double>(n * 3, &err);
    L.d_last_kps = dev_alloc<cmos_keypoint>(n, &err);
    L.d_last_counts = dev_alloc<int>(params->chunk_frames, &err);
    L.d_match = dev_alloc<int>(n, &err);
    L.d_nm = dev_alloc<int>(params->chunk_frames, &err);
    L.d_flags = dev_alloc<uint8_t>(n, &err);
    L.d_last_desc = dev_alloc<uint8_t>(n * 32, &err);
    L.d_points = dev_alloc<cmos_last_point>(n, &err);
    L.d_point_start = dev_alloc<int>(params->chunk_frames + 1, &err);
    L.d_pidx = dev_alloc<int>(n, &err);
    if (err != cudaSuccess) { set_error("device allocation failed: %s", cudaGetErrorString(err)); rc = CMOS_ERR_CUDA; break; }
  }
  for (Slot& sl : h->slots) {
    sl.done.assign(h->lanes.size(), nullptr);
    for (cudaEvent_t& e : sl.done)
      if (!rc && cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) rc = CMOS_ERR_CUDA;
  }
  if (rc) { cmos_track_destroy(h); return rc; }
  // the lanes' extractor handles clear their buffers with cudaMemset at creation: asynchronous on the legacy stream, which is
  // not ordered with the lanes' non-blocking streams — wait once here
  if (cudaDeviceSynchronize() != cudaSuccess) { set_error("cudaDeviceSynchronize failed"); cmos_track_destroy(h); return CMOS_ERR_CUDA; }
  *out = h;
  return CMOS_OK;
}

int cmos_track_keypoint_capacity(cmos_track_t h, int32_t* cap) {
  CMOS_REQUIRE(h && cap, "null argument");
  *cap = h->kp_cap;
  return CMOS_OK;
}

}  // extern "C"

// points != nullptr: the last-frame inputs come as packed records (cmos_track_submit_points); assoc != nullptr: as association
// records into the device-resident map-point table (cmos_track_submit_map); otherwise as per-keypoint arrays
static int submit_impl(cmos_track_t h, const uint8_t* images, int64_t frame_stride, int32_t pitch, int32_t width,
                       int32_t height, int32_t n_frames, const double* Tcw, const cmos_keypoint* last_keypoints,
                       const int32_t* last_counts, const uint8_t* last_flags, const double* last_xw,
                       const uint8_t* last_descriptors, int32_t last_stride, const cmos_last_point* points,
                       const cmos_track_assoc* assoc, const int32_t* point_start, float th, int32_t check_orientation,
                       cmos_keypoint* keypoints, uint8_t* descriptors, int32_t* counts, int32_t capacity, int32_t* match,
                       int32_t* nmatches, int64_t* ticket) {
  CMOS_REQUIRE(h && images && Tcw && keypoints && descriptors && counts && match && nmatches && ticket, "null argument");
  CMOS_REQUIRE((points || assoc) ? point_start != nullptr
                                 : (last_keypoints && last_counts && last_flags && last_xw && last_descriptors), "null argument");
  CMOS_REQUIRE(!assoc || h->map_slots > 0, "cmos_track_submit_map needs a map-point table: call cmos_track_map_reserve / _update first");
  const bool records = points || assoc;
  CMOS_REQUIRE(n_frames >= 1, "n_frames must be positive");
  CMOS_REQUIRE(capacity >= h->kp_cap, "capacity %d < cmos_track_keypoint_capacity %d", capacity, h->kp_cap);
  if (records) {
    CMOS_REQUIRE(h->kp_cap <= 65536, "keypoint capacity %d does not fit the 16-bit record index", h->kp_cap);
    last_stride = h->kp_cap;
    for (int f = 0; f < n_frames; f++)
      CMOS_REQUIRE(point_start[f + 1] >= point_start[f] && point_start[f + 1] - point_start[f] <= h->kp_cap,
                   "frame %d: %d records outside 0..%d", f, point_start[f + 1] - point_start[f], h->kp_cap);
  }
  CMOS_REQUIRE(last_stride >= 1 && last_stride <= h->kp_cap, "last_stride %d outside 1..%d", last_stride, h->kp_cap);
  CMOS_CUDA_OK(cudaSetDevice(h->p.orb.device));
  Slot* slot = nullptr;
  for (Slot& sl : h->slots)
    if (!sl.busy) { slot = &sl; break; }
  if (!slot) { set_error("%d batches are already in flight: call cmos_track_wait first", kTrackSlots); return CMOS_ERR_STATE; }
  NvtxRange nvtx_submit("cmos.track.submit");
  const int cf = h->p.chunk_frames, nl = (int)h->lanes.size();
  int launches = 0, rc = CMOS_OK;
  // inside the loop a CUDA error must not return: copies into caller memory may be in flight on other lanes (drained below)
#define TRACK_CUDA(call)                                                                      \
  {                                                                                           \
    const cudaError_t e_ = (call);                                                            \
    if (e_ != cudaSuccess) {                                                                  \
      cmos::set_error("%s failed: %s", #call, cudaGetErrorString(e_));                        \
      rc = CMOS_ERR_CUDA;                                                                     \
      break;                                                                                  \
    }                                                                                         \
  }
  // The lane of a chunk continues round-robin from the previous batch: with more lanes than chunks per batch the chunks
  // of batch k + 1 upload on lanes that are idle while the lanes of batch k compute (inside a lane upload, kernels and
  // download are stream-ordered, so a lane is busy for upload + compute + download of its chunk).
  for (int f0 = 0; f0 < n_frames && !rc; f0 += cf) {
    Lane& L = h->lanes[(size_t)(h->next_chunk++ % (uint64_t)nl)];
    const int n = std::min(cf, n_frames - f0);
    const size_t nq = (size_t)n * last_stride, o = (size_t)f0 * last_stride;
    // images first (cmos_orb_extract_async: upload, kernels, download on the lane's stream) ...
    if ((rc = cmos_orb_extract_async(L.orb, images + (size_t)f0 * frame_stride, frame_stride, pitch, width, height, n,
                                     keypoints + (size_t)f0 * capacity, descriptors + (size_t)f0 * capacity * 32,
                                     counts + f0, capacity, L.st))) break;
    // ... then the inputs of the search on the upload stream, once the lane's previous chunk has stopped reading the buffers
    if (L.used) TRACK_CUDA(cudaStreamWaitEvent(L.st_up, L.ev_done, 0));
    TRACK_CUDA(cudaMemcpyAsync(L.d_T, Tcw + (size_t)f0 * 16, (size_t)n * 16 * sizeof(double), cudaMemcpyHostToDevice, L.st_up));
    const int base = records ? point_start[f0] : 0, cnt = records ? point_start[f0 + n] - base : 0;
    if (records) {
      // packed records of this chunk's frames: ONE copy of exactly the bytes that matter, unpacked on the device below
      TRACK_CUDA(cudaMemcpyAsync(L.d_point_start, point_start + f0, (size_t)(n + 1) * sizeof(int), cudaMemcpyHostToDevice, L.st_up));
      if (cnt > 0) {
        if (points) {
          TRACK_CUDA(cudaMemcpyAsync(L.d_points, points + base, (size_t)cnt * sizeof(cmos_last_point), cudaMemcpyHostToDevice, L.st_up));
        } else {
          TRACK_CUDA(cudaMemcpyAsync(L.d_points, assoc + base, (size_t)cnt * sizeof(cmos_track_assoc), cudaMemcpyHostToDevice, L.st_up));
        }
      }
    } else {
      TRACK_CUDA(cudaMemcpyAsync(L.d_last_kps, last_keypoints + o, nq * sizeof(cmos_keypoint), cudaMemcpyHostToDevice, L.st_up));
      TRACK_CUDA(cudaMemcpyAsync(L.d_last_counts, last_counts + f0, (size_t)n * sizeof(int), cudaMemcpyHostToDevice, L.st_up));
      TRACK_CUDA(cudaMemcpyAsync(L.d_flags, last_flags + o, nq, cudaMemcpyHostToDevice, L.st_up));
      TRACK_CUDA(cudaMemcpyAsync(L.d_xw, last_xw + o * 3, nq * 3 * sizeof(double), cudaMemcpyHostToDevice, L.st_up));
      TRACK_CUDA(cudaMemcpyAsync(L.d_last_desc, last_descriptors + o * 32, nq * 32, cudaMemcpyHostToDevice, L.st_up));
    }
    TRACK_CUDA(cudaEventRecord(L.ev_up, L.st_up));
    TRACK_CUDA(cudaStreamWaitEvent(L.st, L.ev_up, 0));
    if (records) {
      const dim3 ug((h->kp_cap + 255) / 256, n);
      if (points)
        k_unpack_last<<<ug, 256, 0, L.st>>>(L.d_points, L.d_point_start, base, last_stride, L.d_last_kps, L.d_last_counts, L.d_flags,
                                            L.d_xw, L.d_last_desc, L.d_pidx);
      else
        k_unpack_assoc<<<ug, 256, 0, L.st>>>((const cmos_track_assoc*)L.d_points, L.d_point_start, base, last_stride, h->d_map_xw,
                                             h->d_map_desc, h->map_slots, L.d_last_kps, L.d_last_counts, L.d_flags, L.d_xw,
                                             L.d_last_desc, L.d_pidx);
      launches++;
    }
    cmos_keypoint* d_kps; uint8_t* d_desc; int32_t* d_counts;
    if ((rc = cmos_orb_device_results(L.orb, &d_kps, &d_desc, &d_counts, nullptr, nullptr))) break;
    if ((rc = cmos_match_set_frames(L.match, &h->cam, d_kps, d_desc, d_counts, n, h->kp_cap, 1, L.st))) break;
    if ((rc = cmos_match_search_by_projection_frame(L.match, L.d_T, L.d_last_kps, L.d_last_counts, L.d_flags, L.d_xw,
                                                    L.d_last_desc, last_stride, th, check_orientation, nullptr, L.d_match,
                                                    L.d_nm, 1, L.st))) break;
    if (records) {
      const int tot = n * h->kp_cap;
      k_remap_match<<<(tot + 255) / 256, 256, 0, L.st>>>(L.d_match, L.d_pidx, h->kp_cap, tot);
      launches++;
    }
    const size_t row = (size_t)h->kp_cap * sizeof(int);
    if (capacity == h->kp_cap) {
      TRACK_CUDA(cudaMemcpyAsync(match + (size_t)f0 * capacity, L.d_match, row * n, cudaMemcpyDeviceToHost, L.st));
    } else {
      TRACK_CUDA(cudaMemcpy2DAsync(match + (size_t)f0 * capacity, (size_t)capacity * sizeof(int), L.d_match, row, row, n,
                                   cudaMemcpyDeviceToHost, L.st));
    }
    TRACK_CUDA(cudaMemcpyAsync(nmatches + f0, L.d_nm, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, L.st));
    TRACK_CUDA(cudaEventRecord(L.ev_done, L.st));
    L.used = true;
    int a = 0, b = 0;
    cmos_orb_last_launch_count(L.orb, &a);
    cmos_match_last_launch_count(L.match, &b);
    launches += a + 1 + b;   // + the grid kernel of set_frames
  }
#undef TRACK_CUDA
  h->launches = launches;
  if (rc) {   // drain every lane after an error, so no copy into caller memory is still in flight on return
    for (Lane& L : h->lanes) { cudaStreamSynchronize(L.st_up); cmos_orb_finish(L.orb, L.st); }
    return rc;
  }
  for (size_t l = 0; l < h->lanes.size(); l++) {
    if (cudaEventRecord(slot->done[l], h->lanes[l].st) != cudaSuccess) {
      for (Lane& L : h->lanes) { cudaStreamSynchronize(L.st_up); cmos_orb_finish(L.orb, L.st); }
      set_error("cudaEventRecord failed");
      return CMOS_ERR_CUDA;
    }
  }
  slot->busy = true; slot->ticket = h->next_ticket++; slot->match = match; slot->n_frames = n_frames; slot->capacity = capacity;
  *ticket = slot->ticket;
  return CMOS_OK;
}

extern "C" {

int cmos_track_submit(cmos_track_t h, const uint8_t* images, int64_t frame_stride, int32_t pitch, int32_t width,
                      int32_t height, int32_t n_frames, const double* Tcw, const cmos_keypoint* last_keypoints,
                      const int32_t* last_counts, const uint8_t* last_flags, const double* last_xw,
                      const uint8_t* last_descriptors, int32_t last_stride, float th, int32_t check_orientation,
                      cmos_keypoint* keypoints, uint8_t* descriptors, int32_t* counts, int32_t capacity, int32_t* match,
                      int32_t* nmatches, int64_t* ticket) {
  CMOS_REQUIRE(last_keypoints && last_counts && last_flags && last_xw && last_descriptors, "null argument");
  return submit_impl(h, images, frame_stride, pitch, width, height, n_frames, Tcw, last_keypoints, last_counts, last_flags, last_xw,
                     last_descriptors, last_stride, nullptr, nullptr, nullptr, th, check_orientation, keypoints, descriptors, counts, capacity,
                     match, nmatches, ticket);
}

int cmos_track_submit_points(cmos_track_t h, const uint8_t* images, int64_t frame_stride, int32_t pitch, int32_t width,
                             int32_t height, int32_t n_frames, const double* Tcw, const cmos_last_point* points,
                             const int32_t* point_start, float th, int32_t check_orientation, cmos_keypoint* keypoints,
                             uint8_t* descriptors, int32_t* counts, int32_t capacity, int32_t* match, int32_t* nmatches,
                             int64_t* ticket) {
  CMOS_REQUIRE(points && point_start, "null argument");
  return submit_impl(h, images, frame_stride, pitch, width, height, n_frames, Tcw, nullptr, nullptr, nullptr, nullptr, nullptr, 0,
                     points, nullptr, point_start, th, check_orientation, keypoints, descriptors, counts, capacity, match, nmatches,
                     ticket);
}

int cmos_track_submit_map(cmos_track_t h, const uint8_t* images, int64_t frame_stride, int32_t pitch, int32_t width,
                          int32_t height, int32_t n_frames, const double* Tcw, const cmos_track_assoc* assoc,
                          const int32_t* assoc_start, float th, int32_t check_orientation, cmos_keypoint* keypoints,
                          uint8_t* descriptors, int32_t* counts, int32_t capacity, int32_t* match, int32_t* nmatches,
                          int64_t* ticket) {
  CMOS_REQUIRE(assoc && assoc_start, "null argument");
  return submit_impl(h, images, frame_stride, pitch, width, height, n_frames, Tcw, nullptr, nullptr, nullptr, nullptr, nullptr, 0,
                     nullptr, assoc, assoc_start, th, check_orientation, keypoints, descriptors, counts, capacity, match, nmatches,
                     ticket);
}

// Every lane is drained first: a batch in flight reads either the old table or the new one, never a mixture.
static int map_quiesce(cmos_track_t h) {
  CMOS_CUDA_OK(cudaSetDevice(h->p.orb.device));
  for (Lane& L : h->lanes) {
    CMOS_CUDA_OK(cudaStreamSynchronize(L.st_up));
    CMOS_CUDA_OK(cudaStreamSynchronize(L.st));
  }
  return CMOS_OK;
}

int cmos_track_map_reserve(cmos_track_t h, int32_t n_slots) {
  CMOS_REQUIRE(h, "null handle");
  CMOS_REQUIRE(n_slots >= 1, "n_slots must be positive");
  if (n_slots <= h->map_slots) return CMOS_OK;
  int rc = map_quiesce(h);
  if (rc) return rc;
  // on the first lane's stream (idle after the quiesce) and waited for there: a cudaMemset on the legacy stream is asynchronous
  // and NOT ordered with the lanes' non-blocking streams, which read and update the table
  cudaStream_t st = h->lanes[0].st;
  cudaError_t err = cudaSuccess;
  double* xw = dev_alloc<double>((size_t)n_slots * 3, &err);
  uint8_t* desc = dev_alloc<uint8_t>((size_t)n_slots * 32, &err);
  if (err == cudaSuccess) err = cudaMemsetAsync(xw, 0, (size_t)n_slots * 3 * sizeof(double), st);
  if (err == cudaSuccess) err = cudaMemsetAsync(desc, 0, (size_t)n_slots * 32, st);
  if (err == cudaSuccess && h->map_slots > 0) {     // growing keeps the slots already written
    err = cudaMemcpyAsync(xw, h->d_map_xw, (size_t)h->map_slots * 3 * sizeof(double), cudaMemcpyDeviceToDevice, st);
    if (err == cudaSuccess) err = cudaMemcpyAsync(desc, h->d_map_desc, (size_t)h->map_slots * 32, cudaMemcpyDeviceToDevice, st);
  }
  {
    const cudaError_t e2 = cudaStreamSynchronize(st);
    if (err == cudaSuccess) err = e2;
  }
  if (err != cudaSuccess) {
    if (xw) cudaFree(xw);
    if (desc) cudaFree(desc);
    set_error("map-point table of %d slots: %s", n_slots, cudaGetErrorString(err));
    return CMOS_ERR_CUDA;
  }
  if (h->d_map_xw) cudaFree(h->d_map_xw);
  if (h->d_map_desc) cudaFree(h->d_map_desc);
  h->d_map_xw = xw; h->d_map_desc = desc; h->map_slots = n_slots;
  return CMOS_OK;
}

int cmos_track_map_update(cmos_track_t h, int32_t n, const int32_t* slots, int32_t first_slot, const double* xw,
                          const uint8_t* descriptors) {
  CMOS_REQUIRE(h && xw && descriptors, "null argument");
  CMOS_REQUIRE(n >= 0, "n must not be negative");
  CMOS_REQUIRE(h->map_slots > 0, "no map-point table: call cmos_track_map_reserve first");
  if (n == 0) return CMOS_OK;
  if (slots) {
    for (int i = 0; i < n; i++)
      CMOS_REQUIRE(slots[i] >= 0 && slots[i] < h->map_slots, "slots[%d] = %d outside the table of %d slots", i, slots[i], h->map_slots);
  } else {
    CMOS_REQUIRE(first_slot >= 0 && (int64_t)first_slot + n <= h->map_slots, "slots %d..%lld outside the table of %d slots", first_slot,
                 (long long)first_slot + n - 1, h->map_slots);
  }
  int rc = map_quiesce(h);
  if (rc) return rc;
  // everything below runs on the first lane's stream (idle after the quiesce) and is waited for there: no device-wide
  // synchronisation and no legacy-stream copy, so an optimiser thread working on its own stream is not stalled by a map update
  cudaStream_t st = h->lanes[0].st;
  if (!slots) {     // a run of consecutive slots: two plain copies
    CMOS_CUDA_OK(cudaMemcpyAsync(h->d_map_xw + 3 * (size_t)first_slot, xw, (size_t)n * 3 * sizeof(double), cudaMemcpyHostToDevice, st));
    CMOS_CUDA_OK(cudaMemcpyAsync(h->d_map_desc + 32 * (size_t)first_slot, descriptors, (size_t)n * 32, cudaMemcpyHostToDevice, st));
    CMOS_CUDA_OK(cudaStreamSynchronize(st));
    return CMOS_OK;
  }
  cudaError_t err = cudaSuccess;
  int* d_slots = dev_alloc<int>((size_t)n, &err);
  double* d_xw = dev_alloc<double>((size_t)n * 3, &err);
  uint8_t* d_desc = dev_alloc<uint8_t>((size_t)n * 32, &err);
  if (err == cudaSuccess) err = cudaMemcpyAsync(d_slots, slots, (size_t)n * sizeof(int), cudaMemcpyHostToDevice, st);
  if (err == cudaSuccess) err = cudaMemcpyAsync(d_xw, xw, (size_t)n * 3 * sizeof(double), cudaMemcpyHostToDevice, st);
  if (err == cudaSuccess) err = cudaMemcpyAsync(d_desc, descriptors, (size_t)n * 32, cudaMemcpyHostToDevice, st);
  if (err == cudaSuccess) {
    k_map_scatter<<<(n + 255) / 256, 256, 0, st>>>(d_slots, d_xw, d_desc, n, h->map_slots, h->d_map_xw, h->d_map_desc);
    err = cudaGetLastError();
  }
  {
    const cudaError_t e2 = cudaStreamSynchronize(st);      // also before the frees below when something failed
    if (err == cudaSuccess) err = e2;
  }
  if (d_slots) cudaFree(d_slots);
  if (d_xw) cudaFree(d_xw);
  if (d_desc) cudaFree(d_desc);
  if (err != cudaSuccess) { set_error("map update failed: %s", cudaGetErrorString(err)); return CMOS_ERR_CUDA; }
  return CMOS_OK;
}

int cmos_track_wait(cmos_track_t h, int64_t ticket) {
  CMOS_REQUIRE(h, "null handle");
  Slot* slot = nullptr;
  for (Slot& sl : h->slots)
    if (sl.busy && sl.ticket == ticket) { slot = &sl; break; }
  if (!slot) { set_error("ticket %lld is not in flight", (long long)ticket); return CMOS_ERR_STATE; }
  NvtxRange nvtx_wait("cmos.track.wait");
  CMOS_CUDA_OK(cudaSetDevice(h->p.orb.device));
  int rc = CMOS_OK;
  for (cudaEvent_t e : slot->done)
    if (cudaEventSynchronize(e) != cudaSuccess && !rc) { set_error("cudaEventSynchronize failed"); rc = CMOS_ERR_CUDA; }
  // the FAST candidate-overflow flag of a lane is a pinned host word the extractor writes behind its kernels
  for (Lane& L : h->lanes) {
    int r2 = cmos_orb_check_overflow(L.orb);
    if (!rc) rc = r2;
  }
  if (slot->capacity > h->kp_cap)   // rows are kp_cap wide on the device: pad the tail like the unfused call does
    for (int f = 0; f < slot->n_frames; f++)
      for (int i = h->kp_cap; i < slot->capacity; i++) slot->match[(size_t)f * slot->capacity + i] = -1;
  slot->busy = false;
  return rc;
}

int cmos_track_frames(cmos_track_t h, const uint8_t* images, int64_t frame_stride, int32_t pitch, int32_t width,
                      int32_t height, int32_t n_frames, const double* Tcw, const cmos_keypoint* last_keypoints,
                      const int32_t* last_counts, const uint8_t* last_flags, const double* last_xw,
                      const uint8_t* last_descriptors, int32_t last_stride, float th, int32_t check_orientation,
                      cmos_keypoint* keypoints, uint8_t* descriptors, int32_t* counts, int32_t capacity, int32_t* match,
                      int32_t* nmatches) {
  int64_t ticket = -1;
  int rc = cmos_track_submit(h, images, frame_stride, pitch, width, height, n_frames, Tcw, last_keypoints, last_counts,
                             last_flags, last_xw, last_descriptors, last_stride, th, check_orientation, keypoints, descriptors,
                             counts, capacity, match, nmatches, &ticket);
  if (rc) return rc;
  return cmos_track_wait(h, ticket);
}

int cmos_track_last_launch_count(cmos_track_t h, int32_t* n) {
  CMOS_REQUIRE(h && n, "null argument");
  *n = h->launches;
  return CMOS_OK;
}

}  // extern "C"
