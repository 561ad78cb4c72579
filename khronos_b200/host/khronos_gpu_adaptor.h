// Host-side C++ mirror of the reference's integrator / plugin interfaces over the C ABI
// (include/khronos_b200.h). Header-only; compiles against real Hydra/Khronos headers with
// -DKB_HAVE_HYDRA, or against hydra_stub.h (this container: no Hydra/Eigen/OpenCV/ROS).
//
// Class <-> reference interface it stands in for (paths relative to the Khronos checkout):
//   GpuVolumetricMap          hydra::VolumetricMap as owned by ActiveWindow (active_window.h:166-170)
//   GpuProjectiveIntegrator   hydra::ProjectiveIntegrator::updateMap(data, map, allocate, mask)
//                             (call sites active_window.cpp:210, mesh_object_extractor.cpp:242)
//   GpuObjectIntegrator       khronos::ObjectIntegrator (integration/object_integrator.h:51-74)
//   GpuTrackingIntegrator     khronos::TrackingIntegrator::updateBlocks / resetInactive
//                             (integration/tracking_integrator.h:96,102)
//   GpuFreeSpaceMotionDetector khronos::FreeSpaceMotionDetector::processInput
//                             (motion_detection/free_space_motion_detector.h:121; base motion_detector.h:62)
//   GpuConnectedSemantics     khronos::ConnectedSemantics::processInput (object_detection/connected_semantics.h:100)
//   GpuActiveWindowCore       the replaced steps of khronos::ActiveWindow::spinOnce / extractOutputData / finishMapping
//                             (active_window.cpp:118-174, 176-189, 217-240) in one object
//   reconstructStaticObject   the map set-up / integrate / erase steps of MeshObjectExtractor::extractStaticObject
//                             (object_extraction/mesh_object_extractor.cpp:201-264)
//   mirrorBack                repopulates a host hydra::VolumetricMap for MeshIntegrator::generateMesh /
//                             cloneUpdated (active_window.cpp:223,229)
// Error behaviour follows the reference: configuration errors throw at construction
// (config::checkValid), the per-frame path never throws — failures are logged and the frame skipped.
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/khronos_b200.h"

#ifdef KB_HAVE_HYDRA
#include <hydra/input/input_data.h>
#include <hydra/reconstruction/volumetric_map.h>
#include <hydra/utils/timing_utilities.h>
#include <khronos/active_window/data/frame_data.h>
#else
#include "hydra_stub.h"
namespace khronos {
struct Pixel { int u, v; };
struct MeasurementCluster {  // khronos/include/khronos/active_window/data/measurement_clusters.h:63-81
  std::vector<Pixel> pixels;
  float bbox_min[3], bbox_max[3];
  std::vector<std::array<int64_t, 3>> voxels;
  int id = 0;
  int semantic_id = -1;  // stands in for std::optional<SemanticClusterInfo>::category_id (measurement_clusters.h:46-58)
};
struct FrameData {  // khronos/include/khronos/active_window/data/frame_data.h:59-83
  hydra::InputData input;
  std::vector<MeasurementCluster> dynamic_clusters;
  cv::Mat dynamic_image;  // CV_32SC1
  std::vector<MeasurementCluster> semantic_clusters;
  cv::Mat object_image;   // CV_32SC1
};
}  // namespace khronos
#endif

namespace khronos_b200 {

// The reference's timer scopes (khronos `Timer` = hydra::timing::ScopedTimer, common_types.h:130) keep their names, so
// timing/stats.csv and khronos_eval/plotting/timing.py keep their rows. Host-side wall time of the C-ABI call: the calls
// that return results (motion detection, tracking pass, object detection) include the device work; a bare
// kb_integrate_frame only enqueues.
using Timer = hydra::timing::ScopedTimer;

inline void check(int status, kb_handle* h, const char* what) {
  if (status != KB_OK) throw std::runtime_error(std::string(what) + ": " + (h ? kb_last_error(h) : "kb error"));
}

// One GPU-resident map + its integrators (handle-based: extraction workers each own one).
class GpuVolumetricMap {
 public:
  GpuVolumetricMap(const hydra::VolumetricMap::Config& map, const kb_integrator_config& integ,
                   const kb_tracking_config* tracking, const kb_motion_config* motion, int max_blocks,
                   int device = 0)
      : config(map) {
    kb_map_config mc{};
    mc.voxel_size = map.voxel_size;
    mc.voxels_per_side = map.voxels_per_side;
    mc.truncation_distance = map.truncation_distance;
    mc.with_semantics = map.with_semantics;
    mc.with_tracking = map.with_tracking;
    mc.max_blocks = max_blocks;
    num_labels_ = !map.with_semantics ? 0 : integ.semantic_mode == KB_SEMANTICS_MLE ? integ.num_labels
                  : integ.semantic_mode == KB_SEMANTICS_BINARY ? 2 : 0;
    const int st = kb_create(&mc, &integ, tracking, motion, device, &h_);
    if (st != KB_OK) throw std::runtime_error("kb_create failed (status " + std::to_string(st) + "): no GPU or invalid config");
  }
  ~GpuVolumetricMap() { kb_destroy(h_); }
  GpuVolumetricMap(const GpuVolumetricMap&) = delete;
  GpuVolumetricMap& operator=(const GpuVolumetricMap&) = delete;

  kb_handle* handle() const { return h_; }
  int numLabels() const { return num_labels_; }
  void setSensor(const hydra::Camera& c) {
    kb_camera cam{c.width, c.height, c.fx, c.fy, c.cx, c.cy, c.min_range, c.max_range};
    check(kb_set_camera(h_, &cam), h_, "kb_set_camera");
  }
  const hydra::VolumetricMap::Config config;

 private:
  kb_handle* h_ = nullptr;
  int num_labels_ = 0;
};

inline kb_frame makeFrame(const hydra::InputData& data, const cv::Mat* mask, const cv::Mat* object_image,
                          int target_id) {
  kb_frame f{};
  f.depth = data.range_image.empty() ? data.depth_image.ptr<float>() : data.range_image.ptr<float>();
  f.label = data.label_image.empty() ? nullptr : data.label_image.ptr<int32_t>();
  f.mask = (mask && !mask->empty()) ? mask->ptr<int32_t>() : nullptr;
  f.object_image = (object_image && !object_image->empty()) ? object_image->ptr<int32_t>() : nullptr;
  f.color = data.color_image.empty() ? nullptr : data.color_image.ptr<uint8_t>();  // CV_8UC3, RGB
  f.vertex_world = data.vertex_map.empty() ? nullptr : data.vertex_map.ptr<float>();
  const auto T = data.getSensorPose();
#ifdef KB_HAVE_HYDRA
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c) f.world_T_sensor[r * 4 + c] = T.matrix()(r, c);
#else
  std::memcpy(f.world_T_sensor, T.m, sizeof(T.m));
#endif
  f.stamp_ns = data.timestamp_ns;
  f.object_target_id = target_id;
  f.memory = KB_MEM_HOST;
  return f;
}

// hydra::MeshIntegrator stand-in (call sites active_window.cpp:223 `generateMesh(map_, true, true)`,
// mesh_object_extractor.cpp:267 `generateMesh(map, true, false)`): marching cubes run on the device and only the
// triangles of the processed blocks come back — the mesh blocks of `mesh_layer` are replaced block by block, exactly
// what the reference's per-block `mesh_layer.allocateBlock(index)` + clear does. min_weight = hydra's
// MeshIntegratorConfig::min_weight.
class GpuMeshIntegrator {
 public:
  explicit GpuMeshIntegrator(float min_weight = 1e-4f) : min_weight_(min_weight) {}
  // Returns the number of vertices written.
  size_t generateMesh(GpuVolumetricMap& map, hydra::MeshLayer& mesh_layer, bool only_mesh_updated_blocks,
                      bool clear_updated_flag) const {
    Timer timer("active_window/generate_mesh", 0);
    kb_handle* h = map.handle();
    int32_t nb = 0;
    int64_t nv = 0;
    check(kb_generate_mesh(h, only_mesh_updated_blocks, clear_updated_flag, min_weight_, &nb, &nv), h, "kb_generate_mesh");
    std::vector<int32_t> index(3 * static_cast<size_t>(nb));
    std::vector<int64_t> off(static_cast<size_t>(nb) + 1);
    std::vector<float> pts(3 * static_cast<size_t>(nv));
    std::vector<uint8_t> rgb(3 * static_cast<size_t>(nv));
    std::vector<uint32_t> lab(static_cast<size_t>(nv));
    check(kb_get_mesh(h, index.data(), off.data(), pts.data(), rgb.data(), lab.data(), nv), h, "kb_get_mesh");
    for (int32_t b = 0; b < nb; ++b) {
      auto& blk = mesh_layer.allocateBlock({index[3 * b], index[3 * b + 1], index[3 * b + 2]});
      const size_t n = static_cast<size_t>(off[b + 1] - off[b]), o = static_cast<size_t>(off[b]);
      blk.points.resize(n); blk.colors.resize(n); blk.labels.resize(n); blk.faces.resize(n / 3);
      for (size_t i = 0; i < n; ++i) {
        blk.points[i] = {pts[3 * (o + i)], pts[3 * (o + i) + 1], pts[3 * (o + i) + 2]};
        blk.colors[i] = {rgb[3 * (o + i)], rgb[3 * (o + i) + 1], rgb[3 * (o + i) + 2]};
        blk.labels[i] = lab[o + i];
      }
      for (size_t f = 0; f < n / 3; ++f) blk.faces[f] = {3 * f, 3 * f + 1, 3 * f + 2};
    }
    return static_cast<size_t>(nv);
  }

 private:
  float min_weight_;
};

// hydra::ProjectiveIntegrator stand-in. dynamic_image may be passed directly as the mask: the kernel
// tests "!= 0", which is hydra::maskNonZero fused (active_window.cpp:209).
class GpuProjectiveIntegrator {
 public:
  virtual ~GpuProjectiveIntegrator() = default;
  virtual void updateMap(const hydra::InputData& data, GpuVolumetricMap& map, bool allocate_blocks = true,
                         const cv::Mat& integration_mask = cv::Mat()) const {
    Timer timer("active_window/update_map", data.timestamp_ns);  // active_window.cpp:204 (the scope around integrator_.updateMap)
    kb_frame f = makeFrame(data, &integration_mask, objectImage(), targetId());
    map.setSensor(data.getSensor());
    const int st = kb_integrate_frame(map.handle(), &f, allocate_blocks ? 1 : 0, nullptr);
    if (st != KB_OK) std::fprintf(stderr, "[GpuProjectiveIntegrator] frame skipped: %s\n", kb_last_error(map.handle()));
  }

 protected:
  virtual const cv::Mat* objectImage() const { return nullptr; }
  virtual int targetId() const { return 0; }
};

// khronos::ObjectIntegrator: binary semantics with label = (object_image == target id). The map
// handed to updateMap must have been created with KB_SEMANTICS_BINARY (forceBinaryIntegrator,
// object_integrator.cpp:44-48).
class GpuObjectIntegrator : public GpuProjectiveIntegrator {
 public:
  void setFrameData(khronos::FrameData* frame_data, int target_object_id) {
    current_data_ = frame_data;
    current_object_id_ = target_object_id;
  }

 protected:
  const cv::Mat* objectImage() const override { return current_data_ ? &current_data_->object_image : nullptr; }
  int targetId() const override { return current_object_id_; }

 private:
  khronos::FrameData* current_data_ = nullptr;
  int current_object_id_ = -1;
};

class GpuTrackingIntegrator {
 public:
  void updateBlocks(const khronos::FrameData& data, GpuVolumetricMap& map) const {
    Timer timer("integration/tracking", data.input.timestamp_ns);  // tracking_integrator.cpp:72
    const int st = kb_update_tracking(map.handle(), data.input.timestamp_ns);
    if (st != KB_OK) std::fprintf(stderr, "[GpuTrackingIntegrator] %s\n", kb_last_error(map.handle()));
  }
  void resetInactive(GpuVolumetricMap& map, hydra::BlockIndices* removed = nullptr) const {
    int32_t n = 0;
    std::vector<int32_t> buf(static_cast<size_t>(3) * (1 << 20));
    if (kb_reset_inactive(map.handle(), buf.data(), 1 << 20, &n) != KB_OK) return;
    if (removed)
      for (int i = 0; i < n; ++i) removed->push_back({buf[3 * i], buf[3 * i + 1], buf[3 * i + 2]});
  }
};

// khronos::MotionDetector plugin (type string "GpuFreeSpaceMotionDetector" when registered with
// config_utilities). Fills dynamic_image (CV_32SC1) and dynamic_clusters.
class GpuFreeSpaceMotionDetector {
 public:
  void processInput(GpuVolumetricMap& map, khronos::FrameData& data) const {
    Timer timer("motion_detection/all", data.input.timestamp_ns);  // free_space_motion_detector.cpp:75
    kb_frame f = makeFrame(data.input, nullptr, nullptr, 0);
    map.setSensor(data.input.getSensor());
    int32_t n_seeds = 0, n_clusters = 0;
    if (data.dynamic_image.empty()) data.dynamic_image = cv::Mat(data.input.depth_image.rows, data.input.depth_image.cols, 4);
    if (kb_detect_motion(map.handle(), &f, data.dynamic_image.ptr<int32_t>(), &n_seeds, &n_clusters) != KB_OK) {
      std::fprintf(stderr, "[GpuFreeSpaceMotionDetector] %s\n", kb_last_error(map.handle()));
      return;
    }
    data.dynamic_clusters.clear();
    if (n_clusters == 0) return;
    int32_t tp = 0, tv = 0;
    kb_get_motion_clusters(map.handle(), nullptr, nullptr, nullptr, nullptr, &tp, &tv);
    std::vector<int32_t> counts(2 * n_clusters), px(2 * static_cast<size_t>(tp));
    std::vector<int64_t> vx(3 * static_cast<size_t>(tv));
    std::vector<float> bb(6 * n_clusters);
    kb_get_motion_clusters(map.handle(), counts.data(), px.data(), vx.data(), bb.data(), &tp, &tv);
    size_t po = 0, vo = 0;
    for (int c = 0; c < n_clusters; ++c) {
      khronos::MeasurementCluster cl;
      cl.id = c + 1 < 255 ? c + 1 : 255;
      for (int i = 0; i < counts[2 * c]; ++i, ++po) cl.pixels.push_back({px[2 * po], px[2 * po + 1]});
#ifndef KB_HAVE_HYDRA
      for (int i = 0; i < counts[2 * c + 1]; ++i, ++vo) cl.voxels.push_back({vx[3 * vo], vx[3 * vo + 1], vx[3 * vo + 2]});
      std::memcpy(cl.bbox_min, &bb[6 * c], 12);
      std::memcpy(cl.bbox_max, &bb[6 * c + 3], 12);
#else
      for (int i = 0; i < counts[2 * c + 1]; ++i, ++vo) cl.voxels.emplace(vx[3 * vo], vx[3 * vo + 1], vx[3 * vo + 2]);
      cl.bounding_box = khronos::BoundingBox(khronos::Point(bb[6 * c], bb[6 * c + 1], bb[6 * c + 2]),
                                             khronos::Point(bb[6 * c + 3], bb[6 * c + 4], bb[6 * c + 5]));
#endif
      data.dynamic_clusters.push_back(std::move(cl));
    }
  }
};

// khronos::ObjectDetector plugin "ConnectedSemantics" (object_detection/connected_semantics.h:60-160; registered as
// "GpuConnectedSemantics"): fills object_image (CV_32SC1) and semantic_clusters.
class GpuConnectedSemantics {
 public:
  explicit GpuConnectedSemantics(const kb_object_detector_config& config) : config_(config) {}
  void processInput(GpuVolumetricMap& map, khronos::FrameData& data) const {
    Timer timer("object_detection/all", data.input.timestamp_ns);  // connected_semantics.cpp:61
    kb_frame f = makeFrame(data.input, nullptr, nullptr, 0);
    map.setSensor(data.input.getSensor());
    if (data.object_image.empty()) data.object_image = cv::Mat(data.input.depth_image.rows, data.input.depth_image.cols, 4);
    int32_t n = 0;
    if (kb_detect_objects(map.handle(), &config_, &f, data.object_image.ptr<int32_t>(), &n) != KB_OK) {
      std::fprintf(stderr, "[GpuConnectedSemantics] %s\n", kb_last_error(map.handle()));
      return;
    }
    data.semantic_clusters.clear();
    if (n == 0) return;
    int32_t nc = 0, tp = 0;
    kb_get_object_clusters(map.handle(), nullptr, nullptr, &nc, &tp);
    std::vector<int32_t> info(3 * static_cast<size_t>(nc)), px(2 * static_cast<size_t>(tp));
    kb_get_object_clusters(map.handle(), info.data(), px.data(), &nc, &tp);
    size_t po = 0;
    for (int c = 0; c < nc; ++c) {
      khronos::MeasurementCluster cl;
      cl.id = info[3 * c];
#ifndef KB_HAVE_HYDRA
      cl.semantic_id = info[3 * c + 1];
#else
      cl.semantics = khronos::SemanticClusterInfo(info[3 * c + 1]);
#endif
      for (int i = 0; i < info[3 * c + 2]; ++i, ++po) cl.pixels.push_back({px[2 * po], px[2 * po + 1]});
      data.semantic_clusters.push_back(std::move(cl));
    }
  }

 private:
  kb_object_detector_config config_;
};

// khronos::InstanceForwarding stand-in (object_detection/instance_forwarding.cpp:73-149; registered "GpuInstanceForwarding",
// selected by `type:` like the reference's "InstanceForwarding"): object_image = the instance id image, one semantic
// cluster per id that passes the range / background / size / volume filters. `id_is_background` is the host's per-id open-set
// decision (the reference asks an EmbeddingGroup, :96-104); empty = closed set.
class GpuInstanceForwarding {
 public:
  explicit GpuInstanceForwarding(const kb_instance_forwarding_config& config) : config_(config) {}
  void processInput(GpuVolumetricMap& map, khronos::FrameData& data, const std::vector<uint8_t>& id_is_background = {}) const {
    Timer timer("object_detection/all", data.input.timestamp_ns);  // instance_forwarding.cpp:75
    kb_frame f = makeFrame(data.input, nullptr, nullptr, 0);
    map.setSensor(data.input.getSensor());
    if (data.object_image.empty()) data.object_image = cv::Mat(data.input.depth_image.rows, data.input.depth_image.cols, 4);
    int32_t n = 0;
    if (kb_forward_instances(map.handle(), &config_, &f, id_is_background.empty() ? nullptr : id_is_background.data(),
                             static_cast<int32_t>(id_is_background.size()), data.object_image.ptr<int32_t>(), &n) != KB_OK) {
      std::fprintf(stderr, "[GpuInstanceForwarding] %s\n", kb_last_error(map.handle()));
      return;
    }
    data.semantic_clusters.clear();
    if (n == 0) return;
    int32_t nc = 0, tp = 0;
    kb_get_instance_clusters(map.handle(), nullptr, nullptr, nullptr, &nc, &tp);
    std::vector<int32_t> info(2 * static_cast<size_t>(nc)), px(2 * static_cast<size_t>(tp));
    std::vector<float> bbox(6 * static_cast<size_t>(nc));
    kb_get_instance_clusters(map.handle(), info.data(), bbox.data(), px.data(), &nc, &tp);
    size_t po = 0;
    for (int c = 0; c < nc; ++c) {
      khronos::MeasurementCluster cl;
      cl.id = info[2 * c];
#ifndef KB_HAVE_HYDRA
      cl.semantic_id = info[2 * c];  // closed set: SemanticClusterInfo(id) (:138-140)
      for (int a = 0; a < 3; ++a) { cl.bbox_min[a] = bbox[6 * c + a]; cl.bbox_max[a] = bbox[6 * c + 3 + a]; }
#else
      cl.semantics = khronos::SemanticClusterInfo(info[2 * c]);
#endif
      for (int i = 0; i < info[2 * c + 1]; ++i, ++po) cl.pixels.push_back({px[2 * po], px[2 * po + 1]});
      data.semantic_clusters.push_back(std::move(cl));
    }
  }

 private:
  kb_instance_forwarding_config config_;
};

// The measurement step of khronos::MaxIoUTracker with track_by = voxels (tracking/max_iou_tracker.cpp): what
// setupTrackMeasurements (:450-459), computeCentroid (:534-539) and computeIoU (:551-562) compute for the clusters of one
// id image (data.dynamic_image or data.object_image) against the voxel sets of the live tracks — one call per image and
// frame instead of one set construction per cluster and one set probe loop per (cluster, track) pair. The greedy
// association loops (:216-448) keep running on the host over these small matrices.
struct TrackMeasurements {
  int max_id = 0, n_tracks = 0;
  std::vector<int32_t> voxel_counts;   // [max_id]              |cluster.voxels|
  std::vector<int64_t> voxel_sums;     // [max_id * 3]          centroid = (sums / count + 0.5) * voxel_size
  std::vector<int32_t> intersections;  // [max_id * n_tracks]
  std::vector<float> iou;              // [max_id * n_tracks]   computeIoUVoxels
  float iouOf(int row, int track) const { return iou[static_cast<size_t>(row) * n_tracks + track]; }  // row = position in cluster_ids, or id - 1
};

// cluster_ids: the MeasurementCluster::id values of the image's clusters, ascending (2D object images keep their
// creation-order ids), or empty for ids 1..max_id (dynamic clusters, 3D object clusters).
inline TrackMeasurements measureTracks(GpuVolumetricMap& map, const khronos::FrameData& data, const cv::Mat& id_image,
                                       int max_id, const std::vector<int32_t>& cluster_ids, float tracker_voxel_size,
                                       const std::vector<std::vector<int64_t>>& track_last_voxels_xyz) {
  TrackMeasurements out;
  if (!cluster_ids.empty()) max_id = static_cast<int>(cluster_ids.size());
  out.max_id = max_id;
  out.n_tracks = static_cast<int>(track_last_voxels_xyz.size());
  std::vector<int32_t> offsets(1, 0);
  std::vector<int64_t> flat;
  for (const auto& t : track_last_voxels_xyz) {
    flat.insert(flat.end(), t.begin(), t.end());
    offsets.push_back(static_cast<int32_t>(flat.size() / 3));
  }
  if (flat.empty()) flat.resize(3);
  out.voxel_counts.resize(static_cast<size_t>(max_id));
  out.voxel_sums.resize(static_cast<size_t>(max_id) * 3);
  out.intersections.resize(static_cast<size_t>(max_id) * out.n_tracks);
  out.iou.resize(static_cast<size_t>(max_id) * out.n_tracks);
  kb_frame f = makeFrame(data.input, nullptr, nullptr, 0);
  map.setSensor(data.input.getSensor());
  check(kb_track_measurements(map.handle(), &f, id_image.ptr<int32_t>(), max_id, cluster_ids.empty() ? nullptr : cluster_ids.data(),
                              tracker_voxel_size, out.n_tracks,
                              out.n_tracks ? offsets.data() : nullptr, out.n_tracks ? flat.data() : nullptr,
                              out.voxel_counts.data(), out.voxel_sums.data(),
                              out.n_tracks ? out.intersections.data() : nullptr, out.n_tracks ? out.iou.data() : nullptr),
        map.handle(), "kb_track_measurements");
  return out;
}

// K4 wrapper for MeshObjectExtractor::extractStaticObject (mesh_object_extractor.cpp:246-264).
inline int eraseLowConfidence(GpuVolumetricMap& map, float min_confidence, int min_observations) {
  int32_t n = 0;
  check(kb_scan_object_confidence(map.handle(), min_confidence, min_observations, &n), map.handle(), "kb_scan_object_confidence");
  return n;
}

// The GPU part of MeshObjectExtractor::extractStaticObject (mesh_object_extractor.cpp:201-264): private map sized from
// the track's extent (vps 8, truncation 2 voxels, semantics on, tracking off, :201-211), dense allocation of the
// blocks in [centre - dims, centre + dims] (:218-228), projective updates of every semantic frame with the
// ObjectIntegrator's binary labels (:237-243; one kb_integrate_frames call, the frames are fused in batches of 32 on
// the device) and the low-confidence erase (:246-264). The caller mirrors the result back and runs the unchanged
// MeshIntegrator::generateMesh (:267). Returns nullptr where the reference returns "no object" (:177-179, :211-213).
struct ObjectReconstructionConfig {  // the fields of MeshObjectExtractor::Config this step reads (mesh_object_extractor.h:80-100)
  float min_object_reconstruction_confidence = 0.5f;
  int min_object_reconstruction_observations = 10;
  float object_reconstruction_resolution = -0.02f;
  float min_reconstruction_resolution = 0.0f;
  kb_integrator_config projective_integrator{};  // semantic_mode is forced to BINARY (object_integrator.cpp:44-48)
  int max_blocks = 1 << 16;
};

inline std::unique_ptr<GpuVolumetricMap> reconstructStaticObject(
    const float extent_center[3], const float extent_dimensions[3],
    const std::vector<std::pair<const khronos::FrameData*, int>>& frames, const ObjectReconstructionConfig& config,
    int* n_erased = nullptr, int device = 0) {
  if (config.object_reconstruction_resolution == 0.f || frames.empty()) return nullptr;
  hydra::VolumetricMap::Config mc;
  if (config.object_reconstruction_resolution < 0.f) {
    const float max_dim = std::max(extent_dimensions[0], std::max(extent_dimensions[1], extent_dimensions[2]));
    mc.voxel_size = std::max(max_dim * -config.object_reconstruction_resolution, config.min_reconstruction_resolution);
  } else {
    mc.voxel_size = config.object_reconstruction_resolution;
  }
  mc.voxels_per_side = 8;
  mc.truncation_distance = mc.voxel_size * 2;
  mc.with_semantics = true;
  mc.with_tracking = false;
  if (!(mc.voxel_size > 0.f)) return nullptr;  // config::isValid(map_config)
  kb_integrator_config ic = config.projective_integrator;
  ic.semantic_mode = KB_SEMANTICS_BINARY;
  ic.num_labels = 2;
  auto map = std::make_unique<GpuVolumetricMap>(mc, ic, nullptr, nullptr, config.max_blocks, device);
  // tsdf_layer.getBlockIndex(p) = floor(p / block_size) per axis
  const float inv = 1.f / (mc.voxel_size * static_cast<float>(mc.voxels_per_side));
  int32_t lo[3], hi[3];
  for (int a = 0; a < 3; ++a) {
    lo[a] = static_cast<int32_t>(std::floor((extent_center[a] - extent_dimensions[a]) * inv));
    hi[a] = static_cast<int32_t>(std::floor((extent_center[a] + extent_dimensions[a]) * inv));
  }
  check(kb_allocate_box(map->handle(), lo, hi), map->handle(), "kb_allocate_box");
  std::vector<kb_frame> kf;
  kf.reserve(frames.size());
  for (const auto& fr : frames) kf.push_back(makeFrame(fr.first->input, nullptr, &fr.first->object_image, fr.second));
  map->setSensor(frames.front().first->input.getSensor());
  check(kb_integrate_frames(map->handle(), kf.data(), static_cast<int32_t>(kf.size()), /*allocate_blocks=*/0, nullptr),
        map->handle(), "kb_integrate_frames");
  int32_t erased = 0;
  check(kb_scan_object_confidence(map->handle(), config.min_object_reconstruction_confidence,
                                  config.min_object_reconstruction_observations, &erased),
        map->handle(), "kb_scan_object_confidence");
  if (n_erased) *n_erased = erased;
  return map;
}

// The steps of khronos::ActiveWindow::spinOnce / extractOutputData / finishMapping this build replaces, in one object a
// GpuActiveWindow subclass delegates to. Reference order (active_window.cpp:118-174): motion
// detection :127 -> object detection :130 -> tracker :134 (host, unchanged) -> updateMap :137 (= integrate with the
// dynamic mask :209-210 + tracking update :214) -> ... -> every min_output_separation seconds extractOutputData :164
// (generateMesh :223, cloneUpdated :229, resetInactive :235) and clearUpdated :169-171. Here motion detection,
// integration and tracking are one kb_spin_once call (one device round trip); object detection neither reads nor writes
// the map (connected_semantics.cpp:60 ignores it), so running it after that call changes nothing.
class GpuActiveWindowCore {
 public:
  struct Config {
    float min_output_separation = 0.f;            // s (ActiveWindow::Config, active_window.h:78-80)
    bool detect_objects = false;
    kb_object_detector_config object_detector{};
  };

  GpuActiveWindowCore(const Config& config, const hydra::VolumetricMap::Config& map, const kb_integrator_config& integ,
                      const kb_tracking_config& tracking, const kb_motion_config& motion, int max_blocks, int device = 0)
      : config_(config), map_(map, integ, &tracking, &motion, max_blocks, device), object_detector_(config.object_detector) {}

  GpuVolumetricMap& map() { return map_; }

  // Per-frame part of spinOnce. Fills data.dynamic_image (+ object_image / semantic_clusters when enabled; the dynamic
  // cluster lists are fetched lazily by GpuFreeSpaceMotionDetector-style callers via kb_get_motion_clusters).
  // Returns true when an output is due (:158-160), i.e. the caller should now call extractOutputData().
  bool spinOnce(khronos::FrameData& data) {
    Timer timer("active_window/all", data.input.timestamp_ns);  // active_window.cpp:121
    latest_stamp_ = data.input.timestamp_ns;
    kb_frame f = makeFrame(data.input, nullptr, nullptr, 0);
    map_.setSensor(data.input.getSensor());
    if (data.dynamic_image.empty()) data.dynamic_image = cv::Mat(data.input.depth_image.rows, data.input.depth_image.cols, 4);
    int32_t n_seeds = 0, n_clusters = 0;
    if (kb_spin_once(map_.handle(), &f, data.dynamic_image.ptr<int32_t>(), &n_seeds, &n_clusters) != KB_OK) {
      std::fprintf(stderr, "[GpuActiveWindowCore] frame skipped: %s\n", kb_last_error(map_.handle()));
      return false;
    }
    if (config_.detect_objects) object_detector_.processInput(map_, data);
    ++num_frames_processed_;
    const uint64_t sep = static_cast<uint64_t>(static_cast<double>(config_.min_output_separation) * 1e9);  // fromSeconds
    return !(has_output_ && last_full_update_ + sep > latest_stamp_);
  }

  // extractOutputData (:217-240) + the clearUpdated loop of spinOnce (:169-171): the updated blocks are mirrored into
  // the host map for MeshIntegrator::generateMesh / cloneUpdated, inactive blocks are removed and reported.
  void extractOutputData(hydra::VolumetricMap& host_map, hydra::BlockIndices* archived) {
    Timer timer("active_window/extract_output", latest_stamp_);  // active_window.cpp:220
#ifndef KB_HAVE_HYDRA
    mirrorBackImpl(host_map);
#else
    (void)host_map;  // with real Hydra: khronos_b200::mirrorBack(map_, host_map, true) from the including translation unit
#endif
    GpuTrackingIntegrator().resetInactive(map_, archived);
    if (archived)
      for (const auto& b : *archived) host_map.removeBlock(b);
    check(kb_clear_updated(map_.handle()), map_.handle(), "kb_clear_updated");
    last_full_update_ = latest_stamp_;
    has_output_ = true;
  }

  // finishMapping (:176-189): all blocks lose has_active_data, the next resetInactive archives everything.
  void finishMapping() { check(kb_mark_all_inactive(map_.handle()), map_.handle(), "kb_mark_all_inactive"); }

  size_t numFramesProcessed() const { return num_frames_processed_; }

 private:
#ifndef KB_HAVE_HYDRA
  void mirrorBackImpl(hydra::VolumetricMap& host_map);
#endif
  Config config_;
  GpuVolumetricMap map_;
  GpuConnectedSemantics object_detector_;
  uint64_t latest_stamp_ = 0, last_full_update_ = 0;
  bool has_output_ = false;
  size_t num_frames_processed_ = 0;
};

#ifndef KB_HAVE_HYDRA
// Repopulates the host map from the device (updated blocks only at output ticks, everything for
// parity dumps). With real Hydra the same loop writes through TsdfLayer::allocateBlock / getVoxel.
inline void mirrorBack(GpuVolumetricMap& gmap, hydra::VolumetricMap& host, bool updated_only) {
  kb_handle* h = gmap.handle();
  int32_t n = 0;
  const int which = updated_only ? KB_EXPORT_UPDATED : KB_EXPORT_ALL;
  check(kb_num_blocks(h, which, &n), h, "kb_num_blocks");
  if (n == 0) return;
  const size_t V = host.numVoxels(), L = static_cast<size_t>(gmap.numLabels());
  std::vector<int32_t> index(3 * static_cast<size_t>(n));
  std::vector<uint8_t> flags(n), ef(n * V), ac(n * V), tr(n * V), se(n * V);
  std::vector<float> dist(n * V), weight(n * V), lik(n * V * L);
  std::vector<uint64_t> lo(n * V), lc(n * V);
  std::vector<uint32_t> sl(n * V);
  std::vector<uint8_t> rgb(n * V * 3);
  kb_block_export ex{};
  ex.block_index = index.data(); ex.block_flags = flags.data(); ex.distance = dist.data(); ex.weight = weight.data();
  ex.last_observed = lo.data(); ex.last_occupied = lc.data(); ex.ever_free = ef.data(); ex.active = ac.data();
  ex.to_remove = tr.data(); ex.semantic_label = sl.data(); ex.semantic_empty = se.data();
  ex.semantic_likelihoods = L ? lik.data() : nullptr;
  ex.color = rgb.data();
  int32_t nw = 0;
  check(kb_export_blocks(h, which, n, &ex, &nw), h, "kb_export_blocks");
  for (int b = 0; b < nw; ++b) {
    const hydra::BlockIndex bi{index[3 * b], index[3 * b + 1], index[3 * b + 2]};
    auto tb = host.getTsdfLayer().allocateBlock(bi, V);
    tb->updated = flags[b] & KB_FLAG_UPDATED; tb->mesh_updated = flags[b] & KB_FLAG_MESH_UPDATED;
    tb->esdf_updated = flags[b] & KB_FLAG_ESDF_UPDATED; tb->tracking_updated = flags[b] & KB_FLAG_TRACKING_UPDATED;
    for (size_t i = 0; i < V; ++i) {
      auto& v = tb->voxels[i];
      v.distance = dist[b * V + i]; v.weight = weight[b * V + i];
      std::memcpy(v.color, &rgb[(b * V + i) * 3], 3);
    }
    if (auto* tl = host.getTrackingLayer()) {
      auto kb_ = tl->allocateBlock(bi, V);
      kb_->has_active_data = flags[b] & KB_FLAG_HAS_ACTIVE_DATA;
      for (size_t i = 0; i < V; ++i) {
        auto& v = kb_->voxels[i];
        v.last_observed = lo[b * V + i]; v.last_occupied = lc[b * V + i];
        v.ever_free = ef[b * V + i]; v.active = ac[b * V + i]; v.to_remove = tr[b * V + i];
      }
    }
    if (auto* slayer = host.getSemanticLayer()) {
      auto sb = slayer->allocateBlock(bi, V);
      for (size_t i = 0; i < V; ++i) {
        auto& v = sb->voxels[i];
        v.empty = se[b * V + i]; v.semantic_label = sl[b * V + i];
        if (!v.empty) v.semantic_likelihoods.assign(lik.begin() + (b * V + i) * L, lik.begin() + (b * V + i + 1) * L);
      }
    }
  }
}
inline void GpuActiveWindowCore::mirrorBackImpl(hydra::VolumetricMap& host_map) { mirrorBack(map_, host_map, /*updated_only=*/true); }
#endif

// khronos::RayVerificator's measurement store (backend/change_detection/ray_verificator.h:58-258) over a kb_ray_index:
// same method names and result type; the scene-graph lookups (agent layer -> pose arrays, mesh -> vertex arrays) are done
// by the caller, who passes plain arrays.
class GpuRayVerificator {
 public:
  struct CheckResult {  // RayVerificator::CheckResult (:104-115)
    std::vector<uint64_t> absent;
    std::vector<uint64_t> present;
    void merge(const CheckResult& other) {
      absent.insert(absent.end(), other.absent.begin(), other.absent.end());
      present.insert(present.end(), other.present.begin(), other.present.end());
    }
  };

  explicit GpuRayVerificator(const kb_ray_config& config, int device = 0) {
    const int st = kb_rays_create(&config, device, &h_);
    if (st != KB_OK) throw std::runtime_error(st == KB_ERR_NO_DEVICE ? "GpuRayVerificator: no CUDA device" : "GpuRayVerificator: invalid config");
  }
  ~GpuRayVerificator() { kb_rays_destroy(h_); }
  GpuRayVerificator(const GpuRayVerificator&) = delete;
  GpuRayVerificator& operator=(const GpuRayVerificator&) = delete;

  void clear() { rcheck(kb_rays_clear(h_), "kb_rays_clear"); }  // setDsg (:150-166)

  // addVertices (:222-276) for the mesh vertices [first_vertex, first_vertex + n): returns the observed blocks (x, y, z triples).
  std::vector<int32_t> addVertices(int policy, float active_window_duration, const std::vector<uint64_t>& pose_stamps,
                                   const std::vector<float>& pose_positions_xyz, int first_vertex, int n, const float* vertices_xyz,
                                   const uint64_t* first_seen, const uint64_t* last_seen) {
    std::vector<int32_t> observed(3 * 1024);
    int32_t n_obs = 0, n_added = 0;
    for (;;) {
      const int st = kb_rays_add_vertices(h_, policy, active_window_duration, static_cast<int32_t>(pose_stamps.size()), pose_stamps.data(),
                                          pose_positions_xyz.data(), n, first_vertex, vertices_xyz, first_seen, last_seen, observed.data(),
                                          static_cast<int32_t>(observed.size() / 3), &n_obs, &n_added);
      if (st == KB_ERR_CAPACITY && static_cast<size_t>(n_obs) > observed.size() / 3) { observed.resize(3 * static_cast<size_t>(n_obs)); continue; }
      rcheck(st, "kb_rays_add_vertices");
      break;
    }
    observed.resize(3 * static_cast<size_t>(n_obs));
    return observed;
  }

  // The rays' scene-graph ids, to gather deformed endpoints with; then setEndpoints (what RayLookup reads at :88-100).
  void rayIds(std::vector<int32_t>* pose_index, std::vector<int32_t>* vertex_index) const {
    int32_t n = 0;
    kb_rays_size(h_, &n, nullptr);
    pose_index->resize(n);
    vertex_index->resize(n);
    rcheck(kb_rays_get_ray_ids(h_, pose_index->data(), vertex_index->data(), nullptr, n), "kb_rays_get_ray_ids");
  }
  void setEndpoints(const std::vector<float>& sources_xyz, const std::vector<float>& targets_xyz) {
    rcheck(kb_rays_set_endpoints(h_, static_cast<int32_t>(sources_xyz.size() / 3), sources_xyz.data(), targets_xyz.data()), "kb_rays_set_endpoints");
  }
  void recomputeHash() { rcheck(kb_rays_rehash(h_), "kb_rays_rehash"); }  // :314-324

  // check (:66-146) for a batch of points, each with its own stamp window (the per-vertex loops of
  // RayObjectChangeDetector :127-137 and RayBackgroundChangeDetector::checkVertex :90-93 in one call).
  std::vector<CheckResult> check(const std::vector<float>& points_xyz, const std::vector<uint64_t>& earliest,
                                 const std::vector<uint64_t>& latest) const {
    const int32_t n = static_cast<int32_t>(points_xyz.size() / 3);
    std::vector<int32_t> counts(2 * static_cast<size_t>(n));
    int64_t total = 0;
    rcheck(kb_rays_check(h_, n, points_xyz.data(), earliest.data(), latest.data(), counts.data(), &total), "kb_rays_check");
    std::vector<uint64_t> stamps(static_cast<size_t>(total));
    rcheck(kb_rays_get_stamps(h_, stamps.data(), total), "kb_rays_get_stamps");
    std::vector<CheckResult> out(static_cast<size_t>(n));
    size_t o = 0;
    for (int32_t i = 0; i < n; ++i) {
      out[i].absent.assign(stamps.begin() + o, stamps.begin() + o + counts[2 * i]);
      o += counts[2 * i];
      out[i].present.assign(stamps.begin() + o, stamps.begin() + o + counts[2 * i + 1]);
      o += counts[2 * i + 1];
    }
    return out;
  }

  kb_ray_index* handle() const { return h_; }

 private:
  void rcheck(int st, const char* what) const {
    if (st != KB_OK) throw std::runtime_error(std::string(what) + ": " + kb_rays_last_error(h_));
  }
  kb_ray_index* h_ = nullptr;
};

}  // namespace khronos_b200
