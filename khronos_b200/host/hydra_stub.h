// Minimal stand-ins for the Hydra / OpenCV / spatial_hash types the Khronos active window touches,
// used ONLY when the real headers are absent (this container has no Hydra, Eigen, OpenCV or ROS 2).
// Member names and meanings mirror what Khronos uses in-tree:
//   hydra::InputData fields     khronos/src/active_window/active_window.cpp:283,
//                               motion_detection/free_space_motion_detector.cpp:80,169-175
//   VolumetricMap / layers      integration/tracking_integrator.cpp:75-77,109,128,142-147
//   voxel structs               tracking_integrator.cpp:156-158,229-245; mesh_object_extractor.cpp:342-356
// With -DKB_HAVE_HYDRA the adaptor includes the real headers instead and this file is not used.
#pragma once

#include <array>
#include <cstdint>
#include <map>
#include <memory>
#include <chrono>
#include <string>
#include <vector>

namespace cv {
// Row-major image with the subset of cv::Mat the path uses.
struct Mat {
  int rows = 0, cols = 0, elem = 0;
  std::vector<uint8_t> storage;
  Mat() = default;
  Mat(int r, int c, int elem_size) : rows(r), cols(c), elem(elem_size), storage(static_cast<size_t>(r) * c * elem_size, 0) {}
  bool empty() const { return storage.empty(); }
  template <typename T> T* ptr() { return reinterpret_cast<T*>(storage.data()); }
  template <typename T> const T* ptr() const { return reinterpret_cast<const T*>(storage.data()); }
  template <typename T> T& at(int v, int u) { return ptr<T>()[static_cast<size_t>(v) * cols + u]; }
  template <typename T> const T& at(int v, int u) const { return ptr<T>()[static_cast<size_t>(v) * cols + u]; }
};
}  // namespace cv

namespace hydra {

using TimeStamp = uint64_t;
struct BlockIndex {
  int x, y, z;
  bool operator<(const BlockIndex& o) const { return x != o.x ? x < o.x : (y != o.y ? y < o.y : z < o.z); }
};
using BlockIndices = std::vector<BlockIndex>;

struct Isometry3d {  // row-major 4x4 rigid transform (Eigen::Isometry3d stand-in)
  double m[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
};

struct Camera {  // hydra::Camera pinhole model (UP App. A.3)
  int width = 0, height = 0;
  float fx = 0, fy = 0, cx = 0, cy = 0;
  float min_range = 0, max_range = 0;
};

struct InputData {
  TimeStamp timestamp_ns = 0;
  Isometry3d world_T_body;
  Isometry3d body_T_sensor;
  cv::Mat depth_image;  // CV_32FC1
  cv::Mat range_image;  // CV_32FC1
  cv::Mat label_image;  // CV_32SC1
  cv::Mat color_image;  // CV_8UC3
  cv::Mat vertex_map;   // CV_32FC3, world frame
  Camera sensor;
  const Camera& getSensor() const { return sensor; }
  Isometry3d getSensorPose() const;  // world_T_body * body_T_sensor
};

struct TsdfVoxel { float distance = 0.f, weight = 0.f; uint8_t color[4] = {0, 0, 0, 0}; };
struct TrackingVoxel { TimeStamp last_observed = 0, last_occupied = 0; bool ever_free = false, active = false, to_remove = false; };
struct SemanticVoxel { uint32_t semantic_label = 0; std::vector<float> semantic_likelihoods; bool empty = true; };

template <typename VoxelT>
struct Block {
  BlockIndex index{0, 0, 0};
  std::vector<VoxelT> voxels;
  bool updated = false, mesh_updated = false, esdf_updated = false, tracking_updated = false;
  bool has_active_data = false;
  size_t numVoxels() const { return voxels.size(); }
  VoxelT& getVoxel(size_t i) { return voxels[i]; }
  const VoxelT& getVoxel(size_t i) const { return voxels[i]; }
};
using TsdfBlock = Block<TsdfVoxel>;
using TrackingBlock = Block<TrackingVoxel>;
using SemanticBlock = Block<SemanticVoxel>;

template <typename BlockT>
struct Layer {
  std::map<BlockIndex, std::shared_ptr<BlockT>> blocks;
  std::shared_ptr<BlockT> getBlockPtr(const BlockIndex& i) const { auto it = blocks.find(i); return it == blocks.end() ? nullptr : it->second; }
  std::shared_ptr<BlockT> allocateBlock(const BlockIndex& i, size_t nvox) {
    auto& b = blocks[i];
    if (!b) { b = std::make_shared<BlockT>(); b->index = i; b->voxels.resize(nvox); }
    return b;
  }
  void removeBlock(const BlockIndex& i) { blocks.erase(i); }
  size_t numBlocks() const { return blocks.size(); }
};

struct VolumetricMap {
  struct Config {
    float voxel_size = 0.1f;
    int voxels_per_side = 16;
    float truncation_distance = 0.3f;
    bool with_semantics = false;
    bool with_tracking = false;
  } config;
  explicit VolumetricMap(const Config& c) : config(c) {}
  Layer<TsdfBlock> tsdf;
  Layer<TrackingBlock> tracking;
  Layer<SemanticBlock> semantics;
  Layer<TsdfBlock>& getTsdfLayer() { return tsdf; }
  const Layer<TsdfBlock>& getTsdfLayer() const { return tsdf; }
  Layer<TrackingBlock>* getTrackingLayer() { return config.with_tracking ? &tracking : nullptr; }
  Layer<SemanticBlock>* getSemanticLayer() { return config.with_semantics ? &semantics : nullptr; }
  void removeBlock(const BlockIndex& i) { tsdf.removeBlock(i); tracking.removeBlock(i); semantics.removeBlock(i); }
  size_t numVoxels() const { return static_cast<size_t>(config.voxels_per_side) * config.voxels_per_side * config.voxels_per_side; }
};

inline Isometry3d InputData::getSensorPose() const {
  Isometry3d r;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      double s = 0;
      for (int k = 0; k < 4; ++k) s += world_T_body.m[i * 4 + k] * body_T_sensor.m[k * 4 + j];
      r.m[i * 4 + j] = s;
    }
  return r;
}

// hydra::MeshLayer stand-in (spark_dsg::Mesh per block: points, colors, labels, faces), what MeshIntegrator::generateMesh
// fills and khronos::utils::combineMeshLayer (khronos/src/utils/geometry_utils.cpp:61-86) concatenates.
struct MeshBlock {
  BlockIndex index{0, 0, 0};
  std::vector<std::array<float, 3>> points;
  std::vector<std::array<uint8_t, 3>> colors;
  std::vector<uint32_t> labels;
  std::vector<std::array<size_t, 3>> faces;
};
struct MeshLayer {
  std::map<BlockIndex, MeshBlock> blocks;
  MeshBlock& allocateBlock(const BlockIndex& i) { auto& b = blocks[i]; b.index = i; return b; }
  void removeBlock(const BlockIndex& i) { blocks.erase(i); }
  size_t numBlocks() const { return blocks.size(); }
};

// hydra::timing::ScopedTimer / ElapsedTimeRecorder stand-ins (khronos aliases `Timer`, common_types.h:130): the adaptor
// opens the reference's timer names so timing/stats.csv keeps its rows.
namespace timing {
struct ElapsedTimeRecorder {
  struct Entry { std::string name; uint64_t stamp; double seconds; };
  std::vector<Entry> entries;
  static ElapsedTimeRecorder& instance() { static ElapsedTimeRecorder r; return r; }
};
struct ScopedTimer {
  ScopedTimer(const std::string& name, uint64_t stamp) : name_(name), stamp_(stamp), start_(std::chrono::steady_clock::now()) {}
  ~ScopedTimer() {
    const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - start_).count();
    ElapsedTimeRecorder::instance().entries.push_back({name_, stamp_, s});
  }
  std::string name_;
  uint64_t stamp_;
  std::chrono::steady_clock::time_point start_;
};
}  // namespace timing

}  // namespace hydra
