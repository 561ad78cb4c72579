// TEST INFRASTRUCTURE — NOT PRODUCT CODE.
//
// CPU oracle: a dependency-free C++17 restatement of Khronos' active-window fusion hot path. Only
// tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it.
//
// PARITY UNPINNED: the reference ships no tests / golden vectors for this path, and the per-voxel
// arithmetic lives in MIT-SPARK/Hydra @ main (floating, un-vendored: install/https.rosinstall:5-8),
// which is absent from /root/reference. In-tree code (tracking_integrator.cpp,
// free_space_motion_detector.cpp, object_integrator.cpp, mesh_object_extractor.cpp) is followed
// line by line; upstream behaviour is restated per SURVEY.md Appendix A / docs/ORACLE_SPEC.md.
//
// All citations are relative to /root/reference/.
#pragma once

#include <array>
#include <atomic>
#include <cstdint>
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "../include/khronos_b200.h"  // plain-C config/frame/export structs (declarations only)

namespace ko {

struct Idx3 {
  int32_t x = 0, y = 0, z = 0;
  bool operator==(const Idx3& o) const { return x == o.x && y == o.y && z == o.z; }
  bool operator<(const Idx3& o) const {
    return x != o.x ? x < o.x : (y != o.y ? y < o.y : z < o.z);
  }
};
// spatial_hash block hash (UP): x + y*17191 + z*17191^2. Only affects iteration order.
struct Idx3Hash {
  size_t operator()(const Idx3& i) const {
    return static_cast<size_t>(static_cast<int64_t>(i.x) + static_cast<int64_t>(i.y) * 17191 +
                               static_cast<int64_t>(i.z) * 17191 * 17191);
  }
};
struct GIdx {
  int64_t x = 0, y = 0, z = 0;
  bool operator==(const GIdx& o) const { return x == o.x && y == o.y && z == o.z; }
};
struct GIdxHash {
  size_t operator()(const GIdx& i) const {
    return static_cast<size_t>(i.x + i.y * 17191 + i.z * 17191 * 17191);
  }
};
// Deterministic seed order: ascending (z, y, x).
struct GIdxZyxLess {
  bool operator()(const GIdx& a, const GIdx& b) const {
    return a.z != b.z ? a.z < b.z : (a.y != b.y ? a.y < b.y : a.x < b.x);
  }
};

struct Pixel {
  int32_t u, v;
};

// One voxel block of all layers (hydra TsdfBlock + TrackingBlock + SemanticBlock, UP App. A.2),
// stored SoA. Linear voxel index = x + vps*(y + vps*z).
struct Block {
  Idx3 index;
  std::vector<float> distance, weight;                   // TsdfVoxel
  std::vector<uint8_t> color;                            // 3 per voxel
  std::vector<uint64_t> last_observed, last_occupied;    // TrackingVoxel
  std::vector<uint8_t> ever_free, active, to_remove;
  std::vector<uint32_t> semantic_label;                  // SemanticVoxel
  std::vector<uint8_t> semantic_empty;
  std::vector<float> likelihoods;                        // V * L
  bool updated = false, mesh_updated = false, esdf_updated = false, tracking_updated = false;
  bool has_active_data = false;
};

struct Cluster {
  std::vector<Pixel> pixels;
  std::unordered_set<GIdx, GIdxHash> voxels;
  int id = 0;
  float bbox_min[3] = {0, 0, 0}, bbox_max[3] = {0, 0, 0};
};

class Oracle {
 public:
  Oracle(const kb_map_config& map, const kb_integrator_config& integ, const kb_tracking_config* trk,
         const kb_motion_config* mot);

  void setCamera(const kb_camera& cam) { cam_ = cam; has_cam_ = true; }

  // K0+K1: hydra::ProjectiveIntegrator::updateMap (UP; call site active_window.cpp:210).
  void integrateFrame(const kb_frame& f, bool allocate_blocks, kb_frame_stats* stats);
  // K2+K3: TrackingIntegrator::updateBlocks (tracking_integrator.cpp:71-104).
  void updateTracking(uint64_t stamp_ns);
  // K2r: TrackingIntegrator::resetInactive (tracking_integrator.cpp:106-131).
  void resetInactive(std::vector<Idx3>* removed);
  void markAllInactive();
  void clearUpdated();
  // M1-M4: FreeSpaceMotionDetector::processInput (free_space_motion_detector.cpp:73-103).
  void detectMotion(const kb_frame& f, int32_t* dynamic_image, int32_t* n_seeds, int32_t* n_clusters);
  const std::vector<Cluster>& clusters() const { return clusters_; }
  // E0 / K4 (mesh_object_extractor.cpp:220-228, :246-264, :342-356).
  void allocateBox(const int32_t mn[3], const int32_t mx[3]);
  int scanObjectConfidence(float min_confidence, int min_observations);

  // Object detection: khronos::ConnectedSemantics::processInput (object_detection/connected_semantics.cpp:60-217).
  struct ObjectCluster { int id = 0, semantic_id = 0; std::vector<Pixel> pixels; };
  void detectObjects(const kb_object_detector_config& cfg, const kb_frame& f, int32_t* object_image,
                     std::vector<ObjectCluster>* clusters);
  const std::vector<ObjectCluster>& objectClusters() const { return object_clusters_; }

  // khronos::InstanceForwarding::extractSemanticClusters (object_detection/instance_forwarding.cpp:80-149).
  struct InstanceCluster { int id = 0; std::vector<Pixel> pixels; float bbox[6] = {0, 0, 0, 0, 0, 0}; };
  void forwardInstances(const kb_instance_forwarding_config& cfg, const kb_frame& f, const uint8_t* background, int n_background,
                        int32_t* object_image);
  const std::vector<InstanceCluster>& instanceClusters() const { return instance_clusters_; }

  // Track measurements: khronos::MaxIoUTracker, track_by = voxels (tracking/max_iou_tracker.cpp:450-459, :534-539,
  // :551-562) for the clusters of an id image (pixel values cluster_ids[0..max_id), or 1..max_id when cluster_ids is null). Results: per id the voxel set (ordered z, y, x), and per
  // (id, track) the intersection count and the IoU float of :562.
  struct TrackMeasurements {
    std::vector<std::vector<GIdx>> voxels;  // [max_id]
    std::vector<int32_t> intersections;     // [max_id * n_tracks]
    std::vector<float> iou;                 // [max_id * n_tracks]
  };
  void trackMeasurements(const kb_frame& f, const int32_t* id_image, int max_id, const int32_t* cluster_ids, float voxel_size, int n_tracks,
                         const int32_t* track_offsets, const int64_t* track_voxels_xyz);
  const TrackMeasurements& trackResult() const { return track_result_; }
  // World-frame vertex map of a depth image (upstream parseInputPacket; docs/ORACLE_SPEC.md §8 "vertex map").
  bool computeVertexMap(const kb_frame& f, float* out);

  // ---- block-hash sharded protocol (our multi-GPU design, not in the reference). The oracle
  // implements it on host buffers with the layouts of csrc/kb_kernels.cuh::ShardExchange so that world-size-2
  // gloo tests can prove "union of the shards == the unsharded map" on CPU.
  void setShard(int rank, int nranks) { rank_ = rank; nranks_ = nranks; cell_ = 0; table_.clear(); }
  // kb_set_shard_cells: periodic tiling of cell x cell block cells (x/y) over a gx x gy grid of ranks.
  void setShardCells(int rank, int nranks, int cell, int gx, int gy) { rank_ = rank; nranks_ = nranks; cell_ = cell; gx_ = gx; gy_ = gy; table_.clear(); }
  static int cellOwner(int bx, int by, int cell, int gx, int gy, int nranks);
  // kb_set_shard_table: explicit cell -> rank table over [ox, ox + w) x [oy, oy + h) cells, periodic tiling outside.
  void setShardTable(int rank, int nranks, int cell, int ox, int oy, int w, int h, const uint8_t* owners);
  // kb_frame_cells: cells touched by the frame's frustum selection.
  void frameCells(const kb_frame& f, int cell, int ox, int oy, int w, int h, uint8_t* touched) const;
  int owner(const Idx3& b) const;
  // kb_frame_owners: ranks owning at least one block the frame's frustum test selects.
  uint32_t frameOwners(const kb_frame& f) const;
  int rank() const { return rank_; }
  int nranks() const { return nranks_; }
  static int blockOwner(const Idx3& b, int nranks);
  void motionLookupLocal(const kb_frame& f, uint8_t* flags);
  void motionClusterGlobal(const uint8_t* flags, int32_t* dynamic_image, int32_t* n_seeds, int32_t* n_clusters);
  void trackingBegin(uint64_t stamp_ns, int32_t* pending_out, int cap_pending);
  void packHalo(const int32_t* all_pending, int cap_pending, int32_t* halo_out, int cap_halo);
  void trackingFinish(const int32_t* all_pending, int cap_pending, const int32_t* all_halo, int cap_halo);

  // Marching cubes over the TSDF: hydra::MeshIntegrator::generateMesh(map, only_mesh_updated_blocks, clear_updated_flag)
  // (UPSTREAM; call sites active_window.cpp:223, mesh_object_extractor.cpp:267; voxblox MeshIntegrator / MarchingCubes
  // heritage, restated in docs/ORACLE_SPEC.md §13). One MeshBlock per processed block, blocks ascending (x, y, z);
  // vertices are not shared: triangle k of a block is (points[3k], points[3k+1], points[3k+2]).
  struct MeshBlock {
    Idx3 index;
    std::vector<float> points;     // 3 per vertex, world frame
    std::vector<uint8_t> colors;   // 3 per vertex
    std::vector<uint32_t> labels;  // 1 per vertex
  };
  void generateMesh(bool only_mesh_updated, bool clear_updated_flag, float min_weight);
  const std::vector<MeshBlock>& mesh() const { return mesh_; }

  std::vector<const Block*> sortedBlocks(int which) const;
  int V() const { return V_; }
  int L() const { return L_; }
  bool ok() const { return error_.empty(); }
  const std::string& error() const { return error_; }

  // Exposed for known-answer unit tests.
  struct Weights {
    bool valid = false, bilinear = false;
    int u = 0, v = 0;
    float w[4] = {0, 0, 0, 0};
  };
  Weights computeWeights(float u, float v, const float* range) const;
  float interpolateRange(const float* range, const Weights& w) const;
  int32_t interpolateID(const int32_t* img, const Weights& w) const;
  void interpolateColor(const uint8_t* rgb, const Weights& w, uint8_t out[3]) const;
  bool project(const float p_C[3], float* u, float* v) const;
  float computeWeight(float depth, float sdf) const;
  bool pointInFrustum(const float p_C[3], float inflation) const;

 private:
  Block* allocateBlock(const Idx3& idx);
  Block* getBlock(const Idx3& idx) const;
  void updateBlock(Block& b, const kb_frame& f, const float R[9], const float t[3],
                   std::atomic<int>* counters);
  void updateBlockTracking(Block& b, uint64_t stamp, float thr);
  using GhostMap = std::unordered_map<Idx3, const uint32_t*, Idx3Hash>;  // remote block -> its published free mask
  void updateBlockEverFree(const Block& b, uint64_t stamp, std::vector<int>* to_set, const GhostMap* ghosts = nullptr) const;
  std::vector<Block*> trackingPassLocal(uint64_t stamp);  // K2 on all blocks; returns the tracking_updated ones
  void applyEverFree(const std::vector<Block*>& updated, uint64_t stamp, const GhostMap* ghosts);
  // M1 split: per-pixel voxel keys (range / z / index validity, no block lookup), then flags, then clustering.
  struct PixKey { bool valid = false; Idx3 block; int lin = 0; GIdx g; };
  void computePixelKeys(const kb_frame& f);
  void clusterFromFlags(const uint8_t* flags, int32_t* dynamic_image, int32_t* n_seeds, int32_t* n_clusters);
  bool voxelIsFree(const Block& b, int lin, uint64_t stamp) const;

  kb_map_config map_;
  kb_integrator_config integ_;
  kb_tracking_config trk_{};
  kb_motion_config mot_{};
  bool has_trk_ = false, has_mot_ = false, has_cam_ = false;
  kb_camera cam_{};
  int vps_, V_, L_;
  float block_size_, voxel_size_inv_, block_size_inv_;
  float mle_diag_ = 0, mle_off_ = 0, mle_init_ = 0;
  std::unordered_map<Idx3, std::unique_ptr<Block>, Idx3Hash> blocks_;
  std::vector<Cluster> clusters_;
  std::vector<float> vertex_scratch_;
  const float* vertex_ = nullptr;          // world-frame vertex map of the frame whose keys are in pix_keys_
  std::vector<PixKey> pix_keys_;
  std::vector<uint8_t> flags_scratch_;
  int rank_ = 0, nranks_ = 1;
  int cell_ = 0, gx_ = 1, gy_ = 1;
  std::vector<uint8_t> table_;
  int tab_ox_ = 0, tab_oy_ = 0, tab_w_ = 0, tab_h_ = 0;
  std::vector<ObjectCluster> object_clusters_;
  TrackMeasurements track_result_;
  std::vector<MeshBlock> mesh_;
  std::vector<InstanceCluster> instance_clusters_;
  std::vector<Block*> open_pending_;       // ever-free work list between trackingBegin and trackingFinish
  uint64_t open_stamp_ = 0;
  std::string error_;
};

}  // namespace ko
