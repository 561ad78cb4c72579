// Host side of the motion detector: M2-M4 (cluster growing, merging, filtering, image write-back) on
// the per-pixel voxel keys produced by the M1 kernel. Serial, data-dependent graph walks over the few
// voxels that contain seeds; SURVEY.md §8(f) row 4 lists a device version as a later step.
#pragma once

#include <cstdint>
#include <vector>

namespace kb {

struct MotionHostParams {
  int W, H;
  float fx, fy, cx, cy;
  float Rw[9], tw[3];
  int connectivity;
  int min_cluster_size, max_cluster_size;
  float min_separation_distance;
};

struct MotionCluster {
  std::vector<int32_t> pixels;  // (u, v) pairs, duplicates preserved
  std::vector<int64_t> voxels;  // (x, y, z) global voxel indices, ascending (z, y, x)
  float bbox[6];                // min xyz, max xyz of the cluster's world vertices
};

struct MotionResult {
  int n_seeds = 0;
  std::vector<MotionCluster> clusters;
};

// pixel_gidx: 3 ints per pixel (x == INT_MIN: pixel dropped); pixel_seed: 1 if its voxel is ever-free.
// depth (host) is used for bounding boxes when vertex_world is null. Writes cluster ids into
// dynamic_image (pre-zeroed by the caller).
void clusterMotion(const MotionHostParams& p, const int32_t* pixel_gidx, const uint8_t* pixel_seed,
                   const float* depth, const float* vertex_world, int32_t* dynamic_image,
                   MotionResult* out);

}  // namespace kb
