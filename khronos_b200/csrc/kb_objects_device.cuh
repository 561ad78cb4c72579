// Device-side semantic object detection (khronos::ConnectedSemantics), see kb_objects_device.cu.
#pragma once

#include "kb_motion_device.cuh"

namespace kb {

struct ObjectParams {
  int W, H;
  float fx, fy, cx, cy;
  float Rw[9], tw[3];      // world_T_sensor (float)
  const float* depth;      // device
  const int* label;        // device
  const float* vertex;     // device world-frame vertex map or null (computed from depth + pose)
  float inv_grid, max_range;
  unsigned long long object_mask;  // bit l set: label l is an object class
  int full;                // use_full_connectivity
  int min_size, max_size;
  int32_t* image;          // device out: H*W cluster ids
};

// Both modes reuse the motion detector's table memory (the two never run concurrently on one handle's stream) and
// leave the number of kept clusters in t.scalars[kMsClusters] and the number of components in t.scalars[kMsRoots].
void launchObjectClustering3D(const MotionTable& t, const ObjectParams& p, cudaStream_t s);
void launchObjectClustering2D(const MotionTable& t, const ObjectParams& p, cudaStream_t s);

// InstanceForwarding: per-id pixel counts + world AABBs (ordered-bit encoded floats, min xyz then max xyz) of the pixels that
// pass the range / background tests; keep[px] = 1 for those pixels. Uses ObjectParams' camera, pose, depth, label, vertex, max_range.
void launchInstanceForward(const ObjectParams& p, const uint8_t* background, int max_ids, int* counts, unsigned int* bbox,
                           uint8_t* keep, int* bad_id, cudaStream_t s);

}  // namespace kb
