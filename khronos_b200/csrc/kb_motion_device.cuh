// Device-side motion clustering (M2-M4), see kb_motion_device.cu.
#pragma once

#include "kb_device.cuh"

namespace kb {

constexpr uint8_t kMvSeed = 1;
enum MotionScalar { kMsSeeds = 0, kMsRoots = 1, kMsClusters = 2, kMsOccupied = 3, kMsCount = 4 };

// Open-addressed table of the voxels that contain points of the current frame; a slot index is the voxel's
// id in all per-voxel arrays (capacity = mask + 1 >= 2 * pixels).
struct MotionTable {
  unsigned long long* keys;       // packed (z,y,x), order preserving; ~0 = empty
  uint32_t mask;
  uint32_t* count;                // pixels in the voxel
  uint8_t* flags;                 // kMvSeed
  int* deg;                       // seeds: 1; absorbed voxels: number of adjacent seeds; 0: not in a cluster
  int* parent;                    // union-find
  unsigned long long* pix_total;  // per root: pixel multiset size of the cluster
  unsigned long long* min_seed;   // per root: smallest seed key (z,y,x order)
  int* cluster_id;                // per root: 0 = filtered, else 1..255
  int* occupied;                  // compact list of occupied slots (capacity = pixels)
  int* roots;                     // compact list of roots
  int max_roots;
  int* scalars;                   // MotionScalar
  int* pix_slot;                  // [pixels] slot of the pixel's voxel or -1
  const int* gate;                // device counter of seed pixels of this frame: 0 => every stage is a no-op
};

// Runs C1-C6 on `s`; writes the dynamic image (device) and scalars (seeds, roots, clusters). All stages read the
// device-side seed-pixel counter t.gate, so the sequence can be enqueued before the host knows whether M1 found seeds.
void launchMotionClustering(const MotionTable& t, const int3* gidx, const uint8_t* seed, int P, int conn, int D,
                            int min_size, int max_size, int32_t* image, cudaStream_t s);

// KB_MOTION_SPARSE experiment: same result; the table is reset slot by slot after use instead of wholesale before use
// (table_dirty: another user — the object detector — left entries behind, do one full reset).
void launchMotionClusteringSparse(const MotionTable& t, const int3* gidx, const uint8_t* seed, int P, int conn, int D,
                                  int min_size, int max_size, int32_t* image, bool table_dirty, cudaStream_t s);

}  // namespace kb
