// Device-side track measurements (khronos::MaxIoUTracker, voxel mode), see kb_tracks_device.cu.
#pragma once

#include "kb_motion_device.cuh"

namespace kb {

constexpr int kTrackMaxIds = 1022;    // clusters per call: row + 1 takes 10 bits of the table key (1023 is left out so no key is ~0)
constexpr int kTrackCoordBits = 18;   // +-131072 tracker voxels per axis (13 km at the default 0.1 m)

struct TrackParams {
  int W, H;
  float fx, fy, cx, cy;
  float Rw[9], tw[3];          // world_T_sensor (float)
  const float* depth;          // device
  const float* vertex;         // device world-frame vertex map or null (computed from depth + pose)
  const int32_t* ids;          // device H*W cluster-id image (dynamic_image / object_image)
  int n_ids;                   // number of clusters (rows), <= kTrackMaxIds
  const int* id_list;          // device, strictly ascending pixel values of the clusters, or null: values 1..n_ids
  float inv_voxel;             // 1 / MaxIoUTracker::Config::voxel_size
  int* voxel_counts;           // device [n_ids]   : |cluster.voxels|
  unsigned long long* sums;    // device [n_ids*3] : sum of the voxel indices (two's complement)
};

// T0-T1: fills the (shared) table with the unique (cluster row + 1, voxel) pairs of the id image, per-id counts / index sums.
// Leaves the number of entries in t.scalars[kMsOccupied] and the entry slots in t.occupied.
void launchTrackVoxelize(const MotionTable& t, const TrackParams& p, cudaStream_t s);

// T2: intersections[row*n_tracks + track] = |cluster(row).voxels ∩ track.last_voxels|. track_keys: one packed voxel key
// per track voxel (trackVoxelKey), track_of: its track index; present_ids: row + 1 of the clusters with voxel_counts > 0.
void launchTrackIntersect(const MotionTable& t, const unsigned long long* track_keys, const int* track_of, int n_track_voxels,
                          const int* present_ids, int n_present, int n_tracks, int* intersections, cudaStream_t s);

// World-frame vertex map (H*W*3 floats, device) of p.depth with p's camera and pose (the other fields are unused).
void launchVertexMap(const TrackParams& p, float* out, cudaStream_t s);

// T3: copies the occupied keys (id, z, y, x order preserving) to `out` (device, capacity = pixels); the caller sorts.
void launchTrackExportKeys(const MotionTable& t, unsigned long long* out, cudaStream_t s);

// Host + device: packed, order preserving (z, y, x) voxel key without an id; false if a coordinate is out of range.
__host__ __device__ inline bool trackVoxelKey(long long x, long long y, long long z, unsigned long long* key) {
  const long long bias = 1ll << (kTrackCoordBits - 1);
  if (x < -bias || x >= bias || y < -bias || y >= bias || z < -bias || z >= bias) return false;
  const unsigned long long m = (1ull << kTrackCoordBits) - 1ull;
  *key = ((static_cast<unsigned long long>(z + bias) & m) << (2 * kTrackCoordBits)) |
         ((static_cast<unsigned long long>(y + bias) & m) << kTrackCoordBits) | (static_cast<unsigned long long>(x + bias) & m);
  return true;
}

__host__ __device__ inline void trackKeyDecode(unsigned long long k, int* id, int* x, int* y, int* z) {
  const long long bias = 1ll << (kTrackCoordBits - 1);
  const unsigned long long m = (1ull << kTrackCoordBits) - 1ull;
  *x = static_cast<int>(static_cast<long long>(k & m) - bias);
  *y = static_cast<int>(static_cast<long long>((k >> kTrackCoordBits) & m) - bias);
  *z = static_cast<int>(static_cast<long long>((k >> (2 * kTrackCoordBits)) & m) - bias);
  *id = static_cast<int>(k >> (3 * kTrackCoordBits));
}

}  // namespace kb
