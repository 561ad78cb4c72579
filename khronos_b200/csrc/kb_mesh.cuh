// Marching cubes over the TSDF blocks of the device map (kb_generate_mesh).
#pragma once

#include "kb_device.cuh"

namespace kb {

struct MeshParams {
  const int* slots;          // [n_blocks] block slots to mesh, in output order
  int n_blocks;
  float voxel_size, block_size, min_weight;
  unsigned char* cases;      // [n_blocks][V] scratch: marching-cubes case of every cube in emission order (0 = nothing)
  int* tri_count;            // [n_blocks] triangles per block (pass 1)
  const long long* tri_base; // [n_blocks] exclusive prefix of tri_count (pass 2)
  float* points;             // [3 * 3 * triangles]
  unsigned char* colors;     // [3 * 3 * triangles]
  unsigned int* labels;      // [3 * triangles]
  int clear_flag;            // clear KB_FLAG_MESH_UPDATED on the processed blocks (pass 2)
};

void launchMeshCount(const DeviceMap& m, const MeshParams& p, cudaStream_t s);
void launchMeshScan(const int* tri_count, long long* tri_base, int n, cudaStream_t s);  // tri_base[n] = total
void launchMeshEmit(const DeviceMap& m, const MeshParams& p, cudaStream_t s);

}  // namespace kb
