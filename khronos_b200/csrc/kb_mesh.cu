// Marching cubes on the device: hydra::MeshIntegrator::generateMesh (UPSTREAM; call sites
// khronos/src/active_window/active_window.cpp:223, khronos/src/active_window/object_extraction/mesh_object_extractor.cpp:267)
// over the block pool, so that an output tick returns triangles instead of mirroring every updated block (68 KiB TSDF +
// 328 KiB semantics each) back to a host hydra::VolumetricMap. Behaviour: docs/ORACLE_SPEC.md §13; parity oracle:
// oracle/oracle_mesh.cpp.
//
// One CTA per block, two passes (count, emit) around a prefix sum, so that the vertex order is exactly the reference's
// serial order (inside cubes x-major, then the max-x / max-y / max-z border planes) without any host work per cube.
// The block's own TSDF (V x 8 B, contiguous in the pool) is staged into shared memory by one cp.async.bulk (TMA unit,
// mbarrier complete_tx): 7 of the 8 corner reads of an interior cube then hit shared memory; corners in the +x/+y/+z
// neighbour blocks (border cubes only) are read from global memory through 8 neighbour slots resolved once per CTA.
// Compile with -fmad=false like the rest of the library (vertex interpolation must round like the oracle).
#include "kb_mesh.cuh"

#include "../../include/khronos_b200.h"
#include "kb_mc_tables.h"

namespace kb {

namespace {

__constant__ signed char c_tri[256][16];
__constant__ unsigned char c_ntri[256];
__constant__ unsigned char c_edge[12][2];
bool g_tables_ready[64] = {false};  // per device

constexpr int kMeshThreads = 256;

__device__ __forceinline__ uint32_t smem32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// Emission order index -> cube coordinates (oracle_mesh.cpp: inside block x-major, then the three border planes).
template <int VPS>
__device__ __forceinline__ void cubeOfOrder(int t, int& x, int& y, int& z) {
  constexpr int M = VPS - 1, M3 = M * M * M;
  if (t < M3) { x = t / (M * M); y = (t / M) % M; z = t % M; return; }
  t -= M3;
  if (t < VPS * VPS) { x = M; z = t / VPS; y = t % VPS; return; }
  t -= VPS * VPS;
  if (t < VPS * M) { y = M; z = t / M; x = t % M; return; }
  t -= VPS * M;
  z = M; y = t / M; x = t % M;
}

template <int VPS>
struct BlockView {
  const float2* s_tsdf;   // the block's own voxels in shared memory
  const int* s_nbr;       // 8 neighbour slots: bit0 +x, bit1 +y, bit2 +z (index 0 = the block itself)
  const DeviceMap* m;
  // corner c of cube (x, y, z): returns false if its block is missing
  __device__ __forceinline__ bool corner(int x, int y, int z, int c, float2& v, int& slot, int& lin, int& nsel) const {
    int vx = x + ((c == 1 || c == 2 || c == 5 || c == 6) ? 1 : 0);
    int vy = y + ((c == 2 || c == 3 || c == 6 || c == 7) ? 1 : 0);
    int vz = z + (c >= 4 ? 1 : 0);
    nsel = 0;
    if (vx == VPS) { vx = 0; nsel |= 1; }
    if (vy == VPS) { vy = 0; nsel |= 2; }
    if (vz == VPS) { vz = 0; nsel |= 4; }
    lin = vx + VPS * (vy + VPS * vz);
    slot = s_nbr[nsel];
    if (slot < 0) return false;
    v = nsel == 0 ? s_tsdf[lin] : m->tsdf[static_cast<size_t>(slot) * (VPS * VPS * VPS) + lin];
    return true;
  }
};

// Stages the block's TSDF into shared memory with one bulk copy and resolves the neighbour slots.
template <int VPS>
__device__ __forceinline__ void stageBlock(const DeviceMap& m, int slot, float2* s_tsdf, int* s_nbr, unsigned long long* s_bar) {
  constexpr int V = VPS * VPS * VPS;
  if (threadIdx.x == 0) {
    const uint32_t bar = smem32(s_bar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const uint32_t bytes = V * sizeof(float2);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem32(s_tsdf)),
                 "l"(m.tsdf + static_cast<size_t>(slot) * V), "r"(bytes), "r"(bar)
                 : "memory");
  }
  if (threadIdx.x < 8) {
    const int3 bi = m.block_index[slot];
    const int k = threadIdx.x;
    s_nbr[k] = k == 0 ? slot : hashLookup(m, bi.x + (k & 1), bi.y + ((k >> 1) & 1), bi.z + ((k >> 2) & 1));
  }
  __syncthreads();  // barrier initialised + neighbour slots visible
  const uint32_t bar = smem32(s_bar);
  uint32_t done = 0;
  while (!done) {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(done)
                 : "r"(bar), "r"(0u)
                 : "memory");
  }
}

// Exclusive scan of one int per thread over the CTA (256 threads); returns the thread's offset, *total = CTA sum.
__device__ __forceinline__ int ctaExclusiveScan(int v, int* s_warp, int* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int n = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += n;
  }
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int w = lane < kMeshThreads / 32 ? s_warp[lane] : 0;
#pragma unroll
    for (int o = 1; o < kMeshThreads / 32; o <<= 1) {
      const int n = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += n;
    }
    if (lane < kMeshThreads / 32) s_warp[lane] = w;
  }
  __syncthreads();
  *total = s_warp[kMeshThreads / 32 - 1];
  return inc - v + (warp ? s_warp[warp - 1] : 0);
}

template <int VPS>
__global__ void __launch_bounds__(kMeshThreads) meshCountKernel(const DeviceMap m, const MeshParams p) {
  constexpr int V = VPS * VPS * VPS, PER = V / kMeshThreads;
  __shared__ __align__(128) float2 s_tsdf[V];
  __shared__ __align__(8) unsigned long long s_bar;
  __shared__ int s_nbr[8];
  __shared__ int s_warp[kMeshThreads / 32];
  const int slot = p.slots[blockIdx.x];
  stageBlock<VPS>(m, slot, s_tsdf, s_nbr, &s_bar);
  BlockView<VPS> bv{s_tsdf, s_nbr, &m};
  int count = 0;
  unsigned char* __restrict__ cases = p.cases + static_cast<size_t>(blockIdx.x) * V;
#pragma unroll 1
  for (int k = 0; k < PER; ++k) {
    const int t = threadIdx.x * PER + k;
    int x, y, z;
    cubeOfOrder<VPS>(t, x, y, z);
    int index = 0;
    bool ok = true;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      float2 v;
      int cs, cl, ns;
      if (!bv.corner(x, y, z, c, v, cs, cl, ns) || !(v.y >= p.min_weight)) { ok = false; break; }
      if (v.x < 0.f) index |= 1 << c;
    }
    if (!ok || index == 255) index = 0;
    cases[t] = static_cast<unsigned char>(index);
    count += c_ntri[index];
  }
  int total;
  ctaExclusiveScan(count, s_warp, &total);
  if (threadIdx.x == 0) p.tri_count[blockIdx.x] = total;
}

template <int VPS>
__global__ void __launch_bounds__(kMeshThreads) meshEmitKernel(const DeviceMap m, const MeshParams p) {
  constexpr int V = VPS * VPS * VPS, PER = V / kMeshThreads;
  __shared__ __align__(128) float2 s_tsdf[V];
  __shared__ __align__(8) unsigned long long s_bar;
  __shared__ int s_nbr[8];
  __shared__ int s_warp[kMeshThreads / 32];
  const int slot = p.slots[blockIdx.x];
  stageBlock<VPS>(m, slot, s_tsdf, s_nbr, &s_bar);
  BlockView<VPS> bv{s_tsdf, s_nbr, &m};
  const unsigned char* __restrict__ cases = p.cases + static_cast<size_t>(blockIdx.x) * V;
  int count = 0;
#pragma unroll 1
  for (int k = 0; k < PER; ++k) count += c_ntri[cases[threadIdx.x * PER + k]];
  int total;
  long long tri = p.tri_base[blockIdx.x] + ctaExclusiveScan(count, s_warp, &total);
  if (threadIdx.x == 0 && p.clear_flag) atomicAnd(&m.block_flags[slot], ~static_cast<uint32_t>(KB_FLAG_MESH_UPDATED));
  if (count == 0) return;
#pragma unroll 1
  for (int k = 0; k < PER; ++k) {
    const int t = threadIdx.x * PER + k;
    const int index = cases[t];
    if (c_ntri[index] == 0) continue;
    int x, y, z;
    cubeOfOrder<VPS>(t, x, y, z);
    float sdf[8], px[8], py[8], pz[8];
    int cslot[8], clin[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      float2 v;
      int ns;
      bv.corner(x, y, z, c, v, cslot[c], clin[c], ns);
      sdf[c] = v.x;
      const int3 bi = m.block_index[cslot[c]];
      const int vx = clin[c] % VPS, vy = (clin[c] / VPS) % VPS, vz = clin[c] / (VPS * VPS);
      px[c] = static_cast<float>(bi.x) * p.block_size + (static_cast<float>(vx) + 0.5f) * p.voxel_size;
      py[c] = static_cast<float>(bi.y) * p.block_size + (static_cast<float>(vy) + 0.5f) * p.voxel_size;
      pz[c] = static_cast<float>(bi.z) * p.block_size + (static_cast<float>(vz) + 0.5f) * p.voxel_size;
    }
    const signed char* row = c_tri[index];
    for (int q = 0; row[q] != -1; q += 3, ++tri) {
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const int e = row[q + 2 - j];  // voxblox meshCube emits (row[q+2], row[q+1], row[q])
        const int c0 = c_edge[e][0], c1 = c_edge[e][1];
        const float diff = sdf[c0] - sdf[c1];
        float t01 = 0.5f, vx, vy, vz;
        if (fabsf(diff) >= 1e-6f) {
          t01 = sdf[c0] / diff;
          vx = px[c0] + t01 * (px[c1] - px[c0]);
          vy = py[c0] + t01 * (py[c1] - py[c0]);
          vz = pz[c0] + t01 * (pz[c1] - pz[c0]);
        } else {
          vx = 0.5f * (px[c0] + px[c1]);
          vy = 0.5f * (py[c0] + py[c1]);
          vz = 0.5f * (pz[c0] + pz[c1]);
        }
        const int cn = t01 < 0.5f ? c0 : c1;  // attributes of the nearer corner voxel
        const long long vi = tri * 3 + j;
        p.points[vi * 3 + 0] = vx;
        p.points[vi * 3 + 1] = vy;
        p.points[vi * 3 + 2] = vz;
        const size_t gi = static_cast<size_t>(cslot[cn]) * V + clin[cn];
        uchar4 col = make_uchar4(0, 0, 0, 0);
        if (m.color) col = m.color[gi];
        p.colors[vi * 3 + 0] = col.x;
        p.colors[vi * 3 + 1] = col.y;
        p.colors[vi * 3 + 2] = col.z;
        unsigned int label = 0;
        const int sem = m.sem_label ? m.block_sem[cslot[cn]] : -1;
        if (sem >= 0) {
          const uint16_t lb = m.sem_label[static_cast<size_t>(sem) * V + clin[cn]];
          if (lb != kSemEmpty) label = lb;
        }
        p.labels[vi] = label;
      }
    }
  }
}

// Exclusive prefix sum of the per-block triangle counts (one CTA; n is a few 10^4 at most); base[n] = total.
__global__ void __launch_bounds__(1024) meshScanKernel(const int* __restrict__ cnt, long long* __restrict__ base, int n) {
  __shared__ long long s_part[1024];
  const int per = (n + 1023) / 1024;
  const int lo = min(threadIdx.x * per, n), hi = min(lo + per, n);
  long long s = 0;
  for (int i = lo; i < hi; ++i) s += cnt[i];
  s_part[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    long long run = 0;
    for (int i = 0; i < 1024; ++i) { const long long v = s_part[i]; s_part[i] = run; run += v; }
    base[n] = run;
  }
  __syncthreads();
  long long run = s_part[threadIdx.x];
  for (int i = lo; i < hi; ++i) { base[i] = run; run += cnt[i]; }
}

void ensureTables() {
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 64 && g_tables_ready[dev]) return;
  unsigned char ntri[256];
  for (int i = 0; i < 256; ++i) {
    int n = 0;
    while (n < 16 && kMcTriangles[i][n] != -1) ++n;
    ntri[i] = static_cast<unsigned char>(n / 3);
  }
  cudaMemcpyToSymbol(c_tri, kMcTriangles, sizeof(kMcTriangles));
  cudaMemcpyToSymbol(c_ntri, ntri, sizeof(ntri));
  cudaMemcpyToSymbol(c_edge, kMcEdgeCorners, sizeof(kMcEdgeCorners));
  if (dev < 64) g_tables_ready[dev] = true;
}

}  // namespace

void launchMeshCount(const DeviceMap& m, const MeshParams& p, cudaStream_t s) {
  if (p.n_blocks <= 0) return;
  ensureTables();
  if (m.vps == 16) meshCountKernel<16><<<p.n_blocks, kMeshThreads, 0, s>>>(m, p);
  else meshCountKernel<8><<<p.n_blocks, kMeshThreads, 0, s>>>(m, p);
}

void launchMeshScan(const int* tri_count, long long* tri_base, int n, cudaStream_t s) {
  meshScanKernel<<<1, 1024, 0, s>>>(tri_count, tri_base, n);
}

void launchMeshEmit(const DeviceMap& m, const MeshParams& p, cudaStream_t s) {
  if (p.n_blocks <= 0) return;
  ensureTables();
  if (m.vps == 16) meshEmitKernel<16><<<p.n_blocks, kMeshThreads, 0, s>>>(m, p);
  else meshEmitKernel<8><<<p.n_blocks, kMeshThreads, 0, s>>>(m, p);
}

}  // namespace kb
