// Hand-written sm_100a kernels of the active-window fusion hot path.
// Compile with -fmad=false: the per-voxel arithmetic must round exactly like the fp32 reference
// (no FMA contraction), integer outputs (labels, indices, flags) are bit-exact by construction.
//
//   K0     selectBlocksKernel   per batch: frustum test (+ conservative depth culling), block hash
//                               insert, semantic slot assignment, compaction into a work list
//   K1     fuseKernel           persistent CTAs over (block, z-slab) items: projective TSDF + semantic
//                               fusion of up to 32 frames with the voxel state held in registers
//   K2     trackingKernel       per-voxel last_occupied / active / to_remove, block has_active_data
//   K3     everFreeKernel       ever-free labelling with 6/18/26 neighbourhood across blocks
//   K2r    resetInactiveKernel  block removal + slot recycling
//   M1     motionLookupKernel   per-pixel endpoint voxel lookup + ever-free seed test
//   K4     scanConfidenceKernel object-extraction low-confidence erase
#include <limits.h>

#include <algorithm>

#include "../../include/khronos_b200.h"
#include "kb_kernels.cuh"

namespace kb {

namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ void xform(const float* R, const float* t, float x, float y, float z,
                                      float& ox, float& oy, float& oz) {
  ox = ((R[0] * x + R[1] * y) + R[2] * z) + t[0];
  oy = ((R[3] * x + R[4] * y) + R[5] * z) + t[1];
  oz = ((R[6] * x + R[7] * y) + R[8] * z) + t[2];
}

// hydra::Camera::pointIsInViewFrustum restated (oracle.cpp pointInFrustum): z, range and 4 planes.
__device__ __forceinline__ bool inFrustum(const BatchParams& p, float x, float y, float z) {
  const float infl = p.infl;
  if (z < -infl) return false;
  const float r = sqrtf((x * x + y * y) + z * z);
  if (r < p.min_range - infl || r > p.max_range + infl) return false;
  if (p.pl[0][0] * x + p.pl[0][1] * z < -infl) return false;
  if (p.pl[1][0] * x + p.pl[1][1] * z < -infl) return false;
  if (p.pl[2][0] * y + p.pl[2][1] * z < -infl) return false;
  if (p.pl[3][0] * y + p.pl[3][1] * z < -infl) return false;
  return true;
}

struct Taps {
  bool valid, bilinear;
  int u, v;
  float w0, w1, w2, w3;
};

__device__ __forceinline__ Taps nearestTaps(const BatchParams& p, const float* __restrict__ depth, float u, float v) {
  Taps t;
  t.bilinear = false;
  t.u = static_cast<int>(roundf(u));
  t.v = static_cast<int>(roundf(v));
  t.w0 = t.w1 = t.w2 = t.w3 = 0.f;
  t.valid = t.u >= 0 && t.u < p.W && t.v >= 0 && t.v < p.H && __ldg(&depth[t.v * p.W + t.u]) > 0.f;
  return t;
}

// ProjectionInterpolator{Nearest,Bilinear,Adaptive}::computeWeights (UP).
// Returns the interpolated range through `range` when valid.
__device__ __forceinline__ Taps computeTaps(const BatchParams& p, const float* __restrict__ depth, float u, float v, float& range) {
  Taps t;
  t.valid = false;
  if (p.interp == KB_INTERP_NEAREST) {
    t = nearestTaps(p, depth, u, v);
    if (t.valid) range = __ldg(&depth[t.v * p.W + t.u]);
    return t;
  }
  const int u0 = static_cast<int>(floorf(u)), v0 = static_cast<int>(floorf(v));
  const bool inside = u0 >= 0 && v0 >= 0 && u0 + 1 < p.W && v0 + 1 < p.H;
  bool use_nearest = !inside;
  float r0 = 0.f, r1 = 0.f, r2 = 0.f, r3 = 0.f;
  if (inside) {
    const float* row0 = depth + v0 * p.W + u0;
    r0 = __ldg(row0);
    r2 = __ldg(row0 + 1);
    r1 = __ldg(row0 + p.W);
    r3 = __ldg(row0 + p.W + 1);
    const bool all_valid = r0 > 0.f && r1 > 0.f && r2 > 0.f && r3 > 0.f;
    if (p.interp == KB_INTERP_ADAPTIVE) {
      const float mx = fmaxf(fmaxf(r0, r1), fmaxf(r2, r3));
      const float mn = fminf(fminf(r0, r1), fminf(r2, r3));
      use_nearest = !all_valid || !(mx - mn < p.adaptive_thr);
    } else if (!all_valid) {
      return t;  // bilinear: invalid
    }
  } else if (p.interp != KB_INTERP_ADAPTIVE) {
    return t;
  }
  if (use_nearest) {
    t = nearestTaps(p, depth, u, v);
    if (t.valid) range = __ldg(&depth[t.v * p.W + t.u]);
    return t;
  }
  const float du = u - static_cast<float>(u0), dv = v - static_cast<float>(v0);
  t.valid = true;
  t.bilinear = true;
  t.u = u0;
  t.v = v0;
  t.w0 = (1.f - du) * (1.f - dv);
  t.w1 = (1.f - du) * dv;
  t.w2 = du * (1.f - dv);
  t.w3 = du * dv;
  range = ((t.w0 * r0 + t.w1 * r1) + t.w2 * r2) + t.w3 * r3;
  return t;
}

// interpolateID: value at the tap with the largest weight, ties -> lowest tap index
// (taps ordered (u,v), (u,v+1), (u+1,v), (u+1,v+1)).
__device__ __forceinline__ int tapID(const BatchParams& p, const int* __restrict__ img, const Taps& t) {
  int du = 0, dv = 0;
  if (t.bilinear) {
    int best = 0;
    float bw = t.w0;
    if (t.w1 > bw) { best = 1; bw = t.w1; }
    if (t.w2 > bw) { best = 2; bw = t.w2; }
    if (t.w3 > bw) { best = 3; }
    du = best >> 1;
    dv = best & 1;
  }
  return __ldg(&img[(t.v + dv) * p.W + t.u + du]);
}

__device__ __forceinline__ float measurementWeight(const BatchParams& p, float depth, float sdf) {
  float w = (p.fx * p.fy) * (p.voxel_size * p.voxel_size) / (depth * depth);
  if (!p.constant_weight) w = w / (depth * depth);
  if (p.use_dropoff && sdf < -p.dropoff_eps) {
    w = w * ((p.trunc + sdf) / (p.trunc - p.dropoff_eps));
    w = fmaxf(w, 0.f);
  }
  return w;
}

__device__ __forceinline__ int warpSum(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---- per-frame 16x16 tile maxima of the depth image (input of the conservative block culling) -------
constexpr int kTile = 16;
__global__ void __launch_bounds__(256) tileMaxKernel(const __grid_constant__ BatchParams p) {
  const int b = blockIdx.y;
  const int tx = blockIdx.x % p.tiles_x, ty = blockIdx.x / p.tiles_x;
  const int u = tx * kTile + (threadIdx.x & 15), v = ty * kTile + (threadIdx.x >> 4);
  float d = 0.f;
  if (u < p.W && v < p.H) d = __ldg(&p.f[b].depth[v * p.W + u]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) d = fmaxf(d, __shfl_xor_sync(0xffffffffu, d, o));
  __shared__ float s[8];
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = d;
  __syncthreads();
  if (threadIdx.x == 0) {
    float m = s[0];
#pragma unroll
    for (int i = 1; i < 8; ++i) m = fmaxf(m, s[i]);
    const_cast<float*>(p.f[b].tile_max)[blockIdx.x] = m;
  }
}

// True if NO voxel of the block can receive a valid measurement from frame b, so the whole
// (block, frame) pair can be skipped without changing any result (SURVEY §7 hard part 4): either the
// block projects entirely outside the image, or every depth pixel its voxels could tap is invalid, or
// every voxel lies more than the truncation distance behind the farthest of those depths
// (sdf < -trunc). Uses margins (1 mm, 2 px) far above the fp32 rounding of the per-voxel arithmetic.
__device__ __forceinline__ bool blockCulled(const BatchParams& p, const FrameView& f, float ox, float oy, float oz) {
  const float lo = 0.5f * p.voxel_size, hi = p.block_size - 0.5f * p.voxel_size;
  float zmin = 3.0e38f, umin = 3.0e38f, umax = -3.0e38f, vmin = 3.0e38f, vmax = -3.0e38f;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    float x, y, z;
    xform(f.R, f.t, ox + ((c & 1) ? hi : lo), oy + ((c & 2) ? hi : lo), oz + ((c & 4) ? hi : lo), x, y, z);
    if (z < 1e-2f) return false;  // block reaches behind / near the camera plane: keep
    const float u = p.fx * x / z + p.cx, v = p.fy * y / z + p.cy;
    zmin = fminf(zmin, z);
    umin = fminf(umin, u); umax = fmaxf(umax, u);
    vmin = fminf(vmin, v); vmax = fmaxf(vmax, v);
  }
  if (umax < -0.5f || vmax < -0.5f || umin > static_cast<float>(p.W - 1) + 0.5f || vmin > static_cast<float>(p.H - 1) + 0.5f)
    return true;  // projects entirely outside the image
  const int u0 = max(static_cast<int>(floorf(umin)) - 2, 0), u1 = min(static_cast<int>(floorf(umax)) + 3, p.W - 1);
  const int v0 = max(static_cast<int>(floorf(vmin)) - 2, 0), v1 = min(static_cast<int>(floorf(vmax)) + 3, p.H - 1);
  const int tx0 = u0 / kTile, tx1 = u1 / kTile, ty0 = v0 / kTile, ty1 = v1 / kTile;
  if ((tx1 - tx0 + 1) * (ty1 - ty0 + 1) > 96) return false;  // large footprint (near block): keep
  float dmax = 0.f;
  for (int ty = ty0; ty <= ty1; ++ty)
    for (int tx = tx0; tx <= tx1; ++tx) dmax = fmaxf(dmax, __ldg(&f.tile_max[ty * p.tiles_x + tx]));
  if (!(dmax > 0.f)) return true;                 // no valid depth anywhere in the footprint
  return zmin - p.trunc - 1e-3f > dmax;           // everything is beyond the truncation band
}

// ---- K0: block selection for a batch of frames ---------------------------------------------------------
// One thread per candidate block of the batch's AABB (allocate mode; hydra findBlocksInViewFrustum,
// SURVEY App. A.5) or per pool slot (allocate == 0: all allocated blocks, mesh_object_extractor.cpp:242).
__global__ void __launch_bounds__(128) selectBlocksKernel(const DeviceMap m, const __grid_constant__ BatchParams p) {
  const int c0 = blockIdx.x * blockDim.x + threadIdx.x;
  if (c0 == 0) m.counters[kCtrWork0 + (p.parity ^ 1)] = 0;  // the next batch's work counter
  int slot = -1, bx = 0, by = 0, bz = 0;
  uint32_t mask = 0;
  int created = 0;
  if (p.allocate) {
    if (c0 >= p.dims[0] * p.dims[1] * p.dims[2]) return;
    int c = c0;
    bx = p.lo[0] + c % p.dims[0];
    c /= p.dims[0];
    by = p.lo[1] + c % p.dims[1];
    bz = p.lo[2] + c / p.dims[1];
    const float cx = (static_cast<float>(bx) + 0.5f) * p.block_size;
    const float cy = (static_cast<float>(by) + 0.5f) * p.block_size;
    const float cz = (static_cast<float>(bz) + 0.5f) * p.block_size;
    for (int b = 0; b < p.n_frames; ++b) {
      float x, y, z;
      xform(p.f[b].R, p.f[b].t, cx, cy, cz, x, y, z);
      if (inFrustum(p, x, y, z)) mask |= 1u << b;
    }
    if (!mask) return;
    if (p.nranks > 1 && blockOwner(bx, by, bz, p.nranks) != p.rank) return;
    slot = hashFindOrInsert(m, bx, by, bz, &created);
    if (slot < 0) return;
  } else {
    if (c0 >= p.n_slots || !(m.block_flags[c0] & kFlagAllocated)) return;
    slot = c0;
    const int3 bi = m.block_index[slot];
    bx = bi.x; by = bi.y; bz = bi.z;
    mask = p.n_frames >= 32 ? 0xffffffffu : ((1u << p.n_frames) - 1u);
  }
  atomicAdd(&m.counters[kCtrFrustum], __popc(mask));
  if (created) atomicAdd(&m.counters[kCtrAllocated], 1);
  if (p.cull) {
    const float ox = static_cast<float>(bx) * p.block_size, oy = static_cast<float>(by) * p.block_size,
                oz = static_cast<float>(bz) * p.block_size;
    uint32_t rem = mask;
    while (rem) {
      const int b = __ffs(rem) - 1;
      rem &= rem - 1;
      if (blockCulled(p, p.f[b], ox, oy, oz)) mask &= ~(1u << b);
    }
    if (!mask) return;
  }
  // Blocks that may receive measurements get their semantic slot here (one thread per block, so no
  // allocation race inside the fuse kernel); never-measured blocks cost no semantic memory.
  if (p.L > 0 && m.block_sem[slot] < 0) {
    m.block_sem[slot] = allocSlot(m.counters, kCtrSemHwm, kCtrSemFreeCount, m.sem_free_list, m.max_sem);
  }
  const int i = atomicAdd(&m.counters[kCtrWork0 + p.parity], 1);
  if (i < p.max_work) {
    p.work_slots[i] = slot;
    p.work_masks[i] = mask;
    p.work_upd[i] = 0;
  } else {
    atomicExch(&m.counters[kCtrCapacityExceeded], 1);
  }
}

// ---- K1: projective TSDF + semantic fusion ----------------------------------------------------------------
// Persistent CTAs stride over work items = (selected block, z-slab). Thread (x, y) of the 16x16 slab
// face owns NV voxels stacked in z and keeps their {distance, weight, last_observed} in registers
// while it walks the frames of the batch in order, so a voxel's TSDF is read and written once per
// batch, warp accesses are 256 B coalesced, and the NV independent gather chains give ILP.
// ProjectiveIntegrator::updateBlock / getVoxelMeasurement / computeLabel / updateVoxel (UP App. A.6;
// computeLabel structure pinned by khronos/src/active_window/integration/object_integrator.cpp:58-81).
template <int VPS>
__global__ void __launch_bounds__(kThreads) fuseKernel(const DeviceMap m, const __grid_constant__ BatchParams p) {
  constexpr int V = VPS * VPS * VPS;
  constexpr int PARTS = VPS == 16 ? 4 : 1;
  constexpr int NV = V / PARTS / kThreads;  // voxels per thread: 4 (16^3) or 2 (8^3)
  __shared__ int s_cnt[3];
  __shared__ uint32_t s_upd;
  const int tid = threadIdx.x;
  const int n_items = min(m.counters[kCtrWork0 + p.parity], p.max_work) * PARTS;
  const bool binary = p.sem_mode == KB_SEMANTICS_BINARY;

  for (int w = blockIdx.x; w < n_items; w += gridDim.x) {
    const int wi = w / PARTS, part = w % PARTS;
    const int slot = p.work_slots[wi];
    const uint32_t fmask = p.work_masks[wi];
    const int3 bi = m.block_index[slot];
    const int sem = p.L > 0 ? m.block_sem[slot] : -1;
    if (tid == 0) { s_cnt[0] = s_cnt[1] = s_cnt[2] = 0; s_upd = 0; }
    __syncthreads();

    const float ox = static_cast<float>(bi.x) * p.block_size;
    const float oy = static_cast<float>(bi.y) * p.block_size;
    const float oz = static_cast<float>(bi.z) * p.block_size;
    float2* __restrict__ tsdf = m.tsdf + static_cast<size_t>(slot) * V;
    float wx[NV], wy[NV], wz[NV];
    float2 st[NV];
    uint32_t lobs[NV];
    uint32_t have = 0, touched = 0;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int lin = part * (V / PARTS) + k * kThreads + tid;
      const int vx = lin % VPS, vy = (lin / VPS) % VPS, vz = lin / (VPS * VPS);
      wx[k] = ox + (static_cast<float>(vx) + 0.5f) * p.voxel_size;
      wy[k] = oy + (static_cast<float>(vy) + 0.5f) * p.voxel_size;
      wz[k] = oz + (static_cast<float>(vz) + 0.5f) * p.voxel_size;
      st[k] = make_float2(0.f, 0.f);
      lobs[k] = 0;
    }
    int n_valid = 0, n_band = 0, n_sem = 0;
    uint32_t upd_frames = 0;

    uint32_t rem = fmask;
    while (rem) {
      const int b = __ffs(rem) - 1;
      rem &= rem - 1;
      const FrameView& f = p.f[b];
      const bool has_sem = sem >= 0 && (binary ? f.object_image != nullptr : f.label != nullptr);
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        float x, y, z;
        xform(f.R, f.t, wx[k], wy[k], wz[k], x, y, z);
        if (z <= 0.f) continue;
        const float u = p.fx * x / z + p.cx;
        const float v = p.fy * y / z + p.cy;
        if (u < 0.f || u > static_cast<float>(p.W - 1) || v < 0.f || v > static_cast<float>(p.H - 1)) continue;
        float range = 0.f;
        const Taps taps = computeTaps(p, f.depth, u, v, range);
        if (!taps.valid) continue;
        const float sdf = range - z;
        if (sdf < -p.trunc) continue;
        const bool in_band = fabsf(sdf) < p.trunc;
        uint32_t label = 0;
        bool have_label = false;
        if (in_band) {
          if (f.mask != nullptr && tapID(p, f.mask, taps) != 0) continue;
          if (p.L > 0 && (binary ? f.object_image != nullptr : f.label != nullptr)) {
            if (binary) {
              label = tapID(p, f.object_image, taps) == f.target_id ? 1u : 0u;
            } else {
              label = static_cast<uint32_t>(tapID(p, f.label, taps));
              if (label < static_cast<uint32_t>(KB_MAX_LABELS) && ((p.blocked_mask >> label) & 1ull)) continue;
            }
            have_label = true;
          }
        }
        const float wm = measurementWeight(p, z, sdf);
        const int lin = part * (V / PARTS) + k * kThreads + tid;
        if (!((have >> k) & 1u)) { st[k] = tsdf[lin]; have |= 1u << k; }
        const float sdf_c = fminf(fmaxf(sdf, -p.trunc), p.trunc);
        const float2 old = st[k];
        st[k].x = (old.x * old.y + sdf_c * wm) / (old.y + wm);
        st[k].y = fminf(old.y + wm, p.max_weight);
        lobs[k] = f.frame_idx;
        touched |= 1u << k;
        upd_frames |= 1u << b;
        ++n_valid;
        if (!in_band) continue;
        ++n_band;
        if (has_sem && have_label && label < static_cast<uint32_t>(p.L)) {
          // SemanticIntegrator::updateLikelihoods (UP App. A.8) on the voxel's likelihood row
          uint16_t* __restrict__ slabel = m.sem_label + static_cast<size_t>(sem) * V;
          const bool empty = slabel[lin] == kSemEmpty;
          int best = 0;
          if (binary) {
            float2* lk = reinterpret_cast<float2*>(m.sem_lik + (static_cast<size_t>(sem) * V + lin) * 2);
            float2 c = empty ? make_float2(0.f, 0.f) : *lk;
            if (label) c.y = c.y + 1.f; else c.x = c.x + 1.f;
            *lk = c;
            best = c.y > c.x ? 1 : 0;
          } else {
            float4* lk = reinterpret_cast<float4*>(m.sem_lik + (static_cast<size_t>(sem) * V + lin) * m.Lp);
            float bestv = 0.f;
            for (int k4 = 0; k4 < m.Lp; k4 += 4) {
              float4 c = empty ? make_float4(p.mle_init, p.mle_init, p.mle_init, p.mle_init) : lk[k4 >> 2];
              float* cf = reinterpret_cast<float*>(&c);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const int kk = k4 + j;
                if (kk < p.L) {
                  cf[j] = cf[j] + (static_cast<uint32_t>(kk) == label ? p.mle_diag : p.mle_off);
                  if (kk == 0 || cf[j] > bestv) { bestv = cf[j]; best = kk; }
                }
              }
              lk[k4 >> 2] = c;
            }
          }
          slabel[lin] = static_cast<uint16_t>(best);
          ++n_sem;
        }
      }
    }

    // ---- write the voxel state back once, then block flags + counters ----
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      if ((touched >> k) & 1u) {
        const int lin = part * (V / PARTS) + k * kThreads + tid;
        tsdf[lin] = st[k];
        if (p.with_tracking) m.last_obs[static_cast<size_t>(slot) * V + lin] = lobs[k];
      }
    }
    const int wv = warpSum(n_valid);
    if (wv) {  // warp-uniform
      const int wb = warpSum(n_band), ws = warpSum(n_sem);
      uint32_t uf = upd_frames;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) uf |= __shfl_xor_sync(0xffffffffu, uf, o);
      if ((tid & 31) == 0) {
        atomicAdd(&s_cnt[0], wv);
        if (wb) atomicAdd(&s_cnt[1], wb);
        if (ws) atomicAdd(&s_cnt[2], ws);
        atomicOr(&s_upd, uf);
      }
    }
    __syncthreads();
    if (tid == 0 && s_cnt[0] > 0) {
      atomicOr(&m.block_flags[slot], static_cast<uint32_t>(KB_FLAG_UPDATED | KB_FLAG_MESH_UPDATED | KB_FLAG_ESDF_UPDATED | KB_FLAG_TRACKING_UPDATED));
      // blocks_updated counts (block, frame) pairs once even though several z-slabs report them
      const uint32_t prev = atomicOr(&p.work_upd[wi], s_upd);
      const int fresh = __popc(s_upd & ~prev);
      if (fresh) atomicAdd(&m.counters[kCtrBlocksUpdated], fresh);
      atomicAdd(&m.counters[kCtrVoxelsUpdated], s_cnt[0]);
      if (s_cnt[1]) atomicAdd(&m.counters[kCtrVoxelsBand], s_cnt[1]);
      if (s_cnt[2]) atomicAdd(&m.counters[kCtrVoxelsSemantic], s_cnt[2]);
    }
    __syncthreads();
  }
}

// ---- K2: TrackingIntegrator::updateBlockTracking (tracking_integrator.cpp:133-166, :224-246) --------
__global__ void __launch_bounds__(kThreads) trackingKernel(const DeviceMap m, const TrackingParams p) {
  const int slot = blockIdx.x;
  const uint32_t flags = m.block_flags[slot];
  if (!(flags & kFlagAllocated)) return;
  const int V = m.V;
  const float2* __restrict__ tsdf = m.tsdf + static_cast<size_t>(slot) * V;
  const uint32_t* __restrict__ last_obs = m.last_obs + static_cast<size_t>(slot) * V;
  uint32_t* __restrict__ last_occ = m.last_occ + static_cast<size_t>(slot) * V;
  uint8_t* __restrict__ vf = m.vflags + static_cast<size_t>(slot) * V;
  int any_active = 0;
  for (int lin = threadIdx.x; lin < V; lin += kThreads) {
    if (tsdf[lin].x < p.occupancy_thr) last_occ[lin] = p.frame_idx;
    const uint32_t lo = last_obs[lin];
    const bool now_active = lo == 0 ? (p.zero_active != 0) : (lo >= p.active_min_idx);
    uint8_t f = vf[lin];
    const bool was_active = f & kVoxActive;
    uint8_t nf = (f & ~kVoxActive) | (now_active ? kVoxActive : 0);
    if (was_active && !now_active) nf |= kVoxToRemove;
    if (nf != f) vf[lin] = nf;
    any_active |= now_active;
  }
  any_active = __syncthreads_or(any_active);
  if (threadIdx.x == 0) {
    uint32_t f = flags & ~(static_cast<uint32_t>(KB_FLAG_TRACKING_UPDATED) | KB_FLAG_HAS_ACTIVE_DATA | kFlagEverFreePending);
    if (flags & KB_FLAG_TRACKING_UPDATED) f |= kFlagEverFreePending;  // latch for K3
    if (any_active) f |= KB_FLAG_HAS_ACTIVE_DATA;
    m.block_flags[slot] = f;
  }
}

// ---- K3: TrackingIntegrator::updateBlockEverFree (tracking_integrator.cpp:168-222) ------------------
// "free(v)" = ever_free(v) || voxelIsFree(v). Neighbours set ever_free concurrently, but a voxel set
// in this pass necessarily satisfies voxelIsFree, so the predicate is stable under the race.
__device__ __forceinline__ bool voxelFreeOrEverFree(const DeviceMap& m, const TrackingParams& p, size_t idx) {
  if (m.vflags[idx] & kVoxEverFree) return true;
  const uint32_t lo = m.last_obs[idx];
  if (lo == 0) return false;
  const uint32_t oc = m.last_occ[idx];
  return oc == 0 ? (p.zero_free != 0) : (oc < p.free_max_idx);
}

__global__ void __launch_bounds__(kThreads) everFreeKernel(const DeviceMap m, const TrackingParams p) {
  const int slot = blockIdx.x;
  __shared__ int s_nbr[27];
  const uint32_t flags = m.block_flags[slot];
  if (!(flags & kFlagAllocated) || !(flags & kFlagEverFreePending)) return;
  const int vps = m.vps, V = m.V;
  if (threadIdx.x < 27) {
    const int3 bi = m.block_index[slot];
    const int dx = threadIdx.x % 3 - 1, dy = (threadIdx.x / 3) % 3 - 1, dz = threadIdx.x / 9 - 1;
    s_nbr[threadIdx.x] = (dx == 0 && dy == 0 && dz == 0) ? slot : hashLookup(m, bi.x + dx, bi.y + dy, bi.z + dz);
  }
  __syncthreads();
  const size_t base = static_cast<size_t>(slot) * V;
  for (int lin = threadIdx.x; lin < V; lin += kThreads) {
    const uint8_t f = m.vflags[base + lin];
    if (f & kVoxEverFree) continue;
    {
      const uint32_t lo = m.last_obs[base + lin];
      if (lo == 0) continue;
      const uint32_t oc = m.last_occ[base + lin];
      const bool is_free = oc == 0 ? (p.zero_free != 0) : (oc < p.free_max_idx);
      if (!is_free) continue;
    }
    const int vx = lin % vps, vy = (lin / vps) % vps, vz = lin / (vps * vps);
    bool blocked = false;
    for (int dz = -1; dz <= 1 && !blocked; ++dz)
      for (int dy = -1; dy <= 1 && !blocked; ++dy)
        for (int dx = -1; dx <= 1; ++dx) {
          const int nnz = (dx != 0) + (dy != 0) + (dz != 0);
          if (nnz == 0 || (p.connectivity == 6 && nnz > 1) || (p.connectivity == 18 && nnz > 2)) continue;
          int nx = vx + dx, ny = vy + dy, nz = vz + dz;
          int bx = 1, by = 1, bz = 1;
          if (nx < 0) { nx += vps; bx = 0; } else if (nx >= vps) { nx -= vps; bx = 2; }
          if (ny < 0) { ny += vps; by = 0; } else if (ny >= vps) { ny -= vps; by = 2; }
          if (nz < 0) { nz += vps; bz = 0; } else if (nz >= vps) { nz -= vps; bz = 2; }
          const int ns = s_nbr[bx + 3 * by + 9 * bz];
          if (ns < 0) { blocked = true; break; }
          const size_t nidx = static_cast<size_t>(ns) * V + (nx + vps * (ny + vps * nz));
          if (!voxelFreeOrEverFree(m, p, nidx)) { blocked = true; break; }
        }
    if (!blocked) m.vflags[base + lin] = f | kVoxEverFree;
  }
  __syncthreads();
  if (threadIdx.x == 0) m.block_flags[slot] = flags & ~kFlagEverFreePending;
}

// ---- K2r: TrackingIntegrator::resetInactive (tracking_integrator.cpp:106-131) -----------------------
__global__ void __launch_bounds__(kThreads) resetInactiveKernel(const DeviceMap m, int3* removed, int max_removed) {
  const int slot = blockIdx.x;
  const uint32_t flags = m.block_flags[slot];
  if (!(flags & kFlagAllocated)) return;
  const int V = m.V;
  const size_t base = static_cast<size_t>(slot) * V;
  int all_remove = 1;
  for (int lin = threadIdx.x; lin < V; lin += kThreads) all_remove &= (m.vflags[base + lin] & kVoxToRemove) ? 1 : 0;
  all_remove = __syncthreads_and(all_remove);
  if ((flags & KB_FLAG_HAS_ACTIVE_DATA) && !all_remove) return;
  // Remove: scrub the slot so that a later allocation starts from the default voxel state.
  const int sem = m.block_sem[slot];
  for (int lin = threadIdx.x; lin < V; lin += kThreads) {
    m.tsdf[base + lin] = make_float2(0.f, 0.f);
    m.last_obs[base + lin] = 0;
    m.last_occ[base + lin] = 0;
    m.vflags[base + lin] = 0;
    if (sem >= 0) m.sem_label[static_cast<size_t>(sem) * V + lin] = kSemEmpty;
  }
  if (threadIdx.x == 0) {
    const int3 bi = m.block_index[slot];
    const unsigned long long key = packKey(bi.x, bi.y, bi.z);
    uint32_t h = static_cast<uint32_t>(mix64(key)) & m.hash_mask;
    for (uint32_t probe = 0; probe <= m.hash_mask; ++probe) {
      const unsigned long long k = m.hash_keys[h];
      if (k == key) { m.hash_keys[h] = kTombKey; break; }
      if (k == kEmptyKey) break;
      h = (h + 1) & m.hash_mask;
    }
    m.block_flags[slot] = 0;
    m.block_sem[slot] = -1;
    m.free_list[atomicAdd(&m.counters[kCtrFreeCount], 1)] = slot;
    if (sem >= 0) m.sem_free_list[atomicAdd(&m.counters[kCtrSemFreeCount], 1)] = sem;
    atomicSub(&m.counters[kCtrLiveBlocks], 1);
    const int r = atomicAdd(&m.counters[kCtrRemoved], 1);
    if (r < max_removed) removed[r] = bi;
  }
}

__global__ void markAllInactiveKernel(const DeviceMap m, int n) {
  const int slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot < n && (m.block_flags[slot] & kFlagAllocated)) m.block_flags[slot] &= ~static_cast<uint32_t>(KB_FLAG_HAS_ACTIVE_DATA);
}

__global__ void clearUpdatedKernel(const DeviceMap m, int n) {
  const int slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot < n && (m.block_flags[slot] & kFlagAllocated)) m.block_flags[slot] &= ~static_cast<uint32_t>(KB_FLAG_UPDATED);
}

// ---- M1: FreeSpaceMotionDetector::setUpPointMapPart (free_space_motion_detector.cpp:158-203) --------
__global__ void motionLookupKernel(const DeviceMap m, const __grid_constant__ MotionParams p) {
  const int px = blockIdx.x * blockDim.x + threadIdx.x;
  if (px >= p.W * p.H) return;
  int3 g = make_int3(INT_MIN, 0, 0);
  uint8_t seed = 0;
  const float range = __ldg(&p.depth[px]);
  if (range > 0.f && range <= p.max_range) {
    float wx, wy, wz;
    if (p.vertex) {
      wx = __ldg(&p.vertex[3 * px]); wy = __ldg(&p.vertex[3 * px + 1]); wz = __ldg(&p.vertex[3 * px + 2]);
    } else {
      const int u = px % p.W, v = px / p.W;
      const float cxn = (static_cast<float>(u) - p.cx) / p.fx * range;
      const float cyn = (static_cast<float>(v) - p.cy) / p.fy * range;
      xform(p.Rw, p.tw, cxn, cyn, range, wx, wy, wz);
    }
    if (!(wz < p.min_z_world)) {
      const int bx = static_cast<int>(floorf(wx * p.block_size_inv));
      const int by = static_cast<int>(floorf(wy * p.block_size_inv));
      const int bz = static_cast<int>(floorf(wz * p.block_size_inv));
      const int slot = hashLookup(m, bx, by, bz);
      if (slot >= 0) {
        const int vps = m.vps;
        const int vx = static_cast<int>(floorf((wx - static_cast<float>(bx) * p.block_size) * p.voxel_size_inv));
        const int vy = static_cast<int>(floorf((wy - static_cast<float>(by) * p.block_size) * p.voxel_size_inv));
        const int vz = static_cast<int>(floorf((wz - static_cast<float>(bz) * p.block_size) * p.voxel_size_inv));
        if (vx >= 0 && vy >= 0 && vz >= 0 && vx < vps && vy < vps && vz < vps) {
          g = make_int3(bx * vps + vx, by * vps + vy, bz * vps + vz);
          seed = (m.vflags[static_cast<size_t>(slot) * m.V + (vx + vps * (vy + vps * vz))] & kVoxEverFree) ? 1 : 0;
        }
      }
    }
  }
  p.pixel_gidx[px] = g;
  p.pixel_seed[px] = seed;
  // one counter update per warp
  const unsigned ballot = __ballot_sync(__activemask(), seed != 0);
  if (ballot && (threadIdx.x & 31) == (__ffs(ballot) - 1)) atomicAdd(&m.counters[kCtrSeeds], __popc(ballot));
}

// ---- E0: dense allocation (mesh_object_extractor.cpp:220-228) ----------------------------------------
__global__ void allocateBoxKernel(const DeviceMap m, int3 lo, int3 dims, int rank, int nranks) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= dims.x * dims.y * dims.z) return;
  const int bx = lo.x + c % dims.x;
  c /= dims.x;
  const int by = lo.y + c % dims.y, bz = lo.z + c / dims.y;
  if (nranks > 1 && blockOwner(bx, by, bz, nranks) != rank) return;
  int created;
  hashFindOrInsert(m, bx, by, bz, &created);
}

// ---- K4: low-confidence erase (mesh_object_extractor.cpp:246-264, computeConfidence :342-356) -------
__global__ void __launch_bounds__(kThreads) scanConfidenceKernel(const DeviceMap m, float min_conf, float min_obs, float trunc) {
  const int slot = blockIdx.x;
  if (!(m.block_flags[slot] & kFlagAllocated)) return;
  const int V = m.V;
  const int sem = m.block_sem[slot];
  int erased = 0;
  for (int lin = threadIdx.x; lin < V; lin += kThreads) {
    float2 t = m.tsdf[static_cast<size_t>(slot) * V + lin];
    if (t.x > 0.f) continue;
    float conf = 0.f;  // empty voxel
    if (sem >= 0 && m.sem_label[static_cast<size_t>(sem) * V + lin] != kSemEmpty) {
      const float2 c = *reinterpret_cast<const float2*>(m.sem_lik + (static_cast<size_t>(sem) * V + lin) * 2);
      const float total = c.x + c.y;
      conf = total < min_obs ? -1.f : c.y / total;
    }
    if (conf < min_conf) {
      t.x = trunc;
      m.tsdf[static_cast<size_t>(slot) * V + lin] = t;
      ++erased;
    }
  }
  erased = warpSum(erased);
  if ((threadIdx.x & 31) == 0 && erased) atomicAdd(&m.counters[kCtrErased], erased);
}

// ---- export gathers -------------------------------------------------------------------------------------
__global__ void gatherTsdfKernel(const DeviceMap m, const int* slots, float* dist, float* weight) {
  const int V = m.V, slot = slots[blockIdx.x];
  for (int lin = threadIdx.x; lin < V; lin += blockDim.x) {
    const float2 t = m.tsdf[static_cast<size_t>(slot) * V + lin];
    dist[static_cast<size_t>(blockIdx.x) * V + lin] = t.x;
    weight[static_cast<size_t>(blockIdx.x) * V + lin] = t.y;
  }
}

__global__ void gatherTrackingKernel(const DeviceMap m, const int* slots, const unsigned long long* stamps,
                                     unsigned long long* last_obs, unsigned long long* last_occ,
                                     uint8_t* ever_free, uint8_t* active, uint8_t* to_remove) {
  const int V = m.V, slot = slots[blockIdx.x];
  for (int lin = threadIdx.x; lin < V; lin += blockDim.x) {
    const size_t src = static_cast<size_t>(slot) * V + lin, dst = static_cast<size_t>(blockIdx.x) * V + lin;
    last_obs[dst] = stamps[m.last_obs[src]];
    last_occ[dst] = stamps[m.last_occ[src]];
    const uint8_t f = m.vflags[src];
    ever_free[dst] = (f & kVoxEverFree) ? 1 : 0;
    active[dst] = (f & kVoxActive) ? 1 : 0;
    to_remove[dst] = (f & kVoxToRemove) ? 1 : 0;
  }
}

__global__ void gatherSemanticKernel(const DeviceMap m, const int* slots, int L, uint32_t* label,
                                     uint8_t* empty, float* lik) {
  const int V = m.V, slot = slots[blockIdx.x];
  const int sem = m.block_sem[slot];
  for (int lin = threadIdx.x; lin < V; lin += blockDim.x) {
    const size_t dst = static_cast<size_t>(blockIdx.x) * V + lin;
    uint16_t lb = kSemEmpty;
    if (sem >= 0) lb = m.sem_label[static_cast<size_t>(sem) * V + lin];
    const bool is_empty = lb == kSemEmpty;
    label[dst] = is_empty ? 0u : lb;
    empty[dst] = is_empty ? 1 : 0;
    if (lik) {
      for (int k = 0; k < L; ++k)
        lik[dst * L + k] = is_empty ? 0.f : m.sem_lik[(static_cast<size_t>(sem) * V + lin) * m.Lp + k];
    }
  }
}

}  // namespace

void launchTileMax(const BatchParams& p, cudaStream_t s) {
  dim3 grid(p.tiles_x * p.tiles_y, p.n_frames);
  tileMaxKernel<<<grid, 256, 0, s>>>(p);
}
void launchSelectBlocks(const DeviceMap& m, const BatchParams& p, cudaStream_t s) {
  const int n = p.allocate ? p.dims[0] * p.dims[1] * p.dims[2] : p.n_slots;
  selectBlocksKernel<<<(std::max(n, 1) + 127) / 128, 128, 0, s>>>(m, p);
}
void launchFuse(const DeviceMap& m, const BatchParams& p, int grid, cudaStream_t s) {
  if (grid <= 0) return;
  if (m.vps == 16) fuseKernel<16><<<grid, kThreads, 0, s>>>(m, p);
  else fuseKernel<8><<<grid, kThreads, 0, s>>>(m, p);
}
void launchTracking(const DeviceMap& m, const TrackingParams& p, int n, cudaStream_t s) {
  if (n > 0) trackingKernel<<<n, kThreads, 0, s>>>(m, p);
}
void launchEverFree(const DeviceMap& m, const TrackingParams& p, int n, cudaStream_t s) {
  if (n > 0) everFreeKernel<<<n, kThreads, 0, s>>>(m, p);
}
void launchResetInactive(const DeviceMap& m, int n, int3* removed, int max_removed, cudaStream_t s) {
  if (n > 0) resetInactiveKernel<<<n, kThreads, 0, s>>>(m, removed, max_removed);
}
void launchMarkAllInactive(const DeviceMap& m, int n, cudaStream_t s) {
  if (n > 0) markAllInactiveKernel<<<(n + 255) / 256, 256, 0, s>>>(m, n);
}
void launchClearUpdated(const DeviceMap& m, int n, cudaStream_t s) {
  if (n > 0) clearUpdatedKernel<<<(n + 255) / 256, 256, 0, s>>>(m, n);
}
void launchMotionLookup(const DeviceMap& m, const MotionParams& p, cudaStream_t s) {
  const int n = p.W * p.H;
  motionLookupKernel<<<(n + 255) / 256, 256, 0, s>>>(m, p);
}
void launchAllocateBox(const DeviceMap& m, int3 lo, int3 dims, int rank, int nranks, cudaStream_t s) {
  const int n = dims.x * dims.y * dims.z;
  if (n > 0) allocateBoxKernel<<<(n + 127) / 128, 128, 0, s>>>(m, lo, dims, rank, nranks);
}
void launchScanConfidence(const DeviceMap& m, float min_conf, float min_obs, float trunc, int n, cudaStream_t s) {
  if (n > 0) scanConfidenceKernel<<<n, kThreads, 0, s>>>(m, min_conf, min_obs, trunc);
}
void launchGatherTsdf(const DeviceMap& m, const int* slots, int n, float* dist, float* weight, cudaStream_t s) {
  if (n > 0) gatherTsdfKernel<<<n, 256, 0, s>>>(m, slots, dist, weight);
}
void launchGatherTracking(const DeviceMap& m, const int* slots, int n, const unsigned long long* stamps,
                          unsigned long long* last_obs, unsigned long long* last_occ, uint8_t* ever_free,
                          uint8_t* active, uint8_t* to_remove, cudaStream_t s) {
  if (n > 0) gatherTrackingKernel<<<n, 256, 0, s>>>(m, slots, stamps, last_obs, last_occ, ever_free, active, to_remove);
}
void launchGatherSemantic(const DeviceMap& m, const int* slots, int n, int L, uint32_t* label, uint8_t* empty,
                          float* lik, cudaStream_t s) {
  if (n > 0) gatherSemanticKernel<<<n, 256, 0, s>>>(m, slots, L, label, empty, lik);
}

}  // namespace kb
