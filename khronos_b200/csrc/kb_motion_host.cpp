// Host part of the motion detector (M2-M4), consuming the M1 kernel's per-pixel voxel keys.
// Restates FreeSpaceMotionDetector::clusterDynamicVoxels / mergeClusters / applyClusterLevelFilters /
// writeClustersToData (khronos/src/active_window/motion_detection/free_space_motion_detector.cpp
// :205-399) on flat sorted arrays + union-find instead of nested hash maps; results are identical
// (cluster membership, duplicate-pixel multiplicities, ids, image) — see tests/test_motion_*.py.
// Compile with -ffp-contract=off (bounding boxes re-evaluate the device's vertex arithmetic).
#include "kb_motion_host.h"

#include <algorithm>
#include <array>
#include <climits>
#include <cmath>
#include <numeric>
#include <unordered_map>

namespace kb {

namespace {

struct Vox {
  int x, y, z;
};
inline bool zyxLess(const Vox& a, const Vox& b) {
  return a.z != b.z ? a.z < b.z : (a.y != b.y ? a.y < b.y : a.x < b.x);
}
inline bool voxEq(const Vox& a, const Vox& b) { return a.x == b.x && a.y == b.y && a.z == b.z; }

struct Groups {
  std::vector<Vox> vox;        // unique voxels that contain pixels, ascending (z,y,x)
  std::vector<int> start, count;  // range into pix
  std::vector<uint8_t> seed;
  std::vector<int> pix;        // pixel indices grouped by voxel
  int find(const Vox& v) const {
    auto it = std::lower_bound(vox.begin(), vox.end(), v, zyxLess);
    return (it != vox.end() && voxEq(*it, v)) ? static_cast<int>(it - vox.begin()) : -1;
  }
};

struct RawCluster {
  std::vector<int> pix_groups;  // group ids whose pixels belong to the cluster (with multiplicity)
  std::vector<int> vox_groups;  // unique group ids (voxels)
};

int findRoot(std::vector<int>& parent, int i) {
  while (parent[i] != i) {
    parent[i] = parent[parent[i]];
    i = parent[i];
  }
  return i;
}

}  // namespace

void clusterMotion(const MotionHostParams& p, const int32_t* pixel_gidx, const uint8_t* pixel_seed,
                   const float* depth, const float* vertex_world, int32_t* dynamic_image,
                   MotionResult* out) {
  const int P = p.W * p.H;
  out->clusters.clear();
  out->n_seeds = 0;

  // ---- group pixels by voxel (the reference's BlockToPointsMap) ----
  Groups g;
  std::vector<int> order;
  order.reserve(P);
  for (int i = 0; i < P; ++i)
    if (pixel_gidx[3 * i] != INT_MIN) order.push_back(i);
  auto voxOf = [&](int i) { return Vox{pixel_gidx[3 * i], pixel_gidx[3 * i + 1], pixel_gidx[3 * i + 2]}; };
  std::sort(order.begin(), order.end(), [&](int a, int b) {
    const Vox va = voxOf(a), vb = voxOf(b);
    if (!voxEq(va, vb)) return zyxLess(va, vb);
    return a < b;
  });
  g.pix = order;
  for (size_t i = 0; i < order.size();) {
    size_t j = i;
    const Vox v = voxOf(order[i]);
    uint8_t sd = 0;
    while (j < order.size() && voxEq(voxOf(order[j]), v)) { sd |= pixel_seed[order[j]]; ++j; }
    g.vox.push_back(v);
    g.start.push_back(static_cast<int>(i));
    g.count.push_back(static_cast<int>(j - i));
    g.seed.push_back(sd);
    i = j;
  }
  const int G = static_cast<int>(g.vox.size());
  for (int i = 0; i < G; ++i) out->n_seeds += g.seed[i];
  if (out->n_seeds == 0) return;

  // ---- M2: grow clusters from seeds in ascending (z,y,x) order ----
  std::vector<std::array<int, 3>> offs;
  for (int dz = -1; dz <= 1; ++dz)
    for (int dy = -1; dy <= 1; ++dy)
      for (int dx = -1; dx <= 1; ++dx) {
        const int nnz = (dx != 0) + (dy != 0) + (dz != 0);
        if (nnz == 0 || (p.connectivity == 6 && nnz > 1) || (p.connectivity == 18 && nnz > 2)) continue;
        offs.push_back({dx, dy, dz});
      }
  std::vector<uint8_t> closed(G, 0);
  std::vector<int> vox_mark(G, -1);
  std::vector<RawCluster> raw;
  std::vector<int> stack;
  for (int s = 0; s < G; ++s) {
    if (!g.seed[s] || closed[s]) continue;
    const int cid = static_cast<int>(raw.size());
    RawCluster c;
    stack.assign(1, s);
    while (!stack.empty()) {
      const int v = stack.back();
      stack.pop_back();
      if (closed[v]) continue;
      closed[v] = 1;
      c.pix_groups.push_back(v);
      if (vox_mark[v] != cid) { vox_mark[v] = cid; c.vox_groups.push_back(v); }
      for (const auto& o : offs) {
        const int n = g.find(Vox{g.vox[v].x + o[0], g.vox[v].y + o[1], g.vox[v].z + o[2]});
        if (n < 0) continue;
        if (g.seed[n]) {
          stack.push_back(n);
        } else {
          // occupied non-seed neighbour: absorbed once per adjacent seed voxel, then closed (:255-265)
          c.pix_groups.push_back(n);
          if (vox_mark[n] != cid) { vox_mark[n] = cid; c.vox_groups.push_back(n); }
          closed[n] = 1;
        }
      }
    }
    raw.emplace_back(std::move(c));
  }

  // ---- M3: merge clusters whose voxels are closer than min_separation_distance ----
  // Eigen's integer norm() truncates: int(sqrt(s)) < d  <=>  s < ceil(d)^2.
  const int C = static_cast<int>(raw.size());
  std::vector<int> parent(C);
  std::iota(parent.begin(), parent.end(), 0);
  const int D = static_cast<int>(std::ceil(p.min_separation_distance));
  if (D >= 1 && C > 1) {
    std::vector<std::vector<int>> owners(G);  // clusters containing each voxel group
    for (int c = 0; c < C; ++c)
      for (int v : raw[c].vox_groups) owners[v].push_back(c);
    const long long D2 = static_cast<long long>(D) * D;
    for (int c = 0; c < C; ++c)
      for (int v : raw[c].vox_groups)
        for (int dz = -(D - 1); dz <= D - 1; ++dz)
          for (int dy = -(D - 1); dy <= D - 1; ++dy)
            for (int dx = -(D - 1); dx <= D - 1; ++dx) {
              if (static_cast<long long>(dx) * dx + static_cast<long long>(dy) * dy + static_cast<long long>(dz) * dz >= D2) continue;
              const int n = g.find(Vox{g.vox[v].x + dx, g.vox[v].y + dy, g.vox[v].z + dz});
              if (n < 0) continue;
              for (int oc : owners[n]) {
                int a = findRoot(parent, c), b = findRoot(parent, oc);
                if (a != b) parent[std::max(a, b)] = std::min(a, b);  // lowest index survives
              }
            }
  }
  std::vector<int> merged_of(C, -1);
  std::vector<RawCluster> merged;
  for (int c = 0; c < C; ++c) {
    const int r = findRoot(parent, c);
    if (merged_of[r] < 0) {
      merged_of[r] = static_cast<int>(merged.size());
      merged.emplace_back();
    }
    RawCluster& m = merged[merged_of[r]];
    m.pix_groups.insert(m.pix_groups.end(), raw[c].pix_groups.begin(), raw[c].pix_groups.end());
    m.vox_groups.insert(m.vox_groups.end(), raw[c].vox_groups.begin(), raw[c].vox_groups.end());
  }

  // ---- M4: size filter, ids, image, bounding boxes ----
  int id = 1;
  for (auto& m : merged) {
    long long npx = 0;
    for (int v : m.pix_groups) npx += g.count[v];
    if (npx < p.min_cluster_size || npx > p.max_cluster_size) continue;
    std::sort(m.vox_groups.begin(), m.vox_groups.end());
    m.vox_groups.erase(std::unique(m.vox_groups.begin(), m.vox_groups.end()), m.vox_groups.end());
    MotionCluster oc;
    oc.pixels.reserve(static_cast<size_t>(npx) * 2);
    bool first = true;
    for (int v : m.pix_groups)
      for (int k = 0; k < g.count[v]; ++k) {
        const int px = g.pix[g.start[v] + k];
        const int u = px % p.W, vv = px / p.W;
        oc.pixels.push_back(u);
        oc.pixels.push_back(vv);
        dynamic_image[px] = id;
        float w[3];
        if (vertex_world) {
          w[0] = vertex_world[3 * px]; w[1] = vertex_world[3 * px + 1]; w[2] = vertex_world[3 * px + 2];
        } else {
          const float d = depth[px];
          const float x = (static_cast<float>(u) - p.cx) / p.fx * d;
          const float y = (static_cast<float>(vv) - p.cy) / p.fy * d;
          w[0] = ((p.Rw[0] * x + p.Rw[1] * y) + p.Rw[2] * d) + p.tw[0];
          w[1] = ((p.Rw[3] * x + p.Rw[4] * y) + p.Rw[5] * d) + p.tw[1];
          w[2] = ((p.Rw[6] * x + p.Rw[7] * y) + p.Rw[8] * d) + p.tw[2];
        }
        for (int a = 0; a < 3; ++a) {
          oc.bbox[a] = first ? w[a] : std::min(oc.bbox[a], w[a]);
          oc.bbox[3 + a] = first ? w[a] : std::max(oc.bbox[3 + a], w[a]);
        }
        first = false;
      }
    for (int v : m.vox_groups) {  // group ids ascend in (z,y,x)
      oc.voxels.push_back(g.vox[v].x);
      oc.voxels.push_back(g.vox[v].y);
      oc.voxels.push_back(g.vox[v].z);
    }
    out->clusters.emplace_back(std::move(oc));
    if (id < 255) ++id;  // ids saturate at 255 (:390-395)
  }
}

}  // namespace kb
