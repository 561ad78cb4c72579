// TEST INFRASTRUCTURE — CPU restatement of khronos::RayVerificator (khronos/src/backend/change_detection/
// ray_verificator.cpp), the oracle of the product's ray index (csrc/kb_rays.cu). Only tests/ may use it.
// Parity: the file is fully in-tree; the reference has no tests for it, so this restatement is checked against an
// independent numpy restatement (tests/test_ray_index_oracle.py) and hand-computed cases.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <set>
#include <string>
#include <vector>

#include "../include/khronos_b200.h"

namespace {

struct V3 { float x, y, z; };
struct BIdx {
  int64_t x, y, z;
  bool operator<(const BIdx& o) const { return z != o.z ? z < o.z : (y != o.y ? y < o.y : x < o.x); }
};

V3 sub(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
// Eigen, unvectorised 3-vectors: squaredNorm = (x^2 + y^2) + z^2; norm = sqrt; normalized = v / norm if squaredNorm > 0
float sqn(V3 a) { return (a.x * a.x + a.y * a.y) + a.z * a.z; }
float norm(V3 a) { return std::sqrt(sqn(a)); }
V3 normalized(V3 a) {
  const float n = sqn(a);
  if (n > 0.f) { const float s = std::sqrt(n); return {a.x / s, a.y / s, a.z / s}; }
  return a;
}
float dot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }

}  // namespace

struct ko_ray_index {
  kb_ray_config cfg{};
  std::vector<V3> src, dst;            // lookup.getSource / getTarget of every ray (current scene graph)
  std::vector<uint64_t> stamp;         // Ray::timestamp
  std::map<BIdx, std::set<size_t>> block_seen_by_rays;  // unordered in the reference; order does not matter
  std::vector<int32_t> ray_pose, ray_vertex;  // Ray::source_node (pose index) / Ray::target_index, -1 for plain rays
  std::vector<uint64_t> result;
  bool have_result = false;
  int64_t entries = 0;

  BIdx toIndex(V3 p) const {  // spatial_hash::Grid(block_size).toIndex: floor(p * (1.f / block_size))
    const float inv = 1.f / cfg.block_size;
    return {static_cast<int64_t>(std::floor(p.x * inv)), static_cast<int64_t>(std::floor(p.y * inv)),
            static_cast<int64_t>(std::floor(p.z * inv))};
  }

  // addRayToHash (:326-350)
  void addRayToHash(size_t ray_index, std::set<BIdx>* observed) {
    const V3 source = src[ray_index], target = dst[ray_index];
    const V3 direction = normalized(sub(target, source));
    const float max_depth = norm(sub(target, source));
    const float ray_step = cfg.block_size / 4;
    float ray_distance = 0.f;
    while (ray_distance <= max_depth) {
      ray_distance += ray_step;
      const V3 ray_point{source.x + ray_distance * direction.x, source.y + ray_distance * direction.y,
                         source.z + ray_distance * direction.z};
      const BIdx index = toIndex(ray_point);
      if (block_seen_by_rays[index].insert(ray_index).second) ++entries;
      if (observed) observed->insert(index);
    }
  }
};

extern "C" {

int ko_rays_create(const kb_ray_config* config, int, ko_ray_index** out) {
  if (!config || !out) return KB_ERR_INVALID;
  *out = nullptr;
  if (!(config->block_size > 0.f) || !(config->radial_tolerance > 0.f) || !(config->depth_tolerance > 0.f)) return KB_ERR_INVALID;
  *out = new ko_ray_index();
  (*out)->cfg = *config;
  return KB_OK;
}

int ko_rays_destroy(ko_ray_index* h) { delete h; return KB_OK; }
const char* ko_rays_last_error(const ko_ray_index*) { return ""; }

int ko_rays_clear(ko_ray_index* h) {
  if (!h) return KB_ERR_INVALID;
  h->src.clear(); h->dst.clear(); h->stamp.clear(); h->block_seen_by_rays.clear(); h->entries = 0; h->have_result = false;
  h->ray_pose.clear(); h->ray_vertex.clear();
  return KB_OK;
}

int ko_rays_size(ko_ray_index* h, int32_t* n_rays, int64_t* n_block_entries) {
  if (!h) return KB_ERR_INVALID;
  if (n_rays) *n_rays = static_cast<int32_t>(h->src.size());
  if (n_block_entries) *n_block_entries = h->entries;
  return KB_OK;
}

int ko_rays_add(ko_ray_index* h, int32_t n, const float* s, const float* t, const uint64_t* ts, int32_t* observed_xyz,
                int32_t max_observed, int32_t* n_observed) {
  if (!h || n < 0 || (n > 0 && (!s || !t || !ts))) return KB_ERR_INVALID;
  if (n_observed) *n_observed = 0;
  for (int i = 0; i < 3 * n; ++i)
    if (!std::isfinite(s[i]) || !std::isfinite(t[i])) return KB_ERR_INVALID;
  // dry run for the observed-block count (the product refuses before adding when the buffer is too small)
  ko_ray_index probe;
  probe.cfg = h->cfg;
  std::set<BIdx> observed;
  for (int i = 0; i < n; ++i) {
    probe.src.push_back({s[3 * i], s[3 * i + 1], s[3 * i + 2]});
    probe.dst.push_back({t[3 * i], t[3 * i + 1], t[3 * i + 2]});
    probe.addRayToHash(static_cast<size_t>(i), &observed);
  }
  if (n_observed) *n_observed = static_cast<int32_t>(observed.size());
  if (observed_xyz) {
    if (static_cast<size_t>(std::max(max_observed, 0)) < observed.size()) return KB_ERR_CAPACITY;
    size_t k = 0;
    for (const BIdx& b : observed) {
      observed_xyz[3 * k] = static_cast<int32_t>(b.x); observed_xyz[3 * k + 1] = static_cast<int32_t>(b.y);
      observed_xyz[3 * k + 2] = static_cast<int32_t>(b.z);
      ++k;
    }
  }
  for (int i = 0; i < n; ++i) {
    h->src.push_back({s[3 * i], s[3 * i + 1], s[3 * i + 2]});
    h->dst.push_back({t[3 * i], t[3 * i + 1], t[3 * i + 2]});
    h->stamp.push_back(ts[i]);
    h->ray_pose.push_back(-1);
    h->ray_vertex.push_back(-1);
    h->addRayToHash(h->src.size() - 1, nullptr);
  }
  h->have_result = false;
  return KB_OK;
}

// addVertices (:222-276) + computeVertexSources (:278-330)
int ko_rays_add_vertices(ko_ray_index* h, int32_t policy, float active_window_duration, int32_t n_poses, const uint64_t* pose_stamps,
                         const float* pose_positions, int32_t n_vertices, int32_t vertex_index_base, const float* vertices,
                         const uint64_t* first_seen_in, const uint64_t* last_seen_in, int32_t* observed_xyz, int32_t max_observed,
                         int32_t* n_observed, int32_t* n_rays_added) {
  if (!h || n_poses < 0 || n_vertices < 0 || policy < KB_RAYS_FIRST || policy > KB_RAYS_ALL) return KB_ERR_INVALID;
  if ((n_poses > 0 && (!pose_stamps || !pose_positions)) || (n_vertices > 0 && (!vertices || !first_seen_in || !last_seen_in))) return KB_ERR_INVALID;
  const std::vector<uint64_t> timestamps_(pose_stamps, pose_stamps + n_poses);
  if (!std::is_sorted(timestamps_.begin(), timestamps_.end())) return KB_ERR_INVALID;
  if (n_observed) *n_observed = 0;
  if (n_rays_added) *n_rays_added = 0;
  std::vector<uint64_t> last_seen(last_seen_in, last_seen_in + n_vertices);
  if (active_window_duration > 0) {
    const uint64_t offset_ns = active_window_duration * 1e9;   // :247
    for (auto& stamp : last_seen) stamp -= offset_ns;
  }
  std::vector<float> s, t;
  std::vector<uint64_t> ts;
  std::vector<int32_t> pose_of, vertex_of;
  for (int i = 0; i < n_vertices; ++i) {
    const size_t first = first_seen_in[i], last = last_seen[i];
    std::set<size_t> result;  // unordered_set in the reference
    if (policy == KB_RAYS_FIRST || policy == KB_RAYS_FIRST_AND_LAST) {
      const auto it = std::upper_bound(timestamps_.begin(), timestamps_.end(), first);
      if (it != timestamps_.end()) result.insert(it - timestamps_.begin());
    }
    if (policy == KB_RAYS_LAST || policy == KB_RAYS_FIRST_AND_LAST) {
      const auto it = std::lower_bound(timestamps_.begin(), timestamps_.end(), last);
      if (it != timestamps_.end()) result.insert(it - timestamps_.begin());
    }
    if (policy == KB_RAYS_MIDDLE) {
      const size_t stamp = (last + first) / 2;
      const auto it = std::lower_bound(timestamps_.begin(), timestamps_.end(), stamp);
      if (it != timestamps_.end()) result.insert(it - timestamps_.begin());
    }
    if (policy == KB_RAYS_ALL) {
      const auto it_lower = std::upper_bound(timestamps_.begin(), timestamps_.end(), first);
      const auto it_upper = std::lower_bound(timestamps_.begin(), timestamps_.end(), last);
      for (auto it = it_lower; it < it_upper; ++it) result.insert(it - timestamps_.begin());
    }
    for (const size_t source_index : result) {
      for (int a = 0; a < 3; ++a) { s.push_back(pose_positions[3 * source_index + a]); t.push_back(vertices[3 * static_cast<size_t>(i) + a]); }
      ts.push_back(timestamps_[source_index]);
      pose_of.push_back(static_cast<int32_t>(source_index));
      vertex_of.push_back(vertex_index_base + i);
    }
  }
  const size_t before = h->src.size();
  const int st = ko_rays_add(h, static_cast<int32_t>(ts.size()), s.data(), t.data(), ts.data(), observed_xyz, max_observed, n_observed);
  if (st != KB_OK) return st;
  for (size_t k = 0; k < ts.size(); ++k) { h->ray_pose[before + k] = pose_of[k]; h->ray_vertex[before + k] = vertex_of[k]; }
  if (n_rays_added) *n_rays_added = static_cast<int32_t>(ts.size());
  return KB_OK;
}

int ko_rays_get_ray_ids(ko_ray_index* h, int32_t* pose_index, int32_t* vertex_index, uint64_t* timestamps, int32_t capacity) {
  if (!h) return KB_ERR_INVALID;
  const size_t n = h->src.size();
  if (static_cast<size_t>(std::max(capacity, 0)) < n) return KB_ERR_CAPACITY;
  if (pose_index) std::copy(h->ray_pose.begin(), h->ray_pose.end(), pose_index);
  if (vertex_index) std::copy(h->ray_vertex.begin(), h->ray_vertex.end(), vertex_index);
  if (timestamps) std::copy(h->stamp.begin(), h->stamp.end(), timestamps);
  return KB_OK;
}

int ko_rays_set_endpoints(ko_ray_index* h, int32_t n_rays, const float* s, const float* t) {
  if (!h || !s || !t || static_cast<size_t>(n_rays) != h->src.size()) return KB_ERR_INVALID;
  for (int i = 0; i < n_rays; ++i) {
    h->src[i] = {s[3 * i], s[3 * i + 1], s[3 * i + 2]};
    h->dst[i] = {t[3 * i], t[3 * i + 1], t[3 * i + 2]};
  }
  h->have_result = false;
  return KB_OK;
}

int ko_rays_rehash(ko_ray_index* h) {  // recomputeHash (:314-324)
  if (!h) return KB_ERR_INVALID;
  h->block_seen_by_rays.clear();
  h->entries = 0;
  for (size_t i = 0; i < h->src.size(); ++i) h->addRayToHash(i, nullptr);
  return KB_OK;
}

int ko_rays_check(ko_ray_index* h, int32_t n_points, const float* p, const uint64_t* earliest, const uint64_t* latest,
                  int32_t* counts, int64_t* total) {
  if (!h || n_points < 0 || (n_points > 0 && (!p || !earliest || !latest || !counts))) return KB_ERR_INVALID;
  h->result.clear();
  for (int i = 0; i < n_points; ++i) {
    // check (:66-146)
    std::vector<uint64_t> absent, present;
    const V3 point{p[3 * i], p[3 * i + 1], p[3 * i + 2]};
    if (!h->src.empty()) {
      const auto it = h->block_seen_by_rays.find(h->toIndex(point));
      if (it != h->block_seen_by_rays.end()) {
        for (size_t ray_index : it->second) {
          const uint64_t timestamp = h->stamp[ray_index];
          if (timestamp < earliest[i] || timestamp > latest[i]) continue;
          const V3 source = h->src[ray_index];
          const V3 direction = normalized(sub(point, source));
          const float depth = norm(sub(point, source));
          const V3 vertex = h->dst[ray_index];
          const float radial_distance = norm(cross(sub(point, source), sub(source, vertex))) / depth;
          if (radial_distance > h->cfg.radial_tolerance) continue;   // no overlap
          const float depth_distance = dot(sub(vertex, source), direction);
          if (depth - depth_distance > h->cfg.depth_tolerance) continue;  // occluded
          if (depth_distance - depth > h->cfg.depth_tolerance) { absent.push_back(timestamp); continue; }
          present.push_back(timestamp);
        }
      }
    }
    std::sort(absent.begin(), absent.end());
    std::sort(present.begin(), present.end());
    counts[2 * i] = static_cast<int32_t>(absent.size());
    counts[2 * i + 1] = static_cast<int32_t>(present.size());
    h->result.insert(h->result.end(), absent.begin(), absent.end());
    h->result.insert(h->result.end(), present.begin(), present.end());
  }
  if (total) *total = static_cast<int64_t>(h->result.size());
  h->have_result = true;
  return KB_OK;
}

int ko_rays_get_stamps(ko_ray_index* h, uint64_t* stamps, int64_t capacity) {
  if (!h) return KB_ERR_INVALID;
  if (!h->have_result) return KB_ERR_STATE;
  if (capacity < static_cast<int64_t>(h->result.size())) return KB_ERR_CAPACITY;
  if (!h->result.empty()) std::memcpy(stamps, h->result.data(), sizeof(uint64_t) * h->result.size());
  return KB_OK;
}

}  // extern "C"
