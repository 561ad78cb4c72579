// TEST INFRASTRUCTURE — NOT PRODUCT CODE. C API of the CPU oracle, mirroring include/khronos_b200.h
// one-to-one with a ko_ prefix so the same Python harness can drive either side.
#include <algorithm>
#include <cstring>
#include <vector>

#include "oracle.hpp"

using ko::Oracle;

struct ko_handle {
  Oracle* o;
  std::string last_error;
  kb_frame_stats totals{};
  kb_totals64 totals64{};
  size_t pixels = 0;
  std::vector<float> depth_f;
  std::vector<int32_t> label_i;
  // sharded protocol state
  int cap_pending = 1024, cap_halo = 2048;
  std::vector<int32_t> last_dynamic;  // dynamic image of the last detection (KB_MASK_LAST_DETECTION)
  int32_t last_seeds = 0, last_clusters = 0;
  bool have_dynamic = false;
};

static int fail(ko_handle* h, int code) {
  if (h && h->o) h->last_error = h->o->error();
  return code;
}

extern "C" {

int ko_create(const kb_map_config* map, const kb_integrator_config* integ,
              const kb_tracking_config* trk, const kb_motion_config* mot, int /*device*/,
              ko_handle** out) {
  if (!map || !integ || !out) return KB_ERR_INVALID;
  auto* h = new ko_handle();
  h->o = new Oracle(*map, *integ, trk, mot);
  if (!h->o->ok()) {
    delete h->o;
    delete h;
    return KB_ERR_INVALID;
  }
  *out = h;
  return KB_OK;
}

int ko_destroy(ko_handle* h) {
  if (!h) return KB_OK;
  delete h->o;
  delete h;
  return KB_OK;
}

const char* ko_last_error(const ko_handle* h) { return h ? h->last_error.c_str() : "null handle"; }
int ko_abi_version(void) { return KB_ABI_VERSION; }
int ko_set_stream(ko_handle*, void*) { return KB_OK; }
int ko_synchronize(ko_handle*) { return KB_OK; }

int ko_set_camera(ko_handle* h, const kb_camera* cam) {
  if (!h || !cam) return KB_ERR_INVALID;
  h->o->setCamera(*cam);
  h->pixels = static_cast<size_t>(cam->width) * cam->height;
  return KB_OK;
}

// Compact sensor formats (kb_frame.depth_u16 / label_u8) are expanded exactly like the device does:
// depth = float(u16) * scale, label = int32(u8).
static const kb_frame* expandCompact(ko_handle* h, const kb_frame* f, kb_frame* tmp) {
  if (!f->depth_u16 && !f->label_u8) return f;
  const size_t px = h->pixels;
  *tmp = *f;
  if (f->depth_u16) {
    h->depth_f.resize(px);
    for (size_t i = 0; i < px; ++i) h->depth_f[i] = static_cast<float>(f->depth_u16[i]) * f->depth_u16_scale;
    tmp->depth = h->depth_f.data();
  }
  if (f->label_u8) {
    h->label_i.resize(px);
    for (size_t i = 0; i < px; ++i) h->label_i[i] = static_cast<int32_t>(f->label_u8[i]);
    tmp->label = h->label_i.data();
  }
  return tmp;
}

int ko_integrate_frame(ko_handle* h, const kb_frame* f_in, int allocate_blocks, kb_frame_stats* stats) {
  if (!h || !f_in || (!f_in->depth && !f_in->depth_u16)) return KB_ERR_INVALID;
  kb_frame tmp, tmp2;
  const kb_frame* f = expandCompact(h, f_in, &tmp);
  if (f->mask == KB_MASK_LAST_DETECTION) {  // dynamic image of the last detection kept by the handle
    tmp2 = *f;
    tmp2.mask = h->have_dynamic ? h->last_dynamic.data() : nullptr;
    f = &tmp2;
  }
  kb_frame_stats local{};
  h->o->integrateFrame(*f, allocate_blocks != 0, &local);
  h->totals.blocks_in_frustum += local.blocks_in_frustum;
  h->totals.blocks_allocated += local.blocks_allocated;
  h->totals.blocks_updated += local.blocks_updated;
  h->totals.voxels_updated += local.voxels_updated;
  h->totals.voxels_in_band += local.voxels_in_band;
  h->totals.voxels_semantic += local.voxels_semantic;
  h->totals64.blocks_in_frustum += static_cast<uint64_t>(local.blocks_in_frustum);
  h->totals64.blocks_allocated += static_cast<uint64_t>(local.blocks_allocated);
  h->totals64.blocks_updated += static_cast<uint64_t>(local.blocks_updated);
  h->totals64.voxels_updated += static_cast<uint64_t>(local.voxels_updated);
  h->totals64.voxels_in_band += static_cast<uint64_t>(local.voxels_in_band);
  h->totals64.voxels_semantic += static_cast<uint64_t>(local.voxels_semantic);
  h->totals64.frames += 1;
  if (stats) *stats = local;
  return h->o->ok() ? KB_OK : fail(h, KB_ERR_INVALID);
}

int ko_integrate_frames(ko_handle* h, const kb_frame* frames, int32_t n, int allocate_blocks, kb_frame_stats* stats) {
  if (!h || !frames) return KB_ERR_INVALID;
  kb_frame_stats sum{};
  for (int i = 0; i < n; ++i) {
    kb_frame_stats s{};
    const int st = ko_integrate_frame(h, frames + i, allocate_blocks, &s);
    if (st != KB_OK) return st;
    sum.blocks_in_frustum += s.blocks_in_frustum; sum.blocks_allocated += s.blocks_allocated;
    sum.blocks_updated += s.blocks_updated; sum.voxels_updated += s.voxels_updated;
    sum.voxels_in_band += s.voxels_in_band; sum.voxels_semantic += s.voxels_semantic;
    sum.total_blocks = s.total_blocks;
  }
  if (stats) *stats = sum;
  return KB_OK;
}

int ko_set_culling(ko_handle*, int) { return KB_OK; }
int ko_get_debug_counters(ko_handle*, int32_t* out, int32_t n) { for (int i = 0; i < n; ++i) out[i] = 0; return KB_OK; }

int ko_get_totals(ko_handle* h, kb_frame_stats* t) {
  if (!h || !t) return KB_ERR_INVALID;
  *t = h->totals;
  t->total_blocks = static_cast<int32_t>(h->o->sortedBlocks(KB_EXPORT_ALL).size());
  return KB_OK;
}

int ko_get_totals64(ko_handle* h, kb_totals64* t) {
  if (!h || !t) return KB_ERR_INVALID;
  *t = h->totals64;
  t->block_frame_pairs = t->blocks_in_frustum;  // the oracle visits every selected block (no culling)
  t->total_blocks = h->o->sortedBlocks(KB_EXPORT_ALL).size();
  return KB_OK;
}

// kb_map_checksum restated on the oracle's blocks (definition: include/khronos_b200.h).
static inline uint64_t koMix64(uint64_t k) {
  k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
  return k;
}
int ko_map_checksum(ko_handle* h, uint64_t out[4]) {
  if (!h || !out) return KB_ERR_INVALID;
  uint64_t sum = 0, xr = 0, nb = 0, seen = 0;
  const uint64_t o = 1ull << 20, m = (1ull << 21) - 1ull;
  const int V = h->o->V();
  for (const ko::Block* b : h->o->sortedBlocks(KB_EXPORT_ALL)) {
    const uint64_t key = ((static_cast<uint64_t>(static_cast<int64_t>(b->index.x) + static_cast<int64_t>(o)) & m)) |
                         ((static_cast<uint64_t>(static_cast<int64_t>(b->index.y) + static_cast<int64_t>(o)) & m) << 21) |
                         ((static_cast<uint64_t>(static_cast<int64_t>(b->index.z) + static_cast<int64_t>(o)) & m) << 42);
    ++nb;
    for (int lin = 0; lin < V; ++lin) {
      uint32_t db, wb;
      std::memcpy(&db, &b->distance[lin], 4);
      std::memcpy(&wb, &b->weight[lin], 4);
      const bool has_sem = !b->semantic_empty.empty() && !b->semantic_empty[lin];
      const uint32_t label = has_sem ? b->semantic_label[lin] : 0xFFFFFFFFu;
      const uint64_t stamp = b->last_observed.empty() ? 0ull : b->last_observed[lin];
      uint64_t v = koMix64(key ^ koMix64(static_cast<uint64_t>(lin) + 1ull));
      v = koMix64(v ^ (static_cast<uint64_t>(db) | (static_cast<uint64_t>(wb) << 32)));
      v = koMix64(v ^ static_cast<uint64_t>(label));
      v = koMix64(v ^ stamp);
      sum += v;
      xr ^= v;
      seen += (stamp != 0ull || b->weight[lin] > 0.f) ? 1ull : 0ull;
    }
  }
  out[0] = sum; out[1] = xr; out[2] = nb; out[3] = seen;
  return KB_OK;
}

int ko_forward_instances(ko_handle* h, const kb_instance_forwarding_config* cfg, const kb_frame* f_in, const uint8_t* id_is_background,
                         int32_t n_background, int32_t* object_image_out, int32_t* n_clusters) {
  if (!h || !cfg || !f_in || !object_image_out || (!f_in->depth && !f_in->depth_u16)) return KB_ERR_INVALID;
  kb_frame tmp;
  const kb_frame* f = expandCompact(h, f_in, &tmp);
  const size_t P = h->pixels;
  if (f->label)
    for (size_t i = 0; i < P; ++i)
      if (f->label[i] < 0 || f->label[i] >= KB_MAX_INSTANCE_IDS) return KB_ERR_INVALID;
  h->o->forwardInstances(*cfg, *f, id_is_background, n_background, object_image_out);
  if (n_clusters) *n_clusters = static_cast<int32_t>(h->o->instanceClusters().size());
  return h->o->ok() ? KB_OK : fail(h, KB_ERR_INVALID);
}

int ko_get_instance_clusters(ko_handle* h, int32_t* id_count, float* bbox_min_max, int32_t* pixels_uv, int32_t* n_clusters, int32_t* total_pixels) {
  if (!h) return KB_ERR_INVALID;
  int32_t total = 0;
  size_t c = 0;
  for (const auto& cl : h->o->instanceClusters()) {
    if (id_count) { id_count[2 * c] = cl.id; id_count[2 * c + 1] = static_cast<int32_t>(cl.pixels.size()); }
    if (bbox_min_max) std::memcpy(bbox_min_max + 6 * c, cl.bbox, sizeof(cl.bbox));
    if (pixels_uv)
      for (const auto& p : cl.pixels) { pixels_uv[2 * total] = p.u; pixels_uv[2 * total + 1] = p.v; ++total; }
    else total += static_cast<int32_t>(cl.pixels.size());
    ++c;
  }
  if (n_clusters) *n_clusters = static_cast<int32_t>(c);
  if (total_pixels) *total_pixels = total;
  return KB_OK;
}

int ko_generate_mesh(ko_handle* h, int only_mesh_updated, int clear_updated_flag, float min_weight, int32_t* n_blocks, int64_t* n_vertices) {
  if (!h) return KB_ERR_INVALID;
  h->o->generateMesh(only_mesh_updated != 0, clear_updated_flag != 0, min_weight);
  int64_t nv = 0;
  for (const auto& mb : h->o->mesh()) nv += static_cast<int64_t>(mb.labels.size());
  if (n_blocks) *n_blocks = static_cast<int32_t>(h->o->mesh().size());
  if (n_vertices) *n_vertices = nv;
  return KB_OK;
}

int ko_get_mesh(ko_handle* h, int32_t* block_index_xyz, int64_t* block_vertex_offsets, float* points_xyz, uint8_t* colors_rgb,
                uint32_t* labels, int64_t capacity_vertices) {
  if (!h) return KB_ERR_INVALID;
  int64_t off = 0;
  size_t i = 0;
  for (const auto& mb : h->o->mesh()) {
    const int64_t n = static_cast<int64_t>(mb.labels.size());
    if (off + n > capacity_vertices) return KB_ERR_CAPACITY;
    if (block_index_xyz) { block_index_xyz[3 * i] = mb.index.x; block_index_xyz[3 * i + 1] = mb.index.y; block_index_xyz[3 * i + 2] = mb.index.z; }
    if (block_vertex_offsets) block_vertex_offsets[i] = off;
    if (points_xyz && n) std::memcpy(points_xyz + 3 * off, mb.points.data(), sizeof(float) * 3 * n);
    if (colors_rgb && n) std::memcpy(colors_rgb + 3 * off, mb.colors.data(), 3 * n);
    if (labels && n) std::memcpy(labels + off, mb.labels.data(), sizeof(uint32_t) * n);
    off += n;
    ++i;
  }
  if (block_vertex_offsets) block_vertex_offsets[i] = off;
  return KB_OK;
}

int ko_update_tracking(ko_handle* h, uint64_t stamp_ns) {
  if (!h) return KB_ERR_INVALID;
  h->o->updateTracking(stamp_ns);
  return h->o->ok() ? KB_OK : fail(h, KB_ERR_INVALID);
}

int ko_detect_objects(ko_handle* h, const kb_object_detector_config* config, const kb_frame* f_in,
                      int32_t* object_image_out, int32_t* n_clusters) {
  if (!h || !config || !f_in || !object_image_out || (!f_in->depth && !f_in->depth_u16)) return KB_ERR_INVALID;
  kb_frame tmp;
  const kb_frame* f = expandCompact(h, f_in, &tmp);
  h->o->detectObjects(*config, *f, object_image_out, nullptr);
  if (n_clusters) *n_clusters = static_cast<int32_t>(h->o->objectClusters().size());
  return h->o->ok() ? KB_OK : fail(h, KB_ERR_INVALID);
}

int ko_get_object_clusters(ko_handle* h, int32_t* id_semantic_count, int32_t* pixels_uv, int32_t* n_clusters,
                           int32_t* total_pixels) {
  if (!h) return KB_ERR_INVALID;
  const auto& cl = h->o->objectClusters();
  size_t tp = 0;
  for (size_t c = 0; c < cl.size(); ++c) {
    if (id_semantic_count) {
      id_semantic_count[c * 3] = cl[c].id; id_semantic_count[c * 3 + 1] = cl[c].semantic_id;
      id_semantic_count[c * 3 + 2] = static_cast<int32_t>(cl[c].pixels.size());
    }
    if (pixels_uv)
      for (size_t i = 0; i < cl[c].pixels.size(); ++i) { pixels_uv[(tp + i) * 2] = cl[c].pixels[i].u; pixels_uv[(tp + i) * 2 + 1] = cl[c].pixels[i].v; }
    tp += cl[c].pixels.size();
  }
  if (n_clusters) *n_clusters = static_cast<int32_t>(cl.size());
  if (total_pixels) *total_pixels = static_cast<int32_t>(tp);
  return KB_OK;
}

int ko_track_measurements(ko_handle* h, const kb_frame* f_in, const int32_t* id_image, int32_t max_id,
                          const int32_t* cluster_ids, float voxel_size, int32_t n_tracks, const int32_t* track_offsets, const int64_t* track_voxels_xyz,
                          int32_t* voxel_counts, int64_t* voxel_sums, int32_t* intersections, float* iou) {
  if (!h || !f_in || !id_image || (!f_in->depth && !f_in->depth_u16 && !f_in->vertex_world)) return KB_ERR_INVALID;
  if (max_id < 1 || max_id > 1022 || !(voxel_size > 0.f) || n_tracks < 0) return KB_ERR_INVALID;
  if (n_tracks > 0 && (!track_offsets || !track_voxels_xyz || track_offsets[0] != 0)) return KB_ERR_INVALID;
  for (int t = 0; t < n_tracks; ++t)
    if (track_offsets[t + 1] < track_offsets[t]) return KB_ERR_INVALID;
  kb_frame tmp;
  const kb_frame* f = expandCompact(h, f_in, &tmp);
  for (int i = 1; cluster_ids && i < max_id; ++i)
    if (cluster_ids[i] <= cluster_ids[i - 1]) return KB_ERR_INVALID;
  h->o->trackMeasurements(*f, id_image, max_id, cluster_ids, voxel_size, n_tracks, track_offsets, track_voxels_xyz);
  if (!h->o->ok()) return fail(h, KB_ERR_STATE);
  const auto& r = h->o->trackResult();
  for (int i = 0; i < max_id; ++i) {
    if (voxel_counts) voxel_counts[i] = static_cast<int32_t>(r.voxels[i].size());
    if (voxel_sums) {
      int64_t s[3] = {0, 0, 0};
      for (const auto& g : r.voxels[i]) { s[0] += g.x; s[1] += g.y; s[2] += g.z; }
      voxel_sums[3 * i] = s[0]; voxel_sums[3 * i + 1] = s[1]; voxel_sums[3 * i + 2] = s[2];
    }
  }
  if (intersections) std::copy(r.intersections.begin(), r.intersections.end(), intersections);
  if (iou) std::copy(r.iou.begin(), r.iou.end(), iou);
  return KB_OK;
}

int ko_compute_vertex_map(ko_handle* h, const kb_frame* f_in, float* out) {
  if (!h || !f_in || !out || (!f_in->depth && !f_in->depth_u16)) return KB_ERR_INVALID;
  kb_frame tmp;
  const kb_frame* f = expandCompact(h, f_in, &tmp);
  return h->o->computeVertexMap(*f, out) ? KB_OK : fail(h, KB_ERR_STATE);
}

int ko_get_cluster_voxels(ko_handle* h, int32_t* offsets, int64_t* voxels_xyz, int32_t capacity, int32_t* total) {
  if (!h) return KB_ERR_INVALID;
  const auto& r = h->o->trackResult();
  if (r.voxels.empty()) return fail(h, KB_ERR_STATE);
  size_t n = 0;
  for (const auto& v : r.voxels) n += v.size();
  if (total) *total = static_cast<int32_t>(n);
  if (offsets) {
    offsets[0] = 0;
    for (size_t i = 0; i < r.voxels.size(); ++i) offsets[i + 1] = offsets[i] + static_cast<int32_t>(r.voxels[i].size());
  }
  if (!voxels_xyz) return KB_OK;
  if (static_cast<size_t>(std::max(capacity, 0)) < n) return fail(h, KB_ERR_CAPACITY);
  size_t k = 0;
  for (const auto& v : r.voxels)
    for (const auto& g : v) { voxels_xyz[3 * k] = g.x; voxels_xyz[3 * k + 1] = g.y; voxels_xyz[3 * k + 2] = g.z; ++k; }
  return KB_OK;
}

// ---- block-hash sharded protocol on host buffers (same layouts as the product's device buffers) ----------

int ko_set_shard(ko_handle* h, int rank, int nranks) {
  if (!h || nranks < 1 || rank < 0 || rank >= nranks) return KB_ERR_INVALID;
  h->o->setShard(rank, nranks);
  return KB_OK;
}

int ko_set_shard_cells(ko_handle* h, int rank, int nranks, int cell_blocks, int grid_x, int grid_y) {
  if (!h || nranks < 1 || rank < 0 || rank >= nranks || cell_blocks < 0 || (cell_blocks > 0 && (grid_x < 1 || grid_y < 1))) return KB_ERR_INVALID;
  if (cell_blocks == 0) h->o->setShard(rank, nranks); else h->o->setShardCells(rank, nranks, cell_blocks, grid_x, grid_y);
  return KB_OK;
}
int ko_set_shard_table(ko_handle* h, int rank, int nranks, int cell_blocks, int32_t origin_cx, int32_t origin_cy, int32_t width,
                       int32_t height, const uint8_t* owners) {
  if (!h || nranks < 1 || nranks > 255 || rank < 0 || rank >= nranks || cell_blocks < 1 || width < 1 || height < 1 || !owners) return KB_ERR_INVALID;
  for (size_t i = 0; i < static_cast<size_t>(width) * height; ++i)
    if (owners[i] >= nranks) return KB_ERR_INVALID;
  h->o->setShardTable(rank, nranks, cell_blocks, origin_cx, origin_cy, width, height, owners);
  return KB_OK;
}
int ko_frame_cells(ko_handle* h, const kb_frame* frames, int32_t n, int cell_blocks, int32_t origin_cx, int32_t origin_cy, int32_t width,
                   int32_t height, uint8_t* touched) {
  if (!h || !frames || !touched || n < 0 || cell_blocks < 1 || width < 1 || height < 1) return KB_ERR_INVALID;
  const size_t cells = static_cast<size_t>(width) * height;
  std::memset(touched, 0, cells * static_cast<size_t>(n));
  for (int i = 0; i < n; ++i) h->o->frameCells(frames[i], cell_blocks, origin_cx, origin_cy, width, height, touched + cells * i);
  return KB_OK;
}
// Handle-free mirrors: a throw-away oracle carries the camera and the layout.
static Oracle* layoutOracle(const kb_camera* camera, float voxel_size, int32_t vps, const kb_shard_layout* layout) {
  kb_map_config mc{};
  mc.voxel_size = voxel_size; mc.voxels_per_side = vps; mc.truncation_distance = 3.f * voxel_size; mc.max_blocks = 1;
  kb_integrator_config ic{};
  ic.max_weight = 1e5f; ic.interpolation_method = KB_INTERP_ADAPTIVE; ic.adaptive_max_depth_difference = 0.2f; ic.num_threads = 1;
  Oracle* o = new Oracle(mc, ic, nullptr, nullptr);
  o->setCamera(*camera);
  if (layout) {
    if (layout->cell_blocks <= 0) o->setShard(0, layout->nranks);
    else if (layout->table) o->setShardTable(0, layout->nranks, layout->cell_blocks, layout->table_origin_cx, layout->table_origin_cy,
                                             layout->table_width, layout->table_height, layout->table);
    else o->setShardCells(0, layout->nranks, layout->cell_blocks, layout->grid_x, layout->grid_y);
  }
  return o;
}
int ko_frame_owners_host(const kb_camera* camera, float voxel_size, int32_t voxels_per_side, const kb_shard_layout* layout,
                         const kb_frame* frames, int32_t n, uint32_t* owner_mask) {
  if (!camera || !layout || !frames || !owner_mask || n < 0 || layout->nranks < 1 || layout->nranks > 32) return KB_ERR_INVALID;
  Oracle* o = layoutOracle(camera, voxel_size, voxels_per_side, layout);
  for (int i = 0; i < n; ++i) owner_mask[i] = o->frameOwners(frames[i]);
  delete o;
  return KB_OK;
}
int ko_frame_cells_host(const kb_camera* camera, float voxel_size, int32_t voxels_per_side, const kb_frame* frames, int32_t n,
                        int cell_blocks, int32_t origin_cx, int32_t origin_cy, int32_t width, int32_t height, uint8_t* touched) {
  if (!camera || !frames || !touched || n < 0 || cell_blocks < 1 || width < 1 || height < 1) return KB_ERR_INVALID;
  Oracle* o = layoutOracle(camera, voxel_size, voxels_per_side, nullptr);
  const size_t cells = static_cast<size_t>(width) * height;
  std::memset(touched, 0, cells * static_cast<size_t>(n));
  for (int i = 0; i < n; ++i) o->frameCells(frames[i], cell_blocks, origin_cx, origin_cy, width, height, touched + cells * i);
  delete o;
  return KB_OK;
}
int ko_cell_owner(int32_t bx, int32_t by, int cell_blocks, int grid_x, int grid_y, int nranks) {
  return Oracle::cellOwner(bx, by, cell_blocks, grid_x, grid_y, nranks);
}
int ko_frame_owners(ko_handle* h, const kb_frame* frames, int32_t n, uint32_t* owner_mask) {
  if (!h || !frames || !owner_mask || n < 0 || h->o->nranks() > 32) return KB_ERR_INVALID;
  for (int i = 0; i < n; ++i) owner_mask[i] = h->o->frameOwners(frames[i]);
  return KB_OK;
}

int ko_block_owner(int32_t bx, int32_t by, int32_t bz, int nranks) { return Oracle::blockOwner(ko::Idx3{bx, by, bz}, nranks); }

int ko_set_shard_capacity(ko_handle* h, int32_t pending_capacity, int32_t halo_capacity) {
  if (!h || pending_capacity <= 0 || halo_capacity <= 0) return KB_ERR_INVALID;
  h->cap_pending = pending_capacity;
  h->cap_halo = halo_capacity;
  return KB_OK;
}

int ko_shard_buffer_sizes(ko_handle* h, int64_t* pending_bytes, int64_t* halo_bytes, int64_t* pixel_flag_bytes) {
  if (!h) return KB_ERR_INVALID;
  if (pending_bytes) *pending_bytes = static_cast<int64_t>(4 + 3 * h->cap_pending) * 4;
  if (halo_bytes) *halo_bytes = static_cast<int64_t>(4 + static_cast<int64_t>(h->cap_halo) * (4 + h->o->V() / 32)) * 4;
  if (pixel_flag_bytes) *pixel_flag_bytes = static_cast<int64_t>(h->pixels);
  return KB_OK;
}

int ko_tracking_begin(ko_handle* h, uint64_t stamp_ns, void* pending_out) {
  if (!h || !pending_out) return KB_ERR_INVALID;
  h->o->trackingBegin(stamp_ns, static_cast<int32_t*>(pending_out), h->cap_pending);
  return h->o->ok() ? KB_OK : fail(h, KB_ERR_INVALID);
}

int ko_tracking_pack_halo(ko_handle* h, const void* all_pending, void* halo_out) {
  if (!h || !all_pending || !halo_out) return KB_ERR_INVALID;
  h->o->packHalo(static_cast<const int32_t*>(all_pending), h->cap_pending, static_cast<int32_t*>(halo_out), h->cap_halo);
  return h->o->ok() ? KB_OK : fail(h, KB_ERR_INVALID);
}

int ko_tracking_finish(ko_handle* h, const void* all_pending, const void* all_halo) {
  if (!h || !all_pending || !all_halo) return KB_ERR_INVALID;
  h->o->trackingFinish(static_cast<const int32_t*>(all_pending), h->cap_pending, static_cast<const int32_t*>(all_halo), h->cap_halo);
  return h->o->ok() ? KB_OK : fail(h, KB_ERR_CAPACITY);
}

int ko_multicast_copy(void* dst, const void* src, size_t bytes, void*) { std::memcpy(dst, src, bytes); return KB_OK; }

// Peer-memory variants on host buffers: compute this shard's part, then store it into slot `rank` of every rank's buffer.
int ko_tracking_begin_peers(ko_handle* h, uint64_t stamp_ns, void* const* peers, int32_t n) {
  if (!h || !peers || n != h->o->nranks()) return KB_ERR_INVALID;
  const size_t stride = 4 + 3 * static_cast<size_t>(h->cap_pending);
  std::vector<int32_t> mine(stride, 0);
  h->o->trackingBegin(stamp_ns, mine.data(), h->cap_pending);
  for (int q = 0; q < n; ++q) std::memcpy(static_cast<int32_t*>(peers[q]) + h->o->rank() * stride, mine.data(), stride * 4);
  return h->o->ok() ? KB_OK : fail(h, KB_ERR_INVALID);
}

int ko_tracking_pack_halo_peers(ko_handle* h, const void* all_pending, void* const* peers, int32_t n) {
  if (!h || !all_pending || !peers || n != h->o->nranks()) return KB_ERR_INVALID;
  const size_t stride = 4 + static_cast<size_t>(h->cap_halo) * (4 + h->o->V() / 32);
  std::vector<int32_t> mine(stride, 0);
  h->o->packHalo(static_cast<const int32_t*>(all_pending), h->cap_pending, mine.data(), h->cap_halo);
  for (int q = 0; q < n; ++q) std::memcpy(static_cast<int32_t*>(peers[q]) + h->o->rank() * stride, mine.data(), stride * 4);
  return h->o->ok() ? KB_OK : fail(h, KB_ERR_INVALID);
}

int ko_motion_lookup_peers(ko_handle* h, const kb_frame* f_in, uint8_t* const* peers, int32_t n) {
  if (!h || !f_in || !peers || n != h->o->nranks()) return KB_ERR_INVALID;
  kb_frame tmp;
  const kb_frame* f = expandCompact(h, f_in, &tmp);
  std::vector<uint8_t> mine(h->pixels, 0);
  h->o->motionLookupLocal(*f, mine.data());
  for (int q = 0; q < n; ++q)
    for (size_t px = 0; px < h->pixels; ++px)
      if (mine[px]) peers[q][px] = mine[px];
  return h->o->ok() ? KB_OK : fail(h, KB_ERR_INVALID);
}

int ko_motion_lookup_local(ko_handle* h, const kb_frame* f_in, uint8_t* pixel_flags) {
  if (!h || !f_in || !pixel_flags) return KB_ERR_INVALID;
  kb_frame tmp;
  const kb_frame* f = expandCompact(h, f_in, &tmp);
  h->o->motionLookupLocal(*f, pixel_flags);
  return h->o->ok() ? KB_OK : fail(h, KB_ERR_INVALID);
}

int ko_motion_cluster_global(ko_handle* h, const uint8_t* pixel_flags) {
  if (!h || !pixel_flags) return KB_ERR_INVALID;
  h->last_dynamic.assign(h->pixels, 0);
  h->o->motionClusterGlobal(pixel_flags, h->last_dynamic.data(), &h->last_seeds, &h->last_clusters);
  h->have_dynamic = true;
  return h->o->ok() ? KB_OK : fail(h, KB_ERR_INVALID);
}

int ko_motion_result(ko_handle* h, int32_t* dynamic_image_out, int32_t* n_seeds, int32_t* n_clusters) {
  if (!h || !h->have_dynamic) return KB_ERR_STATE;
  if (dynamic_image_out) std::memcpy(dynamic_image_out, h->last_dynamic.data(), sizeof(int32_t) * h->pixels);
  if (n_seeds) *n_seeds = h->last_seeds;
  if (n_clusters) *n_clusters = h->last_clusters;
  return KB_OK;
}

int ko_reset_inactive(ko_handle* h, int32_t* removed_xyz, int32_t max_removed, int32_t* n_removed) {
  if (!h) return KB_ERR_INVALID;
  std::vector<ko::Idx3> removed;
  h->o->resetInactive(&removed);
  if (n_removed) *n_removed = static_cast<int32_t>(removed.size());
  if (removed_xyz) {
    for (int i = 0; i < std::min<int>(max_removed, removed.size()); ++i) {
      removed_xyz[i * 3 + 0] = removed[i].x;
      removed_xyz[i * 3 + 1] = removed[i].y;
      removed_xyz[i * 3 + 2] = removed[i].z;
    }
  }
  return KB_OK;
}

int ko_mark_all_inactive(ko_handle* h) { h->o->markAllInactive(); return KB_OK; }
int ko_clear_updated(ko_handle* h) { h->o->clearUpdated(); return KB_OK; }

int ko_detect_motion(ko_handle* h, const kb_frame* f_in, int32_t* dynamic_image_out, int32_t* n_seeds,
                     int32_t* n_clusters) {
  if (!h || !f_in || !dynamic_image_out) return KB_ERR_INVALID;
  kb_frame tmp;
  const kb_frame* f = expandCompact(h, f_in, &tmp);
  h->o->detectMotion(*f, dynamic_image_out, &h->last_seeds, &h->last_clusters);
  h->last_dynamic.assign(dynamic_image_out, dynamic_image_out + h->pixels);
  h->have_dynamic = true;
  if (n_seeds) *n_seeds = h->last_seeds;
  if (n_clusters) *n_clusters = h->last_clusters;
  return h->o->ok() ? KB_OK : fail(h, KB_ERR_INVALID);
}

int ko_spin_once(ko_handle* h, const kb_frame* f, int32_t* dynamic_image_out, int32_t* n_seeds, int32_t* n_clusters) {
  if (!h || !f) return KB_ERR_INVALID;
  std::vector<int32_t> tmp;
  int32_t* img = dynamic_image_out;
  if (!img) { tmp.assign(h->pixels, 0); img = tmp.data(); }
  int st = ko_detect_motion(h, f, img, n_seeds, n_clusters);
  if (st != KB_OK) return st;
  kb_frame g = *f;
  g.mask = img;
  if ((st = ko_integrate_frame(h, &g, 1, nullptr)) != KB_OK) return st;
  return ko_update_tracking(h, f->stamp_ns);
}

int ko_get_motion_clusters(ko_handle* h, int32_t* counts, int32_t* pixels_uv, int64_t* voxels_xyz,
                           float* bbox_min_max, int32_t* total_pixels, int32_t* total_voxels) {
  if (!h) return KB_ERR_INVALID;
  const auto& cl = h->o->clusters();
  int tp = 0, tv = 0;
  for (size_t c = 0; c < cl.size(); ++c) {
    if (counts) {
      counts[c * 2 + 0] = static_cast<int32_t>(cl[c].pixels.size());
      counts[c * 2 + 1] = static_cast<int32_t>(cl[c].voxels.size());
    }
    if (pixels_uv) {
      for (size_t i = 0; i < cl[c].pixels.size(); ++i) {
        pixels_uv[(tp + i) * 2 + 0] = cl[c].pixels[i].u;
        pixels_uv[(tp + i) * 2 + 1] = cl[c].pixels[i].v;
      }
    }
    if (voxels_xyz) {
      std::vector<ko::GIdx> vs(cl[c].voxels.begin(), cl[c].voxels.end());
      std::sort(vs.begin(), vs.end(), ko::GIdxZyxLess());
      for (size_t i = 0; i < vs.size(); ++i) {
        voxels_xyz[(tv + i) * 3 + 0] = vs[i].x;
        voxels_xyz[(tv + i) * 3 + 1] = vs[i].y;
        voxels_xyz[(tv + i) * 3 + 2] = vs[i].z;
      }
    }
    if (bbox_min_max) {
      for (int a = 0; a < 3; ++a) {
        bbox_min_max[c * 6 + a] = cl[c].bbox_min[a];
        bbox_min_max[c * 6 + 3 + a] = cl[c].bbox_max[a];
      }
    }
    tp += static_cast<int>(cl[c].pixels.size());
    tv += static_cast<int>(cl[c].voxels.size());
  }
  if (total_pixels) *total_pixels = tp;
  if (total_voxels) *total_voxels = tv;
  return KB_OK;
}

int ko_allocate_box(ko_handle* h, const int32_t mn[3], const int32_t mx[3]) {
  if (!h) return KB_ERR_INVALID;
  h->o->allocateBox(mn, mx);
  return KB_OK;
}

int ko_scan_object_confidence(ko_handle* h, float min_confidence, int32_t min_observations,
                              int32_t* n_erased) {
  if (!h) return KB_ERR_INVALID;
  const int n = h->o->scanObjectConfidence(min_confidence, min_observations);
  if (n_erased) *n_erased = n;
  return h->o->ok() ? KB_OK : fail(h, KB_ERR_INVALID);
}

int ko_num_blocks(ko_handle* h, int which, int32_t* n) {
  if (!h || !n) return KB_ERR_INVALID;
  *n = static_cast<int32_t>(h->o->sortedBlocks(which).size());
  return KB_OK;
}

int ko_export_blocks(ko_handle* h, int which, int32_t max_blocks, kb_block_export* out,
                     int32_t* n_written) {
  if (!h || !out) return KB_ERR_INVALID;
  const auto blocks = h->o->sortedBlocks(which);
  const int n = std::min<int>(max_blocks, blocks.size());
  const size_t V = h->o->V(), L = h->o->L();
  for (int i = 0; i < n; ++i) {
    const ko::Block& b = *blocks[i];
    if (out->block_index) {
      out->block_index[i * 3 + 0] = b.index.x;
      out->block_index[i * 3 + 1] = b.index.y;
      out->block_index[i * 3 + 2] = b.index.z;
    }
    if (out->block_flags) {
      out->block_flags[i] = (b.updated ? KB_FLAG_UPDATED : 0) | (b.mesh_updated ? KB_FLAG_MESH_UPDATED : 0) |
                            (b.esdf_updated ? KB_FLAG_ESDF_UPDATED : 0) |
                            (b.tracking_updated ? KB_FLAG_TRACKING_UPDATED : 0) |
                            (b.has_active_data ? KB_FLAG_HAS_ACTIVE_DATA : 0);
    }
    if (out->distance) std::memcpy(out->distance + i * V, b.distance.data(), V * 4);
    if (out->weight) std::memcpy(out->weight + i * V, b.weight.data(), V * 4);
    if (out->color) std::memcpy(out->color + i * V * 3, b.color.data(), V * 3);
    const bool trk = !b.last_observed.empty();
    if (out->last_observed) trk ? (void)std::memcpy(out->last_observed + i * V, b.last_observed.data(), V * 8) : (void)std::memset(out->last_observed + i * V, 0, V * 8);
    if (out->last_occupied) trk ? (void)std::memcpy(out->last_occupied + i * V, b.last_occupied.data(), V * 8) : (void)std::memset(out->last_occupied + i * V, 0, V * 8);
    if (out->ever_free) trk ? (void)std::memcpy(out->ever_free + i * V, b.ever_free.data(), V) : (void)std::memset(out->ever_free + i * V, 0, V);
    if (out->active) trk ? (void)std::memcpy(out->active + i * V, b.active.data(), V) : (void)std::memset(out->active + i * V, 0, V);
    if (out->to_remove) trk ? (void)std::memcpy(out->to_remove + i * V, b.to_remove.data(), V) : (void)std::memset(out->to_remove + i * V, 0, V);
    const bool sem = L > 0;
    if (out->semantic_label) sem ? (void)std::memcpy(out->semantic_label + i * V, b.semantic_label.data(), V * 4) : (void)std::memset(out->semantic_label + i * V, 0, V * 4);
    if (out->semantic_empty) sem ? (void)std::memcpy(out->semantic_empty + i * V, b.semantic_empty.data(), V) : (void)std::memset(out->semantic_empty + i * V, 1, V);
    if (out->semantic_likelihoods && sem) {
      if (b.likelihoods.empty()) std::memset(out->semantic_likelihoods + i * V * L, 0, V * L * 4);
      else std::memcpy(out->semantic_likelihoods + i * V * L, b.likelihoods.data(), V * L * 4);
    }
  }
  if (n_written) *n_written = n;
  return KB_OK;
}

}  // extern "C"
