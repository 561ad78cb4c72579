// Compile-and-link check of the host adaptor against the stub Hydra types. With a GPU present it also
// runs one flat-wall frame through GpuProjectiveIntegrator + GpuTrackingIntegrator + mirrorBack and
// prints a few voxel values that tests/test_host_adaptor.py compares with the oracle.
#include <cstdio>
#include <cmath>
#include <cstdlib>
#include <algorithm>

#include "../../khronos_b200/host/khronos_gpu_adaptor.h"

int main(int argc, char** argv) {
  using namespace khronos_b200;
  hydra::VolumetricMap::Config mc;
  mc.voxel_size = 0.1f; mc.voxels_per_side = 8; mc.truncation_distance = 0.3f;
  mc.with_semantics = true; mc.with_tracking = true;
  kb_integrator_config ic{};
  ic.use_weight_dropoff = 1; ic.weight_dropoff_epsilon = -1.f; ic.max_weight = 1e5f;
  ic.interpolation_method = KB_INTERP_NEAREST; ic.adaptive_max_depth_difference = 0.2f;
  ic.semantic_mode = KB_SEMANTICS_MLE; ic.num_labels = 5; ic.label_confidence = 0.9f;
  kb_tracking_config tc{1.f, 1.f, -1.5f, 18, 3.f, 1};
  kb_motion_config mo{26, 0, 1000000, 1.f, 10000.f, -10000.f, 1};
  khronos::FrameData frame;
  auto& in = frame.input;
  in.sensor = hydra::Camera{64, 48, 32.f, 32.f, 31.5f, 23.5f, 0.1f, 3.f};
  in.timestamp_ns = 1000000000ull;
  in.depth_image = cv::Mat(48, 64, 4);
  in.range_image = cv::Mat(48, 64, 4);
  in.label_image = cv::Mat(48, 64, 4);
  for (int i = 0; i < 48 * 64; ++i) { in.depth_image.ptr<float>()[i] = 2.f; in.range_image.ptr<float>()[i] = 2.f; in.label_image.ptr<int32_t>()[i] = 3; }
  try {
    GpuVolumetricMap gmap(mc, ic, &tc, &mo, 4096);
    GpuProjectiveIntegrator integrator;
    GpuTrackingIntegrator tracking;
    GpuFreeSpaceMotionDetector detector;
    detector.processInput(gmap, frame);
    integrator.updateMap(frame.input, gmap, true, frame.dynamic_image);
    tracking.updateBlocks(frame, gmap);
    hydra::VolumetricMap host(mc);
    mirrorBack(gmap, host, false);
    auto blk = host.getTsdfLayer().getBlockPtr({0, 0, 2});
    if (!blk) { std::printf("missing block\n"); return 2; }
    const auto& v = blk->getVoxel(0 + 8 * (0 + 8 * 3));
    // mesh of the fused wall (GpuMeshIntegrator = hydra::MeshIntegrator::generateMesh on the device): all vertices on z = 2
    hydra::MeshLayer mesh_layer;
    const size_t mesh_vertices = GpuMeshIntegrator().generateMesh(gmap, mesh_layer, true, true);
    float mesh_err = 0.f;
    for (const auto& kv : mesh_layer.blocks)
      for (const auto& p : kv.second.points) mesh_err = std::max(mesh_err, std::fabs(p[2] - 2.f));
    const size_t mesh_again = GpuMeshIntegrator().generateMesh(gmap, mesh_layer, true, true);  // flags were cleared
    std::printf("mesh_vertices=%zu mesh_blocks=%zu mesh_err=%.4f mesh_again=%zu ", mesh_vertices, mesh_layer.numBlocks(), mesh_err, mesh_again);
    // object detector: the whole wall has label 3 -> one cluster when 3 is an object class
    kb_object_detector_config dc{};
    dc.use_full_connectivity = 1; dc.max_cluster_size = -1; dc.use_3d = 1; dc.grid_size = 0.1f; dc.is_object[3] = 1;
    if (argc > 3) {  // third argument: drive GpuActiveWindowCore for a few frames (tests/test_zz_object_detection.py)
      GpuActiveWindowCore::Config cc;
      cc.min_output_separation = 0.25f;
      cc.detect_objects = true;
      cc.object_detector = dc;
      GpuActiveWindowCore core(cc, mc, ic, tc, mo, 4096);
      hydra::VolumetricMap out_map(mc);
      int outputs = 0;
      size_t archived_total = 0;
      for (int k = 0; k < 8; ++k) {
        khronos::FrameData fr = frame;
        fr.dynamic_image = cv::Mat();
        fr.object_image = cv::Mat();
        fr.input.timestamp_ns = 1000000000ull + static_cast<uint64_t>(k) * 100000000ull;  // 10 Hz
        if (core.spinOnce(fr)) {
          hydra::BlockIndices archived;
          core.extractOutputData(out_map, &archived);
          archived_total += archived.size();
          ++outputs;
        }
      }
      core.finishMapping();
      hydra::BlockIndices archived;
      core.extractOutputData(out_map, &archived);
      std::printf("core_frames=%zu core_outputs=%d core_blocks_after_finish=%zu core_archived=%zu ", core.numFramesProcessed(), outputs,
                  out_map.getTsdfLayer().numBlocks(), archived_total + archived.size());
    }
    if (argc > 4) {  // fourth argument: tracker measurements + ray verificator (tests/test_zz_ray_index.py)
      // the left half of the wall is cluster 1, the right half cluster 2; the only track holds cluster 1's voxels
      khronos::FrameData fr = frame;
      cv::Mat ids(48, 64, 4);
      for (int i = 0; i < 48 * 64; ++i) ids.ptr<int32_t>()[i] = (i % 64) < 32 ? 1 : 2;
      TrackMeasurements m0 = measureTracks(gmap, fr, ids, 2, {}, 0.1f, {});
      int32_t offs[3]; int32_t total = 0;
      kb_get_cluster_voxels(gmap.handle(), offs, nullptr, 0, &total);
      std::vector<int64_t> vox(3 * static_cast<size_t>(total));
      kb_get_cluster_voxels(gmap.handle(), offs, vox.data(), total, &total);
      std::vector<int64_t> track(vox.begin(), vox.begin() + 3 * offs[1]);
      TrackMeasurements m1 = measureTracks(gmap, fr, ids, 2, {}, 0.1f, {track});
      std::printf("track_counts=%d,%d track_iou=%.3f,%.3f ", m0.voxel_counts[0], m0.voxel_counts[1], m1.iouOf(0, 0), m1.iouOf(1, 0));
      // one ray from the camera to the wall centre: the hit is present, a point half way is absent, one behind is occluded
      GpuRayVerificator rays(kb_ray_config{1.f, 0.1f, 0.1f});
      const float vertex[3] = {0.f, 0.f, 2.f};
      const uint64_t first_seen[1] = {900000000ull}, last_seen[1] = {1100000000ull};
      auto observed = rays.addVertices(KB_RAYS_MIDDLE, 0.f, {1000000000ull}, {0.f, 0.f, 0.f}, 0, 1, vertex, first_seen, last_seen);
      auto res = rays.check({0.f, 0.f, 2.f, 0.f, 0.f, 1.f, 0.f, 0.f, 2.5f}, {0, 0, 0}, {~0ull, ~0ull, ~0ull});
      std::printf("ray_blocks=%zu ray_verdicts=%zu%zu,%zu%zu,%zu%zu ", observed.size() / 3, res[0].absent.size(), res[0].present.size(),
                  res[1].absent.size(), res[1].present.size(), res[2].absent.size(), res[2].present.size());
    }
    GpuConnectedSemantics object_detector(dc);
    const GpuInstanceForwarding instance_forwarding(kb_instance_forwarding_config{0.f, 0, -1, 0.0, -1.0});  // compiled, not run here
    (void)instance_forwarding;
    if (argc > 2) {  // second argument: also run the object detector (tests/test_zz_object_detection.py)
      object_detector.processInput(gmap, frame);
      std::printf("semantic_clusters=%zu cluster_pixels=%zu ", frame.semantic_clusters.size(),
                  frame.semantic_clusters.empty() ? size_t(0) : frame.semantic_clusters[0].pixels.size());
    }
    // extractor path: private vps-8 binary map around the wall patch in front of the camera, same frame 12 times
    frame.object_image = cv::Mat(48, 64, 4);
    for (int i = 0; i < 48 * 64; ++i) frame.object_image.ptr<int32_t>()[i] = (i % 64) < 32 ? 7 : 0;
    ObjectReconstructionConfig oc;
    oc.projective_integrator = ic;
    oc.min_object_reconstruction_observations = 3;
    const float centre[3] = {0.f, 0.f, 2.f}, dims[3] = {0.5f, 0.5f, 0.25f};
    std::vector<khronos::FrameData> copies(12, frame);
    std::vector<std::pair<const khronos::FrameData*, int>> obs;
    for (size_t k = 0; k < copies.size(); ++k) { copies[k].input.timestamp_ns += k + 1; obs.emplace_back(&copies[k], 7); }
    int erased = -1;
    auto omap = reconstructStaticObject(centre, dims, obs, oc, &erased);
    int32_t oblocks = 0;
    if (omap) kb_num_blocks(omap->handle(), KB_EXPORT_ALL, &oblocks);
    std::printf("object_blocks=%d erased=%d ", oblocks, erased);
    {  // the reference's timer names were opened (timing/stats.csv rows)
      int hits = 0;
      for (const char* name : {"motion_detection/all", "active_window/update_map", "integration/tracking"})
        for (const auto& e : hydra::timing::ElapsedTimeRecorder::instance().entries)
          if (e.name == name) { ++hits; break; }
      std::printf("timers=%d ", hits);
    }
    std::printf("blocks=%zu distance=%.9g weight=%.9g label=%u\n", host.getTsdfLayer().numBlocks(), v.distance, v.weight,
                host.getSemanticLayer()->getBlockPtr({0, 0, 2})->getVoxel(0 + 8 * (0 + 8 * 3)).semantic_label);
  } catch (const std::exception& e) {
    std::printf("no-gpu: %s\n", e.what());
    return argc > 1 ? 1 : 0;  // pass any argument to require a GPU
  }
  return 0;
}
