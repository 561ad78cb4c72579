// CPU check of khronos_b200/host/khronos_gpu_replay.h against khronos_b200/replay.py (tests/test_replay_cpp.py):
// reads a scenario from stdin, prints the layout table, the chunk homes and every rank's step plan.
#include <cstdio>
#include <vector>

#include "../../khronos_b200/host/khronos_gpu_replay.h"

int main() {
  using namespace khronos_b200;
  int F, H, W, world, stripe, n_step;
  if (std::scanf("%d %d %d %d %d %d", &F, &H, &W, &world, &stripe, &n_step) != 6) return 2;
  std::vector<uint8_t> touched(static_cast<size_t>(F) * H * W);
  for (auto& t : touched) { int v; if (std::scanf("%d", &v) != 1) return 2; t = static_cast<uint8_t>(v); }
  std::vector<uint32_t> masks(F);
  for (auto& m : masks) { unsigned v; if (std::scanf("%u", &v) != 1) return 2; m = v; }
  std::vector<int> step(n_step);
  for (auto& s : step) if (std::scanf("%d", &s) != 1) return 2;
  int gx, gy;
  rankGrid(world, &gx, &gy);
  std::printf("grid %d %d\n", gx, gy);
  const auto table = bisectLayout(touched.data(), F, H, W, world);
  std::printf("table");
  for (auto v : table) std::printf(" %d", v);
  std::printf("\n");
  const auto homes = routeHomes(masks.data(), F, world, stripe);
  std::printf("homes");
  for (auto v : homes) std::printf(" %d", v);
  std::printf("\n");
  for (int routed = 0; routed < 2; ++routed)
    for (int r = 0; r < world; ++r) {
      StripedSchedule s(world, r, stripe, routed ? homes : std::vector<int32_t>{});
      const StepPlan p = s.plan(step, masks.data());
      std::printf("plan %d %d %d mine", routed, r, p.n_remote);
      for (const auto& m : p.mine) std::printf(" %d,%d,%d", m.position, m.frame, m.slot);
      std::printf(" ranges");
      for (const auto& g : p.ranges) std::printf(" %d,%d,%d,%d", g.src_rank, g.src_local, g.dst_slot, g.count);
      std::printf(" resident %zu\n", s.resident(F).size());
    }
  return 0;
}
