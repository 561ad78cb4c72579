// Host-side scheduler of the cell-sharded replay (C++17, header only): the same three pure functions as
// khronos_b200/replay.py — which rank owns which cells (bisectLayout), where the chunks of the stream live (routeHomes) and
// what every rank integrates and pulls per step (StripedSchedule::plan) — for a C++ host such as an offline Khronos replay /
// khronos_eval run on several GPUs (the reference has no multi-GPU path). Inputs are what the C ABI returns:
// kb_frame_cells (touched) and kb_frame_owners (owner masks). Outputs feed kb_set_shard_table, kb_gather_plan_create and the
// per-rank kb_integrate_frames calls. tests/test_replay_cpp.py checks these against the Python versions.
#pragma once

#include <algorithm>
#include <cstdint>
#include <vector>

namespace khronos_b200 {

// gx x gy = world with gx >= gy as square as possible (8 -> 4 x 2, 6 -> 3 x 2).
inline void rankGrid(int world, int* gx, int* gy) {
  int g = 1;
  for (int k = 1; k * k <= world; ++k)
    if (world % k == 0) g = k;
  *gy = g;
  *gx = world / g;
}

// Trajectory-aware cell -> rank table: touched[f * H * W + cy * W + cx] != 0 iff frame f touches cell (cx, cy). A rectangle
// that gets k ranks is cut, along x or y, where the busier side's frames-per-rank is smallest (a frame on the cut counts on
// both sides); recursion until every rectangle has one rank. Returns table[cy * W + cx].
inline std::vector<uint8_t> bisectLayout(const uint8_t* touched, int F, int H, int W, int world) {
  std::vector<uint8_t> table(static_cast<size_t>(H) * W, 0);
  auto load = [&](int x0, int x1, int y0, int y1) {
    int n = 0;
    for (int f = 0; f < F; ++f) {
      const uint8_t* t = touched + static_cast<size_t>(f) * H * W;
      bool any = false;
      for (int y = y0; y < y1 && !any; ++y)
        for (int x = x0; x < x1; ++x)
          if (t[y * W + x]) { any = true; break; }
      n += any ? 1 : 0;
    }
    return n;
  };
  struct Split {
    static void run(const decltype(load)& ld, std::vector<uint8_t>& tab, int W, int x0, int x1, int y0, int y1, int base, int k) {
      if (k == 1 || (x1 - x0 <= 1 && y1 - y0 <= 1)) {
        for (int y = y0; y < y1; ++y)
          for (int x = x0; x < x1; ++x) tab[static_cast<size_t>(y) * W + x] = static_cast<uint8_t>(base);
        return;
      }
      const int k1 = k / 2, k2 = k - k1;
      double best = -1.0;
      int best_axis = 0, best_c = 0;
      for (int axis = 0; axis < 2; ++axis) {
        const int lo = axis == 0 ? x0 : y0, hi = axis == 0 ? x1 : y1;
        for (int c = lo + 1; c < hi; ++c) {
          const int a = axis == 0 ? ld(x0, c, y0, y1) : ld(x0, x1, y0, c);
          const int b = axis == 0 ? ld(c, x1, y0, y1) : ld(x0, x1, c, y1);
          const double cost = std::max(static_cast<double>(a) / k1, static_cast<double>(b) / k2);
          if (best < 0.0 || cost < best) { best = cost; best_axis = axis; best_c = c; }
        }
      }
      if (best_axis == 0) {
        run(ld, tab, W, x0, best_c, y0, y1, base, k1);
        run(ld, tab, W, best_c, x1, y0, y1, base + k1, k2);
      } else {
        run(ld, tab, W, x0, x1, y0, best_c, base, k1);
        run(ld, tab, W, x0, x1, best_c, y1, base + k1, k2);
      }
    }
  };
  Split::run(load, table, W, 0, W, 0, H, 0, world);
  return table;
}

// Pose-aware placement: every chunk of `stripe` consecutive frames goes to the rank most of its frames touch (ties: the
// rank holding the fewest frames so far, then the lowest rank).
inline std::vector<int32_t> routeHomes(const uint32_t* owner_mask, int n, int world, int stripe) {
  std::vector<int32_t> homes(static_cast<size_t>(n), 0);
  std::vector<int64_t> held(static_cast<size_t>(world), 0);
  for (int c0 = 0; c0 < n; c0 += stripe) {
    const int c1 = std::min(n, c0 + stripe);
    int best_votes = -1, best_rank = 0;
    for (int r = 0; r < world; ++r) {
      int votes = 0;
      for (int g = c0; g < c1; ++g) votes += (owner_mask[g] >> r) & 1u;
      if (votes > best_votes || (votes == best_votes && held[r] < held[best_rank])) { best_votes = votes; best_rank = r; }
    }
    for (int g = c0; g < c1; ++g) homes[g] = best_rank;
    held[best_rank] += c1 - c0;
  }
  return homes;
}

// One rank's share of a step.
struct StepPlan {
  struct Mine { int position, frame, slot; };           // slot >= 0: receive buffer; < 0: own pool, local index = -slot - 1
  struct Range { int src_rank, src_local, dst_slot, count; };
  std::vector<Mine> mine;
  std::vector<Range> ranges;
  int n_remote = 0;
};

class StripedSchedule {
 public:
  // homes empty: chunks of `stripe` frames dealt round robin
  StripedSchedule(int world, int rank, int stripe, std::vector<int32_t> homes = {})
      : world_(world), rank_(rank), stripe_(stripe), homes_(std::move(homes)) {
    if (!homes_.empty()) {
      local_.resize(homes_.size());
      std::vector<int64_t> cnt(static_cast<size_t>(world), 0);
      for (size_t g = 0; g < homes_.size(); ++g) local_[g] = cnt[homes_[g]]++;
    }
  }
  int home(int g) const { return homes_.empty() ? (g / stripe_) % world_ : homes_[g]; }
  int localIndex(int g) const {
    return homes_.empty() ? (g / (stripe_ * world_)) * stripe_ + g % stripe_ : static_cast<int>(local_[g]);
  }
  std::vector<int> resident(int lap) const {
    std::vector<int> out;
    for (int g = 0; g < lap; ++g)
      if (home(g) == rank_) out.push_back(g);
    return out;
  }
  // step_frames[j] = global frame of position j of the step; owner_mask[g] = kb_frame_owners bit mask of frame g
  StepPlan plan(const std::vector<int>& step_frames, const uint32_t* owner_mask) const {
    StepPlan p;
    int slot = 0;
    for (size_t j = 0; j < step_frames.size(); ++j) {
      const int g = step_frames[j];
      if (!((owner_mask[g] >> rank_) & 1u)) continue;
      const int src = home(g), li = localIndex(g);
      if (src == rank_) { p.mine.push_back({static_cast<int>(j), g, -li - 1}); continue; }
      if (!p.ranges.empty() && p.ranges.back().src_rank == src && p.ranges.back().src_local + p.ranges.back().count == li &&
          p.ranges.back().dst_slot + p.ranges.back().count == slot) {
        ++p.ranges.back().count;
      } else {
        p.ranges.push_back({src, li, slot, 1});
      }
      p.mine.push_back({static_cast<int>(j), g, slot});
      ++slot;
    }
    p.n_remote = slot;
    return p;
  }

 private:
  int world_, rank_, stripe_;
  std::vector<int32_t> homes_;
  std::vector<int64_t> local_;
};

}  // namespace khronos_b200
