// TEST TOOLING — NOT PRODUCT CODE. Fiber scheduler of the CUDA-on-CPU shim (see include/cuda_runtime.h).
#include <cstdio>
#include <vector>

#include "cuda_runtime.h"

uint3 threadIdx, blockIdx;
dim3 blockDim, gridDim;

namespace emu {
namespace {

enum State { kRunnable, kWaitWarp, kWaitBlock, kDone };
constexpr size_t kStack = 256 * 1024;

// Minimal x86-64 System V context switch (callee-saved registers + stack pointer): ucontext's swapcontext makes two
// signal-mask system calls per switch, which dominated the run time.
extern "C" void emu_switch(void** save_sp, void* load_sp);
asm(R"(
.text
.globl emu_switch
.type emu_switch,@function
emu_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size emu_switch,.-emu_switch
)");

struct Fiber {
  void* sp = nullptr;
  char* stack = nullptr;
  State state = kDone;
  uint3 tid{0, 0, 0};
  // A released lane may reach its next exchange before a sibling has read this one's value: values are double
  // buffered by the lane's exchange count (a lane can never be more than one exchange ahead of its warp / block).
  unsigned long long slot[2] = {0, 0};  // values offered to warp exchanges
  unsigned wgen = 0;
  int red[2] = {0, 0};                  // values offered to block reductions
  unsigned bgen = 0;
};

std::vector<Fiber> g_fibers;
void* g_sched_sp = nullptr;
int g_current = -1, g_nthreads = 0;
std::function<void()>* g_body = nullptr;
std::vector<char> g_smem;
std::vector<unsigned> g_warp_mask;  // participants of the exchange a warp was last released from
// Scheduling order of the fibers within a pass (and of the blocks): KB_EMU_ORDER = forward (default) | reverse |
// shuffle (a fresh pseudo-random permutation per pass, seeded by KB_EMU_SEED): different extremes of what a GPU may do.
const bool g_reverse = [] { const char* e = std::getenv("KB_EMU_ORDER"); return e && e[0] == 'r'; }();
const bool g_shuffle = [] { const char* e = std::getenv("KB_EMU_ORDER"); return e && e[0] == 's'; }();
unsigned long long g_rng = [] { const char* e = std::getenv("KB_EMU_SEED"); return 0x9E3779B97F4A7C15ull ^ (e ? std::strtoull(e, nullptr, 10) : 1ull); }();
inline unsigned rnd(unsigned n) {  // xorshift64*
  g_rng ^= g_rng >> 12; g_rng ^= g_rng << 25; g_rng ^= g_rng >> 27;
  return static_cast<unsigned>(((g_rng * 0x2545F4914F6CDD1Dull) >> 33) % n);
}
std::vector<int> g_order;

void trampoline() {
  (*g_body)();
  g_fibers[g_current].state = kDone;
  emu_switch(&g_fibers[g_current].sp, g_sched_sp);
  std::abort();  // a finished fiber is never resumed
}

void yield(State st) {
  Fiber& f = g_fibers[g_current];
  f.state = st;
  emu_switch(&f.sp, g_sched_sp);
}

// Prepares a fresh stack so that the first switch into the fiber "returns" into trampoline().
void prepare(Fiber& f) {
  uintptr_t top = (reinterpret_cast<uintptr_t>(f.stack) + kStack) & ~uintptr_t(15);
  void** sp = reinterpret_cast<void**>(top);
  *--sp = nullptr;                                   // fake return address of trampoline (keeps rsp % 16 == 8 at entry)
  *--sp = reinterpret_cast<void*>(&trampoline);      // popped by `ret` in emu_switch
  for (int i = 0; i < 6; ++i) *--sp = nullptr;       // rbp, rbx, r12-r15
  f.sp = sp;
}

void run_block() {
  const int n = g_nthreads, n_warps = (n + 31) / 32;
  g_warp_mask.assign(n_warps, 0);
  for (int i = 0; i < n; ++i) {
    Fiber& f = g_fibers[i];
    if (!f.stack) f.stack = static_cast<char*>(std::malloc(kStack));
    prepare(f);
    f.state = kRunnable;
    f.wgen = f.bgen = 0;
    f.tid = {static_cast<unsigned>(i % blockDim.x), static_cast<unsigned>((i / blockDim.x) % blockDim.y),
             static_cast<unsigned>(i / (blockDim.x * blockDim.y))};
  }
  int remaining = n;
  while (remaining > 0) {
    bool progress = false;
    if (g_shuffle) {
      g_order.resize(n);
      for (int k = 0; k < n; ++k) g_order[k] = k;
      for (int k = n - 1; k > 0; --k) std::swap(g_order[k], g_order[rnd(k + 1)]);
    }
    for (int k = 0; k < n; ++k) {
      const int i = g_shuffle ? g_order[k] : (g_reverse ? n - 1 - k : k);
      if (g_fibers[i].state != kRunnable) continue;
      g_current = i;
      threadIdx = g_fibers[i].tid;
      emu_switch(&g_sched_sp, g_fibers[i].sp);
      progress = true;
      if (g_fibers[i].state == kDone) --remaining;
    }
    // release warps whose live lanes have all arrived at an exchange
    for (int w = 0; w < n_warps; ++w) {
      unsigned live = 0, waiting = 0;
      for (int l = 0; l < 32 && w * 32 + l < n; ++l) {
        const State st = g_fibers[w * 32 + l].state;
        if (st != kDone) live |= 1u << l;
        if (st == kWaitWarp) waiting |= 1u << l;
      }
      if (waiting && waiting == live) {
        g_warp_mask[w] = waiting;
        for (int l = 0; l < 32 && w * 32 + l < n; ++l)
          if (g_fibers[w * 32 + l].state == kWaitWarp) g_fibers[w * 32 + l].state = kRunnable;
        progress = true;
      }
    }
    // release the block barrier when every live thread waits at it
    int live = 0, at_block = 0;
    for (int i = 0; i < n; ++i) {
      if (g_fibers[i].state != kDone) ++live;
      if (g_fibers[i].state == kWaitBlock) ++at_block;
    }
    if (live > 0 && at_block == live) {
      for (int i = 0; i < n; ++i)
        if (g_fibers[i].state == kWaitBlock) g_fibers[i].state = kRunnable;
      progress = true;
    }
    if (!progress && remaining > 0) {
      std::fprintf(stderr, "cuda_emu: deadlock (divergent synchronisation) in block (%u,%u)\n", blockIdx.x, blockIdx.y);
      std::abort();
    }
  }
}

}  // namespace

void launch_(std::function<void()> body, dim3 grid, dim3 block, size_t smem, cudaStream_t) {
  g_body = &body;
  gridDim = grid;
  blockDim = block;
  g_nthreads = static_cast<int>(block.x * block.y * block.z);
  if (static_cast<int>(g_fibers.size()) < g_nthreads) g_fibers.resize(g_nthreads);
  g_smem.assign(smem + 16, 0);
  const unsigned long long n_blocks = static_cast<unsigned long long>(grid.x) * grid.y * grid.z;
  for (unsigned long long k = 0; k < n_blocks; ++k) {
    const unsigned long long b = g_reverse ? n_blocks - 1 - k : k;
    blockIdx = {static_cast<unsigned>(b % grid.x), static_cast<unsigned>((b / grid.x) % grid.y),
                static_cast<unsigned>(b / (static_cast<unsigned long long>(grid.x) * grid.y))};
    run_block();
  }
  g_body = nullptr;
}

void* dyn_smem() { return g_smem.data(); }

unsigned active_mask() {
  const int w = g_current / 32;
  unsigned live = 0;
  for (int l = 0; l < 32 && w * 32 + l < g_nthreads; ++l)
    if (g_fibers[w * 32 + l].state != kDone) live |= 1u << l;
  return live;
}

unsigned long long warp_exchange(unsigned long long v, int src_lane, int mode, int arg) {
  const int me = g_current, w = me / 32, lane = me % 32;
  const unsigned gen = g_fibers[me].wgen++ & 1u;
  g_fibers[me].slot[gen] = v;
  yield(kWaitWarp);
  const unsigned part = g_warp_mask[w];
  auto slot_of = [&](int l) { return g_fibers[w * 32 + l].slot[gen]; };
  switch (mode) {
    case kShfl: return slot_of(src_lane & 31);
    case kShflXor: { const int s = lane ^ arg; return (w * 32 + s < g_nthreads) ? slot_of(s) : v; }
    case kShflUp: { const int s = lane - arg; return s >= 0 ? slot_of(s) : v; }
    case kBallot: {
      unsigned m = 0;
      for (int l = 0; l < 32; ++l)
        if (((part >> l) & 1u) && slot_of(l)) m |= 1u << l;
      return m;
    }
    default: {  // kAny
      for (int l = 0; l < 32; ++l)
        if (((part >> l) & 1u) && slot_of(l)) return 1;
      return 0;
    }
  }
}

void block_barrier() { yield(kWaitBlock); }

int block_reduce(int v, int mode) {
  const unsigned gen = g_fibers[g_current].bgen++ & 1u;
  g_fibers[g_current].red[gen] = v;
  yield(kWaitBlock);
  // every thread of the block takes part in these reductions (CUDA requires it)
  int r = mode == 0 ? 1 : 0;
  for (int i = 0; i < g_nthreads; ++i) {
    if (mode == 0) r = r && (g_fibers[i].red[gen] != 0);
    else r = r || (g_fibers[i].red[gen] != 0);
  }
  return r;
}

}  // namespace emu
