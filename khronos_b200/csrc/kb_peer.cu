// Peer-memory frame exchange for the spatially sharded map (no counterpart in the reference, which is
// single-process CPU code). With cell sharding (kb_set_shard_cells) a frame is needed only by the few ranks whose cells
// its frustum touches, and the stream enters the box striped over the GPUs (one PCIe link each), so the exchange is a
// sparse all-to-all: every rank PULLS the frames it needs straight out of the other ranks' frame pools over NVLink.
// Pools are cudaMalloc allocations shared through CUDA IPC handles; they are read-only during the exchange, so the only
// ordering needed is local (gather -> fusion on this rank's streams): no collective, no cross-rank barrier per step.
//
// Three transports behind one plan (a list of contiguous ranges), selectable per run for A/B measurements:
//   KB_GATHER_CE    cudaMemcpyAsync per range: copy engines, no SM time
//   KB_GATHER_SM    persistent CTAs, 16 B loads from the peer mapping / stores to local HBM
//   KB_GATHER_BULK  one elected thread per CTA drives a 4-stage cp.async.bulk pipeline (peer global -> shared, mbarrier
//                   complete_tx; shared -> local global, bulk groups): ~64 KB in flight per CTA with one warp, so a few
//                   dozen single-warp CTAs keep an NVLink port busy while the fusion kernel owns the SMs
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/khronos_b200.h"

namespace {

constexpr int kChunk = 16 * 1024;  // bytes per pipeline stage / work unit
constexpr int kStages = 4;

struct Range {
  const char* src;
  char* dst;
  unsigned long long bytes;        // multiple of 16
  unsigned long long first_chunk;  // prefix sum of chunk counts
};

__device__ __forceinline__ int findRange(const Range* __restrict__ r, int n, unsigned long long chunk) {
  int lo = 0, hi = n - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (r[mid].first_chunk <= chunk) lo = mid; else hi = mid - 1;
  }
  return lo;
}

__global__ void __launch_bounds__(256) gatherSmKernel(const Range* __restrict__ ranges, int n, unsigned long long total_chunks) {
  for (unsigned long long c = blockIdx.x; c < total_chunks; c += gridDim.x) {
    const int i = findRange(ranges, n, c);
    const Range r = ranges[i];
    const unsigned long long off = (c - r.first_chunk) * kChunk;
    const int bytes = static_cast<int>(min(static_cast<unsigned long long>(kChunk), r.bytes - off));
    const uint4* __restrict__ s = reinterpret_cast<const uint4*>(r.src + off);
    uint4* __restrict__ d = reinterpret_cast<uint4*>(r.dst + off);
    const int nv = bytes >> 4;
    // 4 independent 16 B loads per thread in flight (kChunk / 16 = 1024 vectors = 4 per thread)
    uint4 v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int j = threadIdx.x + k * 256;
      if (j < nv) v[k] = __ldcs(s + j);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int j = threadIdx.x + k * 256;
      if (j < nv) d[j] = v[k];
    }
  }
}

__device__ __forceinline__ uint32_t smemAddr(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// One thread per CTA runs the whole pipeline; the other 31 lanes of the warp idle (the copy is done by the TMA unit).
__global__ void __launch_bounds__(32) gatherBulkKernel(const Range* __restrict__ ranges, int n, unsigned long long total_chunks) {
  extern __shared__ __align__(128) unsigned char s_buf[];  // kStages x kChunk
  __shared__ __align__(8) unsigned long long s_bar[kStages];
  if (threadIdx.x != 0) return;
  for (int st = 0; st < kStages; ++st)
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smemAddr(&s_bar[st])));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  // work units of this CTA: chunk ids blockIdx.x, +gridDim.x, ...
  unsigned long long issue = blockIdx.x, drain = blockIdx.x;
  unsigned long long n_issued = 0, n_drained = 0;
  auto load = [&](unsigned long long c, int st) {
    const int i = findRange(ranges, n, c);
    const unsigned long long off = (c - ranges[i].first_chunk) * kChunk;
    const uint32_t bytes = static_cast<uint32_t>(min(static_cast<unsigned long long>(kChunk), ranges[i].bytes - off));
    const uint32_t bar = smemAddr(&s_bar[st]);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smemAddr(s_buf + static_cast<size_t>(st) * kChunk)),
                 "l"(ranges[i].src + off), "r"(bytes), "r"(bar)
                 : "memory");
  };
  // prologue: fill the stages
  for (; n_issued < kStages && issue < total_chunks; ++n_issued, issue += gridDim.x) load(issue, static_cast<int>(n_issued % kStages));
  while (drain < total_chunks) {
    const int st = static_cast<int>(n_drained % kStages);
    const uint32_t parity = static_cast<uint32_t>((n_drained / kStages) & 1ull);
    const uint32_t bar = smemAddr(&s_bar[st]);
    uint32_t done = 0;
    while (!done) {
      asm volatile(
          "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
          : "=r"(done)
          : "r"(bar), "r"(parity)
          : "memory");
    }
    const int i = findRange(ranges, n, drain);
    const unsigned long long off = (drain - ranges[i].first_chunk) * kChunk;
    const uint32_t bytes = static_cast<uint32_t>(min(static_cast<unsigned long long>(kChunk), ranges[i].bytes - off));
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(ranges[i].dst + off),
                 "r"(smemAddr(s_buf + static_cast<size_t>(st) * kChunk)), "r"(bytes)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    ++n_drained;
    drain += gridDim.x;
    if (issue < total_chunks) {
      // the stage being refilled is the one just handed to the store: wait until the store has READ it
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      load(issue, static_cast<int>(n_issued % kStages));
      ++n_issued;
      issue += gridDim.x;
    }
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

thread_local std::string g_err;
int fail(int code, const char* what, cudaError_t e) {
  g_err = std::string(what) + ": " + cudaGetErrorString(e);
  return code;
}
#define KBP(call)                                              \
  do {                                                         \
    cudaError_t e_ = (call);                                   \
    if (e_ != cudaSuccess) return fail(KB_ERR_CUDA, #call, e_); \
  } while (0)

}  // namespace

struct kb_gather_plan {
  int device = 0;
  std::vector<Range> host;
  Range* dev = nullptr;
  unsigned long long total_chunks = 0, total_bytes = 0;
};

extern "C" {

const char* kb_peer_last_error(void) { return g_err.c_str(); }

int kb_peer_alloc(int device, size_t bytes, void** ptr) {
  if (!ptr) return KB_ERR_INVALID;
  KBP(cudaSetDevice(device));
  KBP(cudaMalloc(ptr, std::max<size_t>(bytes, 16)));
  return KB_OK;
}

int kb_peer_free(int device, void* ptr) {
  KBP(cudaSetDevice(device));
  KBP(cudaFree(ptr));
  return KB_OK;
}

int kb_peer_export(int device, void* ptr, uint8_t handle[64]) {
  if (!ptr || !handle) return KB_ERR_INVALID;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
  KBP(cudaSetDevice(device));
  cudaIpcMemHandle_t hdl;
  KBP(cudaIpcGetMemHandle(&hdl, ptr));
  std::memcpy(handle, &hdl, 64);
  return KB_OK;
}

int kb_peer_open(int device, const uint8_t handle[64], void** mapped) {
  if (!handle || !mapped) return KB_ERR_INVALID;
  KBP(cudaSetDevice(device));
  cudaIpcMemHandle_t hdl;
  std::memcpy(&hdl, handle, 64);
  KBP(cudaIpcOpenMemHandle(mapped, hdl, cudaIpcMemLazyEnablePeerAccess));
  return KB_OK;
}

int kb_peer_close(int device, void* mapped) {
  KBP(cudaSetDevice(device));
  KBP(cudaIpcCloseMemHandle(mapped));
  return KB_OK;
}

int kb_peer_enable_access(int device, int peer_device) {
  KBP(cudaSetDevice(device));
  int can = 0;
  KBP(cudaDeviceCanAccessPeer(&can, device, peer_device));
  if (!can) { g_err = "devices are not P2P capable"; return KB_ERR_NO_DEVICE; }
  const cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
  if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail(KB_ERR_CUDA, "cudaDeviceEnablePeerAccess", e);
  (void)cudaGetLastError();
  return KB_OK;
}

int kb_gather_plan_create(int device, int32_t n, const void* const* src, void* const* dst, const uint64_t* bytes, kb_gather_plan** out) {
  if (!out || n < 0 || (n > 0 && (!src || !dst || !bytes))) return KB_ERR_INVALID;
  auto* p = new kb_gather_plan();
  p->device = device;
  for (int i = 0; i < n; ++i) {
    if (bytes[i] == 0) continue;
    if ((bytes[i] & 15ull) || (reinterpret_cast<uintptr_t>(src[i]) & 15u) || (reinterpret_cast<uintptr_t>(dst[i]) & 15u)) {
      delete p;
      g_err = "gather ranges must be 16-byte aligned and sized";
      return KB_ERR_INVALID;
    }
    Range r;
    r.src = static_cast<const char*>(src[i]);
    r.dst = static_cast<char*>(dst[i]);
    r.bytes = bytes[i];
    r.first_chunk = p->total_chunks;
    p->total_chunks += (bytes[i] + kChunk - 1) / kChunk;
    p->total_bytes += bytes[i];
    p->host.push_back(r);
  }
  if (!p->host.empty()) {
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&p->dev), sizeof(Range) * p->host.size());
    if (e == cudaSuccess) e = cudaMemcpy(p->dev, p->host.data(), sizeof(Range) * p->host.size(), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { delete p; return fail(KB_ERR_CUDA, "gather plan upload", e); }
  }
  *out = p;
  return KB_OK;
}

int kb_gather_plan_destroy(kb_gather_plan* p) {
  if (!p) return KB_OK;
  cudaSetDevice(p->device);
  cudaFree(p->dev);
  delete p;
  return KB_OK;
}

uint64_t kb_gather_plan_bytes(const kb_gather_plan* p) { return p ? p->total_bytes : 0; }

int kb_gather_run(kb_gather_plan* p, int mode, int max_ctas, void* cuda_stream) {
  if (!p) return KB_ERR_INVALID;
  if (p->host.empty()) return KB_OK;
  KBP(cudaSetDevice(p->device));
  cudaStream_t s = static_cast<cudaStream_t>(cuda_stream);
  if (mode == KB_GATHER_CE) {
    for (const Range& r : p->host) KBP(cudaMemcpyAsync(r.dst, r.src, r.bytes, cudaMemcpyDefault, s));
    return KB_OK;
  }
  const int n = static_cast<int>(p->host.size());
  const int ctas = static_cast<int>(std::min<unsigned long long>(p->total_chunks, static_cast<unsigned long long>(std::max(1, max_ctas))));
  if (mode == KB_GATHER_SM) {
    gatherSmKernel<<<ctas, 256, 0, s>>>(p->dev, n, p->total_chunks);
  } else if (mode == KB_GATHER_BULK) {
    static bool attr_set = false;
    if (!attr_set) {
      KBP(cudaFuncSetAttribute(gatherBulkKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kStages * kChunk));
      attr_set = true;
    }
    gatherBulkKernel<<<ctas, 32, kStages * kChunk, s>>>(p->dev, n, p->total_chunks);
  } else {
    g_err = "unknown gather mode";
    return KB_ERR_INVALID;
  }
  KBP(cudaGetLastError());
  return KB_OK;
}

}  // extern "C"
