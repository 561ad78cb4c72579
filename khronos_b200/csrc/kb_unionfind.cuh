// Lock-free union-find shared by the device-side connected-component passes (motion clustering, kb_motion_device.cu;
// semantic object detection, kb_objects_device.cu).
#pragma once

namespace kb {

__device__ __forceinline__ int ufFind(int* parent, int i) {
  int p = parent[i];
  while (p != i) {
    const int g = parent[p];
    parent[i] = g;  // path halving (benign race: always points to an ancestor)
    i = p;
    p = g;
  }
  return i;
}

__device__ __forceinline__ void ufUnion(int* parent, int a, int b) {
  for (;;) {
    a = ufFind(parent, a);
    b = ufFind(parent, b);
    if (a == b) return;
    if (a < b) { const int t = a; a = b; b = t; }  // hook the larger root under the smaller
    if (atomicCAS(&parent[a], a, b) == a) return;
  }
}

}  // namespace kb
