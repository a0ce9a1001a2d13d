// tcgen05 (5th-gen tensor core) path — placeholder until the kernel lands; nothing is eligible yet.
#pragma once
#include <cuda_runtime.h>
#include <vector>
#include "fv_kernels.cuh"
#include "fv_model.h"

namespace fv {
struct TcLayer { bool eligible = false; };
struct TcWeights {
  std::vector<TcLayer> layers;
  int build(const std::vector<Layer>& ls, const float* derived, cudaStream_t st) { layers.assign(ls.size(), TcLayer{}); return 0; }
  const TcLayer* layer(int i) const { return i < (int)layers.size() ? &layers[i] : nullptr; }
  void release() { layers.clear(); }
};
inline int launch_conv_tc(const ConvArgs&, const TcLayer&, cudaStream_t) { return 1; }
}  // namespace fv
