// tests/emu/cooperative_groups.h -- TEST INFRASTRUCTURE ONLY (see cuda_runtime.h in this directory)
#pragma once
#include "cuda_runtime.h"
namespace cooperative_groups {
struct grid_group {
  void sync() const { emu::sync_grid(); }
  unsigned long long size() const { return (unsigned long long)gridDim.x * blockDim.x; }
  unsigned long long thread_rank() const { return (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; }
};
inline grid_group this_grid() { return grid_group(); }
struct thread_block {
  void sync() const { emu::sync_block(); }
  unsigned size() const { return blockDim.x; }
  unsigned thread_rank() const { return threadIdx.x; }
};
inline thread_block this_thread_block() { return thread_block(); }
}   // namespace cooperative_groups
