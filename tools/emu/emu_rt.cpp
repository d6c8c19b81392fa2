// tools/emu runtime: see emu_cuda.h. DEVELOPMENT TOOL ONLY.
#include "emu_cuda.h"

thread_local uint3 threadIdx, blockIdx;
thread_local dim3 blockDim, gridDim;

namespace emu {
namespace {
struct WarpCtx {
  std::barrier<> bar{32};
  float f[32];
  int i[32];
};
thread_local WarpCtx* tlWarp = nullptr;
thread_local std::barrier<>* tlBlockBar = nullptr;
thread_local int tlLane = 0;

void runThread(dim3 grid, dim3 block, unsigned bx, unsigned by, unsigned bz, unsigned t, const std::function<void()>& body) {
  gridDim = grid; blockDim = block;
  blockIdx = {bx, by, bz};
  threadIdx = {t % block.x, (t / block.x) % block.y, t / (block.x * block.y)};
  tlLane = (int)(t & 31);
  body();
}
}  // namespace

void syncthreads() { tlBlockBar->arrive_and_wait(); }
float shfl_up(float v, unsigned d) {
  WarpCtx* w = tlWarp;
  w->f[tlLane] = v;
  w->bar.arrive_and_wait();
  const float r = tlLane >= (int)d ? w->f[tlLane - d] : v;
  w->bar.arrive_and_wait();
  return r;
}
int shfl_idx(int v, int lane) {
  WarpCtx* w = tlWarp;
  w->i[tlLane] = v;
  w->bar.arrive_and_wait();
  const int r = w->i[lane & 31];
  w->bar.arrive_and_wait();
  return r;
}
float2 ld_volatile(const float2* p) {
  std::this_thread::yield();
  const uint64_t u = __atomic_load_n(reinterpret_cast<const uint64_t*>(p), __ATOMIC_ACQUIRE);
  float2 v;
  std::memcpy(&v, &u, 8);
  return v;
}
void st_volatile(float2* p, float2 v) {
  uint64_t u;
  std::memcpy(&u, &v, 8);
  __atomic_store_n(reinterpret_cast<uint64_t*>(p), u, __ATOMIC_RELEASE);
}

void launch(const char* name, dim3 grid, dim3 block, Mode mode, const std::function<void()>& body) {
  const unsigned nt = block.x * block.y * block.z;
  if (grid.x == 0 || grid.y == 0 || grid.z == 0 || nt == 0 || nt > 1024) {
    std::fprintf(stderr, "emu: bad launch configuration for %s\n", name);
    std::abort();
  }
  if (mode == SEQUENTIAL) {
    for (unsigned bz = 0; bz < grid.z; ++bz)
      for (unsigned by = 0; by < grid.y; ++by)
        for (unsigned bx = 0; bx < grid.x; ++bx)
          for (unsigned t = 0; t < nt; ++t) runThread(grid, block, bx, by, bz, t, body);
    return;
  }
  if ((nt & 31) != 0) { std::fprintf(stderr, "emu: %s needs a block size that is a multiple of 32\n", name); std::abort(); }
  const unsigned nw = nt / 32;
  auto runBlocks = [&](unsigned b0, unsigned b1) {   // blocks [b0, b1) of the linearised grid, all concurrently
    const unsigned nb = b1 - b0;
    std::vector<std::unique_ptr<WarpCtx>> warps(nb * nw);
    std::vector<std::unique_ptr<std::barrier<>>> bars(nb);
    for (auto& w : warps) w = std::make_unique<WarpCtx>();
    for (auto& b : bars) b = std::make_unique<std::barrier<>>(nt);
    std::vector<std::thread> th;
    th.reserve((size_t)nb * nt);
    for (unsigned b = b0; b < b1; ++b) {
      const unsigned bx = b % grid.x, by = (b / grid.x) % grid.y, bz = b / (grid.x * grid.y);
      for (unsigned t = 0; t < nt; ++t)
        th.emplace_back([&, b, bx, by, bz, t] {
          tlWarp = warps[(b - b0) * nw + t / 32].get();
          tlBlockBar = bars[b - b0].get();
          runThread(grid, block, bx, by, bz, t, body);
          tlBlockBar->arrive_and_drop();   // a thread that has left the kernel no longer takes part in barriers
        });
    }
    for (auto& t : th) t.join();
  };
  const unsigned nBlocks = grid.x * grid.y * grid.z;
  if (mode == GRID_WARPS) {
    runBlocks(0, nBlocks);   // inter-block dependencies: everything has to be resident
  } else {
    for (unsigned b = 0; b < nBlocks; ++b) runBlocks(b, b + 1);   // static __shared__ storage: one block at a time
  }
}
}  // namespace emu
