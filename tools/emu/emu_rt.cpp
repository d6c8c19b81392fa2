// tools/emu runtime: see emu_cuda.h. DEVELOPMENT TOOL ONLY.
// CUDA threads are fibers (ucontext): a block with __syncthreads runs its threads round-robin on one OS thread and
// switches at the barrier; a kernel with warp shuffles / inter-block spin-waits (k_sweep*) gets one OS thread per warp
// whose 32 lane-fibers advance in lock-step at the shuffles, every warp of the grid resident at once.
#include "emu_cuda.h"

#include <ucontext.h>
#include <mutex>

thread_local uint3 threadIdx, blockIdx;
thread_local dim3 blockDim, gridDim;

namespace emu {
namespace {
constexpr size_t kStackBytes = 192 * 1024;
enum State { READY, AT_BARRIER, AT_SHFL, DONE };
struct Fiber {
  ucontext_t ctx;
  std::unique_ptr<char[]> stack;
  State state = READY;
  uint3 tid{}, bid{};
  int lane = 0;
  int warp = 0;          // warp index inside its group
  float fval = 0;        // shuffle operand / result
  int ival = 0;
  unsigned delta = 0;
  int srcLane = 0;
  bool isFloatShfl = true;
  bool down = false;     // __shfl_down_sync
};
struct Group {           // fibers scheduled together on one OS thread: a block (BLOCK_BARRIER) or a warp (GRID_WARPS)
  std::vector<Fiber> fibers;
  ucontext_t sched;
  const std::function<void()>* body = nullptr;
  dim3 grid, block;
};
thread_local Group* tlGroup = nullptr;
thread_local Fiber* tlFiber = nullptr;

void fiberEntry() {
  Fiber* f = tlFiber;
  (*tlGroup->body)();
  f->state = DONE;
  swapcontext(&f->ctx, &tlGroup->sched);
}
void switchToScheduler(State s) {
  Fiber* f = tlFiber;
  f->state = s;
  swapcontext(&f->ctx, &tlGroup->sched);
}
void runGroup(Group& g) {
  tlGroup = &g;
  for (Fiber& f : g.fibers) {
    f.stack.reset(new char[kStackBytes]);
    getcontext(&f.ctx);
    f.ctx.uc_stack.ss_sp = f.stack.get();
    f.ctx.uc_stack.ss_size = kStackBytes;
    f.ctx.uc_link = &g.sched;
    makecontext(&f.ctx, fiberEntry, 0);
  }
  for (;;) {
    bool progressed = false, anyAlive = false;
    for (Fiber& f : g.fibers) {
      if (f.state != READY) { anyAlive |= f.state != DONE; continue; }
      anyAlive = true;
      progressed = true;
      tlFiber = &f;
      gridDim = g.grid; blockDim = g.block; blockIdx = f.bid; threadIdx = f.tid;
      swapcontext(&g.sched, &f.ctx);
    }
    if (!anyAlive) break;
    if (progressed) continue;
    // nobody can run: release a barrier (all live fibers of the block wait) or complete the shuffles of full warps
    bool released = false;
    bool allAtBarrier = true;
    for (Fiber& f : g.fibers) if (f.state != DONE && f.state != AT_BARRIER) allAtBarrier = false;
    if (allAtBarrier) {
      for (Fiber& f : g.fibers) if (f.state == AT_BARRIER) f.state = READY;
      released = true;
    } else {
      // shuffles: per warp, when every live lane waits at a shuffle
      const int nw = ((int)g.fibers.size() + 31) / 32;
      for (int w = 0; w < nw; ++w) {
        const int b = w * 32, e = std::min<int>(b + 32, (int)g.fibers.size());
        bool all = true, any = false;
        for (int i = b; i < e; ++i) {
          if (g.fibers[i].state == DONE) continue;
          any = true;
          if (g.fibers[i].state != AT_SHFL) all = false;
        }
        if (!any || !all) continue;
        float fv[32]; int iv[32];
        for (int i = b; i < e; ++i) { fv[i - b] = g.fibers[i].fval; iv[i - b] = g.fibers[i].ival; }
        for (int i = b; i < e; ++i) {
          Fiber& f = g.fibers[i];
          if (f.state != AT_SHFL) continue;
          const int l = i - b;
          if (f.isFloatShfl) f.fval = f.down ? (l + (int)f.delta < e - b ? fv[l + f.delta] : fv[l]) : (l >= (int)f.delta ? fv[l - f.delta] : fv[l]);
          else f.ival = iv[f.srcLane & 31];
          f.state = READY;
        }
        released = true;
      }
    }
    if (!released) { std::fprintf(stderr, "emu: deadlock inside a block/warp (divergent barrier or shuffle)\n"); std::abort(); }
  }
  tlGroup = nullptr;
  tlFiber = nullptr;
}
Group makeGroup(dim3 grid, dim3 block, const std::function<void()>& body, unsigned bx, unsigned by, unsigned bz, unsigned t0, unsigned t1) {
  Group g;
  g.body = &body; g.grid = grid; g.block = block;
  g.fibers.resize(t1 - t0);
  for (unsigned t = t0; t < t1; ++t) {
    Fiber& f = g.fibers[t - t0];
    f.bid = {bx, by, bz};
    f.tid = {t % block.x, (t / block.x) % block.y, t / (block.x * block.y)};
    f.lane = (int)(t & 31);
  }
  return g;
}
}  // namespace

void syncthreads() { switchToScheduler(AT_BARRIER); }
float shfl_down(float v, unsigned d) {
  Fiber* f = tlFiber;
  f->fval = v; f->delta = d; f->isFloatShfl = true; f->down = true;
  switchToScheduler(AT_SHFL);
  return f->fval;
}
float shfl_up(float v, unsigned d) {
  Fiber* f = tlFiber;
  f->fval = v; f->delta = d; f->isFloatShfl = true; f->down = false;
  switchToScheduler(AT_SHFL);
  return f->fval;
}
int shfl_idx(int v, int lane) {
  Fiber* f = tlFiber;
  f->ival = v; f->srcLane = lane; f->isFloatShfl = false;
  switchToScheduler(AT_SHFL);
  return f->ival;
}
float2 ld_volatile(const float2* p) {
  std::this_thread::yield();   // spin-waits across warps: let the producer's OS thread run
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
  // one launch at a time: the execution model keeps per-launch state in globals (gridDim, static __shared__ arrays), so host threads that
  // drive several contexts at once (the batch driver's chains) take turns
  static std::mutex oneLaunch;
  std::lock_guard<std::mutex> lock(oneLaunch);
  const unsigned nt = block.x * block.y * block.z;
  if (grid.x == 0 || grid.y == 0 || grid.z == 0 || nt == 0 || nt > 1024) {
    std::fprintf(stderr, "emu: bad launch configuration for %s\n", name);
    std::abort();
  }
  if (mode == SEQUENTIAL) {
    gridDim = grid; blockDim = block;
    for (unsigned bz = 0; bz < grid.z; ++bz)
      for (unsigned by = 0; by < grid.y; ++by)
        for (unsigned bx = 0; bx < grid.x; ++bx) {
          blockIdx = {bx, by, bz};
          for (unsigned t = 0; t < nt; ++t) {
            threadIdx = {t % block.x, (t / block.x) % block.y, t / (block.x * block.y)};
            body();
          }
        }
    return;
  }
  if (mode == BLOCK_BARRIER) {   // static __shared__ storage: one block at a time
    for (unsigned bz = 0; bz < grid.z; ++bz)
      for (unsigned by = 0; by < grid.y; ++by)
        for (unsigned bx = 0; bx < grid.x; ++bx) {
          Group g = makeGroup(grid, block, body, bx, by, bz, 0, nt);
          runGroup(g);
        }
    return;
  }
  // GRID_WARPS: one OS thread per warp, every warp of the grid resident (inter-block dependencies)
  if ((nt & 31) != 0) { std::fprintf(stderr, "emu: %s needs a block size that is a multiple of 32\n", name); std::abort(); }
  std::vector<std::thread> th;
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx)
        for (unsigned w = 0; w < nt / 32; ++w)
          th.emplace_back([&, bx, by, bz, w] {
            Group g = makeGroup(grid, block, body, bx, by, bz, w * 32, w * 32 + 32);
            runGroup(g);
          });
  for (auto& t : th) t.join();
}
}  // namespace emu
