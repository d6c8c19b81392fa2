// tools/emu driver: PixFlow (FlowBatch::run, the CUDA source rewritten by emu_prep.py) on host threads.
// DEVELOPMENT TOOL ONLY -- used by tools/emu/check_flow.py to compare a kernel change with the oracle
// before GPU time is spent on it. Not part of the package, the C ABI, the tests or bench.py.
#include "../../surround360_b200/csrc/kernels.h"

namespace s360 {
std::atomic<unsigned long long> g_launchCount{0};
// the flow-only emulation library has no warp_kernels.cu: the kernel profiler is off
thread_local KernelProf* t_kernelProf = nullptr;
ProfScope::ProfScope(cudaStream_t s, const char*, double) : st(s), on(false) {}
ProfScope::~ProfScope() {}
KernelProf::~KernelProf() {}
}

extern "C" int emu_flow_batch(const char* alg, int n, const uint8_t* const* i0, const uint8_t* const* i1, const int* hints, int w, int h,
                              float* flowOut /* [n][h][w][2] */) {
  try {
    s360::FlowParams P;
    if (!s360::flowParamsByName(alg, &P)) return 2;
    s360::FlowBatch fb(n, w, h, P);
    if (std::getenv("S360_EMU_LATENCY")) fb.setLatencyMode(true);   // the cp.async-ring variant of the sweep (copies are synchronous here)
    s360::FlowImagePtrs ip{};
    s360::FlowPrevPtrs pv{};
    for (int k = 0; k < n; ++k) {
      ip.i0[k] = reinterpret_cast<const uchar4*>(i0[k]);
      ip.i1[k] = reinterpret_cast<const uchar4*>(i1[k]);
      ip.pitch0[k] = ip.pitch1[k] = w;
      ip.hint[k] = hints[k];
    }
    fb.run(nullptr, ip, pv, reinterpret_cast<float2*>(flowOut));
    return 0;
  } catch (const std::exception& e) {
    std::fprintf(stderr, "emu_flow_batch: %s\n", e.what());
    return 1;
  }
}
