// stand-in for <cuda_runtime.h> when the kernels are compiled for the host emulator (tools/emu)
#pragma once
#include "../emu_cuda.h"
