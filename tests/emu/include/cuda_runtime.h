// TEST INFRASTRUCTURE ONLY: <cuda_runtime.h> of the host emulation (tests/emu/cuda_emu.h)
#pragma once
#include "../cuda_emu.h"
