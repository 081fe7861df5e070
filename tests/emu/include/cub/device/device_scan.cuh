// TEST INFRASTRUCTURE ONLY: host stand-in of the CUB device-wide primitives the library calls (tests/emu/cuda_emu.h)
#pragma once
#include "../cub_emu.cuh"
