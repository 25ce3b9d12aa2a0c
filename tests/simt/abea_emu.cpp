/* abea_emu.cpp — TEST INFRASTRUCTURE: the product's host + kernel sources compiled for the CPU SIMT emulator.
 * Produces tests/simt/libabea_emu.so with the same C ABI as libabea_b200.so; loaded only by "not gpu" tests. */
#define ABEA_SIMT_EMU 1
#include "simt_emu.h"
#include "../../f5c_b200/csrc/abea_host.cu"
