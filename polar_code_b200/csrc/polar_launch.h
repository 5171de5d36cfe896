// polar_launch.h -- kernel lookup across translation units.  Each k_*.cu instantiates one family of kernels and
// returns host-side function pointers; polar_abi.cu launches them with cudaLaunchKernel.
#pragma once
const void* pb_decode_kernel_7(int MP, bool forced, bool metric);
const void* pb_decode_kernel_9(int MP, bool forced, bool metric);
const void* pb_decode_kernel_7s(int MP, bool forced, bool metric);   // N = 128 exactly (static code length)
const void* pb_sweep_kernel_7(int MP, int kind);   // 0 baseline, 1 DL-SCL retry kernel, 2 baseline + trace, 3 binned DL-SCL retry kernel
const void* pb_sweep_kernel_9(int MP, int kind);
const void* pb_sweep_kernel_7s(int MP, int kind);
const void* pb_decode_kernel_7_trace(int MP);    // list decode + leaf-LLR trace (info_llrs output)
const void* pb_decode_kernel_7s_trace(int MP);
const void* pb_decode_kernel_9_trace(int MP);
const void* pb_channel_kernel(int logmax);
