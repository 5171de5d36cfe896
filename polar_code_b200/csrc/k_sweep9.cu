// sweep_kernel / dl_retry_kernel<MP, 9> instantiations (N <= 512).
#include "polar_sweep.cuh"
#include "polar_launch.h"
using namespace pb;
const void* pb_sweep_kernel_9(int MP, bool round) {
    switch (MP) {
        case 1: return round ? (const void*)dl_retry_kernel<1, 9> : (const void*)sweep_kernel<1, 9>;
        case 2: return round ? (const void*)dl_retry_kernel<2, 9> : (const void*)sweep_kernel<2, 9>;
        case 4: return round ? (const void*)dl_retry_kernel<4, 9> : (const void*)sweep_kernel<4, 9>;
        default: return round ? (const void*)dl_retry_kernel<8, 9> : (const void*)sweep_kernel<8, 9>;
    }
}
