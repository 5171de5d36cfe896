// sweep_kernel / dl_retry_kernel<MP, 7> instantiations (N <= 128).
#include "polar_sweep.cuh"
#include "polar_launch.h"
using namespace pb;
const void* pb_sweep_kernel_7(int MP, bool round) {
    switch (MP) {
        case 1: return round ? (const void*)dl_retry_kernel<1, 7> : (const void*)sweep_kernel<1, 7>;
        case 2: return round ? (const void*)dl_retry_kernel<2, 7> : (const void*)sweep_kernel<2, 7>;
        case 4: return round ? (const void*)dl_retry_kernel<4, 7> : (const void*)sweep_kernel<4, 7>;
        default: return round ? (const void*)dl_retry_kernel<8, 7> : (const void*)sweep_kernel<8, 7>;
    }
}
