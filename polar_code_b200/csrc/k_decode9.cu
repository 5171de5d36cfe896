// decode_kernel<MP, 9, FORCED, METRIC> instantiations (N <= 512).
#include "polar_kernels.cuh"
#include "polar_launch.h"
using namespace pb;
const void* pb_decode_kernel_9(int MP, bool forced, bool metric) {
    if (!metric) return forced ? (const void*)decode_kernel<1, 9, true, false> : (const void*)decode_kernel<1, 9, false, false>;
    switch (MP) {
        case 1: return forced ? (const void*)decode_kernel<1, 9, true, true> : (const void*)decode_kernel<1, 9, false, true>;
        case 2: return forced ? (const void*)decode_kernel<2, 9, true, true> : (const void*)decode_kernel<2, 9, false, true>;
        case 4: return forced ? (const void*)decode_kernel<4, 9, true, true> : (const void*)decode_kernel<4, 9, false, true>;
        default: return forced ? (const void*)decode_kernel<8, 9, true, true> : (const void*)decode_kernel<8, 9, false, true>;
    }
}

// trace-recording list decode (info_llrs of all M paths requested): FORCED kernels only, force may be null
const void* pb_decode_kernel_9_trace(int MP) {
    switch (MP) {
        case 1: return (const void*)decode_kernel<1, 9, true, true, 0, true>;
        case 2: return (const void*)decode_kernel<2, 9, true, true, 0, true>;
        case 4: return (const void*)decode_kernel<4, 9, true, true, 0, true>;
        default: return (const void*)decode_kernel<8, 9, true, true, 0, true>;
    }
}
