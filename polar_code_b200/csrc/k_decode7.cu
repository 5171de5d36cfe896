// decode_kernel<MP, 7, FORCED, METRIC> instantiations (N <= 128).
#include "polar_kernels.cuh"
#include "polar_launch.h"
using namespace pb;
const void* pb_decode_kernel_7(int MP, bool forced, bool metric) {
    if (!metric) return forced ? (const void*)decode_kernel<1, 7, true, false> : (const void*)decode_kernel<1, 7, false, false>;
    switch (MP) {
        case 1: return forced ? (const void*)decode_kernel<1, 7, true, true> : (const void*)decode_kernel<1, 7, false, true>;
        case 2: return forced ? (const void*)decode_kernel<2, 7, true, true> : (const void*)decode_kernel<2, 7, false, true>;
        case 4: return forced ? (const void*)decode_kernel<4, 7, true, true> : (const void*)decode_kernel<4, 7, false, true>;
        default: return forced ? (const void*)decode_kernel<8, 7, true, true> : (const void*)decode_kernel<8, 7, false, true>;
    }
}

// trace-recording list decode (info_llrs of all M paths requested): FORCED kernels only, force may be null
const void* pb_decode_kernel_7_trace(int MP) {
    switch (MP) {
        case 1: return (const void*)decode_kernel<1, 7, true, true, 0, true>;
        case 2: return (const void*)decode_kernel<2, 7, true, true, 0, true>;
        case 4: return (const void*)decode_kernel<4, 7, true, true, 0, true>;
        default: return (const void*)decode_kernel<8, 7, true, true, 0, true>;
    }
}
