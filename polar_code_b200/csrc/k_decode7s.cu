// decode_kernel<MP, 7, FORCED, METRIC, NS = 7> instantiations: N = 128 exactly, N and log2 N compile-time constants.
#include "polar_kernels.cuh"
#include "polar_launch.h"
using namespace pb;
const void* pb_decode_kernel_7s(int MP, bool forced, bool metric) {
    if (!metric) return forced ? (const void*)decode_kernel<1, 7, true, false, 7> : (const void*)decode_kernel<1, 7, false, false, 7>;
    switch (MP) {
        case 1: return forced ? (const void*)decode_kernel<1, 7, true, true, 7> : (const void*)decode_kernel<1, 7, false, true, 7>;
        case 2: return forced ? (const void*)decode_kernel<2, 7, true, true, 7> : (const void*)decode_kernel<2, 7, false, true, 7>;
        case 4: return forced ? (const void*)decode_kernel<4, 7, true, true, 7> : (const void*)decode_kernel<4, 7, false, true, 7>;
        default: return forced ? (const void*)decode_kernel<8, 7, true, true, 7> : (const void*)decode_kernel<8, 7, false, true, 7>;
    }
}

// trace-recording list decode (info_llrs of all M paths requested): FORCED kernels only, force may be null
const void* pb_decode_kernel_7s_trace(int MP) {
    switch (MP) {
        case 1: return (const void*)decode_kernel<1, 7, true, true, 7, true>;
        case 2: return (const void*)decode_kernel<2, 7, true, true, 7, true>;
        case 4: return (const void*)decode_kernel<4, 7, true, true, 7, true>;
        default: return (const void*)decode_kernel<8, 7, true, true, 7, true>;
    }
}
