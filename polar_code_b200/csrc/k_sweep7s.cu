// sweep_kernel / dl_retry_kernel<MP, 7, ..., NS = 7> instantiations: N = 128 exactly (static code length).
// kind 0: baseline sweep, 1: DL-SCL retry kernel (frame per group), 2: baseline sweep that records the leaf-LLR trace (retries follow),
// 3: DL-SCL retry kernel, binned by flip position (prefix skipping).
#include "polar_sweep.cuh"
#include "polar_launch.h"
using namespace pb;
template <int MP> static const void* pick(int kind) {
    return kind == 3 ? (const void*)dl_bin_kernel<MP, 7, 7> : kind == 1 ? (const void*)dl_retry_kernel<MP, 7, 7> : kind == 2 ? (const void*)sweep_kernel<MP, 7, true, 7> : (const void*)sweep_kernel<MP, 7, false, 7>;
}
const void* pb_sweep_kernel_7s(int MP, int kind) {
    switch (MP) {
        case 1: return pick<1>(kind);
        case 2: return pick<2>(kind);
        case 4: return pick<4>(kind);
        default: return pick<8>(kind);
    }
}
