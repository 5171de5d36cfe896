// channel_kernel<LOGMAX> instantiations.
#include "polar_sweep.cuh"
#include "polar_launch.h"
using namespace pb;
const void* pb_channel_kernel(int logmax) {
    return logmax <= 7 ? (const void*)channel_kernel<7> : (const void*)channel_kernel<9>;
}
