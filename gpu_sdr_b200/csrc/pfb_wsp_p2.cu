// pfb_fused_wsp_2048_kernel<2, ...>: every instantiation for 2 polyphase tap(s) per channel (see pfb_fused.cuh)
#include "pfb_fused.cuh"

namespace gsdr {
template int pfb_launch_ws<2>(const PfbJob*, int, void*, const float2*, int, cudaStream_t);
}
