// igd_walks.cu -- kernels around the warp-per-call walks of igd_walks.cuh (tick axis across the lanes):
// k_rxarb_walk  = receive liveness walk + gate arbitration of one bridge per warp, header words read straight out
//                 of the received packets (what k_rx_track<packets> + k_gate_arbitrate do in two launches with one
//                 thread per channel / bridge);
// k_plan_walk   = the sender walk of one outgoing call per warp (what k_ed137_plan does with one thread per call).
#include "igd_walks.cuh"
#include "igd_kernels.cuh"

namespace {

// warps (= bridges / senders) per block.  One: a walk is a latency-bound chain per warp, so what matters is that the
// warps spread evenly over the SMs (1024 bridges x 1640 ticks, whole gateway call: 0.68 / 0.70 / 0.72 ms with 1 / 2 / 4)
#ifndef IGD_WALK_WARPS
#define IGD_WALK_WARPS 1
#endif
constexpr int kWalkWarps = IGD_WALK_WARPS;

__global__ void __launch_bounds__(kWalkWarps * 32) k_rxarb_walk(const igd_rxarb_args a)
{
    const int b = (int)(blockIdx.x * kWalkWarps + (threadIdx.x >> 5));
    if (b >= a.B) return;            // whole warps leave: every collective below sees 32 lanes
    igd_rxarb_walk(a, b);
}

__global__ void __launch_bounds__(kWalkWarps * 32) k_plan_walk(const igd_ed137_pack_desc d, igd_tx_plan_rec *__restrict__ plan,
                                                               int32_t *__restrict__ last_src)
{
    const int c = (int)(blockIdx.x * kWalkWarps + (threadIdx.x >> 5));
    if (c >= d.C) return;
    igd_plan_walk(d, plan, last_src, c);
}

}  // namespace

cudaError_t igd_k_rxarb_walk(const igd_launch_cfg &c, const igd_rxarb_args &a)
{
    if (a.B <= 0 || a.F <= 0) return cudaSuccess;
    k_rxarb_walk<<<(unsigned)((a.B + kWalkWarps - 1) / kWalkWarps), kWalkWarps * 32, 0, c.stream>>>(a);
    return cudaGetLastError();
}

cudaError_t igd_k_plan_walk(const igd_launch_cfg &c, const igd_ed137_pack_desc &d, igd_tx_plan_rec *plan, int32_t *last_src)
{
    if (d.C <= 0) return cudaSuccess;
    k_plan_walk<<<(unsigned)((d.C + kWalkWarps - 1) / kWalkWarps), kWalkWarps * 32, 0, c.stream>>>(d, plan, last_src);
    return cudaGetLastError();
}
