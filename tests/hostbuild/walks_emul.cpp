// Host build of igate4xsoftphonedsp_b200/csrc/igd_walks.cuh (-DIGD_HOST_EMUL): the warp-per-call walks run lane for
// lane on 32 fibers (ucontext), ballot / shuffle are rendezvous points of the 32 fibers.  TEST-ONLY: lets the CPU
// suite hold the very code the GPU kernels run against the oracle before a GPU run.  Never loaded by the product.
#define IGD_HOST_EMUL 1
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>

#include <functional>

namespace emu {
constexpr int kLanes = 32;
constexpr size_t kStack = 256 * 1024;
struct Warp {
    ucontext_t main, ctx[kLanes];
    char *stacks;
    bool done[kLanes];
    uint32_t slot[kLanes];
    int cur, arrived;
    unsigned gen;
    const std::function<void()> *body;
};
static Warp *g = nullptr;

static void yield() { swapcontext(&g->ctx[g->cur], &g->main); }
static void rendezvous()
{
    const unsigned gen = g->gen;
    if (++g->arrived == kLanes) { g->arrived = 0; g->gen++; }
    else while (g->gen == gen) yield();
}
static void entry()
{
    (*g->body)();
    g->done[g->cur] = true;
    swapcontext(&g->ctx[g->cur], &g->main);
}
// runs `body` on 32 lanes in lock step at the collectives; returns 0, or -1 when the lanes stop making progress
// together (a collective that not every lane reaches = a divergence bug in the kernel code)
static int run_warp(const std::function<void()> &body)
{
    Warp w;
    memset(&w.done, 0, sizeof w.done);
    w.arrived = 0; w.gen = 0; w.body = &body;
    w.stacks = static_cast<char *>(malloc(kStack * kLanes));
    g = &w;
    for (int l = 0; l < kLanes; l++) {
        getcontext(&w.ctx[l]);
        w.ctx[l].uc_stack.ss_sp = w.stacks + kStack * l;
        w.ctx[l].uc_stack.ss_size = kStack;
        w.ctx[l].uc_link = &w.main;
        makecontext(&w.ctx[l], entry, 0);
    }
    int rc = 0;
    for (;;) {
        int alive = 0, ndone = 0;
        const unsigned gen0 = w.gen;
        const int arrived0 = w.arrived;
        for (int l = 0; l < kLanes; l++) {
            if (w.done[l]) { ndone++; continue; }
            alive++;
            w.cur = l;
            swapcontext(&w.main, &w.ctx[l]);
        }
        if (!alive) break;
        int nd = 0;
        for (int l = 0; l < kLanes; l++) nd += w.done[l];
        if (w.gen == gen0 && w.arrived == arrived0 && nd == ndone) { rc = -1; break; }   // a full round, nothing moved
    }
    free(w.stacks);
    g = nullptr;
    return rc;
}
}  // namespace emu

static inline int igd_w_lane() { return emu::g->cur; }
static inline uint32_t igd_w_ballot(bool p)
{
    emu::g->slot[emu::g->cur] = p ? 1u : 0u;
    emu::rendezvous();
    uint32_t m = 0;
    for (int l = 0; l < emu::kLanes; l++) m |= emu::g->slot[l] << l;
    emu::rendezvous();
    return m;
}
static inline uint32_t igd_w_shfl(uint32_t v, int src)
{
    emu::g->slot[emu::g->cur] = v;
    emu::rendezvous();
    const uint32_t r = emu::g->slot[src & 31];
    emu::rendezvous();
    return r;
}

#include "../../igate4xsoftphonedsp_b200/csrc/igd_walks.cuh"

extern "C" {
int emul_rxarb_walk(int F, int B, int mode, int tick_ms, int period, int wd_ticks, int frame0, long long now0,
                    const uint8_t *pkts, const uint32_t *sizes, const uint8_t *active, igd_rx_state *rx_state,
                    igd_arb_leg *legs, igd_arb_bridge *bridges, igd_rx_event *events, uint16_t *gain)
{
    igd_rxarb_args a;
    a.F = F; a.B = B; a.mode = mode; a.tick_ms = tick_ms; a.r2s_period_ms = period; a.wd_ticks = wd_ticks; a.frame0 = frame0;
    a.now_ms0 = now0; a.pkts = pkts; a.sizes = sizes; a.active = active; a.rx_state = rx_state; a.legs = legs;
    a.bridges = bridges; a.events = events; a.gain_q7 = gain;
    for (int b = 0; b < B; b++)
        if (emu::run_warp([&] { igd_rxarb_walk(a, b); })) return -1;
    return 0;
}

static igd_ed137_pack_desc pack_desc(int F, int C, unsigned flags, unsigned payload_len, int tick_ms, long long now0,
                                     const igd_ed137_ctl *ctl, const uint8_t *payload, igd_ed137_state *state)
{
    igd_ed137_pack_desc d;
    memset(&d, 0, sizeof d);
    d.F = F; d.C = C; d.flags = flags; d.payload_len = payload_len; d.tick_ms = tick_ms; d.now_ms0 = now0;
    d.ctl = ctl; d.payload = payload; d.state = state;
    return d;
}

int emul_plan_walk(int F, int C, unsigned flags, unsigned payload_len, int tick_ms, long long now0, const igd_ed137_ctl *ctl,
                   const uint8_t *payload, igd_ed137_state *state, igd_tx_plan_rec *plan, int32_t *last_src)
{
    const igd_ed137_pack_desc d = pack_desc(F, C, flags, payload_len, tick_ms, now0, ctl, payload, state);
    for (int c = 0; c < C; c++)
        if (emu::run_warp([&] { igd_plan_walk(d, plan, last_src, c); })) return -1;
    return 0;
}

// what the walk must equal: transport_send_rtp (igd_ed137_tx_step, held to the oracle by test_device_math_host.py)
// called tick after tick, nothing skipped
int emul_plan_ref(int F, int C, unsigned flags, unsigned payload_len, int tick_ms, long long now0, const igd_ed137_ctl *ctl,
                  const uint8_t *payload, igd_ed137_state *state, igd_tx_plan_rec *plan, int32_t *last_src)
{
    const bool stuck = payload != nullptr && 12u + payload_len > 60u;
    for (int c = 0; c < C; c++) {
        igd_ed137_state s = state[c];
        int32_t src = -1;
        for (int f = 0; f < F; f++) {
            const size_t i = (size_t)f * C + c;
            if (ctl) {
                const igd_ed137_ctl k = ctl[i];
                s.pttstatus = k.pttstatus; s.pttpriority = k.pttpriority; s.callRecorder = k.callRecorder;
                s.sqlstatus = k.sqlstatus; s.ed137_bssi = k.ed137_bssi; s.pttid = k.pttid;
            }
            if (s.radiostatus && stuck) {
                const uint8_t *pl = payload + i * IGD_FRAME;
                if (pl[28] == pl[38] && pl[28] == pl[48] && pl[28] == 0xd5) s.rtpFalse += 1; else s.rtpFalse = 0;
            }
            const igd_tx_plan t = igd_ed137_tx_step(s, payload_len, now0 + (long long)f * tick_ms);
            if (t.copy_payload) src = f;
            igd_tx_plan_rec r;
            r.word = t.word; r.size = (uint16_t)t.size;
            r.flags = (uint8_t)(t.pt123 | (t.marker << 1) | (t.copy_payload << 2));
            r.reserved = 0;
            r.src_frame = (flags & IGD_F_REF_QUIRKS) ? src : f;
            plan[i] = r;
        }
        state[c] = s;
        last_src[c] = src;
    }
    return 0;
}
}
