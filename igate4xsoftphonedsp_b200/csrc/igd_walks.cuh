// igd_walks.cuh -- the per-call state machines walked with the TICK AXIS ACROSS THE LANES of a warp.
//
// The three walks of the gateway tick (receive liveness: transport_rtp_cb, TransportAdapter.cpp:240-316 +
// detectR2SPacketAndReconn, roip_ed137.cpp:1767-1780; gate arbitration: checkEvents, roip_ed137.cpp:6124-6231 /
// :5627-5719 / :5985-6121; sender: transport_send_rtp, TransportAdapter.cpp:635-874) are sequential per call by
// definition.  The thread-per-channel kernels of igd_packet.cu therefore run one dependent chain of 60-100
// instructions per tick: 0.3-0.5 ms for 1640 ticks however few channels there are.  Here a warp owns one bridge
// (receive side) or one sender and takes 32 consecutive ticks per step, lane = tick:
//   * every piece of receive state is a FORWARD FILL (latched word, r2sPacket stamp, rtpAudio) or a run length
//     between resets (the watchdog strikes): one ballot + one bit scan per lane, no chain at all;
//   * checkEvents() and transport_send_rtp are deterministic functions of (state, input), so once a pass has
//     left the state as it found it every further tick with the same input repeats its output (the steady-state
//     property the thread-per-channel kernels already use).  One ballot marks the ticks whose input differs from
//     the tick before; the full step runs -- warp-uniform, state in registers, input broadcast by shuffle --
//     only on those ticks and on the ticks that follow until the state settles; the steady stretches in between
//     are filled by all lanes at once (the keep-alive throttle of a steady sender is one ballot per emission).
// Bit-exact by construction: a full step is always right, and a skipped tick is one whose step is known to
// change nothing.
//
// The bodies are written against four warp primitives (igd_w_lane / igd_w_ballot / igd_w_shfl / IGD_W_LDG) so that
// tests/hostbuild/walks_emul.cpp can compile THIS FILE with g++ and run it lane for lane on 32 fibers against the
// oracle without a GPU.  The product only runs them inside the kernels of igd_walks.cu.
#pragma once
#include "../../include/igate_dsp.h"
#include "igd_math.cuh"

#ifndef IGD_HOST_EMUL
__device__ __forceinline__ int igd_w_lane() { return (int)(threadIdx.x & 31u); }
__device__ __forceinline__ uint32_t igd_w_ballot(bool p) { return __ballot_sync(0xffffffffu, p); }
__device__ __forceinline__ uint32_t igd_w_shfl(uint32_t v, int src) { return __shfl_sync(0xffffffffu, v, src); }
__device__ __forceinline__ int igd_w_clz(uint32_t v) { return __clz((int)v); }
__device__ __forceinline__ int igd_w_ffs(uint32_t v) { return __ffs((int)v); }
__device__ __forceinline__ int igd_w_popc(uint32_t v) { return __popc(v); }
__device__ __forceinline__ uint32_t igd_w_bswap(uint32_t v) { return __byte_perm(v, 0, 0x0123); }
#define IGD_W_LDG(p) __ldg(p)
#else
// igd_w_lane / igd_w_ballot / igd_w_shfl come from the emulator (declared before this file is included)
static inline int igd_w_clz(uint32_t v) { return v ? __builtin_clz(v) : 32; }
static inline int igd_w_ffs(uint32_t v) { return __builtin_ffs((int)v); }
static inline int igd_w_popc(uint32_t v) { return __builtin_popcount(v); }
static inline uint32_t igd_w_bswap(uint32_t v) { return __builtin_bswap32(v); }
#define IGD_W_LDG(p) (*(p))
#endif

// lanes <= l / lanes < l
IGD_HD uint32_t igd_w_le(int l) { return 0xFFFFFFFFu >> (31 - l); }
IGD_HD uint32_t igd_w_lt(int l) { return (1u << l) - 1u; }
// index of the highest set bit, -1 if none
IGD_HD int igd_w_top(uint32_t m) { return 31 - igd_w_clz(m); }

// ---------------------------------------------------------------------------------------------------------
// receive side of one bridge (4 legs): packets in -> events, gains
struct igd_rxarb_args {
    int32_t F, B, mode, tick_ms, r2s_period_ms, wd_ticks, frame0;
    long long now_ms0;
    const uint8_t *pkts;          // [F][B*4][180]
    const uint32_t *sizes;        // [F][B*4] or NULL (all 180)
    const uint8_t *active;        // [B*4] or NULL
    igd_rx_state *rx_state;       // [B*4] in/out
    igd_arb_leg *legs;            // [B*4] in/out
    igd_arb_bridge *bridges;      // [B] in/out
    igd_rx_event *events;         // [F][B*4] out, or NULL
    uint16_t *gain_q7;            // [F][B*4] out (IGD_GAIN_NO_AUDIO on ticks without a whole audio frame)
};

struct igd_w_raw { uint32_t w0, w3, w4, size; };
// igd_arb_leg with its fields in full registers (the byte-wide fields of the API struct cost a PRMT per update).
// msec keeps its 8 bits: `++lastTxmsec < 6` must wrap where the reference's uint8_t wraps; the tick functions narrow
// what they store into the other fields themselves.
struct igd_w_leg { uint32_t last; uint8_t msec; uint32_t on; int32_t rssi; uint32_t gain_q7, reserved; };

// the header words of the four legs of one bridge on tick t.  Every load is unconditional and independent of the
// others (a slot of the [F][C][180] buffer always exists, whatever the received size says): thirteen loads in flight
// per lane.  What transport_rtp_cb may not look at -- words beyond the received size -- is masked when the tick is
// walked (igd_w_hdr_mask), as k_rx_track<packets> does by not loading it.
IGD_HD void igd_w_fetch_hdr4(const igd_rxarb_args &a, int t, size_t ch0, size_t Cn, igd_w_raw (&r)[4])
{
    IGD_UNROLL
    for (int g = 0; g < 4; g++) { r[g].w0 = r[g].w3 = r[g].w4 = 0u; r[g].size = 0u; }
    if (t < a.F) {
        const size_t i = (size_t)t * Cn + ch0;
        const uint32_t *pw = reinterpret_cast<const uint32_t *>(a.pkts + i * IGD_PKT_MAX);
        IGD_UNROLL
        for (int g = 0; g < 4; g++) {
            r[g].w0 = IGD_W_LDG(pw + g * (IGD_PKT_MAX / 4));
            r[g].w3 = IGD_W_LDG(pw + g * (IGD_PKT_MAX / 4) + 3);
            r[g].w4 = IGD_W_LDG(pw + g * (IGD_PKT_MAX / 4) + 4);
            r[g].size = (uint32_t)IGD_PKT_MAX;
        }
        if (a.sizes) {
            IGD_UNROLL
            for (int g = 0; g < 4; g++) r[g].size = IGD_W_LDG(a.sizes + i + g);
        }
    }
}
IGD_HD igd_w_raw igd_w_hdr_mask(igd_w_raw r)
{
    const uint32_t navail = (r.size < (uint32_t)IGD_PKT_MAX ? r.size : (uint32_t)IGD_PKT_MAX) / 4;
    r.w0 = navail > 0 ? r.w0 : 0u; r.w3 = navail > 3 ? r.w3 : 0u; r.w4 = navail > 4 ? r.w4 : 0u;
    return r;
}

// igd_arb_client_tick (igd_math.cuh; roip_ed137.cpp:6124-6231) for the four register-resident legs of the walk, written
// as selects: the branchy form is a chain of short data-dependent branches per leg, and a full pass is the one
// strictly sequential piece of this kernel (ncu: half of its stall samples were branch resolution and the fixed
// latency behind it).  Same order of updates, leg after leg; held to the oracle by tests/test_walks_host.py.
// Also says what the pass did to the state: `moved` != 0 when anything but a hold-off count changed, `ticked` = the
// legs whose lastTxmsec went up by one (:6141), `room` = how many more passes may do just that before the first
// of them releases (msec enters a pass through `++msec < 6` alone).
IGD_HD void igd_w_client_tick4(igd_arb_bridge &b, igd_w_leg (&legs)[4], const uint32_t (&w)[4], const uint32_t act_mask,
                               uint32_t &moved, uint32_t &ticked, int &room)
{
    const int level0 = b.ptt_level;
    int level = level0;
    const uint32_t g0 = legs[0].gain_q7, g1 = legs[1].gain_q7, g2 = legs[2].gain_q7, g3 = legs[3].gain_q7;
    bool mv = false;
    ticked = 0u; room = 5;
    // (`&` / `|` on the predicates, not `&&` / `||`: the short-circuit forms come out as branches)
    IGD_UNROLL
    for (int i = 0; i < 4; i++) {
        const bool act = ((act_mask >> i) & 1u) != 0u;
        const int ptt_w = (int)(w[i] >> 29);
        const uint32_t last0 = legs[i].last;
        const uint8_t msec0 = legs[i].msec, m1 = (uint8_t)(msec0 + 1);
        // A leg that is where its word says -- same PTT as last tick, no hold-off count, neither beating the level nor
        // about to be switched on or off -- leaves this pass without a trace: most legs on most passes (the steady
        // check finds four of them).  The branch is warp-uniform, one per leg, around ~50 instructions.
        if (!act | (((uint32_t)ptt_w == last0) & (msec0 == 0) & !((ptt_w != 0) & (ptt_w > level)) & ((ptt_w > 0) == (legs[i].on != 0u))))
            continue;
        const bool ne = (uint32_t)ptt_w != last0, rel = ne & (ptt_w == 0);            // :6136-6153
        const uint8_t msec = rel ? m1 : ne ? msec0 : (uint8_t)0;
        const int ptt = rel & (m1 < 6) ? 1 : ptt_w;                                    // released: held for five more ticks
        const bool win = act & (ptt != 0) & (ptt > level);                             // :6155-6175
        level = win ? ptt : level;
        IGD_UNROLL
        for (int j = 0; j < 4; j++)
            if (j != i) legs[j].gain_q7 = win & (legs[j].on != 0u) ? 0u : legs[j].gain_q7;
        const bool on_i = legs[i].on != 0u;
        const bool press = act & (ptt > 0) & !on_i, release = act & (ptt == 0) & on_i; // :6177-6231
        legs[i].gain_q7 = release ? 0u : win ? 256u : legs[i].gain_q7;
        legs[i].on = press ? 1u : release ? 0u : legs[i].on;
        level = release ? 0 : level;
        mv = mv | press | release | (act & (((uint32_t)ptt != last0) | (!ne & (msec0 != 0))));
        const bool tick1 = act & rel;
        const int r = 5 - (int)m1;
        ticked |= tick1 ? 1u << i : 0u;
        room = tick1 & (r < room) ? r : room;
        legs[i].msec = act ? msec : msec0;
        legs[i].last = act ? (uint32_t)ptt : last0;
    }
    b.ptt_level = level;
    moved = (mv | (level != level0) ? 1u : 0u) | (g0 ^ legs[0].gain_q7) | (g1 ^ legs[1].gain_q7) | (g2 ^ legs[2].gain_q7) |
            (g3 ^ legs[3].gain_q7);
}

// `b` = bridge of this warp (warp-uniform); every lane of the warp calls this together
IGD_HD void igd_rxarb_walk(const igd_rxarb_args &a, const int b)
{
    constexpr int G = 4;
    const int lane = igd_w_lane();
    const size_t Cn = (size_t)a.B * G, ch0 = (size_t)b * G;
    // ---- state, warp-uniform in registers
    long long r2sPacket[G];
    uint32_t value[G], paysz[G], rtpAudio[G], r2sCount[G];
    igd_w_leg legs[G];
    IGD_UNROLL
    for (int g = 0; g < G; g++) {
        const igd_rx_state s = a.rx_state[ch0 + g];
        r2sPacket[g] = s.r2sPacket; value[g] = s.ed137_value; paysz[g] = s.payloadsize; rtpAudio[g] = s.rtpAudio; r2sCount[g] = s.r2sCount;
        const uint32_t *lp = reinterpret_cast<const uint32_t *>(a.legs + ch0 + g);     // field by field: the array stays in registers
        const uint32_t l0 = lp[0], l1 = lp[1];
        legs[g].last = l0 & 0xFFu; legs[g].msec = (uint8_t)(l0 >> 8); legs[g].on = (l0 >> 16) & 0xFFu; legs[g].rssi = (int32_t)(int8_t)(l0 >> 24);
        legs[g].gain_q7 = l1 & 0xFFFFu; legs[g].reserved = l1 >> 16;
    }
    igd_arb_bridge br = a.bridges[b];
    uint32_t act_mask = 0xFu;
    if (a.active) {
        act_mask = 0;
        IGD_UNROLL
        for (int g = 0; g < G; g++) act_mask |= (a.active[ch0 + g] != 0 ? 1u : 0u) << g;
    }
    uint32_t prevw[G] = {0u, 0u, 0u, 0u};      // words of the tick before the current step's first
    bool have_prev = false, steady = false, counting = false, hold = false;
    uint32_t hold_legs = 0u;
    int hold_k = 0;
    const int wd_p0 = a.wd_ticks > 0 ? (int)(a.frame0 % a.wd_ticks) : 0;
    const long long late_after = (long long)a.r2s_period_ms * 3;

    igd_w_raw nraw[G];
    igd_w_fetch_hdr4(a, lane, ch0, Cn, nraw);
    for (int t0 = 0; t0 < a.F; t0 += 32) {
        const int nt = a.F - t0 < 32 ? a.F - t0 : 32;
        const int t = t0 + lane;
        const bool valid = lane < nt;
        const long long now = a.now_ms0 + (long long)t * a.tick_ms;
        igd_w_raw raw[G];
        IGD_UNROLL
        for (int g = 0; g < G; g++) raw[g] = igd_w_hdr_mask(nraw[g]);
        igd_w_fetch_hdr4(a, t + 32, ch0, Cn, nraw);                  // the next step's headers are in flight
        bool wd = false;
        if (a.wd_ticks > 0) {
            const int q = t + wd_p0 + 1;
            wd = valid && q >= a.wd_ticks && q % a.wd_ticks == 0;
        }
        uint32_t W[G], noaud = 0u;
        // ---- transport_rtp_cb's view of every header (TransportAdapter.cpp:248-292)
        uint32_t pt[G], word[G], len_raw[G];
        bool present[G], dropped[G], accepted[G], frame[G];
        uint32_t kinds = 0u;                 // bit g: leg g's packet on this tick is forwarded audio (pt != 123)
        bool plain = true;                   // all four legs: a packet, accepted, not dropped
        IGD_UNROLL
        for (int g = 0; g < G; g++) {
            const uint32_t size = raw[g].size;
            pt[g] = (raw[g].w0 >> 8) & 0x7Fu;
            present[g] = valid && size != 0u;
            const bool too_short = size < (uint32_t)IGD_PKT_HDR;
            const uint32_t plen_raw = size - (uint32_t)IGD_PKT_HDR;
            dropped[g] = too_short || plen_raw >= 1024u;
            accepted[g] = !too_short && (pt[g] == 8u || pt[g] == 0u || pt[g] == 18u || pt[g] == 123u);
            word[g] = accepted[g] ? igd_w_bswap(raw[g].w4) : 0u;
            len_raw[g] = accepted[g] ? raw[g].w3 >> 16 : 0u;
            const uint32_t plen = dropped[g] ? 0u : (plen_raw < (uint32_t)IGD_FRAME ? plen_raw : (uint32_t)IGD_FRAME);
            frame[g] = (pt[g] == 0u || pt[g] == 8u) && plen == 160u;
            plain = plain & present[g] & accepted[g] & !dropped[g];
            kinds |= (pt[g] != 123u ? 1u : 0u) << g;
        }
        // A step in which every tick of every leg brings an accepted, undropped packet of one kind per leg (a radio
        // that keeps sending audio, or keeps sending keep-alives -- the usual 640 ms): the forward fills are the
        // identity, an edge can only sit on the step's first tick, and no watchdog tick is late (now - r2sPacket = 0).
        const uint32_t kinds0 = igd_w_shfl(kinds, 0);
        const bool uniform = late_after >= 0 && igd_w_ballot(valid && !(plain && kinds == kinds0)) == 0u;
        if (uniform) {
            const uint32_t m_wd = igd_w_ballot(wd);
            const bool wd_seen = (m_wd & igd_w_le(lane)) != 0u;
            IGD_UNROLL
            for (int g = 0; g < G; g++) {
                const bool kind = ((kinds0 >> g) & 1u) != 0u;
                const bool audio_before = lane > 0 ? kind : rtpAudio[g] != 0u;
                const uint32_t ev = 0x01u | (kind ? 0x02u : 0u) | (kind && frame[g] ? 0x40u : 0u) | (kind != audio_before ? 0x04u : 0u);
                const uint32_t after = wd_seen ? 0u : r2sCount[g];
                W[g] = word[g];
                if (!(ev & 0x40u)) noaud |= 1u << g;
                if (a.events && valid) {
                    uint32_t *ep = reinterpret_cast<uint32_t *>(a.events + (size_t)t * Cn + ch0 + g);   // {word, flags | r2sCount << 8}
                    ep[0] = W[g]; ep[1] = ev | (after << 8);
                }
                value[g] = igd_w_shfl(word[g], nt - 1);
                paysz[g] = igd_w_shfl(len_raw[g], nt - 1);
                r2sPacket[g] = a.now_ms0 + (long long)(t0 + nt - 1) * a.tick_ms;
                rtpAudio[g] = kind ? 1u : 0u;
                r2sCount[g] = m_wd ? 0u : r2sCount[g];
            }
        } else {
        IGD_UNROLL
        for (int g = 0; g < G; g++) {
            // ---- latch :252-256: the word / length of the last accepted packet up to this tick
            const uint32_t m_acc = igd_w_ballot(present[g] && accepted[g]);
            const int s_acc = igd_w_top(m_acc & igd_w_le(lane));
            const uint32_t wv = igd_w_shfl(word[g], s_acc < 0 ? 0 : s_acc), pv = igd_w_shfl(len_raw[g], s_acc < 0 ? 0 : s_acc);
            W[g] = s_acc >= 0 ? wv : value[g];
            const uint32_t P = s_acc >= 0 ? pv : paysz[g];
            // ---- r2sPacket :289,302,311: stamped by every packet
            const uint32_t m_pres = igd_w_ballot(present[g]);
            const int s_pres = igd_w_top(m_pres & igd_w_le(lane));
            const long long r2s = s_pres >= 0 ? a.now_ms0 + (long long)(t0 + s_pres) * a.tick_ms : r2sPacket[g];
            // ---- rtpAudio :298-315: set by the packets that are not dropped; an edge where it flips
            const bool def = present[g] && !dropped[g], aud = def && pt[g] != 123u;
            const uint32_t m_def = igd_w_ballot(def), m_aud = igd_w_ballot(aud);
            const int s_def = igd_w_top(m_def & igd_w_lt(lane));
            const bool audio_before = s_def >= 0 ? ((m_aud >> s_def) & 1u) != 0u : rtpAudio[g] != 0u;
            uint32_t ev = 0u;
            if (present[g]) ev |= 0x01u;
            if (present[g] && dropped[g]) ev |= 0x08u;
            if (aud) {
                ev |= 0x02u;
                if (frame[g]) ev |= 0x40u;
            }
            if (def && (pt[g] != 123u) != audio_before) ev |= 0x04u;
            // ---- watchdog roip_ed137.cpp:1767-1780: strikes = late watchdog ticks since the last one in time
            const bool late = wd && now - r2s > late_after, rst = wd && !late;
            const uint32_t m_late = igd_w_ballot(late), m_rst = igd_w_ballot(rst);
            const int s_rst = igd_w_top(m_rst & igd_w_le(lane));
            const uint32_t since = s_rst >= 0 ? ~igd_w_le(s_rst) : 0xFFFFFFFFu;
            const uint32_t base = s_rst >= 0 ? 0u : r2sCount[g];
            uint32_t before = base + (uint32_t)igd_w_popc(m_late & igd_w_lt(lane) & since);
            before = before < 255u ? before : 255u;
            uint32_t after = before + (late ? 1u : 0u);
            after = after < 255u ? after : 255u;
            if (late) {
                ev |= 0x10u;
                if (before == 5u) ev |= 0x20u;
            }
            if (!(ev & 0x40u)) noaud |= 1u << g;
            if (a.events && valid) {
                uint32_t *ep = reinterpret_cast<uint32_t *>(a.events + (size_t)t * Cn + ch0 + g);   // {word, flags | r2sCount << 8}
                ep[0] = W[g]; ep[1] = ev | (after << 8);
            }
            // ---- carry to the next step: the values after this step's last tick
            value[g] = igd_w_shfl(W[g], nt - 1);
            paysz[g] = igd_w_shfl(P, nt - 1);
            if (m_pres) r2sPacket[g] = a.now_ms0 + (long long)(t0 + igd_w_top(m_pres)) * a.tick_ms;
            if (m_def) rtpAudio[g] = (m_aud >> igd_w_top(m_def)) & 1u;
            r2sCount[g] = igd_w_shfl(after, nt - 1);
        }
        }
        // ---- gate arbitration over this step's ticks: full passes where the words change or the state still moves
        bool differs = false;
        IGD_UNROLL
        for (int g = 0; g < G; g++) {
            const uint32_t up = igd_w_shfl(W[g], lane > 0 ? lane - 1 : 0);
            differs = differs || W[g] != (lane > 0 ? up : prevw[g]);
        }
        const uint32_t chg = igd_w_ballot(valid && (differs || (lane == 0 && !have_prev)));
        uint32_t g01 = 0u, g23 = 0u;        // this lane's tick: the four gains
        int tl = 0;
        while (tl < nt) {
            if (steady || hold) {
                const uint32_t m = chg & ~igd_w_lt(tl);
                int nxt = m ? igd_w_ffs(m) - 1 : nt;
                // CLIENT hold-off (roip_ed137.cpp:6140-6147): the last pass only counted lastTxmsec up on released legs.
                // msec enters a pass through `++msec < 6` alone, so the passes repeat while that stays true: hold_k more
                if (hold && nxt - tl > hold_k) nxt = tl + hold_k;
                if (lane >= tl && lane < nxt) {
                    g01 = legs[0].gain_q7 | (legs[1].gain_q7 << 16);
                    g23 = legs[2].gain_q7 | (legs[3].gain_q7 << 16);
                }
                if (counting) br.sqlStatusCount += nxt - tl;
                if (hold) {
                    IGD_UNROLL
                    for (int g = 0; g < G; g++) legs[g].msec = (uint8_t)(legs[g].msec + (((hold_legs >> g) & 1u) ? nxt - tl : 0));
                    hold = false;           // the next tick takes the full pass (it releases, or re-arms the hold)
                }
                tl = nxt;
                if (tl >= nt) break;
            }
            uint32_t w[G];
            IGD_UNROLL
            for (int g = 0; g < G; g++) w[g] = igd_w_shfl(W[g], tl);
            if (a.mode == IGD_ARB_CLIENT_PTT) {
                uint32_t moved, ticked;
                int room;
                igd_w_client_tick4(br, legs, w, act_mask, moved, ticked, room);
                steady = moved == 0u && ticked == 0u;
                hold = moved == 0u && ticked != 0u && room > 0;
                hold_legs = ticked; hold_k = room;
                counting = false;
            } else {
                igd_w_leg was[G];
                IGD_UNROLL
                for (int g = 0; g < G; g++) was[g] = legs[g];
                const int32_t c0 = br.sqlStatusCount, l0 = br.ptt_level;
                const uint32_t o0 = br.sqlStatusOn;
                auto word = [&](int g) { return w[g]; };
                auto active = [&](int g) { return ((act_mask >> g) & 1u) != 0u; };
                const igd_const_int<G> Gc;
                igd_arb_server_best_tick(br, legs, Gc, word, active);
                uint32_t moved = 0u;
                IGD_UNROLL
                for (int g = 0; g < G; g++)
                    moved |= (was[g].last ^ legs[g].last) | (was[g].on ^ legs[g].on) | (uint32_t)(was[g].rssi ^ legs[g].rssi) |
                             (was[g].gain_q7 ^ legs[g].gain_q7) | (uint32_t)(was[g].msec ^ legs[g].msec);
                const bool legs_same = moved == 0u && br.ptt_level == l0 && br.sqlStatusOn == o0;
                // SERVER mode's second steady form: while a selection is in force a pass only counts sqlStatusCount up
                // (roip_ed137.cpp:6028) and nothing reads the count again (:6029 needs !sqlStatusOn)
                counting = legs_same && br.sqlStatusOn != 0 && br.sqlStatusCount == c0 + 1;
                steady = legs_same && (br.sqlStatusCount == c0 || counting);
                hold = false;
            }
            if (lane == tl) {
                g01 = legs[0].gain_q7 | (legs[1].gain_q7 << 16);
                g23 = legs[2].gain_q7 | (legs[3].gain_q7 << 16);
            }
            tl++;
        }
        have_prev = true;
        IGD_UNROLL
        for (int g = 0; g < G; g++) prevw[g] = igd_w_shfl(W[g], nt - 1);
        if (valid) {
            // a tick without a whole audio frame is silent whatever the gate says (IGD_ARB_F_SILENCE)
            g01 |= ((noaud & 1u) ? (uint32_t)IGD_GAIN_NO_AUDIO : 0u) | ((noaud & 2u) ? (uint32_t)IGD_GAIN_NO_AUDIO << 16 : 0u);
            g23 |= ((noaud & 4u) ? (uint32_t)IGD_GAIN_NO_AUDIO : 0u) | ((noaud & 8u) ? (uint32_t)IGD_GAIN_NO_AUDIO << 16 : 0u);
            uint32_t *gp = reinterpret_cast<uint32_t *>(a.gain_q7 + (size_t)t * Cn + ch0);
            gp[0] = g01; gp[1] = g23;
        }
    }
    // ---- state back
    if (lane < G) {
        IGD_UNROLL
        for (int g = 0; g < G; g++) {
            if (lane == g) {
                igd_rx_state s;
                s.r2sPacket = r2sPacket[g]; s.ed137_value = value[g]; s.payloadsize = (uint16_t)paysz[g];
                s.rtpAudio = (uint8_t)rtpAudio[g]; s.r2sCount = (uint8_t)r2sCount[g];
                a.rx_state[ch0 + g] = s;
                uint32_t *lp = reinterpret_cast<uint32_t *>(a.legs + ch0 + g);
                lp[0] = (legs[g].last & 0xFFu) | ((uint32_t)legs[g].msec << 8) | ((legs[g].on & 0xFFu) << 16) | ((uint32_t)legs[g].rssi << 24);
                lp[1] = (legs[g].gain_q7 & 0xFFFFu) | (legs[g].reserved << 16);
            }
        }
    }
    if (lane == 0) a.bridges[b] = br;
}

// ---------------------------------------------------------------------------------------------------------
// sender walk of one outgoing call: plan records for F ticks (same records as k_ed137_plan, igd_packet.cu)
IGD_HD void igd_plan_walk(const igd_ed137_pack_desc &d, igd_tx_plan_rec *plan, int32_t *last_src, const int c)
{
    const int lane = igd_w_lane();
    igd_ed137_state s = d.state[c];
    int32_t src = -1;
    const bool stuck = d.payload != nullptr && 12u + d.payload_len > 60u;
    igd_tx_plan last;
    last.word = 0; last.size = 0; last.pt123 = 0; last.marker = 0; last.copy_payload = 0;
    uint32_t lc_lo = 0u, lc_hi = 0u;         // setter values of the tick before this step's first
    bool have_last = false, steady = false;
    for (int t0 = 0; t0 < d.F; t0 += 32) {
        const int nt = d.F - t0 < 32 ? d.F - t0 : 32;
        const int t = t0 + lane;
        const bool valid = lane < nt;
        const size_t i = (size_t)(valid ? t : t0) * d.C + c;
        const long long now = d.now_ms0 + (long long)t * d.tick_ms;
        uint32_t k_lo = 0u, k_hi = 0u;
        if (d.ctl) {
            const uint32_t *kp = reinterpret_cast<const uint32_t *>(d.ctl + i);
            k_lo = IGD_W_LDG(kp); k_hi = IGD_W_LDG(kp + 1);
        }
        // stuck-audio detector :657-673: a run length between resets; only the count after the last tick is state
        if (stuck && s.radiostatus) {
            bool cnd = false;
            if (valid) {
                const uint8_t *pl = d.payload + i * IGD_FRAME;
                const uint32_t a40 = pl[40 - 12], a50 = pl[50 - 12], a60 = pl[60 - 12];
                cnd = a40 == a50 && a40 == a60 && a40 == 0xd5u;
            }
            const uint32_t m_c = igd_w_ballot(valid && cnd), m_n = igd_w_ballot(valid && !cnd);
            const int r = igd_w_top(m_n);
            s.rtpFalse = (r >= 0 ? 0 : s.rtpFalse) + igd_w_popc(r >= 0 ? m_c & ~igd_w_le(r) : m_c);
        }
        const uint32_t up_lo = igd_w_shfl(k_lo, lane > 0 ? lane - 1 : 0), up_hi = igd_w_shfl(k_hi, lane > 0 ? lane - 1 : 0);
        const bool differs = lane > 0 ? (k_lo != up_lo || k_hi != up_hi) : (!have_last || k_lo != lc_lo || k_hi != lc_hi);
        const uint32_t chg = igd_w_ballot(valid && differs);
        // this lane's tick
        uint32_t r_word = 0u, r_size = 0u, r_flags = 0u;
        auto put = [&](const igd_tx_plan &p) {
            r_word = p.word; r_size = p.size;
            r_flags = (uint32_t)p.pt123 | ((uint32_t)p.marker << 1) | ((uint32_t)p.copy_payload << 2);
        };
        int tl = 0;
        while (tl < nt) {
            if (steady) {
                const uint32_t m = chg & ~igd_w_lt(tl);
                const int nxt = m ? igd_w_ffs(m) - 1 : nt;
                if (last.copy_payload) {                                   // gated audio: goes out on every tick
                    if (lane >= tl && lane < nxt) put(last);
                    tl = nxt;
                } else {
                    // keep-alive throttle :685-706 over the stretch [tl, nxt): one ballot per packet that is due
                    const unsigned long long ka = (unsigned long long)(long long)s.keepAlivePeroid;
                    int cur = tl;
                    while (cur < nxt) {
                        const bool due = lane >= cur && lane < nxt &&
                                         (unsigned long long)now - (unsigned long long)s.r2sSendtime >= ka;
                        const uint32_t m_due = igd_w_ballot(due);
                        if (!m_due) { cur = nxt; break; }                  // throttled to the end of the stretch (records stay 0)
                        const int u = igd_w_ffs(m_due) - 1;
                        if (last.size == 0) { cur = u; break; }            // no sent header cached: the full step, at tick u (< nxt)
                        s.r2sSendtime = d.now_ms0 + (long long)(t0 + u) * d.tick_ms;
                        if (lane == u) put(last);
                        cur = u + 1;
                    }
                    tl = cur;        // nxt, or the tick that needs the full step
                }
                if (tl >= nt) break;
            }
            // ---- the full step at tick tl: the setters (:135-213), then transport_send_rtp
            if (d.ctl) {
                const uint32_t lo = igd_w_shfl(k_lo, tl), hi = igd_w_shfl(k_hi, tl);
                s.pttstatus = (uint8_t)lo; s.sqlstatus = (uint8_t)(lo >> 8); s.pttpriority = (uint8_t)(lo >> 16);
                s.ed137_bssi = (uint8_t)(lo >> 24); s.pttid = (uint8_t)hi; s.callRecorder = (uint8_t)(hi >> 8);
            }
            if (s.radiostatus && (s.calltype_flags & 1u) && s.callIn) { s.sqlstatus = 0; s.pttstatus = 0; }   // :675-679
            const igd_ed137_state before = s;
            const igd_tx_plan p = igd_ed137_tx_step(s, d.payload_len, d.now_ms0 + (long long)(t0 + tl) * d.tick_ms);
            // steady = the step left everything but the throttle clock as it found it (and the start burst is over)
            steady = before.radiostatus == s.radiostatus && before.pttstatus == s.pttstatus && before.sqlstatus == s.sqlstatus &&
                     before.callIn == s.callIn && before.callRecorder == s.callRecorder && before.pttpriority == s.pttpriority &&
                     before.pttid == s.pttid && before.ed137_bssi == s.ed137_bssi && before.rxSlaveEnable == s.rxSlaveEnable &&
                     before.txSlaveEnable == s.txSlaveEnable && before.rxSlaveEnableChanged == s.rxSlaveEnableChanged &&
                     before.txSlaveEnableChanged == s.txSlaveEnableChanged &&
                     before.trxSlaveEnableChangedCount == s.trxSlaveEnableChangedCount && before.firstR2SPacket == s.firstR2SPacket &&
                     before.calltype_flags == s.calltype_flags && before.sqlpriority == s.sqlpriority &&
                     before.packetCnt == s.packetCnt && before.keepAlivePeroid == s.keepAlivePeroid && !s.firstR2SPacket;
            last = p;
            have_last = true;
            if (lane == tl) put(p);
            tl++;
        }
        lc_lo = igd_w_shfl(k_lo, nt - 1); lc_hi = igd_w_shfl(k_hi, nt - 1);
        // src_frame: the last tick up to this one whose payload was refreshed (quirk Q2), else the tick itself
        const uint32_t m_cp = igd_w_ballot(valid && (r_flags & 4u) != 0u);
        const int s_cp = igd_w_top(m_cp & igd_w_le(lane));
        const int32_t my_src = s_cp >= 0 ? t0 + s_cp : src;
        if (m_cp) src = t0 + igd_w_top(m_cp);
        if (valid) {
            igd_tx_plan_rec r;
            r.word = r_word; r.size = (uint16_t)r_size; r.flags = (uint8_t)r_flags; r.reserved = 0;
            r.src_frame = (d.flags & IGD_F_REF_QUIRKS) ? my_src : t;
            plan[i] = r;
        }
    }
    if (lane == 0) {
        d.state[c] = s;
        last_src[c] = src;
    }
}
