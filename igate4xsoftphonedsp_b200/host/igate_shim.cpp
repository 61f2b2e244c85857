// igate_shim.cpp -- see igate_shim.h.  Host bookkeeping only: the pjmedia_transport plumbing, staging
// buffers, the per-adapter mirror of struct tp_adapter's control fields, and calls into
// libigate_dsp.so.  No codec / meter / header arithmetic lives here.
#include "igate_shim.h"

#include <stdio.h>
#include <string.h>
#include <sys/time.h>
#include <time.h>

#include <mutex>
#include <vector>

namespace {

constexpr int kRxDepth = 4;   // packets one adapter may receive between two flushes (jitter bursts)
constexpr int kTxDepth = 2;   // conference-clock packets between two flushes

struct Staged {
    uint32_t size;            // size as received / as handed to send_rtp
    uint8_t bytes[IGD_PKT_MAX];
};

// One adapter = one `pjmedia_transport` PJSIP can drive + one entry of every per-channel array.
// `base` is the first member, as in the reference's struct tp_adapter (TransportAdapter.h:40-43):
// PJSIP only ever sees &base and calls through base.op.
struct Slot {
    pjmedia_transport base;
    igd_bank *bank;
    int ch;
    pjsua_call_id callID;
    pjmedia_transport *slave_tp;                                   // TransportAdapter.h:49
    pj_bool_t del_base;                                            // :43
    void *stream_user_data;                                        // :45
    void (*stream_rtp_cb)(void *user_data, void *pkt, pj_ssize_t);    // :46
    void (*stream_rtcp_cb)(void *user_data, void *pkt, pj_ssize_t);   // :47
    int keepAlivePeroid;
    Staged rxq[kRxDepth];
    int rxn;
    Staged txq[kTxDepth];
    int txn;
    long long r2sPacket;
    bool rtpAudio;
    igd_ed137_fields last;  // last accepted receive-side fields
    uint32_t ed137_value;   // host order (= ntohl(adapter->ed137_value))
    uint8_t IncomingRTP, OutgoingRTP;
    char calltype[64], trxmode[64], callIndex[64];
};

long long wall_clock_ms(void *)
{
    struct timeval tv;
    gettimeofday(&tv, nullptr);
    return (long long)tv.tv_sec * 1000 + tv.tv_usec / 1000;     // QDateTime::currentMSecsSinceEpoch()
}

}  // namespace

struct igd_bank {
    igd_ctx *ctx;
    int cap, n;                            // n = high-water mark of channel indices in use
    std::mutex mu;                         // PJSIP's ioqueue / clock threads stage while the tick owner flushes
    igd_clock_fn clock;
    void *clock_user;
    std::vector<Slot *> slots;             // [cap], nullptr = free
    std::vector<int> free_ch;
    std::vector<igd_ed137_state> state;    // [cap] sender state (TransportAdapter.h:40-93)
    std::vector<igd_ed137_ctl> ctl;        // [cap] what the setters wrote since the last tick
    std::vector<uint8_t> rtp12, payload;   // [cap][12], [cap][160] staged TX
    std::vector<uint8_t> txpk, bm;         // [cap][180], [cap]
    std::vector<uint32_t> txsz;
    std::vector<uint8_t> sendhdr;          // [cap][20] header region of each adapter's send buffer (send_pkt_buff)
    std::vector<uint32_t> kasz;            // [cap]
    std::vector<uint8_t> rxpk;             // [cap][180] staged RX (zero padded)
    std::vector<uint32_t> rxsz;
    std::vector<igd_ed137_fields> rxf;
    std::vector<uint8_t> rxpay, rxbm;      // [cap][160], [cap]
    std::vector<igd_rx_state> rxstate;     // [cap] receive-side state (igd_rx_track carries it on the GPU)
    std::vector<igd_rx_event> rxev;        // [cap]
    std::vector<uint8_t> rxpresent;        // [cap] a packet was staged this tick
    igd_event_fn on_event;                 // setIncomingED137Value(word, callID) / hang-up request
    void *on_event_user;
    std::vector<uint8_t> txmask;           // channel had a packet staged this tick
    std::vector<uint8_t> stale;            // [cap][160] what each adapter's send buffer still holds (quirk Q2)
};

static igd_bank *g_default_bank = nullptr;

static long long bank_now(igd_bank *b) { return b ? b->clock(b->clock_user) : wall_clock_ms(nullptr); }

igd_bank *igd_bank_open(int device, int max_channels)
{
    igd_ctx *ctx = nullptr;
    if (max_channels <= 0 || igd_init(device, &ctx) != IGD_OK) return nullptr;   // no GPU: no bank, no fallback
    igd_bank *b = new igd_bank();
    b->ctx = ctx;
    b->cap = max_channels;
    b->n = 0;
    b->clock = wall_clock_ms;
    b->clock_user = nullptr;
    b->slots.assign(max_channels, nullptr);
    b->state.resize(max_channels);
    b->ctl.resize(max_channels);
    b->rtp12.assign((size_t)max_channels * 12, 0);
    b->payload.assign((size_t)max_channels * IGD_FRAME, 0);
    b->txpk.assign((size_t)max_channels * IGD_PKT_MAX, 0);
    b->bm.assign(max_channels, 0);
    b->txsz.assign(max_channels, 0);
    b->sendhdr.assign((size_t)max_channels * IGD_PKT_HDR, 0);
    b->kasz.assign(max_channels, 0);
    b->rxpk.assign((size_t)max_channels * IGD_PKT_MAX, 0);
    b->rxsz.assign(max_channels, 0);
    b->rxf.resize(max_channels);
    b->rxpay.assign((size_t)max_channels * IGD_FRAME, 0);
    b->rxbm.assign(max_channels, 0);
    b->rxstate.assign(max_channels, igd_rx_state());
    b->rxev.assign(max_channels, igd_rx_event());
    b->rxpresent.assign(max_channels, 0);
    b->on_event = nullptr;
    b->on_event_user = nullptr;
    b->txmask.assign(max_channels, 0);
    b->stale.assign((size_t)max_channels * IGD_FRAME, 0);
    if (!g_default_bank) g_default_bank = b;
    return b;
}

void igd_bank_close(igd_bank *b)
{
    if (!b) return;
    for (Slot *s : b->slots) delete s;
    igd_shutdown(b->ctx);
    if (g_default_bank == b) g_default_bank = nullptr;
    delete b;
}

void igd_bank_set_default(igd_bank *b) { g_default_bank = b; }
igd_ctx *igd_bank_ctx(igd_bank *b) { return b ? b->ctx : nullptr; }
void igd_bank_set_clock(igd_bank *b, igd_clock_fn fn, void *user)
{
    if (!b) return;
    b->clock = fn ? fn : wall_clock_ms;
    b->clock_user = fn ? user : nullptr;
}

// `base` is the first member: the pointer PJSIP holds is the slot
static Slot *slot_of(pjmedia_transport *tp) { return reinterpret_cast<Slot *>(tp); }

// ------------------------------------------------------------------ the pjmedia_transport_op table
// Same split as the reference's (TransportAdapter.cpp:225-238, 348-406, 635-874, 880-1105): RTP goes
// through the ED-137 path, everything else passes through to the slave transport.
static pj_status_t shim_get_info(pjmedia_transport *tp, pjmedia_transport_info *info)
{
    Slot *s = slot_of(tp);
    return s->slave_tp ? (*s->slave_tp->op->get_info)(s->slave_tp, info) : PJ_SUCCESS;
}

// what the slave (UDP) transport calls for every received RTP packet (transport_rtp_cb, :240)
static void shim_rtp_cb(void *user_data, void *pkt, pj_ssize_t size)
{
    igd_submit_rx(&static_cast<Slot *>(user_data)->base, pkt, size);
}
static void shim_rtcp_cb(void *user_data, void *pkt, pj_ssize_t size)
{
    Slot *s = static_cast<Slot *>(user_data);
    if (s->stream_rtcp_cb) (*s->stream_rtcp_cb)(s->stream_user_data, pkt, size);      // :327-335
}

static pj_status_t shim_attach(pjmedia_transport *tp, void *user_data, const pj_sockaddr_t *rem_addr,
                               const pj_sockaddr_t *rem_rtcp, unsigned addr_len,
                               void (*rtp_cb)(void *, void *, pj_ssize_t), void (*rtcp_cb)(void *, void *, pj_ssize_t))
{
    Slot *s = slot_of(tp);                                                             // :352-386
    s->stream_user_data = user_data;
    s->stream_rtp_cb = rtp_cb;
    s->stream_rtcp_cb = rtcp_cb;
    if (!s->slave_tp) return PJ_SUCCESS;
    const pj_status_t st = (*s->slave_tp->op->attach)(s->slave_tp, s, rem_addr, rem_rtcp, addr_len, &shim_rtp_cb,
                                                      &shim_rtcp_cb);
    if (st != PJ_SUCCESS) {
        s->stream_user_data = nullptr;
        s->stream_rtp_cb = nullptr;
        s->stream_rtcp_cb = nullptr;
    }
    return st;
}

static void shim_detach(pjmedia_transport *tp, void *)
{
    Slot *s = slot_of(tp);                                                             // :392-406
    if (s->stream_user_data != nullptr || s->stream_rtp_cb != nullptr) {
        if (s->slave_tp) (*s->slave_tp->op->detach)(s->slave_tp, s);
        std::lock_guard<std::mutex> g(s->bank->mu);
        s->stream_user_data = nullptr;
        s->stream_rtp_cb = nullptr;
        s->stream_rtcp_cb = nullptr;
    }
}

static pj_status_t shim_send_rtp(pjmedia_transport *tp, const void *pkt, pj_size_t size)
{
    return igd_submit_tx(tp, pkt, size);                                               // :635: staged for the tick
}
static pj_status_t shim_send_rtcp(pjmedia_transport *tp, const void *pkt, pj_size_t size)
{
    Slot *s = slot_of(tp);                                                             // :880-890
    return s->slave_tp ? (*s->slave_tp->op->send_rtcp)(s->slave_tp, pkt, size) : PJ_SUCCESS;
}
static pj_status_t shim_send_rtcp2(pjmedia_transport *tp, const pj_sockaddr_t *addr, unsigned addr_len, const void *pkt,
                                   pj_size_t size)
{
    Slot *s = slot_of(tp);                                                             // :897-906
    return s->slave_tp ? (*s->slave_tp->op->send_rtcp2)(s->slave_tp, addr, addr_len, pkt, size) : PJ_SUCCESS;
}
static pj_status_t shim_media_create(pjmedia_transport *tp, pj_pool_t *sdp_pool, unsigned options,
                                     const pjmedia_sdp_session *rem_sdp, unsigned media_index)
{
    Slot *s = slot_of(tp);                                                             // :912-935
    return s->slave_tp ? (*s->slave_tp->op->media_create)(s->slave_tp, sdp_pool, options, rem_sdp, media_index) : PJ_SUCCESS;
}

static pj_status_t shim_encode_sdp(pjmedia_transport *tp, pj_pool_t *sdp_pool, pjmedia_sdp_session *local_sdp,
                                   const pjmedia_sdp_session *rem_sdp, unsigned media_index)
{
    Slot *s = slot_of(tp);
#ifdef IGD_HAVE_PJSIP
    // the ED-137 media attributes of a radio call, in the reference's order (:961-1037)
    if (s->bank->state[s->ch].radiostatus) {
        char ka[24];
        snprintf(ka, sizeof ka, "%d", s->keepAlivePeroid);
        const bool rxonly = strstr(s->calltype, "Rxonly") != nullptr || strcmp(s->calltype, "Rx") == 0;
        const char *attrs[][2] = {{"rtphe", "1"}, {"type", s->calltype}, {"txrxmode", s->trxmode}, {"bss", "RSSI"},
                                  {"sigtime", "1"}, {"ptt_rep", "0"}, {rxonly ? nullptr : "ptt-id", "1"},
                                  {"R2S-KeepAlivePeriod", ka}, {"R2S-KeepAliveMultiplier", "10"}};
        for (const auto &a : attrs) {
            if (!a[0]) continue;
            pjmedia_sdp_attr *at = PJ_POOL_ALLOC_T(sdp_pool, pjmedia_sdp_attr);
            pj_strdup2(sdp_pool, &at->name, a[0]);
            pj_strdup2(sdp_pool, &at->value, a[1]);
            pjmedia_sdp_attr_add(&local_sdp->media[media_index]->attr_count, local_sdp->media[media_index]->attr, at);
        }
    }
#endif
    return s->slave_tp ? (*s->slave_tp->op->encode_sdp)(s->slave_tp, sdp_pool, local_sdp, rem_sdp, media_index) : PJ_SUCCESS;
}

static pj_status_t shim_media_start(pjmedia_transport *tp, pj_pool_t *pool, const pjmedia_sdp_session *local_sdp,
                                    const pjmedia_sdp_session *rem_sdp, unsigned media_index)
{
    Slot *s = slot_of(tp);                                                             // :1051-1063
    return s->slave_tp ? (*s->slave_tp->op->media_start)(s->slave_tp, pool, local_sdp, rem_sdp, media_index) : PJ_SUCCESS;
}
static pj_status_t shim_media_stop(pjmedia_transport *tp)
{
    Slot *s = slot_of(tp);                                                             // :1068-1076
    return s->slave_tp ? (*s->slave_tp->op->media_stop)(s->slave_tp) : PJ_SUCCESS;
}
static pj_status_t shim_simulate_lost(pjmedia_transport *tp, pjmedia_dir dir, unsigned pct_lost)
{
    Slot *s = slot_of(tp);                                                             // :1081-1087
    return s->slave_tp ? (*s->slave_tp->op->simulate_lost)(s->slave_tp, dir, pct_lost) : PJ_SUCCESS;
}
static pj_status_t shim_destroy(pjmedia_transport *tp)
{
    Slot *s = slot_of(tp);                                                             // :1092-1105
    if (s->del_base && s->slave_tp && s->slave_tp->op->destroy) (*s->slave_tp->op->destroy)(s->slave_tp);
    igd_bank *b = s->bank;
    {
        std::lock_guard<std::mutex> g(b->mu);
        b->slots[s->ch] = nullptr;              // the channel goes back to the bank
        b->free_ch.push_back(s->ch);
        b->state[s->ch].radiostatus = 0;
    }
    delete s;
    return PJ_SUCCESS;
}

static pjmedia_transport_op g_shim_op = {
    &shim_get_info,  &shim_attach,     &shim_detach,      &shim_send_rtp,   &shim_send_rtcp,     &shim_send_rtcp2,
    &shim_media_create, &shim_encode_sdp, &shim_media_start, &shim_media_stop, &shim_simulate_lost, &shim_destroy};

pj_status_t pjmedia_custom_tp_adapter_create(pjmedia_endpt *, const char *name, pjmedia_transport *transport,
                                             pj_bool_t del_base, pj_bool_t radiocall, pj_bool_t callIn,
                                             const char *calltype, pjsua_call_id callId, pjmedia_transport **p_tp,
                                             const char *callIndex, const char *trxmode, int keepAlivePeroid, pj_bool_t,
                                             pj_bool_t)
{
    igd_bank *b = g_default_bank;
    if (!b || !p_tp) return PJ_EINVAL;
    Slot *s = new Slot();
    memset(static_cast<void *>(s), 0, sizeof(*s));
    {
        std::lock_guard<std::mutex> g(b->mu);
        if (!b->free_ch.empty()) {
            s->ch = b->free_ch.back();
            b->free_ch.pop_back();
        } else if (b->n < b->cap) {
            s->ch = b->n++;
        } else {
            delete s;
            return PJ_EINVAL;
        }
        b->slots[s->ch] = s;
    }
    s->bank = b;
    snprintf(s->base.name, sizeof(s->base.name), name ? name : "tpad%p", static_cast<void *>(s));   // :89-100
    s->base.type = (pjmedia_transport_type)(PJMEDIA_TRANSPORT_TYPE_USER + 1);                        // :101-102
    s->base.op = &g_shim_op;                                                                         // :103
    s->slave_tp = transport;                                                                         // :106
    s->del_base = del_base;
    s->callID = callId;
    s->keepAlivePeroid = keepAlivePeroid;
    strncpy(s->calltype, calltype ? calltype : "", sizeof(s->calltype) - 1);
    strncpy(s->callIndex, callIndex ? callIndex : "", sizeof(s->callIndex) - 1);
    strncpy(s->trxmode, trxmode ? trxmode : "", sizeof(s->trxmode) - 1);
    const long long now = bank_now(b);
    s->r2sPacket = now;                                                    // TransportAdapter.cpp:122
    memset(&b->rxstate[s->ch], 0, sizeof(igd_rx_state));
    b->rxstate[s->ch].r2sPacket = now;
    igd_ed137_state_init(&b->state[s->ch], radiocall, callIn, s->calltype, keepAlivePeroid, now);
    memset(&b->ctl[s->ch], 0, sizeof(igd_ed137_ctl));
    memset(&b->sendhdr[(size_t)s->ch * IGD_PKT_HDR], 0, IGD_PKT_HDR);      // PJ_POOL_ZALLOC_T, :97
    memset(&b->stale[(size_t)s->ch * IGD_FRAME], 0, IGD_FRAME);
    *p_tp = &s->base;
    return PJ_SUCCESS;
}

pj_status_t decodeRtp(void *pkt, custom_rtp_hdr **hdr)
{
    *hdr = static_cast<custom_rtp_hdr *>(pkt);      // TransportAdapter.cpp:408-415: a cast, always PJ_SUCCESS
    return PJ_SUCCESS;
}

// The setters only record what the next tick's batched call must see; like the
// reference's they NULL-check and return PJ_SUCCESS regardless (TransportAdapter.cpp:135-223).
pj_status_t setAdapterPtt(pjmedia_transport *tp, bool pttval, int priority, int userRec)
{
    if (Slot *s = slot_of(tp)) {
        igd_ed137_ctl &c = s->bank->ctl[s->ch];
        c.pttstatus = pttval; c.pttpriority = (uint8_t)priority; c.callRecorder = userRec != 0;
    }
    return PJ_SUCCESS;
}
pj_status_t setTxRxSlaveEnable(pjmedia_transport *tp, pj_bool_t rx, pj_bool_t tx)
{
    if (Slot *s = slot_of(tp)) {
        igd_ed137_state &st = s->bank->state[s->ch];
        st.txSlaveEnableChanged = (uint8_t)tx; st.rxSlaveEnableChanged = (uint8_t)rx; st.trxSlaveEnableChangedCount = 0;
    }
    return PJ_SUCCESS;
}
pj_status_t setAdapterRadioModeAndType(pjmedia_transport *tp, char const *type, char const *txrxmode)
{
    if (Slot *s = slot_of(tp)) {
        strncpy(s->calltype, type, sizeof(s->calltype) - 1);
        strncpy(s->trxmode, txrxmode, sizeof(s->trxmode) - 1);
        s->bank->state[s->ch].calltype_flags = (uint8_t)igd_calltype_flags(s->calltype);
    }
    return PJ_SUCCESS;
}
pj_status_t setAdapterQslOn(pjmedia_transport *tp, bool sqlval, int, pj_uint32_t bssi)
{
    if (Slot *s = slot_of(tp)) {
        igd_ed137_ctl &c = s->bank->ctl[s->ch];
        c.sqlstatus = sqlval; c.ed137_bssi = (uint8_t)bssi;
    }
    return PJ_SUCCESS;
}
pj_status_t setAdapterPttId(pjmedia_transport *tp, int pttid)
{
    if (Slot *s = slot_of(tp)) s->bank->ctl[s->ch].pttid = (uint8_t)pttid;
    return PJ_SUCCESS;
}
pj_status_t setcallRecorder(pjmedia_transport *tp, bool val)
{
    if (Slot *s = slot_of(tp)) s->bank->ctl[s->ch].callRecorder = val;
    return PJ_SUCCESS;
}
pj_status_t setCallType(pjmedia_transport *tp, char const *calltype)
{
    if (Slot *s = slot_of(tp)) {
        strncpy(s->calltype, calltype, sizeof(s->calltype) - 1);
        s->bank->state[s->ch].calltype_flags = (uint8_t)igd_calltype_flags(s->calltype);
    }
    return PJ_SUCCESS;
}
pj_uint32_t get_ed137_value(pjmedia_transport *tp) { return tp ? slot_of(tp)->ed137_value : 0; }
long long getR2SStatus(pjmedia_transport *tp)
{
    return tp ? slot_of(tp)->r2sPacket : bank_now(g_default_bank) - 3000;      // TransportAdapter.cpp:317-325
}
int get_IPRadioBss(pjmedia_transport *tp) { return tp ? slot_of(tp)->last.bss : 0; }
int get_IPRadioPttStatus(pjmedia_transport *tp) { return tp ? slot_of(tp)->last.ptt_type : 0; }
int get_IPRadioPttId(pjmedia_transport *tp) { return tp ? slot_of(tp)->last.ptt_id : 0; }
int get_IPRadioSquelch(pjmedia_transport *tp) { return tp ? slot_of(tp)->last.squelch : 0; }
bool get_IPRadioStatus(pjmedia_transport *tp) { return tp && (slot_of(tp)->last.flags & IGD_EDF_ACTIVE); }
uint8_t get_IncomingRTP(pjmedia_transport *tp) { return tp ? slot_of(tp)->IncomingRTP : 0; }
uint8_t get_OutgoingRTP(pjmedia_transport *tp) { return tp ? slot_of(tp)->OutgoingRTP : 0; }

pj_status_t igd_submit_tx(pjmedia_transport *tp, const void *pkt, pj_size_t size)
{
    Slot *s = slot_of(tp);
    if (!s || !pkt || size < 12 || size > 12 + IGD_FRAME) return PJ_EINVAL;
    std::lock_guard<std::mutex> g(s->bank->mu);
    if (s->txn == kTxDepth) {                    // the tick owner is late: the oldest frame is lost
        memmove(&s->txq[0], &s->txq[1], sizeof(Staged) * (kTxDepth - 1));
        s->txn--;
    }
    Staged &q = s->txq[s->txn++];
    q.size = (uint32_t)size;
    memcpy(q.bytes, pkt, size);
    return PJ_SUCCESS;
}

static void send_to_slave(void *, pjmedia_transport *tp, const void *pkt, pj_size_t size)
{
    Slot *s = slot_of(tp);                       // pjmedia_transport_send_rtp(adapter->slave_tp, ...), :848
    if (s->slave_tp) (*s->slave_tp->op->send_rtp)(s->slave_tp, pkt, size);
}

// the 20 bytes the reference's send buffer holds after a transport_send_rtp call: the stamped header when a
// packet was built, else the raw first 20 bytes of the PJSIP packet (copied before the throttle returns, :653)
static void remember_send_header(igd_bank *b, int c)
{
    uint8_t *h = &b->sendhdr[(size_t)c * IGD_PKT_HDR];
    if (b->txsz[c]) {
        memcpy(h, &b->txpk[(size_t)c * IGD_PKT_MAX], IGD_PKT_HDR);
    } else {
        memcpy(h, &b->rtp12[(size_t)c * 12], 12);
        memcpy(h + 12, &b->payload[(size_t)c * IGD_FRAME], 8);
    }
}

int igd_bank_flush_tx(igd_bank *b, long long now_ms, unsigned flags, igd_send_fn send, void *user)
{
    if (!b) return -1;
    if (!send) send = &send_to_slave;
    int sent = 0;
    for (;;) {
        int n, staged = 0;
        {   // take the oldest staged packet of every adapter
            std::lock_guard<std::mutex> g(b->mu);
            n = b->n;
            for (int c = 0; c < n; c++) {
                Slot *s = b->slots[c];
                b->txmask[c] = 0;
                if (!s || s->txn == 0) continue;
                const Staged &q = s->txq[0];
                memcpy(&b->rtp12[(size_t)c * 12], q.bytes, 12);
                memset(&b->payload[(size_t)c * IGD_FRAME], 0, IGD_FRAME);
                memcpy(&b->payload[(size_t)c * IGD_FRAME], q.bytes + 12, q.size - 12);
                memmove(&s->txq[0], &s->txq[1], sizeof(Staged) * (kTxDepth - 1));
                s->txn--;
                b->txmask[c] = 1;
                staged++;
            }
        }
        if (!staged) break;
        // channels without a staged packet this tick must not advance their sender state: mark them
        // non-radio for this call (transport_send_rtp is simply not invoked for them in the reference)
        std::vector<uint8_t> saved(n);
        for (int c = 0; c < n; c++) {
            saved[c] = b->state[c].radiostatus;
            if (!b->txmask[c]) b->state[c].radiostatus = 0;
        }
        igd_ed137_pack_desc d;
        memset(&d, 0, sizeof(d));
        d.struct_size = sizeof(d);
        d.mem = IGD_MEM_HOST;
        d.F = 1; d.C = n;
        d.flags = flags;
        d.payload_len = IGD_FRAME;
        d.out_stride = IGD_PKT_MAX;
        d.tick_ms = 20;
        d.now_ms0 = now_ms;
        d.rtp12 = b->rtp12.data(); d.payload = b->payload.data(); d.ctl = b->ctl.data(); d.state = b->state.data();
        d.pkts = b->txpk.data(); d.sizes = b->txsz.data(); d.bytemean_out = b->bm.data();
        d.stale_payload = b->stale.data();
        const int rc = igd_ed137_pack(b->ctx, &d);
        for (int c = 0; c < n; c++) b->state[c].radiostatus = saved[c];
        if (rc != IGD_OK) return rc;
        for (int c = 0; c < n; c++) {
            Slot *s = b->slots[c];
            if (!s || !b->txmask[c]) continue;
            if (saved[c]) remember_send_header(b, c);
            if (b->txsz[c]) {
                const uint8_t *pk = &b->txpk[(size_t)c * IGD_PKT_MAX];
                if ((pk[1] & 0x7F) != 123) s->OutgoingRTP = b->bm[c];      // setOutgoingRTP
                send(user, &s->base, pk, b->txsz[c]);
                sent++;
            }
        }
    }
    return sent;
}

// sendR2SStatus for the channels of `mask` (nullptr = all): igd_ed137_keepalive over the bank's state and the
// header region of every send buffer; a 20-byte packet leaves where the reference's would (:617-621)
static int keepalive_tick(igd_bank *b, long long now_ms, const Slot *only)
{
    const int n = b->n;
    if (n == 0) return 0;
    std::vector<uint8_t> saved(n);
    for (int c = 0; c < n; c++) {
        saved[c] = b->state[c].radiostatus;
        if (!b->slots[c] || (only && b->slots[c] != only)) b->state[c].radiostatus = 0;   // not called for them
        else {                                     // the setters' values reach the state at the next call
            igd_ed137_state &st = b->state[c];
            const igd_ed137_ctl &k = b->ctl[c];
            st.pttstatus = k.pttstatus; st.sqlstatus = k.sqlstatus; st.pttpriority = k.pttpriority;
            st.ed137_bssi = k.ed137_bssi; st.pttid = k.pttid; st.callRecorder = k.callRecorder;
        }
    }
    const int rc = igd_ed137_keepalive(b->ctx, b->sendhdr.data(), b->state.data(), (size_t)n, now_ms, b->kasz.data(),
                                       IGD_MEM_HOST);
    int sent = 0;
    for (int c = 0; c < n; c++) {
        const bool mine = b->slots[c] && (!only || b->slots[c] == only);
        if (!mine) b->state[c].radiostatus = saved[c];
        if (rc == IGD_OK && mine && b->kasz[c]) {
            send_to_slave(nullptr, &b->slots[c]->base, &b->sendhdr[(size_t)c * IGD_PKT_HDR], b->kasz[c]);
            sent++;
        }
    }
    return rc == IGD_OK ? sent : rc;
}

void sendR2SStatus(pjmedia_transport *tp)
{
    if (Slot *s = slot_of(tp)) keepalive_tick(s->bank, bank_now(s->bank), s);
}
int igd_bank_keepalive(igd_bank *b, long long now_ms) { return b ? keepalive_tick(b, now_ms, nullptr) : -1; }

pj_status_t igd_submit_rx(pjmedia_transport *tp, const void *pkt, pj_ssize_t size)
{
    Slot *s = slot_of(tp);
    if (!s || !pkt || size < 0) return PJ_EINVAL;
    std::lock_guard<std::mutex> g(s->bank->mu);
    if (s->rxn == kRxDepth) {                    // burst longer than the queue: the oldest packet is lost
        memmove(&s->rxq[0], &s->rxq[1], sizeof(Staged) * (kRxDepth - 1));
        s->rxn--;
    }
    Staged &q = s->rxq[s->rxn++];
    const size_t n = (size_t)size < (size_t)IGD_PKT_MAX ? (size_t)size : (size_t)IGD_PKT_MAX;
    memset(q.bytes, 0, IGD_PKT_MAX);
    memcpy(q.bytes, pkt, n);
    q.size = (uint32_t)size;
    return PJ_SUCCESS;
}

void igd_bank_set_event_cb(igd_bank *b, igd_event_fn fn, void *user)
{
    if (b) { b->on_event = fn; b->on_event_user = user; }
}

// one igd_rx_track call over every adapter of the bank: F = 1 tick, `present` = who received a
// packet, run_watchdog selects detectR2SPacketAndReconn's check (roip_ed137.cpp:1767-1780)
static int rx_track_tick(igd_bank *b, int n, long long now_ms, int r2s_period_ms, bool run_watchdog)
{
    igd_rx_track_desc d;
    memset(&d, 0, sizeof(d));
    d.struct_size = sizeof(d);
    d.mem = IGD_MEM_HOST;
    d.F = 1; d.C = n;
    d.tick_ms = 20;
    d.r2s_period_ms = r2s_period_ms;
    d.wd_ticks = run_watchdog ? 1 : 0;
    d.now_ms0 = now_ms;
    d.fields = b->rxf.data();
    d.present = b->rxpresent.data();
    d.state = b->rxstate.data();
    d.events = b->rxev.data();
    return igd_rx_track(b->ctx, &d);
}

int igd_bank_flush_rx(igd_bank *b, long long now_ms, igd_send_fn stream_cb, void *user)
{
    if (!b) return -1;
    int parsed = 0;
    for (;;) {
        int n, staged = 0;
        {   // take the oldest staged packet of every adapter
            std::lock_guard<std::mutex> g(b->mu);
            n = b->n;
            for (int c = 0; c < n; c++) {
                Slot *s = b->slots[c];
                b->rxpresent[c] = 0;
                b->rxsz[c] = 0;                                        // 0: nothing arrived
                if (!s || s->rxn == 0) continue;
                memcpy(&b->rxpk[(size_t)c * IGD_PKT_MAX], s->rxq[0].bytes, IGD_PKT_MAX);
                b->rxsz[c] = s->rxq[0].size;
                b->rxpresent[c] = 1;
                memmove(&s->rxq[0], &s->rxq[1], sizeof(Staged) * (kRxDepth - 1));
                s->rxn--;
                staged++;
            }
        }
        if (!staged) break;
        int rc = igd_ed137_parse(b->ctx, b->rxpk.data(), b->rxsz.data(), (size_t)n, IGD_PKT_MAX, b->rxf.data(),
                                 b->rxpay.data(), IGD_MEM_HOST);
        if (rc == IGD_OK)   // setIncomingRTP: byte-mean of the payload (roip_ed137.cpp:6541-6587), on the GPU
            rc = igd_bytemean(b->ctx, b->rxpay.data(), (size_t)n, IGD_FRAME, IGD_FRAME, 0, b->rxbm.data(), IGD_MEM_HOST);
        if (rc == IGD_OK)   // word latch, r2sPacket stamp, audio <-> keep-alive edge (TransportAdapter.cpp:252-315), on the GPU
            rc = rx_track_tick(b, n, now_ms, 200, false);
        if (rc != IGD_OK) return rc;
        for (int c = 0; c < n; c++) {
            Slot *s = b->slots[c];
            if (!s || !b->rxpresent[c]) continue;
            const igd_ed137_fields &f = b->rxf[c];
            const igd_rx_event &e = b->rxev[c];
            parsed++;
            s->r2sPacket = b->rxstate[c].r2sPacket;                          // :289,302,311
            s->rtpAudio = b->rxstate[c].rtpAudio != 0;
            s->ed137_value = e.word;                                         // :252-256 latch
            if (f.accepted && !(f.flags & IGD_EDF_DROPPED)) s->last = f;
            if (e.flags & IGD_RXE_AUDIO) {                                   // :298-307
                if (f.payload_len == IGD_FRAME) s->IncomingRTP = b->rxbm[c];
                const uint32_t fwd = b->rxsz[c] < (uint32_t)IGD_PKT_MAX ? b->rxsz[c] : (uint32_t)IGD_PKT_MAX;
                if (stream_cb)
                    stream_cb(user, &s->base, &b->rxpk[(size_t)c * IGD_PKT_MAX], fwd);
                else if (s->stream_rtp_cb)                                   // adapter->stream_rtp_cb(...), :301
                    (*s->stream_rtp_cb)(s->stream_user_data, &b->rxpk[(size_t)c * IGD_PKT_MAX], (pj_ssize_t)fwd);
            }
            if ((e.flags & IGD_RXE_EDGE) && b->on_event)                     // :304-306, :312-314
                b->on_event(b->on_event_user, &s->base, IGD_RXE_EDGE, e.word);
        }
    }
    return parsed;
}

int igd_bank_r2s_watchdog(igd_bank *b, long long now_ms, int r2s_period_ms)
{
    if (!b) return -1;
    const int n = b->n;
    if (n == 0) return 0;
    for (int c = 0; c < n; c++) b->rxpresent[c] = 0;
    const int rc = rx_track_tick(b, n, now_ms, r2s_period_ms, true);
    if (rc != IGD_OK) return rc;
    int hangups = 0;
    for (int c = 0; c < n; c++) {
        if (b->slots[c] && (b->rxev[c].flags & IGD_RXE_HANGUP)) {
            hangups++;
            if (b->on_event) b->on_event(b->on_event_user, &b->slots[c]->base, IGD_RXE_HANGUP, b->rxev[c].word);
        }
    }
    return hangups;
}

// ------------------------------------------------------------------ WavWriter
WavWriter::WavWriter() : m_bank(nullptr), m_running(false), m_quirks(true), m_rate(8000), m_law(IGD_LAW_ULAW) {}
WavWriter::~WavWriter() {}
void WavWriter::attach(igd_bank *bank, bool ref_quirks, int law) { m_bank = bank; m_quirks = ref_quirks; m_law = law; }
bool WavWriter::isRunning() { return m_running; }

void WavWriter::start(std::string prefix, int rate)
{
    char name[160];
    time_t now = time(nullptr);
    struct tm *d = localtime(&now);                                          // WavWriter.cpp:197-220
    snprintf(name, sizeof(name), "%s%04d%02d%02d%02d%02d%02d.wav", prefix.c_str(), d->tm_year + 1900, d->tm_mon + 1,
             d->tm_mday, d->tm_hour, d->tm_min, d->tm_sec);
    m_file = name;
    m_rate = rate;
    m_data.clear();
    m_running = true;
}

void WavWriter::wav_write(unsigned char *buf, unsigned int len)
{
    if (m_running) m_data.append(reinterpret_cast<const char *>(buf), len);
}

void WavWriter::writeRTPWav(const char *, const char *payloadbuf, unsigned int, unsigned int payloadlen)
{
    if (m_running) wav_write(reinterpret_cast<unsigned char *>(const_cast<char *>(payloadbuf)), payloadlen);
}

void WavWriter::stop()
{
    if (!m_running) return;
    m_running = false;
    igd_bank *b = m_bank ? m_bank : g_default_bank;
    if (!b) return;                                                           // no GPU context: nothing is written
    std::string out(igd_wav_size(m_data.size(), m_quirks), '\0');
    size_t len = 0;
    if (igd_wav_image(b->ctx, reinterpret_cast<const uint8_t *>(m_data.data()), m_data.size(), m_rate, m_law, m_quirks,
                      reinterpret_cast<uint8_t *>(&out[0]), &len, IGD_MEM_HOST) != IGD_OK)
        return;
    if (FILE *f = fopen(m_file.c_str(), "wb")) {
        fwrite(out.data(), 1, len, f);
        fclose(f);
    }
}

int igd_audio_level_percent(igd_bank *b, const int32_t *raw, size_t n, int32_t *percent)
{
    return b ? igd_level_percent(b->ctx, raw, n, percent, IGD_MEM_HOST) : IGD_ENODEV;
}
