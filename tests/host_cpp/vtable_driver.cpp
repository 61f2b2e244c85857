// vtable_driver.cpp -- plays PJSIP for the product shim (igate4xsoftphonedsp_b200/host/igate_shim.cpp).
//
// Built with -DIGD_HAVE_PJSIP against the same stand-in pjproject headers (oracle/ref_shim) the
// reference's own TransportAdapter.cpp is compiled against for the oracle pins, so the shim sees a
// "real" <pjmedia/transport.h>.  Every adapter is created by the factory and then driven ONLY the way
// PJSIP drives a media transport: tp->op->attach / send_rtp / encode_sdp / destroy ..., and through the
// RTP callback the adapter registered with its (fake) slave transport -- the same moves
// oracle/ref_harness.cpp makes on the reference's adapter, so tests/test_gpu_vtable_shim.py can compare
// the two packet for packet.  Test-only.
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "../../igate4xsoftphonedsp_b200/host/igate_shim.h"

namespace {

long long g_clock = 0;
long long clock_fn(void *) { return g_clock; }

struct Handle;
struct Slave {
    pjmedia_transport base;
    Handle *h;
};
struct Handle {
    pjmedia_transport *tp = nullptr;
    Slave slave;
    void *cb_user = nullptr;
    void (*rtp_cb)(void *, void *, pj_ssize_t) = nullptr;
    void (*rtcp_cb)(void *, void *, pj_ssize_t) = nullptr;
    std::vector<uint8_t> sent;
    int sends = 0, sends_taken = 0;
    int stream_rtp = 0, stream_rtcp = 0, stream_taken = 0;
    size_t stream_last_size = 0;
    std::vector<uint8_t> stream_last;
    int slave_calls[12] = {0};
    unsigned events = 0;
    uint32_t event_word = 0;
};

pj_status_t s_get_info(pjmedia_transport *tp, pjmedia_transport_info *info)
{ ((Slave *)tp)->h->slave_calls[0]++; if (info) info->igd_stub_calls++; return PJ_SUCCESS; }
pj_status_t s_attach(pjmedia_transport *tp, void *user_data, const pj_sockaddr_t *, const pj_sockaddr_t *, unsigned,
                     void (*rtp_cb)(void *, void *, pj_ssize_t), void (*rtcp_cb)(void *, void *, pj_ssize_t))
{
    Handle *h = ((Slave *)tp)->h;
    h->slave_calls[1]++;
    h->cb_user = user_data; h->rtp_cb = rtp_cb; h->rtcp_cb = rtcp_cb;
    return PJ_SUCCESS;
}
void s_detach(pjmedia_transport *tp, void *)
{
    Handle *h = ((Slave *)tp)->h;
    h->slave_calls[2]++;
    h->cb_user = nullptr; h->rtp_cb = nullptr; h->rtcp_cb = nullptr;
}
pj_status_t s_send_rtp(pjmedia_transport *tp, const void *pkt, pj_size_t size)
{
    Handle *h = ((Slave *)tp)->h;
    h->slave_calls[3]++;
    h->sent.assign((const uint8_t *)pkt, (const uint8_t *)pkt + size);
    h->sends++;
    return PJ_SUCCESS;
}
pj_status_t s_send_rtcp(pjmedia_transport *tp, const void *, pj_size_t) { ((Slave *)tp)->h->slave_calls[4]++; return PJ_SUCCESS; }
pj_status_t s_send_rtcp2(pjmedia_transport *tp, const pj_sockaddr_t *, unsigned, const void *, pj_size_t)
{ ((Slave *)tp)->h->slave_calls[5]++; return PJ_SUCCESS; }
pj_status_t s_media_create(pjmedia_transport *tp, pj_pool_t *, unsigned, const pjmedia_sdp_session *, unsigned)
{ ((Slave *)tp)->h->slave_calls[6]++; return PJ_SUCCESS; }
pj_status_t s_encode_sdp(pjmedia_transport *tp, pj_pool_t *, pjmedia_sdp_session *, const pjmedia_sdp_session *, unsigned)
{ ((Slave *)tp)->h->slave_calls[7]++; return PJ_SUCCESS; }
pj_status_t s_media_start(pjmedia_transport *tp, pj_pool_t *, const pjmedia_sdp_session *, const pjmedia_sdp_session *, unsigned)
{ ((Slave *)tp)->h->slave_calls[8]++; return PJ_SUCCESS; }
pj_status_t s_media_stop(pjmedia_transport *tp) { ((Slave *)tp)->h->slave_calls[9]++; return PJ_SUCCESS; }
pj_status_t s_simulate_lost(pjmedia_transport *tp, pjmedia_dir, unsigned) { ((Slave *)tp)->h->slave_calls[10]++; return PJ_SUCCESS; }
pj_status_t s_destroy(pjmedia_transport *tp) { ((Slave *)tp)->h->slave_calls[11]++; return PJ_SUCCESS; }
pjmedia_transport_op g_slave_op = {&s_get_info,     &s_attach,     &s_detach,      &s_send_rtp,   &s_send_rtcp,     &s_send_rtcp2,
                                   &s_media_create, &s_encode_sdp, &s_media_start, &s_media_stop, &s_simulate_lost, &s_destroy};

void stream_rtp_cb(void *user, void *pkt, pj_ssize_t size)
{
    Handle *h = (Handle *)user;
    h->stream_rtp++;
    h->stream_last_size = (size_t)size;
    h->stream_last.assign((const uint8_t *)pkt, (const uint8_t *)pkt + size);
}
void stream_rtcp_cb(void *user, void *, pj_ssize_t) { ((Handle *)user)->stream_rtcp++; }

std::vector<Handle *> g_handles;
void on_event(void *, pjmedia_transport *tp, unsigned flag, pj_uint32_t word)
{
    for (Handle *h : g_handles)
        if (h->tp == tp) { h->events |= flag; h->event_word = word; }
}

}  // namespace

extern "C" {

void *shimta_open(int device, int max_channels)
{
    igd_bank *b = igd_bank_open(device, max_channels);
    if (!b) return nullptr;
    igd_bank_set_default(b);
    igd_bank_set_clock(b, &clock_fn, nullptr);
    igd_bank_set_event_cb(b, &on_event, nullptr);
    return b;
}
void shimta_close(void *bank) { igd_bank_close((igd_bank *)bank); g_handles.clear(); }
void shimta_set_clock(long long ms) { g_clock = ms; }
int shimta_offsetof_base_is_zero(void *hv) { return (void *)((Handle *)hv)->tp != nullptr && ((Handle *)hv)->tp->op != nullptr; }

void *shimta_create(int radiocall, int callIn, const char *calltype, int call_id, const char *callIndex, const char *trxmode,
                    int keepAlivePeroid, int connToRadio, int pttWithPayload, int attach)
{
    Handle *h = new Handle();
    memset(&h->slave.base, 0, sizeof h->slave.base);
    snprintf(h->slave.base.name, sizeof h->slave.base.name, "fakeudp");
    h->slave.base.type = PJMEDIA_TRANSPORT_TYPE_UDP;
    h->slave.base.op = &g_slave_op;
    h->slave.h = h;
    pj_status_t st = pjmedia_custom_tp_adapter_create(nullptr, nullptr, &h->slave.base, PJ_TRUE, radiocall, callIn, calltype,
                                                      call_id, &h->tp, callIndex, trxmode, keepAlivePeroid, connToRadio,
                                                      pttWithPayload);
    if (st != PJ_SUCCESS || !h->tp || !h->tp->op) { delete h; return nullptr; }
    if (attach && pjmedia_transport_attach(h->tp, h, nullptr, nullptr, 0, &stream_rtp_cb, &stream_rtcp_cb) != PJ_SUCCESS) {
        delete h;
        return nullptr;
    }
    g_handles.push_back(h);
    return h;
}

void shimta_destroy(void *hv)
{
    Handle *h = (Handle *)hv;
    pjmedia_transport_detach(h->tp, h);
    pjmedia_transport_close(h->tp);
    for (size_t i = 0; i < g_handles.size(); ++i)
        if (g_handles[i] == h) { g_handles.erase(g_handles.begin() + (long)i); break; }
    delete h;
}

// conference clock thread: pjmedia_transport_send_rtp(tp, ...) == tp->op->send_rtp (the shim stages)
int shimta_send_rtp(void *hv, const uint8_t *pkt, size_t size) { return pjmedia_transport_send_rtp(((Handle *)hv)->tp, pkt, size); }
// tick owner: one batched GPU call; packets leave through every adapter's slave transport
int shimta_flush_tx(void *bank, unsigned flags) { return igd_bank_flush_tx((igd_bank *)bank, g_clock, flags, nullptr, nullptr); }
// what this adapter handed its slave transport since the last take (0 = nothing)
size_t shimta_take_sent(void *hv, uint8_t *out)
{
    Handle *h = (Handle *)hv;
    if (h->sends == h->sends_taken) return 0;
    h->sends_taken = h->sends;
    memcpy(out, h->sent.data(), h->sent.size());
    return h->sent.size();
}
void shimta_sendR2SStatus(void *hv) { sendR2SStatus(((Handle *)hv)->tp); }
int shimta_bank_keepalive(void *bank) { return igd_bank_keepalive((igd_bank *)bank, g_clock); }

// ioqueue worker: the slave transport calls the RTP callback the adapter registered in attach()
int shimta_rx(void *hv, const uint8_t *pkt, size_t size)
{
    Handle *h = (Handle *)hv;
    if (!h->rtp_cb) return -2;
    std::vector<uint8_t> b(pkt, pkt + size);
    b.resize(size + 8, 0);
    (*h->rtp_cb)(h->cb_user, b.data(), (pj_ssize_t)size);
    return 0;
}
int shimta_rtcp(void *hv, const uint8_t *pkt, size_t size)
{
    Handle *h = (Handle *)hv;
    if (!h->rtcp_cb) return -2;
    std::vector<uint8_t> b(pkt, pkt + size);
    int before = h->stream_rtcp;
    (*h->rtcp_cb)(h->cb_user, b.data(), (pj_ssize_t)size);
    return h->stream_rtcp != before;
}
int shimta_flush_rx(void *bank) { return igd_bank_flush_rx((igd_bank *)bank, g_clock, nullptr, nullptr); }
int shimta_watchdog(void *bank, int period) { return igd_bank_r2s_watchdog((igd_bank *)bank, g_clock, period); }
// packets forwarded to the stream since the last take; the last one's bytes in out
int shimta_take_stream(void *hv, uint8_t *out, size_t *size)
{
    Handle *h = (Handle *)hv;
    int n = h->stream_rtp - h->stream_taken;
    h->stream_taken = h->stream_rtp;
    if (n && out) memcpy(out, h->stream_last.data(), h->stream_last.size());
    if (size) *size = n ? h->stream_last_size : 0;
    return n;
}
unsigned shimta_take_events(void *hv, uint32_t *word)
{
    Handle *h = (Handle *)hv;
    unsigned e = h->events;
    if (word) *word = h->event_word;
    h->events = 0;
    return e;
}

void shimta_setAdapterPtt(void *hv, int pttval, int priority, int userRec) { setAdapterPtt(((Handle *)hv)->tp, pttval != 0, priority, userRec); }
void shimta_setTxRxSlaveEnable(void *hv, int rx, int tx) { setTxRxSlaveEnable(((Handle *)hv)->tp, rx, tx); }
void shimta_setAdapterQslOn(void *hv, int sqlval, int priority, uint32_t bssi) { setAdapterQslOn(((Handle *)hv)->tp, sqlval != 0, priority, bssi); }
void shimta_setAdapterPttId(void *hv, int pttid) { setAdapterPttId(((Handle *)hv)->tp, pttid); }
void shimta_setcallRecorder(void *hv, int val) { setcallRecorder(((Handle *)hv)->tp, val != 0); }
void shimta_setCallType(void *hv, const char *calltype) { setCallType(((Handle *)hv)->tp, calltype); }
uint32_t shimta_get_ed137_value(void *hv) { return get_ed137_value(hv ? ((Handle *)hv)->tp : nullptr); }
long long shimta_getR2SStatus(void *hv) { return getR2SStatus(hv ? ((Handle *)hv)->tp : nullptr); }
void shimta_levels(void *hv, int *o)
{
    pjmedia_transport *tp = ((Handle *)hv)->tp;
    o[0] = get_IncomingRTP(tp); o[1] = get_OutgoingRTP(tp);
    o[2] = get_IPRadioBss(tp); o[3] = get_IPRadioPttStatus(tp); o[4] = get_IPRadioPttId(tp);
    o[5] = get_IPRadioSquelch(tp); o[6] = get_IPRadioStatus(tp);
}
void shimta_counters(void *hv, int *out /* [16] */)
{
    Handle *h = (Handle *)hv;
    out[0] = h->sends; out[1] = h->stream_rtp; out[2] = h->stream_rtcp; out[3] = (int)h->stream_last_size;
    for (int i = 0; i < 12; ++i) out[4 + i] = h->slave_calls[i];
}
int shimta_vtable_passthrough(void *hv)
{
    Handle *h = (Handle *)hv;
    pjmedia_transport *tp = h->tp;
    int c0[12]; memcpy(c0, h->slave_calls, sizeof c0);
    pjmedia_transport_info info = {0};
    uint8_t b[8] = {0};
    pj_pool_t *pool = pjmedia_endpt_create_pool(nullptr, "tmp%p", 0, 0);
    (*tp->op->get_info)(tp, &info);
    (*tp->op->send_rtcp)(tp, b, sizeof b);
    (*tp->op->send_rtcp2)(tp, nullptr, 0, b, sizeof b);
    (*tp->op->media_create)(tp, pool, 0, nullptr, 0);
    (*tp->op->media_start)(tp, pool, nullptr, nullptr, 0);
    (*tp->op->media_stop)(tp);
    (*tp->op->simulate_lost)(tp, PJMEDIA_DIR_ENCODING, 10);
    pj_pool_release(pool);
    int bits = 0;
    for (int i = 0; i < 12; ++i) if (h->slave_calls[i] != c0[i]) bits |= 1 << i;
    return bits;
}
size_t shimta_encode_sdp(void *hv, char *out, size_t cap)
{
    Handle *h = (Handle *)hv;
    pj_pool_t *pool = pjmedia_endpt_create_pool(nullptr, "sdp%p", 0, 0);
    pjmedia_sdp_session sdp; memset(&sdp, 0, sizeof sdp);
    pjmedia_sdp_media m; memset(&m, 0, sizeof m);
    sdp.media_count = 1; sdp.media[0] = &m;
    (*h->tp->op->encode_sdp)(h->tp, pool, &sdp, nullptr, 0);
    std::string s;
    for (unsigned i = 0; i < m.attr_count; ++i)
        s += std::string(m.attr[i]->name.ptr, (size_t)m.attr[i]->name.slen) + ":" +
             std::string(m.attr[i]->value.ptr, (size_t)m.attr[i]->value.slen) + "\n";
    pj_pool_release(pool);
    size_t n = s.size() < cap - 1 ? s.size() : cap - 1;
    memcpy(out, s.data(), n); out[n] = 0;
    return n;
}
int shimta_transport_type(void *hv) { return (int)((Handle *)hv)->tp->type; }

}  // extern "C"
