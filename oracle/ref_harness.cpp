/*
 * ref_harness.cpp -- drives the REFERENCE'S OWN hot-path code, compiled where it lies:
 *
 *   /root/reference/TransportAdapter.cpp        (whole file, unmodified)
 *   oracle/_ref/gen/ref_extract.cpp             (line ranges of roip_ed137.cpp / Functions.cpp,
 *                                                unmodified; written by oracle/ref_extract.sh)
 *
 * under the stub Qt / PJSIP headers of oracle/ref_shim/.  The adapter is created by the reference's
 * factory and then driven ONLY the way PJSIP drives it -- through `tp->op->...` (attach, send_rtp,
 * encode_sdp, destroy ...) and through the RTP callback the adapter registers with its slave
 * transport -- plus the reference's public setters.  A fake slave transport captures what the adapter
 * sends; a fake stream captures what it forwards.  The clock behind QDateTime is igd_ref_clock_ms.
 *
 * Output: oracle/_ref/libigd_ref_ta.so (char unsigned, the aarch64 production targets) and
 * oracle/_ref/libigd_ref_ta_sc.so (char signed, x86 linux-g++) -- git-ignored, travel with gpurun.
 * TEST INFRASTRUCTURE ONLY: validates oracle/igd_oracle.c (tests/test_ref_pins.py) and generates the
 * committed goldens (tests/golden/make_golden.py).  The product never loads it.
 */
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include <roip_ed137.h>

extern "C" long long igd_ref_clock_ms = 0;

/* PJLIB / PJMEDIA stub bodies: oracle/ref_shim/igd_pj_stub_impl.cpp (shared with the vtable test of the shim) */
extern "C" {

/* PJSUA: conf slot == call id; the rx level last given to a slot is the leg's gain */
static std::map<int, float> g_rx_level;
static int g_adjust_calls = 0;
pj_status_t pjsua_call_get_info(pjsua_call_id call_id, pjsua_call_info *info)
{
    if (call_id < 0) return PJ_EINVAL;
    info->conf_slot = call_id;
    info->remote_info.ptr = (char *)"";
    info->remote_info.slen = 0;
    return PJ_SUCCESS;
}
pj_status_t pjsua_conf_adjust_rx_level(pjsua_conf_port_id slot, float level)
{
    g_rx_level[slot] = level;
    ++g_adjust_calls;
    return PJ_SUCCESS;
}
}  /* extern "C" */

/* ------------------------------------------------------------------ RoIP_ED137 harness: stubs */
static RoIP_ED137 *g_app = nullptr;
static RoIP_ED137::trx g_radios[4];
static RoIP_ED137::channel g_ch[2];
static std::vector<RoIP_ED137::trx *> g_incall_owned;

RoIP_ED137::RoIP_ED137()
{
    for (auto &r : g_radios) r = trx();
    g_ch[0] = channel(); g_ch[1] = channel();
    g_ch[0].radio1 = &g_radios[0]; g_ch[0].radio2 = &g_radios[1];
    g_ch[1].radio1 = &g_radios[2]; g_ch[1].radio2 = &g_radios[3];
    trx1 = &g_ch[0]; trx2 = &g_ch[1];
}
RoIP_ED137 *RoIP_ED137::instance()
{
    if (!g_app) g_app = new RoIP_ED137();
    return g_app;
}
void RoIP_ED137::scan_call_err() { ++igd_checkEvents_calls; }   /* first statement of checkEvents() */
void RoIP_ED137::updateHomeDisplay(QString, QString, QString, pjsua_call_id, int) {}
QString RoIP_ED137::getTimeDuratio(pjsua_call_id) { return QString(); }
void RoIP_ED137::recorder_pressed(int, QString) {}
void RoIP_ED137::recorder_released(int) {}
void RoIP_ED137::repeat_pressed(int, int) {}
void RoIP_ED137::repeat_released(int) {}
void RoIP_ED137::sqlTest_pressed(int, int) {}
void RoIP_ED137::sendTextMessage(QString m) { igd_last_text_message = m.toStdString(); ++igd_text_messages; }
void RoIP_ED137::cppCommand(QString m) { igd_last_text_message = m.toStdString(); ++igd_text_messages; }
bool RoIP_ED137::getExtFromRemoteInfo(pj_str_t const &, std::string &ext) { ext = "radio"; return true; }

/* ------------------------------------------------------------------ fake slave transport + stream */
struct Handle;
struct Slave {
    pjmedia_transport base;      /* must be first: the adapter sees a pjmedia_transport*            */
    Handle *h;
};
struct Handle {
    pjmedia_transport *tp = nullptr;        /* the reference adapter                               */
    Slave slave;
    void *cb_user = nullptr;                /* what the adapter registered with the slave          */
    void (*rtp_cb)(void *, void *, pj_ssize_t) = nullptr;
    void (*rtcp_cb)(void *, void *, pj_ssize_t) = nullptr;
    std::vector<uint8_t> sent;              /* last packet handed to the slave by send_rtp         */
    int sends = 0;
    int stream_rtp = 0, stream_rtcp = 0;    /* packets forwarded to the (fake) stream              */
    size_t stream_last_size = 0;
    int slave_calls[12] = {0};              /* per vtable entry: calls that reached the slave      */
    int call_id = -1;
    std::vector<uint8_t> rxbuf;             /* keeps the packet alive: the reference stores a
                                               pointer to it in a file-static (TransportAdapter.cpp:76) */
};

static pj_status_t s_get_info(pjmedia_transport *tp, pjmedia_transport_info *info)
{ ((Slave *)tp)->h->slave_calls[0]++; if (info) info->igd_stub_calls++; return PJ_SUCCESS; }
static pj_status_t s_attach(pjmedia_transport *tp, void *user_data, const pj_sockaddr_t *, const pj_sockaddr_t *,
                            unsigned, void (*rtp_cb)(void *, void *, pj_ssize_t),
                            void (*rtcp_cb)(void *, void *, pj_ssize_t))
{
    Handle *h = ((Slave *)tp)->h;
    h->slave_calls[1]++;
    h->cb_user = user_data; h->rtp_cb = rtp_cb; h->rtcp_cb = rtcp_cb;
    return PJ_SUCCESS;
}
static void s_detach(pjmedia_transport *tp, void *)
{
    Handle *h = ((Slave *)tp)->h;
    h->slave_calls[2]++;
    h->cb_user = nullptr; h->rtp_cb = nullptr; h->rtcp_cb = nullptr;
}
static pj_status_t s_send_rtp(pjmedia_transport *tp, const void *pkt, pj_size_t size)
{
    Handle *h = ((Slave *)tp)->h;
    h->slave_calls[3]++;
    h->sent.assign((const uint8_t *)pkt, (const uint8_t *)pkt + size);
    h->sends++;
    return PJ_SUCCESS;
}
static pj_status_t s_send_rtcp(pjmedia_transport *tp, const void *, pj_size_t)
{ ((Slave *)tp)->h->slave_calls[4]++; return PJ_SUCCESS; }
static pj_status_t s_send_rtcp2(pjmedia_transport *tp, const pj_sockaddr_t *, unsigned, const void *, pj_size_t)
{ ((Slave *)tp)->h->slave_calls[5]++; return PJ_SUCCESS; }
static pj_status_t s_media_create(pjmedia_transport *tp, pj_pool_t *, unsigned, const pjmedia_sdp_session *, unsigned)
{ ((Slave *)tp)->h->slave_calls[6]++; return PJ_SUCCESS; }
static pj_status_t s_encode_sdp(pjmedia_transport *tp, pj_pool_t *, pjmedia_sdp_session *, const pjmedia_sdp_session *, unsigned)
{ ((Slave *)tp)->h->slave_calls[7]++; return PJ_SUCCESS; }
static pj_status_t s_media_start(pjmedia_transport *tp, pj_pool_t *, const pjmedia_sdp_session *, const pjmedia_sdp_session *, unsigned)
{ ((Slave *)tp)->h->slave_calls[8]++; return PJ_SUCCESS; }
static pj_status_t s_media_stop(pjmedia_transport *tp) { ((Slave *)tp)->h->slave_calls[9]++; return PJ_SUCCESS; }
static pj_status_t s_simulate_lost(pjmedia_transport *tp, pjmedia_dir, unsigned)
{ ((Slave *)tp)->h->slave_calls[10]++; return PJ_SUCCESS; }
static pj_status_t s_destroy(pjmedia_transport *tp) { ((Slave *)tp)->h->slave_calls[11]++; return PJ_SUCCESS; }
static pjmedia_transport_op g_slave_op = {
    &s_get_info, &s_attach, &s_detach, &s_send_rtp, &s_send_rtcp, &s_send_rtcp2,
    &s_media_create, &s_encode_sdp, &s_media_start, &s_media_stop, &s_simulate_lost, &s_destroy};

static void stream_rtp_cb(void *user, void *, pj_ssize_t size)
{ Handle *h = (Handle *)user; h->stream_rtp++; h->stream_last_size = (size_t)size; }
static void stream_rtcp_cb(void *user, void *, pj_ssize_t) { ((Handle *)user)->stream_rtcp++; }

/* ------------------------------------------------------------------ exported C API */
extern "C" {

/* plain image of the fields of struct tp_adapter (TransportAdapter.h:40-93) the oracle models */
struct refta_state {
    int radiostatus, pttstatus, sqlstatus, callIn, callRecorder;
    int pttpriority, sqlpriority, ed137_bssi, pttid;
    int rxSlaveEnable, txSlaveEnable, rxSlaveEnableChanged, txSlaveEnableChanged;
    int trxSlaveEnableChangedCount;
    int firstR2SPacket, packetCnt;
    int keepAlivePeroid;
    int rtpFalse;
    int rtpAudio;
    long long r2sSendtime, r2sPacket;
    uint32_t ed137_value;
    uint32_t payloadsize;
    char calltype[64];
    uint8_t send_pkt_buff[256];
    uint8_t tmp_payload_buf[256];
    uint8_t payload_buff[256];
    uint64_t payload_bufSize, send_payload_bufSize;
};

int refta_char_is_signed(void) { return (char)0xFF < 0; }
int refta_sizeof_adapter(void) { return (int)sizeof(tp_adapter); }
int refta_offsetof_base(void) { return (int)offsetof(tp_adapter, base); }
void refta_set_clock(long long ms) { igd_ref_clock_ms = ms; }

void *refta_create(int radiocall, int callIn, const char *calltype, int call_id, const char *callIndex,
                   const char *trxmode, int keepAlivePeroid, int connToRadio, int pttWithPayload, int attach)
{
    Handle *h = new Handle();
    memset(&h->slave.base, 0, sizeof h->slave.base);
    snprintf(h->slave.base.name, sizeof h->slave.base.name, "fakeudp");
    h->slave.base.type = PJMEDIA_TRANSPORT_TYPE_UDP;
    h->slave.base.op = &g_slave_op;
    h->slave.h = h;
    h->call_id = call_id;
    pj_status_t st = pjmedia_custom_tp_adapter_create(nullptr, nullptr, &h->slave.base, PJ_TRUE, radiocall, callIn,
                                                      calltype, call_id, &h->tp, callIndex, trxmode,
                                                      keepAlivePeroid, connToRadio, pttWithPayload);
    if (st != PJ_SUCCESS || !h->tp || !h->tp->op) { delete h; return nullptr; }
    if (attach) {   /* what pjmedia_stream_create does with the transport it is given */
        st = pjmedia_transport_attach(h->tp, h, nullptr, nullptr, 0, &stream_rtp_cb, &stream_rtcp_cb);
        if (st != PJ_SUCCESS) { delete h; return nullptr; }
    }
    if (call_id >= 0) RoIP_ED137::instance()->transport_map[call_id] = h->tp;
    return h;
}

void refta_destroy(void *hv)
{
    Handle *h = (Handle *)hv;
    if (h->call_id >= 0) RoIP_ED137::instance()->transport_map.erase(h->call_id);
    pjmedia_transport_detach(h->tp, h);
    pjmedia_transport_close(h->tp);     /* tp->op->destroy: closes the slave (del_base) and frees the pool */
    delete h;
}

/* PJSIP's conference clock thread: pjmedia_transport_send_rtp(tp, pkt, size) == tp->op->send_rtp.
 * Returns the size handed to the slave transport (0 = nothing was sent), bytes in out[]. */
size_t refta_send_rtp(void *hv, const uint8_t *pkt, size_t size, uint8_t *out, int *status)
{
    Handle *h = (Handle *)hv;
    const tp_adapter *a = (const tp_adapter *)h->tp;
    if (!a->radiostatus) {
        /* TransportAdapter.cpp:641,870-874: for a non-radio call the function flows off its end
         * without a return statement (the pass-through is commented out) -- undefined behaviour
         * that an optimising g++ compiles into a fall-through.  Nothing is sent; not executed. */
        if (status) *status = -1;
        return 0;
    }
    int before = h->sends;
    pj_status_t st = pjmedia_transport_send_rtp(h->tp, pkt, size);
    if (status) *status = st;
    if (h->sends == before) return 0;
    memcpy(out, h->sent.data(), h->sent.size());
    return h->sent.size();
}

size_t refta_sendR2SStatus(void *hv, uint8_t *out)
{
    Handle *h = (Handle *)hv;
    int before = h->sends;
    sendR2SStatus(h->tp);
    if (h->sends == before) return 0;
    memcpy(out, h->sent.data(), h->sent.size());
    return h->sent.size();
}

/* PJSIP's ioqueue worker: the slave transport calls the callback the adapter registered in attach().
 * Returns 1 when the packet was forwarded to the stream, 0 otherwise, -2 when nothing is attached. */
int refta_rx(void *hv, const uint8_t *pkt, size_t size, size_t bufsize)
{
    Handle *h = (Handle *)hv;
    if (!h->rtp_cb) return -2;
    h->rxbuf.assign(pkt, pkt + (bufsize > size ? bufsize : size));
    h->rxbuf.resize(h->rxbuf.size() + 1100, 0);   /* the reference may copy up to 1023 B past the header */
    int before = h->stream_rtp;
    (*h->rtp_cb)(h->cb_user, h->rxbuf.data(), (pj_ssize_t)size);
    return h->stream_rtp != before;
}

int refta_rtcp(void *hv, const uint8_t *pkt, size_t size)
{
    Handle *h = (Handle *)hv;
    if (!h->rtcp_cb) return -2;
    std::vector<uint8_t> b(pkt, pkt + size);
    int before = h->stream_rtcp;
    (*h->rtcp_cb)(h->cb_user, b.data(), (pj_ssize_t)size);
    return h->stream_rtcp != before;
}

void refta_setAdapterPtt(void *hv, int pttval, int priority, int userRec) { setAdapterPtt(((Handle *)hv)->tp, pttval != 0, priority, userRec); }
void refta_setTxRxSlaveEnable(void *hv, int rx, int tx) { setTxRxSlaveEnable(((Handle *)hv)->tp, rx, tx); }
void refta_setAdapterQslOn(void *hv, int sqlval, int priority, uint32_t bssi) { setAdapterQslOn(((Handle *)hv)->tp, sqlval != 0, priority, bssi); }
void refta_setAdapterPttId(void *hv, int pttid) { setAdapterPttId(((Handle *)hv)->tp, pttid); }
void refta_setcallRecorder(void *hv, int val) { setcallRecorder(((Handle *)hv)->tp, val != 0); }
void refta_setCallType(void *hv, const char *calltype) { setCallType(((Handle *)hv)->tp, calltype); }
void refta_setAdapterRadioModeAndType(void *hv, const char *type, const char *mode) { setAdapterRadioModeAndType(((Handle *)hv)->tp, type, mode); }
uint32_t refta_get_ed137_value(void *hv) { return get_ed137_value(hv ? ((Handle *)hv)->tp : nullptr); }
long long refta_getR2SStatus(void *hv) { return getR2SStatus(hv ? ((Handle *)hv)->tp : nullptr); }

void refta_get_state(void *hv, refta_state *s)
{
    const tp_adapter *a = (const tp_adapter *)((Handle *)hv)->tp;
    memset(s, 0, sizeof *s);
    s->radiostatus = a->radiostatus; s->pttstatus = a->pttstatus; s->sqlstatus = a->sqlstatus;
    s->callIn = a->callIn; s->callRecorder = a->callRecorder;
    s->pttpriority = a->pttpriority; s->sqlpriority = a->sqlpriority; s->ed137_bssi = a->ed137_bssi; s->pttid = a->pttid;
    s->rxSlaveEnable = a->rxSlaveEnable; s->txSlaveEnable = a->txSlaveEnable;
    s->rxSlaveEnableChanged = a->rxSlaveEnableChanged; s->txSlaveEnableChanged = a->txSlaveEnableChanged;
    s->trxSlaveEnableChangedCount = a->trxSlaveEnableChangedCount;
    s->firstR2SPacket = a->firstR2SPacket; s->packetCnt = a->packetCnt;
    s->keepAlivePeroid = a->keepAlivePeroid; s->rtpFalse = a->rtpFalse; s->rtpAudio = a->rtpAudio;
    s->r2sSendtime = a->r2sSendtime; s->r2sPacket = a->r2sPacket;
    s->ed137_value = a->ed137_value; s->payloadsize = a->payloadsize;
    memcpy(s->calltype, a->calltype, 64);
    memcpy(s->send_pkt_buff, a->send_pkt_buff, 256);
    memcpy(s->tmp_payload_buf, a->tmp_payload_buf, 256);
    memcpy(s->payload_buff, a->payload_buff, 256);
    s->payload_bufSize = a->payload_bufSize; s->send_payload_bufSize = a->send_payload_bufSize;
}

/* test set-up pokes: state a previous (not replayed) history would have left behind */
void refta_poke_send_hdr(void *hv, const uint8_t *hdr20) { memcpy(((tp_adapter *)((Handle *)hv)->tp)->send_pkt_buff, hdr20, 20); }
void refta_poke_word(void *hv, uint32_t host_word) { ((tp_adapter *)((Handle *)hv)->tp)->ed137_value = htonl(host_word); }
void refta_poke_rx(void *hv, long long r2sPacket, int rtpAudio, uint32_t host_word, uint32_t payloadsize)
{
    tp_adapter *a = (tp_adapter *)((Handle *)hv)->tp;
    a->r2sPacket = r2sPacket; a->rtpAudio = rtpAudio; a->ed137_value = htonl(host_word); a->payloadsize = payloadsize;
}

void refta_counters(void *hv, int *out /* [16] */)
{
    Handle *h = (Handle *)hv;
    out[0] = h->sends; out[1] = h->stream_rtp; out[2] = h->stream_rtcp; out[3] = (int)h->stream_last_size;
    for (int i = 0; i < 12; ++i) out[4 + i] = h->slave_calls[i];
}

/* drive every pass-through entry of the vtable once, the way PJSIP would; returns a bit per entry
 * whose call reached the slave transport */
int refta_vtable_passthrough(void *hv)
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

/* tp->op->encode_sdp on an empty one-media SDP: "name:value\n" per attribute the adapter added */
size_t refta_encode_sdp(void *hv, char *out, size_t cap)
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

/* ---------------------------------------------------------------- application side (RoIP_ED137) */
void refapp_reset(int inviteMode, int rxBestSignalEnable)
{
    if (g_app) { g_app->transport_map.clear(); delete g_app; g_app = nullptr; }
    for (auto *t : g_incall_owned) delete t;
    g_incall_owned.clear();
    g_rx_level.clear(); g_adjust_calls = 0;
    RoIP_ED137 *app = RoIP_ED137::instance();
    app->inviteMode = inviteMode;
    app->rxBestSignalEnable = rxBestSignalEnable != 0;
    /* Control-plane side conditions of setvolume(UNMUTE) (Functions.cpp:1679-1691), held the way they stand
     * while a radio is receiving: the gateway is not the radio end (connToRadio false), no local PTT input,
     * no sidetone loop-back, and the display state that updateHomeDisplay (roip_ed137.cpp:825-950, control
     * plane, stubbed) maintains reads RXON.  With these, UNMUTE reaches setSlotVolume with 2.0f. */
    app->connToRadio = false;
    app->pttInput = false;
    app->localSidetoneLoopbackOn = false;
    app->trx1->trxStatus = RXON;
}
/* SERVER: radio slot 0..3 = trx1->radio1, trx1->radio2, trx2->radio1, trx2->radio2 (roip_ed137.cpp:130-139) */
void refapp_server_bind(int slot, int call_id, int callState, const char *trxmode)
{
    RoIP_ED137::instance();
    g_radios[slot].call_id = call_id;
    g_radios[slot].callState = callState != 0;
    g_radios[slot].trxmode = trxmode;
}
/* CLIENT: one inbound call slot (roip_ed137.cpp:141-150) */
int refapp_client_add(int call_id, int callState, const char *trxmode, const char *callName)
{
    RoIP_ED137 *app = RoIP_ED137::instance();
    RoIP_ED137::trx *t = new RoIP_ED137::trx();
    t->call_id = call_id; t->callState = callState != 0; t->trxmode = trxmode; t->callName = callName;
    g_incall_owned.push_back(t);
    app->trx_incall.append(t);
    return app->trx_incall.length() - 1;
}
void refapp_checkEvents(void) { RoIP_ED137::instance()->checkEvents(); }
int refapp_checkEvents_calls(void) { return RoIP_ED137::instance()->igd_checkEvents_calls; }
/* gain last given to the leg through pjsua_conf_adjust_rx_level; -1 when never set */
float refapp_rx_level(int call_id) { auto it = g_rx_level.find(call_id); return it == g_rx_level.end() ? -1.0f : it->second; }
int refapp_adjust_calls(void) { return g_adjust_calls; }

struct refapp_leg { int lastRx, lastTx, lastRxmsec, lastTxmsec, audioSQLOn, rssi, m_PttPressed, pttLevel, SQLOn, IncomingRTP, OutgoingRTP; };
static RoIP_ED137::trx *leg_of(int server_slot_or_client_index)
{
    RoIP_ED137 *app = RoIP_ED137::instance();
    if (app->inviteMode == SERVER) return &g_radios[server_slot_or_client_index];
    return app->trx_incall.at(server_slot_or_client_index);
}
void refapp_get_leg(int idx, refapp_leg *o)
{
    const RoIP_ED137::trx *t = leg_of(idx);
    o->lastRx = t->lastRx; o->lastTx = t->lastTx; o->lastRxmsec = (int)t->lastRxmsec; o->lastTxmsec = (int)t->lastTxmsec;
    o->audioSQLOn = t->audioSQLOn; o->rssi = t->rssi; o->m_PttPressed = t->m_PttPressed; o->pttLevel = t->pttLevel;
    o->SQLOn = t->SQLOn; o->IncomingRTP = t->IncomingRTP; o->OutgoingRTP = t->OutgoingRTP;
}
void refapp_set_leg(int idx, const refapp_leg *o)
{
    RoIP_ED137::trx *t = leg_of(idx);
    t->lastRx = o->lastRx; t->lastTx = o->lastTx; t->lastRxmsec = o->lastRxmsec; t->lastTxmsec = o->lastTxmsec;
    t->audioSQLOn = o->audioSQLOn != 0; t->rssi = o->rssi; t->m_PttPressed = o->m_PttPressed != 0; t->pttLevel = o->pttLevel;
}
void refapp_get_bridge(int *o /* ptt_level, sqlStatusCount, sqlStatusOn */)
{
    RoIP_ED137 *app = RoIP_ED137::instance();
    o[0] = app->ptt_level; o[1] = app->sqlStatusCount; o[2] = app->sqlStatusOn;
}
void refapp_set_bridge(int ptt_level, int sqlStatusCount, int sqlStatusOn)
{
    RoIP_ED137 *app = RoIP_ED137::instance();
    app->ptt_level = ptt_level; app->sqlStatusCount = sqlStatusCount; app->sqlStatusOn = sqlStatusOn != 0;
}
/* field getters by call id (Functions.cpp:1001-1179): bss, ptt_type, ptt_id, squelch, active */
void refapp_fields(int call_id, int *o)
{
    RoIP_ED137 *app = RoIP_ED137::instance();
    o[0] = app->get_IPRadioBss(call_id);
    o[1] = app->get_IPRadioPttStatus(call_id);
    o[2] = app->get_IPRadioPttId(call_id);
    o[3] = app->get_IPRadioSquelch(call_id);
    o[4] = app->get_IPRadioStatus(call_id);
}
/* the by-call-id setters (Functions.cpp:909-999) */
void refapp_setRadioPttbyCallID(int pttval, int call_id, int priority, int userRec) { RoIP_ED137::instance()->setRadioPttbyCallID(pttval != 0, call_id, priority, userRec); }
void refapp_setRadioSqlOnbyCallID(int sqlval, int call_id, int priority, int rssi)
{
    if (rssi < 0) RoIP_ED137::instance()->setRadioSqlOnbyCallID(sqlval != 0, call_id, priority);
    else RoIP_ED137::instance()->setRadioSqlOnbyCallID(sqlval != 0, call_id, priority, (uint16_t)rssi);
}
void refapp_setSlaveEnable(int call_id, int rx, int tx) { RoIP_ED137::instance()->setSlaveEnable(call_id, rx, tx); }
long long refapp_get_R2SStatus(int call_id) { return RoIP_ED137::instance()->get_R2SStatus(call_id); }

/* event logger (Functions.cpp:2126-2230) on SERVER radio slot `slot` */
void refapp_keeplog(int slot, double audioInLevel, int outgoingRTP)
{
    RoIP_ED137 *app = RoIP_ED137::instance();
    app->audioInLevel = audioInLevel;
    g_radios[slot].OutgoingRTP = (uint8_t)outgoingRTP;
    app->keeplogAudioLevel(&g_radios[slot]);
}
size_t refapp_ptt_event(int slot, const char *strEvent, double audioInLevel, const char *url, int softPhoneID,
                        char *out, size_t cap)
{
    RoIP_ED137 *app = RoIP_ED137::instance();
    app->audioInLevel = audioInLevel;
    app->m_softPhoneID = softPhoneID;
    g_radios[slot].url = url;
    int before = app->igd_text_messages;
    app->createPTTEventDataLogger(&g_radios[slot], QString(strEvent));
    if (app->igd_text_messages == before) { if (cap) out[0] = 0; return 0; }
    size_t n = app->igd_last_text_message.size() < cap - 1 ? app->igd_last_text_message.size() : cap - 1;
    memcpy(out, app->igd_last_text_message.data(), n); out[n] = 0;
    return n;
}

}  /* extern "C" */
