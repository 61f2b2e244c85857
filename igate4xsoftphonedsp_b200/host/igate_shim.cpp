// igate_shim.cpp -- see igate_shim.h.  Host bookkeeping only: staging buffers,
// the per-adapter mirror of struct tp_adapter's control fields, and calls into
// libigate_dsp.so.  No codec / meter / header arithmetic lives here.
#include "igate_shim.h"

#include <string.h>
#include <sys/time.h>
#include <time.h>

#include <vector>

namespace {

struct Slot {               // one adapter = one entry of every per-channel array
    igd_bank *bank;
    int ch;
    pjsua_call_id callID;
    bool tx_staged, rx_staged;
    uint32_t rx_size;
    long long r2sPacket;
    bool rtpAudio;
    igd_ed137_fields last;  // last accepted receive-side fields
    uint32_t ed137_value;   // host order (= ntohl(adapter->ed137_value))
    uint8_t IncomingRTP, OutgoingRTP;
    char calltype[64], trxmode[64], callIndex[64];
};

}  // namespace

struct igd_bank {
    igd_ctx *ctx;
    int cap, n;
    std::vector<Slot *> slots;
    std::vector<igd_ed137_state> state;    // [cap] sender state (TransportAdapter.h:40-93)
    std::vector<igd_ed137_ctl> ctl;        // [cap] what the setters wrote since the last tick
    std::vector<uint8_t> rtp12, payload;   // [cap][12], [cap][160] staged TX
    std::vector<uint8_t> txpk, bm;         // [cap][180], [cap]
    std::vector<uint32_t> txsz;
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

static long long now_ms_wall()
{
    struct timeval tv;
    gettimeofday(&tv, nullptr);
    return (long long)tv.tv_sec * 1000 + tv.tv_usec / 1000;     // QDateTime::currentMSecsSinceEpoch()
}

igd_bank *igd_bank_open(int device, int max_channels)
{
    igd_ctx *ctx = nullptr;
    if (max_channels <= 0 || igd_init(device, &ctx) != IGD_OK) return nullptr;   // no GPU: no bank, no fallback
    igd_bank *b = new igd_bank();
    b->ctx = ctx;
    b->cap = max_channels;
    b->n = 0;
    b->state.resize(max_channels);
    b->ctl.resize(max_channels);
    b->rtp12.assign((size_t)max_channels * 12, 0);
    b->payload.assign((size_t)max_channels * IGD_FRAME, 0);
    b->txpk.assign((size_t)max_channels * IGD_PKT_MAX, 0);
    b->bm.assign(max_channels, 0);
    b->txsz.assign(max_channels, 0);
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

static Slot *slot_of(pjmedia_transport *tp) { return reinterpret_cast<Slot *>(tp); }

pj_status_t pjmedia_custom_tp_adapter_create(pjmedia_endpt *, const char *, pjmedia_transport *, pj_bool_t,
                                             pj_bool_t radiocall, pj_bool_t callIn, const char *calltype,
                                             pjsua_call_id callId, pjmedia_transport **p_tp, const char *callIndex,
                                             const char *trxmode, int keepAlivePeroid, pj_bool_t, pj_bool_t)
{
    igd_bank *b = g_default_bank;
    if (!b || !p_tp || b->n >= b->cap) return -1;
    Slot *s = new Slot();
    memset(s, 0, sizeof(*s));
    s->bank = b;
    s->ch = b->n++;
    s->callID = callId;
    strncpy(s->calltype, calltype ? calltype : "", sizeof(s->calltype) - 1);
    strncpy(s->callIndex, callIndex ? callIndex : "", sizeof(s->callIndex) - 1);
    strncpy(s->trxmode, trxmode ? trxmode : "", sizeof(s->trxmode) - 1);
    const long long now = now_ms_wall();
    s->r2sPacket = now;                                                    // TransportAdapter.cpp:122
    memset(&b->rxstate[s->ch], 0, sizeof(igd_rx_state));
    b->rxstate[s->ch].r2sPacket = now;
    igd_ed137_state_init(&b->state[s->ch], radiocall, callIn, s->calltype, keepAlivePeroid, now);
    memset(&b->ctl[s->ch], 0, sizeof(igd_ed137_ctl));
    b->slots.push_back(s);
    *p_tp = reinterpret_cast<pjmedia_transport *>(s);
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
// test hook: scenarios run on their own clock instead of the wall clock the constructor stamps
void igd_test_set_sendtime(pjmedia_transport *tp, long long t)
{
    if (Slot *s = slot_of(tp)) { s->bank->state[s->ch].r2sSendtime = t; s->r2sPacket = t; }
}
pj_uint32_t get_ed137_value(pjmedia_transport *tp) { return tp ? slot_of(tp)->ed137_value : 0; }
void igd_test_set_r2spacket(pjmedia_transport *tp, long long t)      // tests run on their own clock
{
    if (Slot *s = slot_of(tp)) { s->r2sPacket = t; s->bank->rxstate[s->ch].r2sPacket = t; }
}
long long getR2SStatus(pjmedia_transport *tp)
{
    return tp ? slot_of(tp)->r2sPacket : now_ms_wall() - 3000;             // TransportAdapter.cpp:317-325
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
    if (!s || !pkt || size < 12 || size > 12 + IGD_FRAME) return -1;
    igd_bank *b = s->bank;
    memcpy(&b->rtp12[(size_t)s->ch * 12], pkt, 12);
    memset(&b->payload[(size_t)s->ch * IGD_FRAME], 0, IGD_FRAME);
    memcpy(&b->payload[(size_t)s->ch * IGD_FRAME], static_cast<const uint8_t *>(pkt) + 12, size - 12);
    b->txmask[s->ch] = 1;
    return PJ_SUCCESS;
}

int igd_bank_flush_tx(igd_bank *b, long long now_ms, unsigned flags, igd_send_fn send, void *user)
{
    if (!b) return -1;
    const int n = b->n;
    if (n == 0) return 0;
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
    int sent = 0;
    for (int c = 0; c < n; c++) {
        b->state[c].radiostatus = saved[c];
        if (rc == IGD_OK && b->txmask[c] && b->txsz[c]) {
            const uint8_t *pk = &b->txpk[(size_t)c * IGD_PKT_MAX];
            if ((pk[1] & 0x7F) != 123) b->slots[c]->OutgoingRTP = b->bm[c];   // setOutgoingRTP
            if (send) send(user, reinterpret_cast<pjmedia_transport *>(b->slots[c]), pk, b->txsz[c]);
            sent++;
        }
        b->txmask[c] = 0;
    }
    return rc == IGD_OK ? sent : rc;
}

pj_status_t igd_submit_rx(pjmedia_transport *tp, const void *pkt, pj_ssize_t size)
{
    Slot *s = slot_of(tp);
    if (!s || !pkt || size < 0) return -1;
    igd_bank *b = s->bank;
    const size_t n = (size_t)size < (size_t)IGD_PKT_MAX ? (size_t)size : (size_t)IGD_PKT_MAX;
    uint8_t *dst = &b->rxpk[(size_t)s->ch * IGD_PKT_MAX];
    memset(dst, 0, IGD_PKT_MAX);
    memcpy(dst, pkt, n);
    s->rx_staged = true;
    s->rx_size = (uint32_t)size;
    return PJ_SUCCESS;
}

void igd_bank_set_event_cb(igd_bank *b, igd_event_fn fn, void *user)
{
    if (b) { b->on_event = fn; b->on_event_user = user; }
}

// one igd_rx_track call over every adapter of the bank: F = 1 tick, `present` = who received a
// packet, run_watchdog selects detectR2SPacketAndReconn's check (roip_ed137.cpp:1767-1780)
static int rx_track_tick(igd_bank *b, long long now_ms, int r2s_period_ms, bool run_watchdog)
{
    igd_rx_track_desc d;
    memset(&d, 0, sizeof(d));
    d.struct_size = sizeof(d);
    d.mem = IGD_MEM_HOST;
    d.F = 1; d.C = b->n;
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
    const int n = b->n;
    if (n == 0) return 0;
    for (int c = 0; c < n; c++) {
        b->rxsz[c] = b->slots[c]->rx_staged ? b->slots[c]->rx_size : 0;   // 0: nothing arrived
        b->rxpresent[c] = b->slots[c]->rx_staged ? 1 : 0;
    }
    int rc = igd_ed137_parse(b->ctx, b->rxpk.data(), b->rxsz.data(), (size_t)n, IGD_PKT_MAX, b->rxf.data(),
                             b->rxpay.data(), IGD_MEM_HOST);
    if (rc == IGD_OK)   // setIncomingRTP: byte-mean of the payload (roip_ed137.cpp:6541-6587), on the GPU
        rc = igd_bytemean(b->ctx, b->rxpay.data(), (size_t)n, IGD_FRAME, IGD_FRAME, 0, b->rxbm.data(), IGD_MEM_HOST);
    if (rc == IGD_OK)   // word latch, r2sPacket stamp, audio <-> keep-alive edge (TransportAdapter.cpp:252-315), on the GPU
        rc = rx_track_tick(b, now_ms, 200, false);
    int parsed = 0;
    for (int c = 0; c < n; c++) {
        Slot *s = b->slots[c];
        if (rc == IGD_OK && s->rx_staged) {
            const igd_ed137_fields &f = b->rxf[c];
            const igd_rx_event &e = b->rxev[c];
            parsed++;
            s->r2sPacket = b->rxstate[c].r2sPacket;                          // :289,302,311
            s->rtpAudio = b->rxstate[c].rtpAudio != 0;
            s->ed137_value = e.word;                                         // :252-256 latch
            if (f.accepted && !(f.flags & IGD_EDF_DROPPED)) s->last = f;
            if (e.flags & IGD_RXE_AUDIO) {                                   // :298-307
                if (f.payload_len == IGD_FRAME) s->IncomingRTP = b->rxbm[c];
                if (stream_cb)
                    stream_cb(user, reinterpret_cast<pjmedia_transport *>(s), &b->rxpk[(size_t)c * IGD_PKT_MAX], s->rx_size);
            }
            if ((e.flags & IGD_RXE_EDGE) && b->on_event)                     // :304-306, :312-314
                b->on_event(b->on_event_user, reinterpret_cast<pjmedia_transport *>(s), IGD_RXE_EDGE, e.word);
        }
        s->rx_staged = false;
    }
    return rc == IGD_OK ? parsed : rc;
}

int igd_bank_r2s_watchdog(igd_bank *b, long long now_ms, int r2s_period_ms)
{
    if (!b) return -1;
    if (b->n == 0) return 0;
    for (int c = 0; c < b->n; c++) b->rxpresent[c] = 0;
    const int rc = rx_track_tick(b, now_ms, r2s_period_ms, true);
    if (rc != IGD_OK) return rc;
    int hangups = 0;
    for (int c = 0; c < b->n; c++) {
        if (b->rxev[c].flags & IGD_RXE_HANGUP) {
            hangups++;
            if (b->on_event)
                b->on_event(b->on_event_user, reinterpret_cast<pjmedia_transport *>(b->slots[c]), IGD_RXE_HANGUP, b->rxev[c].word);
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
