// igate_shim.h -- thin host shim that keeps the reference's call signatures
// (TransportAdapter.h:18-38, RoIP_ED137 getters Functions.cpp:1001-1179,
// WavWriter.h:44-53, audiometer.cpp:30-31) and routes the per-packet work of a
// whole 20 ms tick through ONE batched call of libigate_dsp.so (include/igate_dsp.h).
//
// The reference does the ED-137 work inside PJSIP callbacks, one packet at a
// time per call (transport_rtp_cb / transport_send_rtp).  Here every adapter is
// a slot of an igd_bank: the callbacks only STAGE their packet, and the owner
// of the tick clock (PJSIP's conference clock thread in the reference,
// roip_ed137.cpp:3031-3036) calls igd_bank_flush_tx / igd_bank_flush_rx once
// per tick.  All arithmetic runs on the GPU; this file holds no codec, meter or
// header math.
//
// The adapter is a REAL `pjmedia_transport`: its struct starts with `pjmedia_transport base` whose
// `op` points at a 12-entry `pjmedia_transport_op` (the table of TransportAdapter.cpp:59-73), so PJSIP
// drives it exactly as it drives the reference's: `tp->op->attach` registers the shim's RTP callback
// with the slave (UDP) transport, `tp->op->send_rtp` stages the conference bridge's packet, the other
// entries pass through to the slave transport, `tp->op->destroy` closes it.  PJ types come from the
// real pjproject headers when IGD_HAVE_PJSIP is defined, else from igate_pj_compat.h.
#pragma once
#include <stddef.h>
#include <stdint.h>

#include <string>

#include "../../include/igate_dsp.h"
#include "igate_pj_compat.h"

// wire image of the reference's struct custom_rtp_hdr (ed137_rtp.h:22-47)
#pragma pack(push, 1)
struct custom_rtp_hdr {
    uint8_t vpxcc;          // V:2 P:1 X:1 CC:4
    uint8_t mpt;            // M:1 PT:7
    uint16_t seq;           // network order
    uint32_t ts;            // network order
    uint32_t ssrc;          // network order
    uint16_t profile_data;  // network order (0x0167)
    uint16_t length;        // network order (1)
    uint32_t ed137;         // network order
};
#pragma pack(pop)
static_assert(sizeof(custom_rtp_hdr) == 20, "ed137_rtp.h:22-47");

struct igd_bank;

// ---- bank = all adapters of one process + the GPU context
igd_bank *igd_bank_open(int device, int max_channels);
void igd_bank_close(igd_bank *bank);
void igd_bank_set_default(igd_bank *bank);    // the bank pjmedia_custom_tp_adapter_create allocates from
// Clock behind the reference's QDateTime::currentMSecsSinceEpoch() calls (adapter creation stamps
// r2sPacket / r2sSendtime, TransportAdapter.cpp:122-123; getR2SStatus(NULL), :324; sendR2SStatus, :426).
// Default: the wall clock.  A gateway that runs on a media clock (or a replay) installs its own.
typedef long long (*igd_clock_fn)(void *user);
void igd_bank_set_clock(igd_bank *bank, igd_clock_fn fn, void *user);
igd_ctx *igd_bank_ctx(igd_bank *bank);

// ---- reference signatures (TransportAdapter.h:18-38), same names and argument meaning
pj_status_t pjmedia_custom_tp_adapter_create(pjmedia_endpt *endpt, const char *name, pjmedia_transport *transport,
                                             pj_bool_t del_base, pj_bool_t radiocall, pj_bool_t callIn,
                                             const char *calltype, pjsua_call_id callId, pjmedia_transport **p_tp,
                                             const char *callIndex, const char *trxmode, int keepAlivePeroid,
                                             pj_bool_t connToRadio, pj_bool_t pttWithPayload);
pj_status_t decodeRtp(void *pkt, custom_rtp_hdr **hdr);
pj_uint32_t get_ed137_value(pjmedia_transport *tp);
pj_status_t setAdapterPtt(pjmedia_transport *tp, bool pttval, int priority, int userRec);
pj_status_t setTxRxSlaveEnable(pjmedia_transport *tp, pj_bool_t rx, pj_bool_t tx);
pj_status_t setAdapterRadioModeAndType(pjmedia_transport *tp, char const *type, char const *txrxmode);
pj_status_t setAdapterQslOn(pjmedia_transport *tp, bool sqlval, int priority, pj_uint32_t bssi);
pj_status_t setAdapterPttId(pjmedia_transport *tp, int pttid);
pj_status_t setcallRecorder(pjmedia_transport *tp, bool val);
pj_status_t setCallType(pjmedia_transport *tp, char const *calltype);
long long getR2SStatus(pjmedia_transport *tp);
// TransportAdapter.h:38 (body TransportAdapter.cpp:422-633): the timer-driven keep-alive of ONE adapter, now:
// one igd_ed137_keepalive launch for its channel; a 20-byte PT-123 packet leaves through the slave transport
// when the reference's would.  igd_bank_keepalive does the same for every adapter of the bank in one launch.
void sendR2SStatus(pjmedia_transport *tp);
int igd_bank_keepalive(igd_bank *bank, long long now_ms);

// ---- RoIP_ED137 field getters (Functions.cpp:1001-1179), per adapter instead of per call id
int get_IPRadioBss(pjmedia_transport *tp);
int get_IPRadioPttStatus(pjmedia_transport *tp);
int get_IPRadioPttId(pjmedia_transport *tp);
int get_IPRadioSquelch(pjmedia_transport *tp);
bool get_IPRadioStatus(pjmedia_transport *tp);
uint8_t get_IncomingRTP(pjmedia_transport *tp);   // trx->IncomingRTP, roip_ed137.cpp:6541-6587
uint8_t get_OutgoingRTP(pjmedia_transport *tp);   // trx->OutgoingRTP, roip_ed137.cpp:6500-6536

// ---- per-tick staging (what the PJSIP callbacks become)
// transport_send_rtp(tp, pkt, size): pkt = 12-byte RTP header + G.711 payload
pj_status_t igd_submit_tx(pjmedia_transport *tp, const void *pkt, pj_size_t size);
// transport_rtp_cb(user_data, pkt, size): a received packet (20-byte header + payload)
pj_status_t igd_submit_rx(pjmedia_transport *tp, const void *pkt, pj_ssize_t size);
typedef void (*igd_send_fn)(void *user, pjmedia_transport *tp, const void *pkt, pj_size_t size);
// runs the staged TX packets of this tick through igd_ed137_pack and hands every packet that the
// reference would have given to pjmedia_transport_send_rtp() to `send`; send == NULL: to the adapter's
// slave transport, pjmedia_transport_send_rtp(slave_tp, pkt, size) as TransportAdapter.cpp:848 does.
// Returns packets sent or <0
int igd_bank_flush_tx(igd_bank *bank, long long now_ms, unsigned flags, igd_send_fn send, void *user);
// runs the staged RX packets through igd_ed137_parse (+ the byte-mean meter); afterwards the getters
// above return the new values; `stream_cb` receives the audio packets (PT != 123) like stream_rtp_cb
// does in the reference; stream_cb == NULL: the callback the stream registered through tp->op->attach
// (adapter->stream_rtp_cb(adapter->stream_user_data, pkt, size), TransportAdapter.cpp:301).
// Returns packets parsed or <0
int igd_bank_flush_rx(igd_bank *bank, long long now_ms, igd_send_fn stream_cb, void *user);

// events the receive side raises (what transport_rtp_cb / the 40 ms watchdog did inline):
//   IGD_RXE_EDGE   -> RoIP_ED137::setIncomingED137Value(word, callID) -> checkEvents()
//                     (TransportAdapter.cpp:304-306, 312-314)
//   IGD_RXE_HANGUP -> trx_call_hangup(call_id, 500, "WG-67 ;cause=2001; text=\"missing R2S KeepAlive\"")
//                     (roip_ed137.cpp:1770-1774)
typedef void (*igd_event_fn)(void *user, pjmedia_transport *tp, unsigned rxe_flag, pj_uint32_t ed137_word);
void igd_bank_set_event_cb(igd_bank *bank, igd_event_fn fn, void *user);
// RoIP_ED137::detectR2SPacketAndReconn's liveness check for every adapter (one igd_rx_track call):
// call it from the 40 ms timer (roip_ed137.cpp:482).  Returns the number of hang-up requests or <0.
int igd_bank_r2s_watchdog(igd_bank *bank, long long now_ms, int r2s_period_ms);

// ---- WavWriter (WavWriter.h:44-53), same public methods; the file image is built on the GPU at stop()
class WavWriter {
public:
    WavWriter();
    ~WavWriter();
    void writeRTPWav(const char *pktbuf, const char *payloadbuf, unsigned int pktlen, unsigned int payloadlen);
    void start(std::string prefix, int rate);
    void stop();
    void wav_write(unsigned char *buf, unsigned int len);
    bool isRunning();
    // additions: which bank/GPU builds the image, the reference's byte-exact header quirks on/off
    void attach(igd_bank *bank, bool ref_quirks, int law);
    const std::string &fileName() const { return m_file; }

private:
    igd_bank *m_bank;
    bool m_running, m_quirks;
    int m_rate, m_law;
    std::string m_file, m_data;
};

// ---- AudioMeter scale (audiometer.cpp:30-31) for a batch of raw levels
int igd_audio_level_percent(igd_bank *bank, const int32_t *raw, size_t n, int32_t *percent);
