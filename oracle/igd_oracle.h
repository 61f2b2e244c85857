/*
 * igd_oracle.h -- CPU ORACLE for the iGate4x per-frame voice path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs may build, link, load or execute anything under
 * oracle/.  The product (igate4xsoftphonedsp_b200/, include/) never does.
 *
 * What it is: a plain-C, scalar restatement of what the reference
 * (piyanon108/iGate4xSoftphoneDSP, paths below are relative to that repo)
 * does per packet / per 20 ms frame, so that the CUDA path can be checked
 * bit-for-bit on the same inputs.
 *
 * Parity pinning status:
 *   - PINNED to the reference's own code compiled from /root/reference by oracle/Makefile into
 *     oracle/_ref/ (TransportAdapter.cpp whole, ed137_rtp.h, WavWriter.cpp, and line-range extracts
 *     of roip_ed137.cpp / Functions.cpp, under the stand-in Qt/PJSIP headers of oracle/ref_shim/):
 *     the ED-137 header layout; the sender state machine transport_send_rtp (:635-874) with
 *     setOutgoingRTP; sendR2SStatus (:422-633); the receive callback transport_rtp_cb (:240-316) with
 *     setIncomingRTP, get_ed137_value, getR2SStatus; the field getters and by-call-id setters
 *     (Functions.cpp:909-1179); checkEvents() gate arbitration (roip_ed137.cpp:5609-6348) with
 *     setvolume / setSlotVolume; keeplogAudioLevel / createPTTEventDataLogger (Functions.cpp:2126-2230);
 *     the WavWriter bytes.  tests/test_ref_pins.py compares oracle and reference on every shared
 *     seeded case; tests/golden/ref_pins.json holds the digests of the REFERENCE run.
 *   - restatement only: the percent scale (audiometer.cpp:30-31, one formula).
 *   - G.711 and the conference-bridge mix : PARITY UNPINNED.  The arithmetic lives in pjproject
 *     (pjmedia/src/pjmedia/alaw_ulaw.c, g711.c, conference.c), which the reference links but does not
 *     vendor and does not version-pin (iGate4xSoftphoneDSP.pro:66-82; "pjsip 2.6 only",
 *     TransportAdapter.cpp:245).  The oracle restates the published ITU-T G.711 / Sun g711.c
 *     algorithm and the pjmedia rx_adj_level + saturating-sum semantics
 *     (Appendix D); it is pinned by the SHA-256 table hashes of Appendix B, by Python's independent
 *     `audioop` decoder (all 256 codes) and encoder (all non-negative inputs), and by the ITU
 *     known-answer codes.
 */
#ifndef IGD_ORACLE_H
#define IGD_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_FRAME 160           /* roip_ed137.h:112-115: 8 kHz, PTIME 20 ms   */
#define ORC_LAW_ALAW 0          /* RTP PT 8, roip_ed137.cpp:3563              */
#define ORC_LAW_ULAW 1          /* RTP PT 0, roip_ed137.cpp:3564              */

/* ---------------- G.711 (pjmedia alaw_ulaw.c) -------- */
int16_t orc_alaw2lin(uint8_t a);
int16_t orc_ulaw2lin(uint8_t u);
uint8_t orc_lin2alaw(int pcm);
uint8_t orc_lin2ulaw(int pcm);
void orc_g711_decode(const uint8_t *codes, int16_t *pcm, size_t n, int law);
void orc_g711_encode(const int16_t *pcm, uint8_t *codes, size_t n, int law);

/* ---------------- meters ----------------------------- */
/* roip_ed137.cpp:6557-6568 (incoming) / :6511-6517 (outgoing): truncated mean
 * of payload bytes.  `signed_char` selects the x86 `char` signedness (quirk
 * Q4); 0 = the aarch64 production target.  n==0 returns 0 (the reference
 * would divide by zero). */
uint8_t orc_bytemean(const uint8_t *payload, int n, int signed_char);
/* Appendix C.2: exact sum of squares and max|x| of one frame of PCM. */
void orc_frame_power(const int16_t *pcm, int n, uint64_t *sumsq, uint32_t *peak);
double orc_rms_dbfs(uint64_t sumsq, int n);     /* 10log10(S/n)-20log10(32768) */
double orc_peak_dbfs(uint32_t peak);            /* 20log10(P/32768)            */
/* audiometer.cpp:30-31 */
int orc_percent(int v);

typedef struct {                /* 16 B, same layout as igd_meter_rec        */
    uint32_t sumsq_lo;
    uint32_t hi;                /* [7:0] sumsq_hi [15:8] bytemean [31:16] peak */
    float rms_dbfs;
    float peak_dbfs;
} orc_meter_rec;

typedef struct {                /* 4 B, same layout as igd_bridge_rec        */
    uint8_t bytemean_out;
    uint8_t n_open;
    uint16_t mix_peak;
} orc_bridge_rec;

/* ---------------- gain + mix -------------------------- */
/* pjsua_conf_adjust_rx_level(level) as called at roip_ed137.cpp:5221:
 * adj = (int)((level-1)*128) + 128  (2.0->256, 0.0->0, 0.1f->13, 0.5->64).  */
int orc_gain_adj(float level);
/* y = clamp16((x*adj)>>7)  (pjmedia conference.c rx_adj_level semantics)    */
int16_t orc_apply_gain(int16_t x, int adj);
/* mix[i] = clamp16( sum_{legs, adj!=0} clamp16((x*adj)>>7) ), legs given as
 * `nlegs` pointers to 160-sample frames.                                   */
void orc_mix_frame(const int16_t *const *legs, const uint16_t *adj, int nlegs,
                   int n, int16_t *mix);

/* ---------------- fused per-frame voice path ------------------------------ */
/* gain_q7 bit 15: no audio frame arrived for the leg on this tick (TransportAdapter.cpp:298-315: keep-alives
 * and dropped packets never reach the stream) -> silent leg-frame, see include/igate_dsp.h IGD_GAIN_NO_AUDIO */
#define ORC_GAIN_NO_AUDIO 0x8000u
typedef struct {
    int F, B, G;                     /* frames, bridges, legs per bridge       */
    const uint8_t *codes;            /* [F][B*G][160]                          */
    const uint8_t *law;              /* [B*G]                                  */
    const uint16_t *gain_q7;         /* [F][B*G]; 0 = gate shut                */
    const uint8_t *out_law;          /* [B]                                    */
    int16_t *mix;                    /* [F][B][160]                            */
    uint8_t *enc;                    /* [F][B][160]                            */
    orc_meter_rec *meter;            /* [F][B*G]                               */
    orc_bridge_rec *bmeter;          /* [F][B]                                 */
    int signed_char;                 /* quirk Q4 for the byte-mean             */
} orc_batch;
/* scalar, frame by frame, channel by channel -- like the reference.          */
void orc_process_batch(const orc_batch *b);
/* same, bridges [b0,b1) only (used by the threaded CPU baseline).            */
void orc_process_batch_range(const orc_batch *b, int b0, int b1);
/* threaded CPU baseline: bridges sharded contiguously over nthreads.          */
void orc_process_batch_mt(const orc_batch *b, int nthreads);

/* ---------------- event summary (Functions.cpp:2126-2230) ----------------- */
typedef struct {                /* 32 B, same layout as igd_summary_rec      */
    uint32_t count;             /* level_in_count                             */
    uint16_t bm_sum;            /* OutgoingRTPSum: uint16_t, wraps (h:742)    */
    uint8_t bm_max;             /* init 0   (Functions.cpp:2166)              */
    uint8_t bm_min;             /* init 255 (Functions.cpp:2167)              */
    uint64_t sum_s;             /* sum over open frames of the frame sumsq    */
    uint64_t max_s;             /* init 0            (Functions.cpp:2162)     */
    uint64_t min_s;             /* init 255*160: reference inits min to 255   */
} orc_summary_rec;
void orc_event_summary(const orc_meter_rec *meter, const uint16_t *gain_q7,
                       int F, int C, orc_summary_rec *out);
/* Functions.cpp:2196-2200: 10*log10(av), 10*log10(max), 10*log10(min), av    */
void orc_summary_db(const orc_summary_rec *s, double *av_db, double *max_db,
                    double *min_db, int *bm_av);

/* ---------------- ED-137 RTP header extension --------- */
typedef struct {
    /* names follow struct tp_adapter (TransportAdapter.h:40-93) */
    int radiostatus, pttstatus, sqlstatus, callIn, callRecorder;
    int pttpriority, sqlpriority, ed137_bssi, pttid;
    int rxSlaveEnable, txSlaveEnable, rxSlaveEnableChanged, txSlaveEnableChanged;
    int trxSlaveEnableChangedCount;
    int firstR2SPacket, packetCnt;
    int keepAlivePeroid;
    int rtpFalse;
    int rtpAudio;
    long long r2sSendtime, r2sPacket;
    uint32_t ed137_value;           /* as received: network byte order        */
    uint32_t payloadsize;           /* hdr->length, un-swapped (:255)         */
    char calltype[64];
    uint8_t send_pkt_buff[256];
    uint8_t tmp_payload_buf[256];
    uint8_t payload_buff[256];
    size_t payload_bufSize, send_payload_bufSize;
    uint8_t IncomingRTP, OutgoingRTP;   /* trx->IncomingRTP / OutgoingRTP     */
    int checkEvents_calls;              /* setIncomingED137Value() count      */
} orc_adapter;

/* pjmedia_custom_tp_adapter_create (TransportAdapter.cpp:82-133)             */
void orc_adapter_init(orc_adapter *a, int radiocall, int callIn,
                      const char *calltype, int keepAlivePeroid, long long now);
/* setters, TransportAdapter.cpp:135-223                                      */
void orc_setAdapterPtt(orc_adapter *a, int pttval, int priority, int userRec);
void orc_setTxRxSlaveEnable(orc_adapter *a, int rx, int tx);
void orc_setAdapterQslOn(orc_adapter *a, int sqlval, int priority, uint32_t bssi);
void orc_setAdapterPttId(orc_adapter *a, int pttid);
void orc_setcallRecorder(orc_adapter *a, int val);
void orc_setCallType(orc_adapter *a, const char *calltype);

/* transport_send_rtp (TransportAdapter.cpp:635-874).  `pkt` = 12-byte RTP
 * header + payload as built by the PJSIP stream.  Returns the number of bytes
 * handed to the slave transport (20 or 20+payloadlen), copied into `out`, or
 * 0 when the keep-alive throttle suppressed the packet.  `inviteServer`!=0
 * enables the SERVER-mode byte-mean (roip_ed137.cpp:6509).                  */
size_t orc_transport_send_rtp(orc_adapter *a, const uint8_t *pkt, size_t size,
                              long long now, uint8_t *out, int inviteServer,
                              int signed_char);
/* sendR2SStatus (TransportAdapter.cpp:422-633), same return convention.      */
size_t orc_sendR2SStatus(orc_adapter *a, long long now, uint8_t *out);
/* transport_rtp_cb (TransportAdapter.cpp:240-316).  Returns 1 when the packet
 * was forwarded to the stream (pt != 123), 0 for a keep-alive, -1 dropped.   */
int orc_transport_rtp_cb(orc_adapter *a, const uint8_t *pkt, size_t size,
                         long long now, int inviteServer, int signed_char);
/* get_ed137_value (TransportAdapter.cpp:337-346)                             */
uint32_t orc_get_ed137_value(const orc_adapter *a);

typedef struct {
    uint32_t word;      /* host order                                         */
    int ptt_type;       /* Functions.cpp:1136-1138                            */
    int ptt_id;         /* Functions.cpp:1148-1150                            */
    int squelch;        /* Functions.cpp:1160-1162                            */
    int bss;            /* Functions.cpp:1018-1020                            */
    int active;         /* Functions.cpp:1172-1178                            */
    int rrc_present;    /* Functions.cpp:1087                                 */
    int main_tx_used;   /* Functions.cpp:1089                                 */
    int main_rx_used;   /* Functions.cpp:1090                                 */
} orc_ed137_fields;
void orc_ed137_fields_from_word(uint32_t host_word, orc_ed137_fields *f);

/* 20-byte header serialiser used by golden-vector generation: writes the wire
 * image of `struct custom_rtp_hdr` (ed137_rtp.h:22-47) from plain fields.     */
void orc_hdr_write(uint8_t out[20], int v, int p, int x, int cc, int m, int pt,
                   uint16_t seq, uint32_t ts, uint32_t ssrc, uint16_t profile,
                   uint16_t length, uint32_t ed137_host);

/* ---------------- R2S liveness watchdog (roip_ed137.cpp:1756-1777) --------- */
/* One timer tick of RoIP_ED137::detectR2SPacketAndReconn for one connected radio:
 *   secDiff = now - r2sPacket; if (secDiff > r2sPeriod*3) { if (r2sCount == 5) hang up
 *   ("WG-67 ;cause=2001"); r2sCount++; } else r2sCount = 0;
 * Returns bit0 = late, bit1 = hang-up requested on this tick.                 */
int orc_r2s_watchdog(long long now, long long r2sPacket, int r2sPeriod, int *r2sCount);

/* ---------------- gate arbitration (RoIP_ED137::checkEvents) --------------- */
/* Per-leg and per-bridge state, names as in the reference's trx / radio
 * structs (roip_ed137.h).  gain_q7 = the slot volume the leg was last given
 * through setSlotVolume (roip_ed137.cpp:5190-5234): 0.0f -> 0, 2.0f -> 256.   */
typedef struct {
    uint8_t last;        /* lastTx (CLIENT) / lastRx (SERVER)                  */
    uint8_t msec;        /* lastTxmsec / lastRxmsec                            */
    uint8_t on;          /* m_PttPressed (CLIENT) / audioSQLOn (SERVER)        */
    int8_t rssi;         /* radio->rssi, -1 when not receiving (SERVER)        */
    uint16_t gain_q7;
    uint16_t reserved;
} orc_arb_leg;
typedef struct {
    int32_t ptt_level;        /* CLIENT: roip_ed137.cpp:6157                    */
    int32_t sqlStatusCount;   /* SERVER: :6028                                  */
    uint8_t sqlStatusOn;      /* SERVER: :6029                                  */
    uint8_t reserved[7];
} orc_arb_bridge;
/* CLIENT mode, one checkEvents() pass over the G legs of one bridge
 * (roip_ed137.cpp:6124-6231): highest ptt_type wins.  words[g] = the leg's
 * latched ED-137 word (host order); active[g] = callState && trxmode != RX.   */
void orc_arb_client_tick(orc_arb_bridge *b, orc_arb_leg *legs, const uint32_t *words,
                         const uint8_t *active, int G);
/* SERVER mode with rxBestSignalEnable, one checkEvents() pass: per-radio
 * squelch bookkeeping (roip_ed137.cpp:5627-5719 and its three twins) followed
 * by the best-signal selection with the 5-tick hold-off (:5985-6121).         */
void orc_arb_server_best_tick(orc_arb_bridge *b, orc_arb_leg *legs, const uint32_t *words,
                              const uint8_t *active, int G);

/* ---------------- PTT event logger message (Functions.cpp:2169-2211) ------- */
/* The "PTTEventDataLogger" JSON exactly as createPTTEventDataLogger lays it out;
 * doubles as QString::arg(double) prints them ('g', 6 digits), ints as arg(int).
 * Returns the length written (without the NUL).                               */
size_t orc_ptt_event_json(char *out, size_t cap, int softPhoneID, const char *strEvent,
                          double level_in_av, double level_in_max, double level_in_min,
                          const char *url, int rtp_av, int rtp_max, int rtp_min);

/* ---------------- WavWriter sink (WavWriter.cpp:63-156, Appendix E) ------- */
/* Writes the 44-byte header exactly as WavWriter::start() lays it out, with
 * the two size fields as WavWriter::stop() patches them for `payload_bytes`
 * bytes passed to wav_write().  Returns 44.                                  */
size_t orc_wav_header(uint8_t out[44], int rate, size_t payload_bytes);
/* WavWriter::wav_write: every payload byte b -> {b, 0x00}. Returns 2*len.    */
size_t orc_wav_body(const uint8_t *buf, size_t len, uint8_t *out);

#ifdef __cplusplus
}
#endif
#endif
