/*
 * igd_oracle.c -- CPU ORACLE (test infrastructure only; see igd_oracle.h).
 *
 * Scalar C restatement of the reference's per-packet voice path.  Every
 * function cites the reference file:line it follows (paths relative to
 * piyanon108/iGate4xSoftphoneDSP).  Nothing in the product may call this.
 */
#include "igd_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ======================================================================
 * G.711 -- SURVEY.md Appendix B.  The reference negotiates PCMA/PCMU with
 * PJSIP (roip_ed137.cpp:3563-3564); the arithmetic is pjmedia's
 * alaw_ulaw.c (Sun Microsystems g711.c derivative), un-vendored.
 * ====================================================================== */

static const int seg_end[8] = {0xFF, 0x1FF, 0x3FF, 0x7FF, 0xFFF, 0x1FFF, 0x3FFF, 0x7FFF};

static int seg_search(int val)
{
    for (int i = 0; i < 8; i++)
        if (val <= seg_end[i])
            return i;
    return 8;
}

int16_t orc_alaw2lin(uint8_t a_val)
{
    int a = a_val ^ 0x55;
    int t = (a & 0x0F) << 4;
    int seg = (a & 0x70) >> 4;
    switch (seg) {
    case 0:  t += 8; break;
    case 1:  t += 0x108; break;
    default: t += 0x108; t <<= seg - 1; break;
    }
    return (int16_t)((a & 0x80) ? t : -t);
}

int16_t orc_ulaw2lin(uint8_t u_val)
{
    int u = (~u_val) & 0xFF;
    int t = ((u & 0x0F) << 3) + 0x84;
    t <<= (u & 0x70) >> 4;
    return (int16_t)((u & 0x80) ? (0x84 - t) : (t - 0x84));
}

uint8_t orc_lin2alaw(int pcm)
{
    int mask, seg, aval;
    if (pcm >= 0) {
        mask = 0xD5;
    } else {
        mask = 0x55;
        pcm = -pcm - 8;
        if (pcm < 0)            /* pjsip ticket #1301 clamp, Appendix B */
            pcm = 0;
    }
    seg = seg_search(pcm);
    if (seg >= 8)
        return (uint8_t)(0x7F ^ mask);
    aval = seg << 4;
    if (seg < 2)
        aval |= (pcm >> 4) & 0x0F;
    else
        aval |= (pcm >> (seg + 3)) & 0x0F;
    return (uint8_t)(aval ^ mask);
}

uint8_t orc_lin2ulaw(int pcm)
{
    int mask, seg;
    if (pcm < 0) {
        pcm = 0x84 - pcm;
        mask = 0x7F;
    } else {
        pcm += 0x84;
        mask = 0xFF;
    }
    seg = seg_search(pcm);
    if (seg >= 8)
        return (uint8_t)(0x7F ^ mask);
    return (uint8_t)(((seg << 4) | ((pcm >> (seg + 3)) & 0x0F)) ^ mask);
}

/* Frame codec loops.  pjmedia's default build (PJMEDIA_HAS_ALAW_ULAW_TABLE) does not run the
 * algorithmic routines per sample: alaw_ulaw_table.c fills lookup tables FROM them once and the
 * codec indexes the tables.  The oracle does the same, so that the CPU baseline bench.py times is
 * the deployed (table) form and not a slower one; the results are those of the routines above by
 * construction (tests/test_oracle_pins.py checks table == routine on every input). */
static int16_t g_dec_tab[2][256];
static uint8_t g_enc_tab[2][65536];
static pthread_once_t g_tab_once = PTHREAD_ONCE_INIT;
static void build_tables(void)
{
    for (int c = 0; c < 256; c++) {
        g_dec_tab[ORC_LAW_ALAW][c] = orc_alaw2lin((uint8_t)c);
        g_dec_tab[ORC_LAW_ULAW][c] = orc_ulaw2lin((uint8_t)c);
    }
    for (int v = -32768; v < 32768; v++) {
        g_enc_tab[ORC_LAW_ALAW][(uint16_t)v] = orc_lin2alaw(v);
        g_enc_tab[ORC_LAW_ULAW][(uint16_t)v] = orc_lin2ulaw(v);
    }
}

void orc_g711_decode(const uint8_t *codes, int16_t *pcm, size_t n, int law)
{
    pthread_once(&g_tab_once, build_tables);
    const int16_t *t = g_dec_tab[law == ORC_LAW_ALAW ? ORC_LAW_ALAW : ORC_LAW_ULAW];
    for (size_t i = 0; i < n; i++) pcm[i] = t[codes[i]];
}

void orc_g711_encode(const int16_t *pcm, uint8_t *codes, size_t n, int law)
{
    pthread_once(&g_tab_once, build_tables);
    const uint8_t *t = g_enc_tab[law == ORC_LAW_ALAW ? ORC_LAW_ALAW : ORC_LAW_ULAW];
    for (size_t i = 0; i < n; i++) codes[i] = t[(uint16_t)pcm[i]];
}

/* ======================================================================
 * Meters -- SURVEY.md Appendix C
 * ====================================================================== */

/* roip_ed137.cpp:6557-6568 / :6511-6517:
 *     int audioLevelSum = 0; for (...) audioLevelSum += payloadbuf[i];
 *     uint8_t audioLevel = audioLevelSum/payloadlen;
 * `payloadbuf` is `const char*`: unsigned on the aarch64 targets
 * (iGate4xSoftphoneDSP.pro:64-80), signed on x86 (quirk Q4). */
uint8_t orc_bytemean(const uint8_t *payload, int n, int signed_char)
{
    int sum = 0;
    if (n <= 0)
        return 0;
    for (int i = 0; i < n; i++)
        sum += signed_char ? (int)(signed char)payload[i] : (int)payload[i];
    return (uint8_t)(sum / n);
}

void orc_frame_power(const int16_t *pcm, int n, uint64_t *sumsq, uint32_t *peak)
{
    uint64_t s = 0;
    uint32_t p = 0;
    for (int i = 0; i < n; i++) {
        int v = pcm[i];
        uint32_t m = (uint32_t)(v < 0 ? -v : v);
        s += (uint64_t)m * m;
        if (m > p) p = m;
    }
    *sumsq = s;
    *peak = p;
}

double orc_rms_dbfs(uint64_t sumsq, int n)
{
    if (sumsq == 0)
        return -INFINITY;
    return 10.0 * log10((double)sumsq / (double)n) - 20.0 * log10(32768.0);
}

double orc_peak_dbfs(uint32_t peak)
{
    if (peak == 0)
        return -INFINITY;
    return 20.0 * log10((double)peak / 32768.0);
}

/* audiometer.cpp:30-31: int(float((v*100.0)/30000.0)) */
int orc_percent(int v)
{
    return (int)((float)((v * 100.0) / 30000.0));
}

static void meter_pack(orc_meter_rec *r, uint64_t s, uint32_t peak, uint8_t bm, int n)
{
    r->sumsq_lo = (uint32_t)s;
    r->hi = (uint32_t)((s >> 32) & 0xFF) | ((uint32_t)bm << 8) | (peak << 16);
    r->rms_dbfs = (float)orc_rms_dbfs(s, n);
    r->peak_dbfs = (float)orc_peak_dbfs(peak);
}

/* ======================================================================
 * Gain + mix -- SURVEY.md Appendix D.  Reference: per-call RX gain set by
 * pjsua_conf_adjust_rx_level(slot, SLOT_VOLUME) (roip_ed137.cpp:5221) with
 * SLOT_VOLUME 0.0 / 2.0 / sidetone (Functions.cpp:1664-1705); every leg is
 * connected to conference slot 0 (roip_ed137.cpp:4907-4920).
 * ====================================================================== */

int orc_gain_adj(float level)
{
    return (int)((level - 1.0f) * 128) + 128;
}

static inline int clamp16(int v)
{
    if (v > 32767) return 32767;
    if (v < -32768) return -32768;
    return v;
}

int16_t orc_apply_gain(int16_t x, int adj)
{
    return (int16_t)clamp16(((int)x * adj) >> 7);
}

void orc_mix_frame(const int16_t *const *legs, const uint16_t *adj, int nlegs,
                   int n, int16_t *mix)
{
    for (int i = 0; i < n; i++) {
        int acc = 0;
        for (int g = 0; g < nlegs; g++)
            if (adj[g] != 0 && !(adj[g] & ORC_GAIN_NO_AUDIO))
                acc += orc_apply_gain(legs[g][i], adj[g]);
        mix[i] = (int16_t)clamp16(acc);
    }
}

/* ======================================================================
 * Fused per-frame voice path: decode -> meter -> gate/gain -> mix -> encode
 * processed the way the reference does it: one packet (channel-frame) at a
 * time.
 * ====================================================================== */

void orc_process_batch_range(const orc_batch *b, int b0, int b1)
{
    const int G = b->G, C = b->B * G, N = ORC_FRAME;
    int16_t pcm[32][ORC_FRAME];
    const int16_t *legp[32];
    for (int f = 0; f < b->F; f++) {
        for (int br = b0; br < b1; br++) {
            const uint16_t *adj = b->gain_q7 + (size_t)f * C + (size_t)br * G;
            int n_open = 0;
            for (int g = 0; g < G; g++) {
                size_t ch = (size_t)br * G + g;
                const uint8_t *codes = b->codes + ((size_t)f * C + ch) * N;
                uint64_t s; uint32_t pk;
                if (adj[g] & ORC_GAIN_NO_AUDIO) {
                    /* no audio frame arrived for this leg on this tick (keep-alive, lost or truncated
                     * packet): transport_rtp_cb never hands such a packet to the stream
                     * (TransportAdapter.cpp:298-315), the bridge hears the stream's silence */
                    memset(pcm[g], 0, sizeof pcm[g]);
                    if (b->meter)
                        meter_pack(&b->meter[(size_t)f * C + ch], 0, 0, 0, N);
                    legp[g] = pcm[g];
                    continue;
                }
                orc_g711_decode(codes, pcm[g], N, b->law[ch]);
                orc_frame_power(pcm[g], N, &s, &pk);
                if (b->meter)
                    meter_pack(&b->meter[(size_t)f * C + ch], s, pk,
                               orc_bytemean(codes, N, b->signed_char), N);
                legp[g] = pcm[g];
                n_open += adj[g] != 0;
            }
            size_t bf = (size_t)f * b->B + br;
            int16_t *mix = b->mix + bf * N;
            uint8_t *enc = b->enc + bf * N;
            orc_mix_frame(legp, adj, G, N, mix);
            orc_g711_encode(mix, enc, N, b->out_law[br]);
            if (b->bmeter) {
                uint64_t s; uint32_t pk;
                orc_frame_power(mix, N, &s, &pk);
                b->bmeter[bf].bytemean_out = orc_bytemean(enc, N, b->signed_char);
                b->bmeter[bf].n_open = (uint8_t)n_open;
                b->bmeter[bf].mix_peak = (uint16_t)pk;
            }
        }
    }
}

void orc_process_batch(const orc_batch *b)
{
    orc_process_batch_range(b, 0, b->B);
}

struct mt_arg { const orc_batch *b; int b0, b1; };

static void *mt_worker(void *p)
{
    struct mt_arg *a = (struct mt_arg *)p;
    orc_process_batch_range(a->b, a->b0, a->b1);
    return NULL;
}

void orc_process_batch_mt(const orc_batch *b, int nthreads)
{
    if (nthreads <= 1) { orc_process_batch(b); return; }
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nthreads);
    struct mt_arg *args = (struct mt_arg *)malloc(sizeof(struct mt_arg) * (size_t)nthreads);
    for (int t = 0; t < nthreads; t++) {
        args[t].b = b;
        args[t].b0 = (int)((long long)b->B * t / nthreads);
        args[t].b1 = (int)((long long)b->B * (t + 1) / nthreads);
        pthread_create(&th[t], NULL, mt_worker, &args[t]);
    }
    for (int t = 0; t < nthreads; t++)
        pthread_join(th[t], NULL);
    free(args);
    free(th);
}

/* ======================================================================
 * Event summary -- Functions.cpp:2126-2145 (keeplogAudioLevel) and
 * :2148-2230 (createPTTEventDataLogger).  "level" for the build is the
 * frame's linear mean square S/160; it is carried as
 * the exact integer S so the reduction is order-independent.
 * ====================================================================== */

void orc_event_summary(const orc_meter_rec *meter, const uint16_t *gain_q7,
                       int F, int C, orc_summary_rec *out)
{
    for (int c = 0; c < C; c++) {
        orc_summary_rec r;
        r.count = 0; r.bm_sum = 0; r.bm_max = 0; r.bm_min = 255;   /* :2159-2167 */
        r.sum_s = 0; r.max_s = 0; r.min_s = 255ull * ORC_FRAME;
        for (int f = 0; f < F; f++) {
            size_t i = (size_t)f * C + c;
            if (gain_q7[i] == 0 || (gain_q7[i] & ORC_GAIN_NO_AUDIO))
                continue;                                  /* eventPttSQL_In_LoggingOn; no level without audio */
            uint64_t s = (uint64_t)meter[i].sumsq_lo | ((uint64_t)(meter[i].hi & 0xFF) << 32);
            uint8_t bm = (uint8_t)(meter[i].hi >> 8);
            r.count += 1;                                  /* :2131 */
            r.sum_s += s;                                  /* :2133 */
            r.bm_sum = (uint16_t)(r.bm_sum + bm);          /* :2134, uint16_t wraps */
            if (s > r.max_s) r.max_s = s;                  /* :2135-2136 */
            if (s < r.min_s) r.min_s = s;                  /* :2137-2138 */
            if (bm > r.bm_max) r.bm_max = bm;              /* :2140-2141 */
            if (bm < r.bm_min) r.bm_min = bm;              /* :2142-2143 */
        }
        out[c] = r;
    }
}

void orc_summary_db(const orc_summary_rec *s, double *av_db, double *max_db,
                    double *min_db, int *bm_av)
{
    /* Functions.cpp:2196-2200 */
    double av = ((double)s->sum_s / ORC_FRAME) / (double)s->count;
    *av_db = 10 * log10(av);
    *max_db = 10 * log10((double)s->max_s / ORC_FRAME);
    *min_db = 10 * log10((double)s->min_s / ORC_FRAME);
    *bm_av = s->count ? (int)(uint8_t)(s->bm_sum / s->count) : 0;
}

/* ======================================================================
 * ED-137 RTP header extension -- SURVEY.md Appendix A
 * ====================================================================== */

static int ct_contains(const orc_adapter *a, const char *needle)
{
    return strstr(a->calltype, needle) != NULL;   /* QString::contains, case sensitive */
}

void orc_adapter_init(orc_adapter *a, int radiocall, int callIn,
                      const char *calltype, int keepAlivePeroid, long long now)
{
    /* TransportAdapter.cpp:97-128 (PJ_POOL_ZALLOC_T then field init) */
    memset(a, 0, sizeof(*a));
    a->radiostatus = radiocall;
    a->pttstatus = 0;
    a->callIn = callIn;
    a->pttpriority = 0;
    a->pttid = 0;
    a->keepAlivePeroid = keepAlivePeroid;
    strncpy(a->calltype, calltype, sizeof(a->calltype) - 1);
    a->ed137_value = 0;
    a->payloadsize = 0;
    a->r2sPacket = now;
    a->r2sSendtime = now;
    a->firstR2SPacket = 1;
    a->packetCnt = 0;
    a->callRecorder = 0;
    a->rxSlaveEnable = 0;
    a->txSlaveEnable = 0;
}

void orc_setAdapterPtt(orc_adapter *a, int pttval, int priority, int userRec)
{   /* TransportAdapter.cpp:135-146 */
    a->pttstatus = pttval != 0;
    a->pttpriority = (uint8_t)priority;
    a->callRecorder = userRec;
}

void orc_setTxRxSlaveEnable(orc_adapter *a, int rx, int tx)
{   /* TransportAdapter.cpp:147-157 */
    a->txSlaveEnableChanged = tx;
    a->rxSlaveEnableChanged = rx;
    a->trxSlaveEnableChangedCount = 0;
}

void orc_setAdapterQslOn(orc_adapter *a, int sqlval, int priority, uint32_t bssi)
{   /* TransportAdapter.cpp:181-192 */
    a->sqlstatus = sqlval != 0;
    a->sqlpriority = (uint8_t)priority;
    a->ed137_bssi = (uint8_t)bssi;
}

void orc_setAdapterPttId(orc_adapter *a, int pttid)
{   /* TransportAdapter.cpp:194-202 */
    a->pttid = (uint8_t)pttid;
}

void orc_setcallRecorder(orc_adapter *a, int val)
{   /* TransportAdapter.cpp:204-213 */
    a->callRecorder = val != 0;
}

void orc_setCallType(orc_adapter *a, const char *calltype)
{   /* TransportAdapter.cpp:215-223 */
    strncpy(a->calltype, calltype, sizeof(a->calltype) - 1);
    a->calltype[sizeof(a->calltype) - 1] = 0;
}

/* wire accessors for the little-endian bit-field branch of
 * struct custom_rtp_hdr (ed137_rtp.h:32-37) */
static void hdr_set_x(uint8_t *h, int x)   { h[0] = (uint8_t)((h[0] & ~0x10) | (x ? 0x10 : 0)); }
static void hdr_set_m(uint8_t *h, int m)   { h[1] = (uint8_t)((h[1] & 0x7F) | (m ? 0x80 : 0)); }
static void hdr_set_pt(uint8_t *h, int pt) { h[1] = (uint8_t)((h[1] & 0x80) | (pt & 0x7F)); }
static int  hdr_get_pt(const uint8_t *h)   { return h[1] & 0x7F; }
static void put_be16(uint8_t *p, uint16_t v) { p[0] = (uint8_t)(v >> 8); p[1] = (uint8_t)v; }
static void put_be32(uint8_t *p, uint32_t v)
{
    p[0] = (uint8_t)(v >> 24); p[1] = (uint8_t)(v >> 16); p[2] = (uint8_t)(v >> 8); p[3] = (uint8_t)v;
}

void orc_hdr_write(uint8_t out[20], int v, int p, int x, int cc, int m, int pt,
                   uint16_t seq, uint32_t ts, uint32_t ssrc, uint16_t profile,
                   uint16_t length, uint32_t ed137_host)
{
    out[0] = (uint8_t)(((v & 3) << 6) | ((p & 1) << 5) | ((x & 1) << 4) | (cc & 15));
    out[1] = (uint8_t)(((m & 1) << 7) | (pt & 0x7F));
    put_be16(out + 2, seq);
    put_be32(out + 4, ts);
    put_be32(out + 8, ssrc);
    put_be16(out + 12, profile);
    put_be16(out + 14, length);
    put_be32(out + 16, ed137_host);
}

/* The ED-137 word.  send_rtp: TransportAdapter.cpp:728-796;
 * sendR2SStatus: :499-565 (does not latch *Changed into the current flags). */
static uint32_t build_word(orc_adapter *a, int latch_changed)
{
    uint32_t w;
    if ((a->txSlaveEnable == a->txSlaveEnableChanged) &
        (a->rxSlaveEnable == a->rxSlaveEnableChanged) &
        (a->trxSlaveEnableChangedCount >= 5)) {
        if ((a->rxSlaveEnable == 0) & (a->txSlaveEnable == 0))      w = 0x00000000;
        else if ((a->rxSlaveEnable == 1) & (a->txSlaveEnable == 1)) w = 0x000131c0;
        else if ((a->rxSlaveEnable == 1) & (a->txSlaveEnable == 0)) w = 0x00013140;
        else if ((a->rxSlaveEnable == 0) & (a->txSlaveEnable == 1)) w = 0x00013180;
        else                                                        w = 0x00000000;
    } else {
        if (latch_changed) {                              /* :744-745 */
            a->txSlaveEnable = a->txSlaveEnableChanged;
            a->rxSlaveEnable = a->rxSlaveEnableChanged;
        }
        a->trxSlaveEnableChangedCount++;                  /* :746-747 */
        if (a->trxSlaveEnableChangedCount >= 5) a->trxSlaveEnableChangedCount = 5;
        if ((a->rxSlaveEnable == 0) & (a->txSlaveEnable == 0))      w = 0x00013100;
        else if ((a->rxSlaveEnable == 1) & (a->txSlaveEnable == 1)) w = 0x000131c0;
        else if ((a->rxSlaveEnable == 1) & (a->txSlaveEnable == 0)) w = 0x00013140;
        else if ((a->rxSlaveEnable == 0) & (a->txSlaveEnable == 1)) w = 0x00013180;
        else                                                        w = 0x00000000;
    }
    if (a->sqlstatus) {                                   /* :763-778 */
        a->sqlpriority = 0;
        w |= 0x10000000u & ((uint32_t)a->sqlstatus << 28);
        w |= 0x0fc00000u & ((uint32_t)a->sqlpriority << 22);
        w |= 0x000000f8u & ((uint32_t)a->ed137_bssi << 3);
    } else if (a->pttstatus == 0) {                       /* :779-784 */
        w |= 0x0fc00000u & ((uint32_t)1 << 22);
    }
    if (a->pttstatus) {                                   /* :786-796 */
        w |= 0x0fc00000u & ((uint32_t)a->pttid << 22);
        w |= 0xe0000000u & ((uint32_t)a->pttpriority << 29);
    }
    return w;
}

/* minimal pjmedia_rtp_decode_rtp (called at TransportAdapter.cpp:651):
 * payload starts after the 12-byte header, the CSRC list and, when X is set,
 * the extension (RFC 3550 5.1/5.3.1); un-vendored pjmedia/rtp.c. */
static size_t rtp_payload_offset(const uint8_t *pkt, size_t size)
{
    size_t off = 12 + 4u * (pkt[0] & 0x0F);
    if ((pkt[0] & 0x10) && off + 4 <= size)
        off += 4 + 4u * (((size_t)pkt[off + 2] << 8) | pkt[off + 3]);
    return off <= size ? off : size;
}

size_t orc_transport_send_rtp(orc_adapter *a, const uint8_t *pkt, size_t size,
                              long long now, uint8_t *out, int inviteServer,
                              int signed_char)
{
    const unsigned long long currenttime = (unsigned long long)now;   /* quint64, :638 */
    if (!a->radiostatus)                                              /* :641, :870-873 */
        return 0;
    if (size > 256) size = 256;                    /* fixed 256-B buffers, TransportAdapter.h:64-73 */
    size_t off = rtp_payload_offset(pkt, size);
    const uint8_t *payload = pkt + off;
    unsigned int payloadlen = (unsigned int)(size - off);
    uint8_t *h = a->send_pkt_buff;

    memcpy(h, pkt, size < 20 ? size : 20);                            /* :653 */
    memcpy(a->tmp_payload_buf, pkt, size);                            /* :654 */

    if (size > 60) {                                                  /* :657-673 */
        const uint8_t *t = a->tmp_payload_buf;
        if ((t[40] == t[50]) & (t[40] == t[60]) & (t[40] == 0xd5)) {
            a->rtpFalse += 1;
        } else {
            a->rtpFalse = 0;
        }
    }
    if (ct_contains(a, "Idle") && a->callIn) {                        /* :675-679 */
        a->sqlstatus = 0;
        a->pttstatus = 0;
    }
    if ((a->pttstatus & (a->callIn == 0)) || (a->sqlstatus & (a->callIn != 0))) {
        if (20 + payloadlen <= sizeof(a->send_pkt_buff))
            memcpy(h + 20, payload, payloadlen);                      /* :680-684 */
    } else {                                                          /* :685-706 */
        unsigned long long since = currenttime - (unsigned long long)a->r2sSendtime;
        if ((since < (unsigned long long)(long long)a->keepAlivePeroid) & (a->firstR2SPacket == 0))
            return 0;
        if (since >= (unsigned long long)(long long)a->keepAlivePeroid)
            a->r2sSendtime = now;
    }

    hdr_set_m(h, (a->firstR2SPacket != 0) & (a->packetCnt == 0));     /* :715-723 */
    hdr_set_x(h, 1);                                                  /* :725 */
    put_be16(h + 12, 0x0167);                                         /* :726 */
    put_be16(h + 14, 0x0001);                                         /* :727 */
    put_be32(h + 16, build_word(a, 1));                               /* :728-800 */

    const int rxonly = ct_contains(a, "Rxonly") || strcmp(a->calltype, "Rx") == 0;
    const int txish = ct_contains(a, "Tx") || ct_contains(a, "TRx");
    if (rxonly & (a->callIn == 0))                                    /* :801-804 */
        hdr_set_pt(h, 123);

    size_t outsize;
    if (!a->pttstatus && !a->sqlstatus) {                             /* :811-815 */
        hdr_set_pt(h, 123); outsize = 20;
    } else if (rxonly && !a->sqlstatus) {                             /* :821-825 */
        hdr_set_pt(h, 123); outsize = 20;
    } else if (txish && !a->pttstatus && !a->sqlstatus) {             /* :826-829 */
        hdr_set_pt(h, 123); outsize = 20;
    } else if (txish && a->pttstatus && a->callIn) {                  /* :830-841 */
        if (a->callRecorder || a->sqlstatus) {
            outsize = 20 + payloadlen;
        } else {
            hdr_set_pt(h, 123); outsize = 20;
        }
    } else {                                                          /* :842-845 */
        outsize = 20 + payloadlen;
    }
    if (outsize > sizeof(a->send_pkt_buff)) outsize = sizeof(a->send_pkt_buff);
    memcpy(out, h, outsize);                                          /* :848 */

    if ((a->firstR2SPacket != 0) & (a->packetCnt < 30))               /* :849-855 */
        a->packetCnt++;
    else if (a->packetCnt >= 30)
        a->firstR2SPacket = 0;

    if (hdr_get_pt(h) != 123) {                                       /* :858-862 */
        a->send_payload_bufSize = payloadlen;
        if (inviteServer)      /* setOutgoingRTP, roip_ed137.cpp:6500-6536 (quirk Q3) */
            a->OutgoingRTP = orc_bytemean(a->tmp_payload_buf, (int)payloadlen, signed_char);
    }
    return outsize;
}

size_t orc_sendR2SStatus(orc_adapter *a, long long now, uint8_t *out)
{
    const unsigned long long currenttime = (unsigned long long)now;   /* :426 */
    if (!a->radiostatus)                                              /* :429 */
        return 0;
    uint8_t *h = a->send_pkt_buff;
    const unsigned int payloadlen = 0;                                /* :436 */

    if (ct_contains(a, "Idle") && a->callIn) {                        /* :444-448 */
        a->sqlstatus = 0;
        a->pttstatus = 0;
    }
    if ((a->pttstatus & (a->callIn == 0)) || (a->sqlstatus & (a->callIn != 0)))
        return 0;                                                     /* :449-454 */
    {
        unsigned long long since = currenttime - (unsigned long long)a->r2sSendtime;
        if ((since < (unsigned long long)(long long)a->keepAlivePeroid) & (a->firstR2SPacket == 0))
            return 0;                                                 /* :457-460 */
        if (since >= (unsigned long long)(long long)a->keepAlivePeroid)
            a->r2sSendtime = now;                                     /* :463-473 */
    }

    hdr_set_m(h, (a->firstR2SPacket != 0) & (a->packetCnt == 0));     /* :485-493 */
    hdr_set_x(h, 1);
    put_be16(h + 12, 0x0167);
    put_be16(h + 14, 0x0001);
    put_be32(h + 16, build_word(a, 0));                               /* :499-569 */

    const int rxonly = ct_contains(a, "Rxonly") || strcmp(a->calltype, "Rx") == 0;
    const int txish = ct_contains(a, "Tx") || ct_contains(a, "TRx");
    if (rxonly & (a->callIn == 0))                                    /* :570-573 */
        hdr_set_pt(h, 123);

    size_t outsize;
    if (!a->pttstatus && !a->sqlstatus) {                             /* :580-584 */
        hdr_set_pt(h, 123); outsize = 20;
    } else if (ct_contains(a, "Idle") && a->callIn) {                 /* :585-589 */
        hdr_set_pt(h, 123); outsize = 20;
    } else if (rxonly && !a->sqlstatus) {                             /* :590-594 */
        hdr_set_pt(h, 123); outsize = 20;
    } else if (txish && !a->pttstatus && !a->sqlstatus) {             /* :595-598 */
        hdr_set_pt(h, 123); outsize = 20;
    } else if (txish && a->pttstatus && a->callIn) {                  /* :599-610 */
        if (a->callRecorder || a->sqlstatus) {
            outsize = 20 + payloadlen;
        } else {
            hdr_set_pt(h, 123); outsize = 20;
        }
    } else {
        outsize = 20 + payloadlen;                                    /* :611-614 */
    }
    if (hdr_get_pt(h) != 123)                                         /* :617 */
        return 0;
    memcpy(out, h, outsize);                                          /* :621 */
    if ((a->firstR2SPacket != 0) & (a->packetCnt < 30))               /* :622-629 */
        a->packetCnt++;
    else if (a->packetCnt >= 30)
        a->firstR2SPacket = 0;
    return outsize;
}

int orc_transport_rtp_cb(orc_adapter *a, const uint8_t *pkt, size_t size,
                         long long now, int inviteServer, int signed_char)
{
    /* decodeRtp (:408-415) is a cast; the reference keeps the header pointer
     * in a file-static shared by all calls (:76), so for non-radio calls the
     * PT test at :298 reads whatever packet a radio call saw last.  The
     * oracle reads this packet's own PT in both cases. */
    const int pt = size >= 2 ? (pkt[1] & 0x7F) : 0;
    if (a->radiostatus && size >= 20) {                               /* :246-262 */
        if (pt == 8 || pt == 0 || pt == 18 || pt == 123) {
            memcpy(&a->ed137_value, pkt + 16, 4);    /* raw, network byte order (:254) */
            uint16_t len_raw;
            memcpy(&len_raw, pkt + 14, 2);           /* un-swapped (:255) */
            a->payloadsize = len_raw;
        }
    }
    unsigned int payloadlen;
    if (!a->radiostatus) {                                            /* :267-275 */
        payloadlen = (unsigned int)(size - 12);
        a->payload_bufSize = payloadlen;
        memcpy(a->payload_buff, pkt + 12, payloadlen < 256 ? payloadlen : 256);
    } else {                                                          /* :276-292 */
        payloadlen = (unsigned int)(size - 20);      /* wraps when size < 20 */
        a->payload_bufSize = payloadlen;
        if (payloadlen < 1024) {
            /* the reference copies up to 1023 B into a 256-B array; capped here */
            memcpy(a->payload_buff, pkt + 20, payloadlen < 256 ? payloadlen : 256);
        } else {
            a->r2sPacket = now;
            return -1;
        }
    }
    if (pt != 123) {                                                  /* :298-307 */
        a->r2sPacket = now;
        if (inviteServer)      /* setIncomingRTP, roip_ed137.cpp:6541-6587 */
            a->IncomingRTP = orc_bytemean(a->payload_buff,
                                          (int)(payloadlen < 256 ? payloadlen : 256), signed_char);
        if (a->rtpAudio == 0)
            a->checkEvents_calls++;                                   /* :304-305 */
        a->rtpAudio = 1;
        return 1;
    }
    a->r2sPacket = now;                                               /* :308-315 */
    if (a->rtpAudio == 1)
        a->checkEvents_calls++;
    a->rtpAudio = 0;
    return 0;
}

uint32_t orc_get_ed137_value(const orc_adapter *a)
{   /* TransportAdapter.cpp:337-346: ntohl(adapter->ed137_value) */
    const uint8_t *p = (const uint8_t *)&a->ed137_value;
    return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3];
}

void orc_ed137_fields_from_word(uint32_t w, orc_ed137_fields *f)
{
    f->word = w;
    f->ptt_type = (int)((w & 0xe0000000u) >> 29);        /* Functions.cpp:1136-1138 */
    f->ptt_id = (int)((w & 0x0fc00000u) >> 22);          /* Functions.cpp:1148-1150 */
    f->squelch = (int)((w & 0x10000000u) >> 28);         /* Functions.cpp:1160-1162 */
    f->bss = (int)((w & 0x000000f8u) >> 3);              /* Functions.cpp:1018-1020 */
    f->active = w > 0;                                   /* Functions.cpp:1172-1178 */
    f->rrc_present = (w & 0x00013100u) == 0x00013100u;   /* Functions.cpp:1087 */
    f->main_tx_used = !((w & 0x00000080u) == 0x80u);     /* Functions.cpp:1089 */
    f->main_rx_used = !((w & 0x00000040u) == 0x40u);     /* Functions.cpp:1090 */
}

/* ======================================================================
 * WavWriter sink -- WavWriter.cpp:63-156
 * ====================================================================== */

static void put_le32(uint8_t *p, uint32_t v)
{
    p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16); p[3] = (uint8_t)(v >> 24);
}
static void put_le16(uint8_t *p, uint16_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); }

size_t orc_wav_header(uint8_t out[44], int rate, size_t payload_bytes)
{
    const unsigned bitspersample = 16, channels = 2;     /* :70-72 */
    const size_t total = 44 + 2 * payload_bytes;         /* wav_write doubles every byte */
    memcpy(out, "RIFF", 4);                              /* :88 */
    put_le32(out + 4, (uint32_t)(total - 8));            /* :119-122 */
    memcpy(out + 8, "WAVE", 4);                          /* :92 */
    memcpy(out + 12, "fmt ", 4);                         /* :93 */
    put_le32(out + 16, 16);                              /* :94-95 */
    put_le16(out + 20, 0x0007);                          /* WAVE_FORMAT_MULAW, 2 bytes (:96-97) */
    put_le16(out + 22, (uint16_t)channels);              /* :98-99 */
    put_le32(out + 24, (uint32_t)rate);                  /* :100-101 */
    put_le32(out + 28, (uint32_t)(rate * bitspersample / 8 * channels));  /* :102-103 */
    put_le16(out + 32, (uint16_t)(bitspersample / 8 * channels));         /* :104-105 */
    put_le16(out + 34, (uint16_t)bitspersample);         /* :106-107 */
    memcpy(out + 36, "data", 4);                         /* first 4 of "data " (:108) */
    put_le32(out + 40, (uint32_t)(total - 44));          /* :123-125 */
    return 44;
}

size_t orc_wav_body(const uint8_t *buf, size_t len, uint8_t *out)
{
    /* wav_write -> write_little_endian(word, channels=2): {b, 0} (:136-156) */
    for (size_t i = 0; i < len; i++) {
        out[2 * i] = buf[i];
        out[2 * i + 1] = 0;
    }
    return 2 * len;
}

/* ------------------------------------------------------------------------- */
/* R2S watchdog: RoIP_ED137::detectR2SPacketAndReconn, roip_ed137.cpp:1767-1780 */
int orc_r2s_watchdog(long long now, long long r2sPacket, int r2sPeriod, int *r2sCount)
{
    int flags = 0;
    long long secDiff = now - r2sPacket;                 /* :1767 */
    if (secDiff > (long long)r2sPeriod * 3) {            /* :1768 */
        flags |= 1;
        if (*r2sCount == 5) flags |= 2;                  /* :1770-1774 trx_call_hangup(...2001...) */
        (*r2sCount)++;                                   /* :1775 */
    } else {
        *r2sCount = 0;                                   /* :1778 */
    }
    return flags;
}

/* ------------------------------------------------------------------------- */
/* CLIENT-mode gate arbitration, roip_ed137.cpp:6124-6231.  setSlotVolume with
 * SLOT_VOLUME = 2.0f / 0.0f becomes gain_q7 = 256 / 0 (orc_gain_adj).          */
void orc_arb_client_tick(orc_arb_bridge *b, orc_arb_leg *legs, const uint32_t *words,
                         const uint8_t *active, int G)
{
    int checkPTT = 0;                                                    /* :6126 */
    for (int i = 0; i < G; i++) {                                        /* :6127 */
        if (active && !active[i]) continue;                              /* :6131-6134 */
        int pttstatus = (int)((words[i] & 0xe0000000u) >> 29);           /* :6138, Functions.cpp:1136 */
        if (pttstatus != legs[i].last) {                                 /* :6140 */
            if (pttstatus == 0) {
                legs[i].msec++;                                          /* :6143 */
                if (legs[i].msec < 6) pttstatus = 1;                     /* :6144-6147 */
            }
        } else {
            legs[i].msec = 0;                                            /* :6152 */
        }
        legs[i].last = (uint8_t)pttstatus;                               /* :6155 */
        if (pttstatus) {                                                 /* :6158 */
            if (pttstatus > b->ptt_level) {                              /* :6159 */
                b->ptt_level = pttstatus;                                /* :6161 */
                legs[i].gain_q7 = 256;                                   /* :6162-6163 SLOT_VOLUME = 2.0f */
                for (int j = 0; j < G; j++)                              /* :6164-6174 */
                    if (legs[j].on && j != i) legs[j].gain_q7 = 0;
            }
        }
        if (pttstatus > 0 && !legs[i].on) {                              /* :6190 */
            legs[i].on = 1;
        } else if (pttstatus == 0 && legs[i].on) {                       /* :6197 */
            legs[i].on = 0;
            for (int j = 0; j < G; j++) if (legs[j].on) checkPTT++;      /* :6201-6205 */
            legs[i].gain_q7 = 0;                                         /* :6217-6218 */
            b->ptt_level = pttstatus;                                    /* :6221 */
        }
    }
    (void)checkPTT;
}

/* SERVER mode with rxBestSignalEnable: roip_ed137.cpp:5627-5719 (x4 radios),
 * then :5985-6121.  setvolume(MUTE) -> gain 0; setvolume(UNMUTE) -> gain 256
 * (the SLOT_VOLUME = 2.0f branch of Functions.cpp:1664-1705; its trxStatus /
 * sidetone side conditions belong to the control plane and are taken as met). */
void orc_arb_server_best_tick(orc_arb_bridge *b, orc_arb_leg *legs, const uint32_t *words,
                              const uint8_t *active, int G)
{
    for (int i = 0; i < G; i++) {
        if (active && !active[i]) continue;                              /* :5614 */
        int sqlon = (int)((words[i] & 0x10000000u) >> 28);               /* :5627, Functions.cpp:1160 */
        legs[i].rssi = (int8_t)((words[i] & 0xf8u) >> 3);                /* :5657, Functions.cpp:1018 */
        if (sqlon != legs[i].last) {                                     /* :5658 */
            if (sqlon == 0) {
                legs[i].msec++;                                          /* :5660 */
                if (legs[i].msec < 1) sqlon = 1;                         /* :5661-5664 */
            }
        } else {
            legs[i].msec = 0;                                            /* :5669 */
        }
        if (legs[i].last != sqlon) {                                     /* :5694 */
            legs[i].last = (uint8_t)sqlon;                               /* :5713 */
            if (!legs[i].last) legs[i].gain_q7 = 0;                      /* :5717-5718 */
        }
    }
    int any = 0;
    for (int i = 0; i < G; i++) {                                        /* :5987-6030 */
        const int callState = !active || active[i];
        if (!callState || legs[i].last == 0) {
            if (legs[i].on) { b->sqlStatusCount = 0; b->sqlStatusOn = 0; }
            legs[i].on = 0;
            legs[i].rssi = -1;
        }
        if (legs[i].last > 0) any = 1;
    }
    if (any) {                                                           /* :6026 */
        b->sqlStatusCount++;                                             /* :6028 */
        if (b->sqlStatusCount >= 5 && !b->sqlStatusOn) {                 /* :6029 */
            for (int i = 0; i < G; i++)                                  /* :6031-6038 */
                if (legs[i].last > 0) legs[i].gain_q7 = 0;
            b->sqlStatusOn = 1;                                          /* :6039 */
            for (int i = 0; i < G; i++) {                                /* :6046-6109: first radio whose */
                int best = legs[i].last != 0;                            /* rssi is >= every other one   */
                for (int j = 0; j < G && best; j++)
                    if (j != i && legs[i].rssi < legs[j].rssi) best = 0;
                if (best) {
                    for (int j = 0; j < G; j++) legs[j].on = (uint8_t)(j == i);
                    if (!active || active[i]) legs[i].gain_q7 = 256;     /* :6056-6057 */
                    break;
                }
            }
        }
    } else {                                                             /* :6112-6120 */
        b->sqlStatusCount = 0;
        b->sqlStatusOn = 0;
        for (int i = 0; i < G; i++) legs[i].on = 0;
    }
}

/* ------------------------------------------------------------------------- */
/* createPTTEventDataLogger's message, Functions.cpp:2169-2181 and :2199-2211:
 * QString("{" "\"menuID\"   :\"PTTEventDataLogger\", " ... "}").arg(...) x9.     */
static void orc_qnum(char *b, size_t n, double v)
{
    if (isnan(v)) snprintf(b, n, "nan");
    else if (isinf(v)) snprintf(b, n, v < 0 ? "-inf" : "inf");
    else snprintf(b, n, "%g", v);                 /* QString::number(v, 'g', 6) */
}
size_t orc_ptt_event_json(char *out, size_t cap, int softPhoneID, const char *strEvent,
                          double av, double mx, double mn, const char *url,
                          int rtp_av, int rtp_max, int rtp_min)
{
    char a[64], b[64], c[64];
    orc_qnum(a, sizeof a, av); orc_qnum(b, sizeof b, mx); orc_qnum(c, sizeof c, mn);
    int n = snprintf(out, cap,
        "{"
        "\"menuID\"                       :\"PTTEventDataLogger\", "
        "\"softPhoneID\"                  :%d, "
        "\"Ptt\"                          :\"%s\", "
        "\"level_in_av\"                  :%s, "
        "\"level_in_max\"                 :%s, "
        "\"level_in_min\"                 :%s, "
        "\"radioUrl \"                    :\"%s\","
        "\"OutgoingRTPAv\"                :%d, "
        "\"OutgoingRTPmax\"               :%d, "
        "\"OutgoingRTPmin\"               :%d "
        "}", softPhoneID, strEvent, a, b, c, url, rtp_av, rtp_max, rtp_min);
    return n < 0 ? 0 : (size_t)n;
}
