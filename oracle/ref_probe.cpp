/*
 * ref_probe.cpp -- thin driver around the REFERENCE'S OWN sources, compiled
 * where they lie under /root/reference (never copied into this repo):
 *   - ed137_rtp.h   : struct custom_rtp_hdr (ed137_rtp.h:22-47), through the
 *                     3-typedef shim in ref_shim/pjmedia/endpoint.h
 *   - WavWriter.cpp : compiled verbatim next to this file by oracle/Makefile
 * Output: oracle/_ref/libigd_ref.so (git-ignored).  Test infrastructure: it
 * validates oracle/igd_oracle.c and generates tests/golden/ fixtures
 * (tests/golden/make_golden.py).  The product never loads it.
 */
#include <arpa/inet.h>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <string>

#include "ed137_rtp.h"
#include "WavWriter.h"

extern "C" {

int ref_hdr_sizeof() { return (int)sizeof(custom_rtp_hdr); }

void ref_hdr_offsets(int *o)
{
    o[0] = (int)offsetof(custom_rtp_hdr, seq);
    o[1] = (int)offsetof(custom_rtp_hdr, ts);
    o[2] = (int)offsetof(custom_rtp_hdr, ssrc);
    o[3] = (int)offsetof(custom_rtp_hdr, profile_data);
    o[4] = (int)offsetof(custom_rtp_hdr, length);
    o[5] = (int)offsetof(custom_rtp_hdr, ed137);
}

/* Fill a header through the reference's bit-fields, with the same byte-order
 * calls TransportAdapter.cpp:725-727,800 uses (ntohs for the 16-bit fields,
 * htonl for the ED-137 word). */
void ref_hdr_build(uint8_t *out20, int v, int p, int x, int cc, int m, int pt,
                   unsigned seq, unsigned ts, unsigned ssrc, unsigned profile,
                   unsigned length, unsigned ed137_host)
{
    memset(out20, 0, sizeof(custom_rtp_hdr));
    custom_rtp_hdr *h = (custom_rtp_hdr *)out20;
    h->v = v; h->p = p; h->x = x; h->cc = cc; h->m = m; h->pt = pt;
    h->seq = htons((uint16_t)seq);
    h->ts = htonl(ts);
    h->ssrc = htonl(ssrc);
    h->profile_data = ntohs((uint16_t)profile);
    h->length = ntohs((uint16_t)length);
    h->ed137 = htonl(ed137_host);
}

/* Read the fields back through the reference struct (transport_rtp_cb reads
 * pt / ed137 / length this way, TransportAdapter.cpp:252-255; get_ed137_value
 * applies ntohl, :342). */
void ref_hdr_parse(const uint8_t *in20, unsigned *f)
{
    const custom_rtp_hdr *h = (const custom_rtp_hdr *)in20;
    f[0] = h->v; f[1] = h->p; f[2] = h->x; f[3] = h->cc; f[4] = h->m; f[5] = h->pt;
    f[6] = ntohs(h->seq); f[7] = ntohl(h->ts); f[8] = ntohl(h->ssrc);
    f[9] = ntohs(h->profile_data); f[10] = h->length /* un-swapped, :255 */;
    f[11] = ntohl(h->ed137);
}

/* Mutate an existing 20-byte buffer the way transport_send_rtp does:
 * m, x=1, profile, length, ed137 word, optional pt override. */
void ref_hdr_stamp(uint8_t *buf20, int m, unsigned ed137_host, int pt_override)
{
    custom_rtp_hdr *h = (custom_rtp_hdr *)buf20;
    h->m = m ? 1 : 0;
    h->x = 1;
    h->profile_data = ntohs(0x0167);
    h->length = ntohs(0x1);
    h->ed137 = ed137_host;
    h->ed137 = htonl(h->ed137);
    if (pt_override >= 0) h->pt = pt_override;
}

/* Drive the reference WavWriter: start(prefix, rate), wav_write in `chunk`
 * byte pieces, stop().  The file lands at <prefix>YYYYMMDDhhmmss.wav. */
int ref_wavwriter_run(const char *prefix, int rate, const uint8_t *payload,
                      unsigned len, unsigned chunk)
{
    WavWriter w;
    w.start(std::string(prefix), rate);
    if (!w.isRunning()) return -1;
    for (unsigned o = 0; o < len; o += chunk) {
        unsigned n = len - o < chunk ? len - o : chunk;
        w.wav_write((unsigned char *)payload + o, n);
    }
    w.stop();
    return 0;
}

}  /* extern "C" */
