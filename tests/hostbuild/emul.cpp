// Host build of igate4xsoftphonedsp_b200/csrc/igd_math.cuh (-DIGD_HOST_EMUL).
// TEST-ONLY: lets the CPU test-suite check the device bit tricks (float-exponent
// G.711 compressor, sender state machine, dB maps) against the oracle without a
// GPU.  It is never loaded by the product.
#define IGD_HOST_EMUL 1
#include <math.h>
#include "../../include/igate_dsp.h"
#include "../../igate4xsoftphonedsp_b200/csrc/igd_math.cuh"

extern "C" {
int emul_alaw2lin(unsigned c) { return igd_alaw2lin(c); }
int emul_ulaw2lin(unsigned c) { return igd_ulaw2lin(c); }
void emul_encode_all(int law, unsigned char *out /*65536*/)
{
    igd_enc_law L = igd_enc_law_make(law);
    for (int v = -32768; v < 32768; v++) out[v + 32768] = (unsigned char)igd_g711_enc1(v, L);
}
float emul_rms_dbfs(unsigned long long s) { return igd_rms_dbfs(s); }
float emul_peak_dbfs(unsigned p) { return igd_peak_dbfs(p); }
unsigned emul_bytemean(int sum, int len) { return igd_bytemean_from_sum(sum, len); }
void emul_fields(unsigned w, unsigned *o)
{
    igd_edf f = igd_ed137_fields_of(w);
    o[0] = f.ptt_type; o[1] = f.ptt_id; o[2] = f.squelch; o[3] = f.bss; o[4] = f.flags;
}
// one sender step; plan out = {word, size, pt123, marker, copy_payload}
void emul_tx_step(igd_ed137_state *s, unsigned payload_len, long long now, unsigned *o)
{
    igd_tx_plan p = igd_ed137_tx_step(*s, payload_len, now);
    o[0] = p.word; o[1] = p.size; o[2] = p.pt123; o[3] = p.marker; o[4] = p.copy_payload;
}
// one sendR2SStatus call; plan out = {word, size, pt123, marker, header_written}
void emul_r2s_step(igd_ed137_state *s, long long now, unsigned stale_pt, unsigned *o)
{
    igd_tx_plan p = igd_ed137_r2s_step(*s, now, stale_pt);
    o[0] = p.word; o[1] = p.size; o[2] = p.pt123; o[3] = p.marker; o[4] = p.copy_payload;
}
// one RX tick: fields record, state in/out; returns IGD_RXE_* bits
unsigned emul_rx_step(igd_rx_state *s, const igd_ed137_fields *f, int present, int wd, long long now, int period)
{
    return igd_rx_step(*s, *f, present != 0, wd != 0, now, period);
}
void emul_arb_tick(int mode, igd_arb_bridge *b, igd_arb_leg *legs, const unsigned *words, const unsigned char *active, int G)
{
    auto word = [&](int g) { return words[g]; };
    auto act = [&](int g) { return active ? active[g] != 0 : true; };
    if (mode == IGD_ARB_CLIENT_PTT) igd_arb_client_tick(*b, legs, G, word, act);
    else igd_arb_server_best_tick(*b, legs, G, word, act);
}
}
