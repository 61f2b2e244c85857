/* igd_tick.c -- the C ABI from plain C (C99): one 20 ms tick of a 4-radio bridge through the GPU path.
 *   gcc -std=c99 -Iinclude examples/igd_tick.c -Ligate4xsoftphonedsp_b200 -ligate_dsp -o igd_tick
 * Packets in -> header parse -> receive-side liveness -> PTT arbitration -> decode/meter/mix/encode straight
 * out of the packets (igd_process_packets; the payload array is never materialised) -> levels. */
#include <stdio.h>
#include <string.h>

#include "igate_dsp.h"

enum { B = 1, G = 4, C = B * G, F = 1 };

int main(void)
{
    igd_ctx *ctx = NULL;
    int rc = igd_init(0, &ctx);
    if (rc != IGD_OK) {                        /* no sm_100 GPU: there is no CPU fallback */
        fprintf(stderr, "igd_init failed: %d\n", rc);
        return 1;
    }
    static uint8_t pkts[C][IGD_PKT_MAX];
    uint32_t sizes[C];
    for (int c = 0; c < C; c++) {              /* radio c keys PTT type c (0 = idle ... 3 = priority) */
        const uint32_t word = (uint32_t)c << 29;
        memset(pkts[c], 0xD5, sizeof pkts[c]); /* A-law silence */
        pkts[c][0] = 0x90; pkts[c][1] = 8;     /* V=2, X=1, PT 8 (PCMA) */
        pkts[c][12] = 0x01; pkts[c][13] = 0x67; pkts[c][14] = 0x00; pkts[c][15] = 0x01;
        pkts[c][16] = (uint8_t)(word >> 24); pkts[c][17] = (uint8_t)(word >> 16);
        pkts[c][18] = (uint8_t)(word >> 8);  pkts[c][19] = (uint8_t)word;
        sizes[c] = IGD_PKT_MAX;
    }
    igd_ed137_fields fields[C];
    rc = igd_ed137_parse(ctx, &pkts[0][0], sizes, C, IGD_PKT_MAX, fields, NULL /* no payload array */, IGD_MEM_HOST);

    igd_rx_state rxs[C];
    igd_rx_event ev[C];
    memset(rxs, 0, sizeof rxs);
    igd_rx_track_desc rx = {sizeof rx, IGD_MEM_HOST, F, C, 20, 200, 2, 0, 0, fields, NULL, rxs, ev, NULL};
    if (rc == IGD_OK) rc = igd_rx_track(ctx, &rx);

    igd_arb_leg legs[C];
    igd_arb_bridge br[B];
    uint16_t gain[C];
    memset(legs, 0, sizeof legs);
    memset(br, 0, sizeof br);
    igd_arb_desc arb = {sizeof arb, IGD_MEM_HOST, F, B, G, IGD_ARB_CLIENT_PTT, 8, 0, ev, NULL, legs, br, gain};
    if (rc == IGD_OK) rc = igd_gate_arbitrate(ctx, &arb);

    const uint8_t law[C] = {IGD_LAW_ALAW, IGD_LAW_ALAW, IGD_LAW_ALAW, IGD_LAW_ALAW}, out_law[B] = {IGD_LAW_ALAW};
    static int16_t mix[B][IGD_FRAME];
    static uint8_t enc[B][IGD_FRAME];
    igd_meter_rec meter[C];
    igd_bridge_rec bm[B];
    igd_packets_desc d = {sizeof d, IGD_MEM_HOST, F, B, G, 0, &pkts[0][0], fields, law, gain, out_law,
                          &mix[0][0], &enc[0][0], meter, bm};
    if (rc == IGD_OK) rc = igd_process_packets(ctx, &d);
    if (rc != IGD_OK) {
        fprintf(stderr, "error %d: %s\n", rc, igd_last_error(ctx));
        igd_shutdown(ctx);
        return 2;
    }
    for (int c = 0; c < C; c++)
        printf("leg %d: ptt_type %u gain_q7 %u bytemean %u peak %u\n", c, (unsigned)fields[c].ptt_type,
               (unsigned)gain[c], (unsigned)IGD_METER_BYTEMEAN(meter[c]), (unsigned)IGD_METER_PEAK(meter[c]));
    printf("bridge: %u legs open, outgoing byte-mean %u, mix[0] %d\n", (unsigned)bm[0].n_open,
           (unsigned)bm[0].bytemean_out, (int)mix[0][0]);
    igd_shutdown(ctx);
    return 0;
}
