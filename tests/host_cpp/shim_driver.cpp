// shim_driver.cpp -- drives the C++ host shim (igate_shim.h) the way the reference's
// RoIP_ED137 / PJSIP callbacks would: create one adapter per call, call the setters,
// hand every tick's packets to the (batched) send path and feed the emitted packets back
// through the receive path.  Reads a scenario file, writes a result file (see
// tests/test_gpu_host_shim.py).  Test-only.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../igate4xsoftphonedsp_b200/host/igate_shim.h"

struct Out {
    std::vector<uint8_t> pkts;
    std::vector<uint32_t> sizes;
    int f, C;
    std::vector<pjmedia_transport *> *tps;
};

static void on_send(void *user, pjmedia_transport *tp, const void *pkt, pj_size_t size)
{
    Out *o = static_cast<Out *>(user);
    int c = 0;
    while ((*o->tps)[c] != tp) c++;
    memcpy(&o->pkts[((size_t)o->f * o->C + c) * 180], pkt, size);
    o->sizes[(size_t)o->f * o->C + c] = (uint32_t)size;
}

struct Ev {
    std::vector<uint8_t> edge;     // [F][C] setIncomingED137Value fired
    std::vector<uint8_t> hang;     // [F][C] watchdog asked for a hang-up
    int f, C;
    std::vector<pjmedia_transport *> *tps;
};

static void on_event(void *user, pjmedia_transport *tp, unsigned flag, pj_uint32_t)
{
    Ev *e = static_cast<Ev *>(user);
    int c = 0;
    while ((*e->tps)[c] != tp) c++;
    if (flag == IGD_RXE_EDGE) e->edge[(size_t)e->f * e->C + c] = 1;
    if (flag == IGD_RXE_HANGUP) e->hang[(size_t)e->f * e->C + c] = 1;
}

int main(int argc, char **argv)
{
    if (argc < 3) return 2;
    FILE *in = fopen(argv[1], "rb");
    if (!in) return 2;
    int32_t F, C, tick, flags;
    int64_t now0;
    if (fread(&F, 4, 1, in) != 1 || fread(&C, 4, 1, in) != 1 || fread(&tick, 4, 1, in) != 1 ||
        fread(&flags, 4, 1, in) != 1 || fread(&now0, 8, 1, in) != 1) return 2;
    igd_bank *bank = igd_bank_open(0, C);
    if (!bank) { fprintf(stderr, "igd_bank_open failed (no sm_100 GPU?)\n"); return 3; }
    // scenarios run on their own clock: the adapter constructor stamps r2sSendtime / r2sPacket from it
    static long long clock_now;
    clock_now = now0;
    igd_bank_set_clock(bank, [](void *u) { return *static_cast<long long *>(u); }, &clock_now);
    std::vector<pjmedia_transport *> tps(C);
    std::vector<int32_t> slave(2 * C);
    for (int c = 0; c < C; c++) {
        int32_t radiocall, callIn, keepalive;
        char calltype[64];
        if (fread(&radiocall, 4, 1, in) != 1 || fread(&callIn, 4, 1, in) != 1 || fread(&keepalive, 4, 1, in) != 1 ||
            fread(&slave[2 * c], 4, 2, in) != 2 || fread(calltype, 1, 64, in) != 64) return 2;
        if (pjmedia_custom_tp_adapter_create(nullptr, "tpad", nullptr, 0, radiocall, callIn, calltype, c, &tps[c], "1",
                                             "", keepalive, 1, 1) != PJ_SUCCESS) return 4;
        if (slave[2 * c] >= 0) setTxRxSlaveEnable(tps[c], slave[2 * c], slave[2 * c + 1]);
    }
    std::vector<igd_ed137_ctl> ctl((size_t)F * C);
    std::vector<uint8_t> rtp12((size_t)F * C * 12), payload((size_t)F * C * 160);
    if (fread(ctl.data(), 8, ctl.size(), in) != ctl.size() || fread(rtp12.data(), 1, rtp12.size(), in) != rtp12.size() ||
        fread(payload.data(), 1, payload.size(), in) != payload.size()) return 2;
    fclose(in);

    Out o;
    o.pkts.assign((size_t)F * C * 180, 0);
    o.sizes.assign((size_t)F * C, 0);
    o.C = C;
    o.tps = &tps;
    std::vector<uint8_t> out_level((size_t)F * C, 0), in_level((size_t)F * C, 0);
    std::vector<uint32_t> rx_word((size_t)F * C, 0);
    Ev ev;
    ev.edge.assign((size_t)F * C, 0);
    ev.hang.assign((size_t)F * C, 0);
    ev.C = C;
    ev.tps = &tps;
    igd_bank_set_event_cb(bank, on_event, &ev);
    for (int f = 0; f < F; f++) {
        o.f = f;
        ev.f = f;
        for (int c = 0; c < C; c++) {
            const igd_ed137_ctl &k = ctl[(size_t)f * C + c];
            setAdapterPtt(tps[c], k.pttstatus, k.pttpriority, k.callRecorder);      // what RoIP_ED137 does on PTT
            setAdapterQslOn(tps[c], k.sqlstatus, 0, k.ed137_bssi);                  // ... and on squelch
            setAdapterPttId(tps[c], k.pttid);
            uint8_t pkt[172];
            memcpy(pkt, &rtp12[((size_t)f * C + c) * 12], 12);
            memcpy(pkt + 12, &payload[((size_t)f * C + c) * 160], 160);
            igd_submit_tx(tps[c], pkt, sizeof(pkt));                                // transport_send_rtp
        }
        if (igd_bank_flush_tx(bank, now0 + (long long)f * tick, (unsigned)flags, on_send, &o) < 0) return 5;
        for (int c = 0; c < C; c++) out_level[(size_t)f * C + c] = get_OutgoingRTP(tps[c]);
        // loop every emitted packet back into the same adapter's receive side (transport_rtp_cb)
        for (int c = 0; c < C; c++) {
            const uint32_t n = o.sizes[(size_t)f * C + c];
            if (n) igd_submit_rx(tps[c], &o.pkts[((size_t)f * C + c) * 180], n);
        }
        if (igd_bank_flush_rx(bank, now0 + (long long)f * tick, nullptr, nullptr) < 0) return 6;
        for (int c = 0; c < C; c++) {
            rx_word[(size_t)f * C + c] = get_ed137_value(tps[c]);
            in_level[(size_t)f * C + c] = get_IncomingRTP(tps[c]);
        }
        // the 40 ms timer of roip_ed137.cpp:482 -> detectR2SPacketAndReconn, every second 20 ms tick
        if ((f & 1) == 1 && igd_bank_r2s_watchdog(bank, now0 + (long long)f * tick, 200) < 0) return 7;
    }
    FILE *out = fopen(argv[2], "wb");
    fwrite(o.sizes.data(), 4, o.sizes.size(), out);
    fwrite(o.pkts.data(), 1, o.pkts.size(), out);
    fwrite(out_level.data(), 1, out_level.size(), out);
    fwrite(rx_word.data(), 4, rx_word.size(), out);
    fwrite(in_level.data(), 1, in_level.size(), out);
    fwrite(ev.edge.data(), 1, ev.edge.size(), out);
    fwrite(ev.hang.data(), 1, ev.hang.size(), out);
    fclose(out);
    // recorder sink: reference-exact file for the first channel's first 3 payloads
    WavWriter w;
    w.attach(bank, true, IGD_LAW_ULAW);
    w.start(std::string(argv[2]) + "_rec_", 8000);
    for (int f = 0; f < 3 && f < F; f++) w.wav_write(&payload[((size_t)f * C) * 160], 160);
    w.stop();
    printf("%s\n", w.fileName().c_str());
    igd_bank_close(bank);
    return 0;
}
