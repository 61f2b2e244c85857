"""Supervisor messages either side of the meters, from the GPU summaries (formatting only).

`ptt_event_logger_json` is the "PTTEventDataLogger" message of
RoIP_ED137::createPTTEventDataLogger (Functions.cpp:2148-2230) with the reference's field
order and spacing; numbers are formatted like QString::arg(double) ('g', 6 digits) and
QString::arg(int).  `vu_meter_json` is the "broadcastVUMeter" message RoIP_ED137 consumes
(roip_ed137.cpp:7686-7716).  The C++ twin is host/igate_eventlog.{h,cpp}.
"""
import math


def _qnum(v):
    v = float(v)
    if math.isnan(v):
        return "nan"
    if math.isinf(v):
        return "-inf" if v < 0 else "inf"
    return "%g" % v


def ptt_event_logger_json(softphone_id, event, level_in_av, level_in_max, level_in_min, radio_url,
                          rtp_av, rtp_max, rtp_min):
    return ("{"
            "\"menuID\"                       :\"PTTEventDataLogger\", "
            f"\"softPhoneID\"                  :{int(softphone_id)}, "
            f"\"Ptt\"                          :\"{event}\", "
            f"\"level_in_av\"                  :{_qnum(level_in_av)}, "
            f"\"level_in_max\"                 :{_qnum(level_in_max)}, "
            f"\"level_in_min\"                 :{_qnum(level_in_min)}, "
            f"\"radioUrl \"                    :\"{radio_url}\","
            f"\"OutgoingRTPAv\"                :{int(rtp_av)}, "
            f"\"OutgoingRTPmax\"               :{int(rtp_max)}, "
            f"\"OutgoingRTPmin\"               :{int(rtp_min)} "
            "}")


def ptt_released_json(softphone_id, rec, db, radio_url):
    """rec / db: one element of VoicePath.event_summary()'s two arrays."""
    return ptt_event_logger_json(softphone_id, "pptTest_released", db["level_av_db"], db["level_max_db"],
                                 db["level_min_db"], radio_url, int(db["bm_av"]) & 0xFF, int(rec["bm_max"]),
                                 int(rec["bm_min"]))


def vu_meter_json(in_level, out_level, in_db, out_db):
    parts = ["\"menuID\":\"broadcastVUMeter\""]
    for i in range(4):
        n = i + 1
        parts.append(f"\"in{n}\":{_qnum(in_level[i])},\"out{n}\":{_qnum(out_level[i])}")
        parts.append(f"\"in{n}dB\":{_qnum(in_db[i])},\"out{n}dB\":{_qnum(out_db[i])}")
    return "{" + ",".join(parts) + "}"


def vu_meter_json_from_records(meter, bmeter):
    """The VU feed of one tick for the four softphones of a box, from the fused kernel's records:
    meter = the igd_meter_rec of the leg each softphone listens to (4 records), bmeter = the
    igd_bridge_rec of each softphone's bridge (4 records).  in<N> / out<N> = peak amplitude (0..32768) of the
    received leg / of the mix, in<N>dB / out<N>dB = the same in dBFS (20*log10(peak/32768), -inf for silence)."""
    def db(p):
        return 20.0 * math.log10(p / 32768.0) if p > 0 else float("-inf")
    pin = [int(m["hi"]) >> 16 for m in meter]
    pout = [int(b["mix_peak"]) for b in bmeter]
    return vu_meter_json(pin, pout, [db(p) for p in pin], [db(p) for p in pout])

