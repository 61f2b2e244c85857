// igate_eventlog.cpp -- see igate_eventlog.h.  Formatting only.
#include "igate_eventlog.h"

#include <math.h>
#include <stdio.h>

namespace {
// QString::arg(double) = QString::number(a, 'g', 6); Qt spells the specials "inf" / "nan"
std::string qnum(double v)
{
    if (isnan(v)) return "nan";
    if (isinf(v)) return v < 0 ? "-inf" : "inf";
    char b[64];
    snprintf(b, sizeof(b), "%g", v);
    return b;
}
}  // namespace

std::string igd_ptt_event_logger_json(int softPhoneID, const char *strEvent, double av, double mx, double mn,
                                      const char *radioUrl, int rtpAv, int rtpMax, int rtpMin)
{
    std::string m = "{";
    m += "\"menuID\"                       :\"PTTEventDataLogger\", ";
    m += "\"softPhoneID\"                  :" + std::to_string(softPhoneID) + ", ";
    m += "\"Ptt\"                          :\"" + std::string(strEvent ? strEvent : "") + "\", ";
    m += "\"level_in_av\"                  :" + qnum(av) + ", ";
    m += "\"level_in_max\"                 :" + qnum(mx) + ", ";
    m += "\"level_in_min\"                 :" + qnum(mn) + ", ";
    m += "\"radioUrl \"                    :\"" + std::string(radioUrl ? radioUrl : "") + "\",";
    m += "\"OutgoingRTPAv\"                :" + std::to_string(rtpAv) + ", ";
    m += "\"OutgoingRTPmax\"               :" + std::to_string(rtpMax) + ", ";
    m += "\"OutgoingRTPmin\"               :" + std::to_string(rtpMin) + " ";
    m += "}";
    return m;
}

std::string igd_ptt_released_json(int softPhoneID, const igd_summary_rec &rec, const igd_summary_db &db,
                                  const char *radioUrl)
{
    return igd_ptt_event_logger_json(softPhoneID, "pptTest_released", db.level_av_db, db.level_max_db,
                                     db.level_min_db, radioUrl, (int)(db.bm_av & 0xFFu), rec.bm_max, rec.bm_min);
}

std::string igd_vu_meter_json(const double in_level[4], const double out_level[4], const double in_db[4],
                              const double out_db[4])
{
    std::string m = "{\"menuID\":\"broadcastVUMeter\"";
    for (int i = 0; i < 4; i++) {
        const std::string n = std::to_string(i + 1);
        m += ",\"in" + n + "\":" + qnum(in_level[i]) + ",\"out" + n + "\":" + qnum(out_level[i]);
        m += ",\"in" + n + "dB\":" + qnum(in_db[i]) + ",\"out" + n + "dB\":" + qnum(out_db[i]);
    }
    m += "}";
    return m;
}

std::string igd_vu_meter_json_from_records(const igd_meter_rec meter[4], const igd_bridge_rec bmeter[4])
{
    double in[4], out[4], indb[4], outdb[4];
    for (int i = 0; i < 4; i++) {
        in[i] = (double)IGD_METER_PEAK(meter[i]);
        out[i] = (double)bmeter[i].mix_peak;
        indb[i] = in[i] > 0 ? 20.0 * log10(in[i] / 32768.0) : -INFINITY;
        outdb[i] = out[i] > 0 ? 20.0 * log10(out[i] / 32768.0) : -INFINITY;
    }
    return igd_vu_meter_json(in, out, indb, outdb);
}

