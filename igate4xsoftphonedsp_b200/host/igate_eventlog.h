// igate_eventlog.h -- the text either side of the meters: the supervisor messages the
// reference builds from its level bookkeeping, produced here from the GPU summaries.
//   * "PTTEventDataLogger" JSON of RoIP_ED137::createPTTEventDataLogger
//     (Functions.cpp:2148-2230; field order, spacing and the "radioUrl " key with its
//     trailing blank are the reference's), numbers formatted like QString::arg(double)
//     ('g', 6 significant digits) and QString::arg(int);
//   * the "broadcastVUMeter" message RoIP_ED137 consumes (roip_ed137.cpp:7686-7716:
//     keys in<N> / out<N> / in<N>dB / out<N>dB for softphone N = 1..4), built from
//     igd_meter_rec levels so that the GPU path can stand in for the external VU process.
// Host-side formatting only: no audio arithmetic lives here.
#pragma once
#include <string>

#include "../../include/igate_dsp.h"

// createPTTEventDataLogger's message, verbatim layout (Functions.cpp:2169-2181 / 2199-2211)
std::string igd_ptt_event_logger_json(int softPhoneID, const char *strEvent, double level_in_av,
                                      double level_in_max, double level_in_min, const char *radioUrl,
                                      int OutgoingRTPAv, int OutgoingRTPmax, int OutgoingRTPmin);
// "pptTest_released" message of one channel from the GPU event summary (igd_event_summary):
// level_in_av/max/min = 10*log10 of av/max/min (Functions.cpp:2196-2198), OutgoingRTPAv = sum/count
// (:2200), max / min as accumulated (:2141-2144)
std::string igd_ptt_released_json(int softPhoneID, const igd_summary_rec &rec, const igd_summary_db &db,
                                  const char *radioUrl);
// {"menuID":"broadcastVUMeter","in1":..,"out1":..,"in1dB":..,"out1dB":.., ... "out4dB":..}
// level = linear level, dB = its dB value, index i = softphone i+1 (roip_ed137.cpp:7688-7712)
std::string igd_vu_meter_json(const double in_level[4], const double out_level[4], const double in_db[4],
                              const double out_db[4]);
// the same message straight from the fused kernel's records of one tick: meter[i] = the leg softphone i+1
// listens to, bmeter[i] = its bridge; level = peak amplitude (0..32768), dB = 20*log10(peak/32768)
std::string igd_vu_meter_json_from_records(const igd_meter_rec meter[4], const igd_bridge_rec bmeter[4]);
