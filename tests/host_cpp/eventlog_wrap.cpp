// C wrappers around host/igate_eventlog.h for the ctypes test. Test-only.
#include <cstring>

#include "../../igate4xsoftphonedsp_b200/host/igate_eventlog.h"

static size_t put(char *out, size_t cap, const std::string &s)
{
    const size_t n = s.size() < cap ? s.size() : cap;
    memcpy(out, s.data(), n);
    return n;
}
extern "C" {
size_t w_ptt_event(char *out, size_t cap, int id, const char *ev, double a, double b, double c, const char *url,
                   int r0, int r1, int r2)
{
    return put(out, cap, igd_ptt_event_logger_json(id, ev, a, b, c, url, r0, r1, r2));
}
size_t w_ptt_released(char *out, size_t cap, int id, const igd_summary_rec *rec, const igd_summary_db *db, const char *url)
{
    return put(out, cap, igd_ptt_released_json(id, *rec, *db, url));
}
size_t w_vu(char *out, size_t cap, const double *a, const double *b, const double *c, const double *d)
{
    return put(out, cap, igd_vu_meter_json(a, b, c, d));
}
size_t w_vu_rec(char *out, size_t cap, const igd_meter_rec *m, const igd_bridge_rec *b)
{
    return put(out, cap, igd_vu_meter_json_from_records(m, b));
}
}
