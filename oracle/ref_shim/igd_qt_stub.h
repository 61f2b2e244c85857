/*
 * igd_qt_stub.h -- the few QtCore types the reference's hot-path sources use (QString with
 * contains / == / number / arg, QDateTime::currentMSecsSinceEpoch, qDebug, QList::length/at),
 * restated from the published Qt 5 API so that the REFERENCE'S OWN TransportAdapter.cpp (and the
 * line-range extracts of roip_ed137.cpp / Functions.cpp that oracle/Makefile generates into
 * oracle/_ref/) compile without Qt.
 *
 * TEST INFRASTRUCTURE ONLY (see igd_pj_stub.h).  The clock is settable so that the reference's
 * keep-alive throttle and r2sPacket stamps are deterministic: igd_ref_clock_ms is what
 * QDateTime::currentMSecsSinceEpoch() returns.
 */
#ifndef IGD_QT_STUB_H
#define IGD_QT_STUB_H

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

typedef int8_t qint8;
typedef uint8_t quint8;
typedef int16_t qint16;
typedef uint16_t quint16;
typedef int32_t qint32;
typedef uint32_t quint32;
typedef long long qint64;
typedef unsigned long long quint64;

extern "C" long long igd_ref_clock_ms;

class QString {
    std::string s_;

public:
    QString() {}
    QString(const char *c) : s_(c ? c : "") {}
    QString(const std::string &s) : s_(s) {}
    static QString fromStdString(const std::string &s) { return QString(s); }
    std::string toStdString() const { return s_; }
    const std::string &igd_str() const { return s_; }
    int length() const { return (int)s_.size(); }
    bool isEmpty() const { return s_.empty(); }
    /* QString::contains(const QString&, Qt::CaseSensitive): plain substring search */
    bool contains(const QString &o) const { return s_.find(o.s_) != std::string::npos; }
    bool operator==(const QString &o) const { return s_ == o.s_; }
    bool operator!=(const QString &o) const { return s_ != o.s_; }
    bool operator==(const char *c) const { return s_ == (c ? c : ""); }
    bool operator!=(const char *c) const { return s_ != (c ? c : ""); }
    QString &operator=(const char *c) { s_ = c ? c : ""; return *this; }

    static QString number(long long v, int base = 10)
    {
        char b[72];
        if (base == 16) snprintf(b, sizeof b, "%llx", v);
        else snprintf(b, sizeof b, "%lld", v);
        return QString(b);
    }
    static QString number(int v, int base = 10) { return number((long long)v, base); }
    static QString number(unsigned v, int base = 10) { return number((long long)v, base); }
    /* QString::number(double, 'g', 6) */
    static QString number(double v)
    {
        char b[64];
        snprintf(b, sizeof b, "%g", v);
        return QString(b);
    }

    /* QString::arg: every occurrence of the LOWEST numbered place marker %1..%99 is replaced */
    QString arg(const QString &a) const
    {
        int low = 100;
        for (size_t i = 0; i + 1 < s_.size(); ++i)
            if (s_[i] == '%' && s_[i + 1] >= '0' && s_[i + 1] <= '9') {
                int n = s_[i + 1] - '0';
                if (i + 2 < s_.size() && s_[i + 2] >= '0' && s_[i + 2] <= '9') n = n * 10 + (s_[i + 2] - '0');
                if (n > 0 && n < low) low = n;
            }
        if (low == 100) return *this; /* Qt warns "Argument missing" and returns the string */
        std::string out;
        for (size_t i = 0; i < s_.size();) {
            if (s_[i] == '%' && i + 1 < s_.size() && s_[i + 1] >= '0' && s_[i + 1] <= '9') {
                int n = s_[i + 1] - '0';
                size_t len = 2;
                if (i + 2 < s_.size() && s_[i + 2] >= '0' && s_[i + 2] <= '9') { n = n * 10 + (s_[i + 2] - '0'); len = 3; }
                if (n == low) { out += a.s_; i += len; continue; }
            }
            out += s_[i++];
        }
        return QString(out);
    }
    QString arg(const char *a) const { return arg(QString(a)); }
    QString arg(int a) const { return arg(number(a)); }
    QString arg(unsigned a) const { return arg(number(a)); }
    QString arg(long long a) const { return arg(number(a)); }
    QString arg(double a) const { return arg(number(a)); } /* arg(double, 0, 'g', -1) */
};

class QDateTime {
public:
    static qint64 currentMSecsSinceEpoch() { return igd_ref_clock_ms; }
};

/* qDebug() << ... : swallowed */
struct IgdQDebugSink {
    template <class T> IgdQDebugSink &operator<<(const T &) { return *this; }
};
inline IgdQDebugSink qDebug() { return IgdQDebugSink(); }

template <class T> class QList {
    std::vector<T> v_;

public:
    int length() const { return (int)v_.size(); }
    int size() const { return (int)v_.size(); }
    const T &at(int i) const { return v_[(size_t)i]; }
    void append(const T &t) { v_.push_back(t); }
    void clear() { v_.clear(); }
};

#define emit

/* the reference prints from its setters ("Set Radio Sql: %d", TransportAdapter.cpp:189,210); muted */
#ifdef IGD_REF_QUIET
#define printf(...) ((void)0)
#endif
#define Q_UNUSED(x) (void)(x)

#endif
