/*
 * roip_ed137.h (HARNESS, not the reference's header) -- declares a `class RoIP_ED137` with exactly
 * the members that the hot-path member functions of the reference name, so that those functions
 * compile UNMODIFIED, by line range, out of /root/reference/roip_ed137.cpp and
 * /root/reference/Functions.cpp (oracle/Makefile writes the extract to oracle/_ref/gen/, never into
 * the repo) next to the reference's own TransportAdapter.cpp:
 *
 *   roip_ed137.cpp:6500-6587  setOutgoingRTP / setIncomingED137Value / setIncomingRTP
 *   roip_ed137.cpp:5609-6348  checkEvents (SERVER squelch + best-signal select, CLIENT PTT priority)
 *   roip_ed137.cpp:5190-5234  setSlotVolume          roip_ed137.cpp:6869-6878  setvolumeSiteTone
 *   Functions.cpp:909-1179    setRadio*byCallID, setSlaveEnable, get_IPRadio*, get_R2SStatus
 *   Functions.cpp:1664-1705   setvolume              Functions.cpp:2126-2230   keeplogAudioLevel,
 *                                                                              createPTTEventDataLogger
 *
 * Everything ELSE those bodies call (display, recorder, repeater, websocket, PJSUA) is a recording
 * stub defined in oracle/ref_harness.cpp.  Member names and types follow the reference's class
 * (roip_ed137.h:246-247, 271-273, 289, 512, 647-800) because the extracted code names them.
 *
 * TEST INFRASTRUCTURE ONLY.
 */
#ifndef IGD_REF_HARNESS_ROIP_ED137_H
#define IGD_REF_HARNESS_ROIP_ED137_H

#include <cmath>
#include <cstdint>
#include <map>
#include <string>

#include "igd_qt_stub.h"
#include "igd_pj_stub.h"
#include "TransportAdapter.h"

using std::string;

#define TRXMODE_TRX "TRx"
#define TRXMODE_TX "Tx"
#define TRXMODE_RX "Rx"
#define CONNECTED "Connected "
#define DISCONNECTED "Disconnected"
#define RPTON "RPT"
#define PTTON "Tx"
#define RXON "Rx"
#define PTTRX "Tx&Rx"
#define STANDBY "--"
#define MUTEAUDIO 0
#define MUTEALL 1
#define MUTEIGNORE 2
#define UNMUTE false
#define MUTE true
#define SERVER 1
#define CLIENT 2
#define CLIENT_RPT 3
#define SERVER_RPT 4

static const float MIN_SLOT_VOLUME = 0.0;
static const float MAX_SLOT_VOLUME = 2.0;

/* the two PJSUA calls the extracted bodies make; defined in ref_harness.cpp */
struct pjsua_call_info {
    pjsua_conf_port_id conf_slot;
    pj_str_t remote_info;
};
extern "C" pj_status_t pjsua_call_get_info(pjsua_call_id call_id, pjsua_call_info *info);
extern "C" pj_status_t pjsua_conf_adjust_rx_level(pjsua_conf_port_id slot, float level);

class RoIP_ED137 {
public:
    static RoIP_ED137 *instance();

    /* ---- bodies taken from the reference by line range (oracle/_ref/gen/ref_extract.cpp) ---- */
    void setIncomingRTP(tp_adapter *adapter);
    void setOutgoingRTP(tp_adapter *adapter);
    void setIncomingED137Value(uint32_t ed137_value, pjsua_acc_id acc_id);
    void checkEvents();
    bool setSlotVolume(int callId, bool increase, bool current);
    void setvolumeSiteTone(pjsua_call_id call_id);
    void setvolume(pjsua_call_id call_id, bool mute);
    void setRadioPttbyCallID(bool pttval, int callId, int priority, int userRec);
    void setAdaptercallRecorder(int callId, int val);
    void setSlaveEnable(pjsua_call_id callId, pj_bool_t rx, pj_bool_t tx);
    void setRadioSqlOnbyCallID(bool sqlval, int callId, int priority);
    void setRadioSqlOnbyCallID(bool sqlval, int callId, int priority, uint16_t rssi);
    int get_IPRadioBss(pjsua_call_id callId);
    void sendRtpType232(pjsua_call_id callId);
    qint64 get_R2SStatus(pjsua_call_id callId);
    int get_IPRadioPttStatus(pjsua_call_id callId);
    int get_IPRadioPttId(pjsua_call_id callId);
    int get_IPRadioSquelch(pjsua_call_id callId);
    bool get_IPRadioStatus(pjsua_call_id callId);

    struct trx;
    void keeplogAudioLevel(trx *radio);
    void createPTTEventDataLogger(trx *radio, QString strEvent);

    /* ---- recording stubs (ref_harness.cpp) ---- */
    void scan_call_err();
    void updateHomeDisplay(QString trxid, QString connState, QString trxState, pjsua_call_id call_id, int radioNodeID);
    QString getTimeDuratio(pjsua_call_id callId);
    void recorder_pressed(int pttLevel, QString txon);
    void recorder_released(int pttLevel);
    void repeat_pressed(int ignorID, int pttLevel);
    void repeat_released(int pttLevel);
    void sqlTest_pressed(int call_id, int level);
    void sendTextMessage(QString message);
    void cppCommand(QString message);
    bool getExtFromRemoteInfo(pj_str_t const &remoteInfo, std::string &extNumber);

    /* ---- state named by the extracted bodies ---- */
    double audioInLevel = 0;
    float SLOT_VOLUME = 2.0f;
    int sqlStatusCount = 0;
    bool sqlStatusOn = false;

    struct trx {
        bool pptTestPressed = false;
        bool sqlTestPressed = false;
        bool callState = false;
        int radioNodeID = 0;
        QString url = "";
        QString trxmode = "";
        QString callName = "";
        int sec_connDuration = 0;
        pjsua_call_id call_id = PJSUA_INVALID_ID;
        QString callIndexName;
        int pttLevel = 0;
        bool m_PttPressed = false;
        int rssi = 0;
        qint64 lastRxmsec = 0;
        qint64 lastTxmsec = 0;
        int lastRx = 0;
        int lastTx = 0;
        bool mainRadioReceiverUsed = false;
        bool mainRadioTransceiverUsed = false;
        uint32_t ed137_val_old = 0;
        bool audioSQLOn = false;
        bool SQLOn = false;
        int SqlGroupDelayCount = 0;

        double level_in = 0;
        int level_in_count = 0;
        double level_in_av = 0;
        double level_in_max = 0;
        double level_in_min = 0;
        bool eventPttSQL_In_LoggingOn = false;

        uint8_t OutgoingRTP = 0;
        uint16_t OutgoingRTPSum = 0;
        uint8_t OutgoingRTPav = 0;
        uint8_t OutgoingRTPmax = 0;
        uint8_t OutgoingRTPmin = 255;
        uint8_t IncomingRTP = 0;
    };
    struct channel {
        trx *radio1;
        trx *radio2;
        QString trxmode = TRXMODE_TRX;
        QString trxStatus = STANDBY;
        bool mainRx = true;
    };

    float sidetone = 0.1f;
    channel *trx1;
    channel *trx2;
    bool rxBestSignalEnable = true;
    QList<trx *> trx_incall;
    QString lastPttOn;
    int ptt_level = 0;
    int inviteMode = SERVER;
    std::map<int, pjmedia_transport *> transport_map;

    bool pttGroupActive = false;
    bool pttGroupActiveTmp = false;
    uint8_t groupMute = MUTEAUDIO;
    bool forceMuteSqlOn = false;
    int SqlGroupDelay = 5;
    int m_PttPressed1 = 0, m_PttPressed2 = 0, m_PttPressed3 = 0, m_PttPressed4 = 0;
    bool m_PttPressed = false;
    bool localSidetoneLoopbackOn = false;
    bool onLocalSidetoneLoopbackChanged = false;
    int m_softPhoneID = 1;
    bool m_radioAutoInactive = false;
    int m_radioMainStandby = 0;
    QString recorderAlloweURI = "recorder";
    bool pttOnStatus = false;
    bool sqlOnStatus = false;
    int delayCount = 0;
    QString trxStatus = STANDBY;
    bool recorderConnected = false;
    bool connToRadio = true;
    bool pttInput = false;

    /* ---- what the stubs recorded (read by ref_harness.cpp's extern "C" API) ---- */
    int igd_checkEvents_calls = 0;
    std::string igd_last_text_message;
    int igd_text_messages = 0;

    RoIP_ED137();
};

#endif
