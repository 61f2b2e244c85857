"""igate4xsoftphonedsp_b200 -- B200 (sm_100a) implementation of the per-frame
voice path of the iGate4x ED-137 RoIP softphone/gateway: G.711 A-/u-law,
per-frame level meters, PTT-gated saturating mix and the ED-137 RTP header
extension, batched over channels x 20 ms frames behind the C ABI of
include/igate_dsp.h (libigate_dsp.so).  No CPU fallback exists."""
from ._native import (CT_IDLE, CT_RXONLY, CT_TXISH, CTL_DT, BRIDGE_DT, EDF_ACTIVE, EDF_DROPPED,
                      EDF_MAIN_RX, EDF_MAIN_TX, EDF_RRC, F_REF_QUIRKS, F_SIGNED_CHAR, FIELDS_DT, FRAME, GAIN_NO_AUDIO,
                      LAW_ALAW, LAW_ULAW, LIB_PATH, METER_DT, PKT_HDR, PKT_MAX, STATE_DT, SUMMARY_DB_DT,
                      SUMMARY_DT, NativeLibraryMissing, load)
from .voicepath import IgdError, VoicePath, calltype_flags, gain_q7, make_state

__all__ = ["VoicePath", "IgdError", "NativeLibraryMissing", "load", "gain_q7", "calltype_flags", "make_state",
           "FRAME", "PKT_HDR", "PKT_MAX", "LAW_ALAW", "LAW_ULAW", "F_SIGNED_CHAR", "F_REF_QUIRKS",
           "METER_DT", "BRIDGE_DT", "SUMMARY_DT", "SUMMARY_DB_DT", "FIELDS_DT", "STATE_DT", "CTL_DT",
           "CT_IDLE", "CT_RXONLY", "CT_TXISH", "EDF_ACTIVE", "EDF_RRC", "EDF_MAIN_TX", "EDF_MAIN_RX",
           "EDF_DROPPED", "LIB_PATH", "GAIN_NO_AUDIO"]
