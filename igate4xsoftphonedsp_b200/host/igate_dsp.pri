# igate_dsp.pri -- include from iGate4xSoftphoneDSP.pro (next to :64-92); see INTEGRATION.md section 1
IGD_ROOT = /opt/igate_dsp                       # include/ + lib/ + host/
INCLUDEPATH += $$IGD_ROOT/include $$IGD_ROOT/host
LIBS        += -L$$IGD_ROOT/lib -ligate_dsp -Wl,-rpath,$$IGD_ROOT/lib
HEADERS     += $$IGD_ROOT/host/igate_shim.h $$IGD_ROOT/host/igate_pj_compat.h $$IGD_ROOT/host/igate_eventlog.h
SOURCES     += $$IGD_ROOT/host/igate_shim.cpp $$IGD_ROOT/host/igate_eventlog.cpp
SOURCES     -= TransportAdapter.cpp             # (:20-36) replaced as a whole by igate_shim.cpp
HEADERS     -= TransportAdapter.h               # its declarations come from igate_shim.h
DEFINES     += IGATE_DSP_GPU=1 IGD_HAVE_PJSIP=1 # igate_shim.cpp then includes the project's real <pjsua.h>
