# igate_dsp.pri -- include from iGate4xSoftphoneDSP.pro (next to its LIBS block, :64-92):
#     include(/opt/igate_dsp/host/igate_dsp.pri)
# Links the B200 voice path (libigate_dsp.so, CUDA runtime linked statically: the host build needs no
# CUDA toolkit) and compiles the two host files with the project's own g++.
IGD_ROOT = /opt/igate_dsp                       # include/ + lib/ + host/
INCLUDEPATH += $$IGD_ROOT/include $$IGD_ROOT/host
LIBS        += -L$$IGD_ROOT/lib -ligate_dsp -Wl,-rpath,$$IGD_ROOT/lib
HEADERS     += $$IGD_ROOT/host/igate_shim.h $$IGD_ROOT/host/igate_eventlog.h
SOURCES     += $$IGD_ROOT/host/igate_shim.cpp $$IGD_ROOT/host/igate_eventlog.cpp
SOURCES     -= TransportAdapter.cpp             # its signatures are provided by igate_shim.cpp
DEFINES     += IGATE_DSP_GPU=1
