/* <pjmedia/transport.h> for the reference-backed oracle build: see igd_pj_stub.h (test infrastructure only). */
#ifndef IGD_REF_SHIM_PJMEDIA_TRANSPORT_H
#define IGD_REF_SHIM_PJMEDIA_TRANSPORT_H
#include "../igd_pj_stub.h"
#endif
