/* <pjmedia.h> for the reference-backed oracle build: see igd_pj_stub.h (test infrastructure only). */
#ifndef IGD_REF_SHIM_PJMEDIA_H
#define IGD_REF_SHIM_PJMEDIA_H
#include "igd_pj_stub.h"
#endif
