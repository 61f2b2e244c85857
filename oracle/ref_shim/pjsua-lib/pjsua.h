/* <pjsua-lib/pjsua.h> for the reference-backed oracle build: see igd_pj_stub.h (test infrastructure only). */
#ifndef IGD_REF_SHIM_PJSUA_LIB_PJSUA_H
#define IGD_REF_SHIM_PJSUA_LIB_PJSUA_H
#include "../igd_pj_stub.h"
#endif
