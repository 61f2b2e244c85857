/* <pjmedia/endpoint.h> for the reference-backed oracle build: the reference's own ed137_rtp.h
 * needs the three PJLIB integer typedefs, TransportAdapter.h the rest -- see igd_pj_stub.h.
 * Used ONLY by oracle/Makefile -> oracle/_ref/ and the vtable test (test infrastructure). */
#ifndef IGD_REF_SHIM_PJMEDIA_ENDPOINT_H
#define IGD_REF_SHIM_PJMEDIA_ENDPOINT_H
#include "../igd_pj_stub.h"
#endif
