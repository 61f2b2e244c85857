/* Shim so that the reference's own ed137_rtp.h (which only needs three PJLIB
 * integer typedefs from <pjmedia/endpoint.h>) compiles without pjproject.
 * Used ONLY by oracle/Makefile -> oracle/_ref/ (test infrastructure). */
#ifndef IGD_REF_SHIM_PJMEDIA_ENDPOINT_H
#define IGD_REF_SHIM_PJMEDIA_ENDPOINT_H
#include <stdint.h>
typedef uint8_t pj_uint8_t;
typedef uint16_t pj_uint16_t;
typedef uint32_t pj_uint32_t;
#endif
