// igate_pj_compat.h -- the PJSIP types the shim's plugin boundary is made of.
//
// Built inside the reference's qmake project define IGD_HAVE_PJSIP: the real
// <pjsua.h> / <pjmedia/transport.h> are used and the shim's adapter is an ordinary
// `pjmedia_transport` that PJSIP drives through `tp->op` (TransportAdapter.cpp:59-73, 95-107).
//
// Built stand-alone (no pjproject on the box) the same ABI is declared here: the 12-entry
// `pjmedia_transport_op` of pjproject <= 2.7 (get_info, attach, detach, send_rtp, send_rtcp,
// send_rtcp2, media_create, encode_sdp, media_start, media_stop, simulate_lost, destroy -- the table
// TransportAdapter.cpp:59-73 fills; "pjsip 2.6 only", :245) and `struct pjmedia_transport`
// {name[32], type, op, user_data}.  Pool / SDP / endpoint types stay opaque: without pjproject the
// shim cannot add SDP attributes (encode_sdp then only passes through).
#pragma once
#include <stddef.h>
#include <stdint.h>

#ifdef IGD_HAVE_PJSIP
#include <pjsua.h>
#include <pjmedia/transport.h>
#include <pjmedia/endpoint.h>
#else
extern "C" {
typedef int pj_status_t;
typedef int pj_bool_t;
typedef uint8_t pj_uint8_t;
typedef uint16_t pj_uint16_t;
typedef uint32_t pj_uint32_t;
typedef size_t pj_size_t;
typedef long pj_ssize_t;
typedef void pj_sockaddr_t;
typedef int pjsua_call_id;
#define PJ_SUCCESS 0
#define PJ_EINVAL 70004
#define PJ_MAX_OBJ_NAME 32
typedef struct pj_pool_t pj_pool_t;
typedef struct pjmedia_endpt pjmedia_endpt;
typedef struct pjmedia_sdp_session pjmedia_sdp_session;
typedef struct pjmedia_transport_info pjmedia_transport_info;
typedef enum pjmedia_dir { PJMEDIA_DIR_NONE = 0, PJMEDIA_DIR_ENCODING = 1, PJMEDIA_DIR_DECODING = 2,
                           PJMEDIA_DIR_ENCODING_DECODING = 3 } pjmedia_dir;
typedef enum pjmedia_transport_type { PJMEDIA_TRANSPORT_TYPE_UDP, PJMEDIA_TRANSPORT_TYPE_ICE,
                                      PJMEDIA_TRANSPORT_TYPE_SRTP, PJMEDIA_TRANSPORT_TYPE_USER } pjmedia_transport_type;
typedef struct pjmedia_transport pjmedia_transport;
typedef struct pjmedia_transport_op {
    pj_status_t (*get_info)(pjmedia_transport *tp, pjmedia_transport_info *info);
    pj_status_t (*attach)(pjmedia_transport *tp, void *user_data, const pj_sockaddr_t *rem_addr,
                          const pj_sockaddr_t *rem_rtcp, unsigned addr_len,
                          void (*rtp_cb)(void *user_data, void *pkt, pj_ssize_t size),
                          void (*rtcp_cb)(void *user_data, void *pkt, pj_ssize_t size));
    void (*detach)(pjmedia_transport *tp, void *user_data);
    pj_status_t (*send_rtp)(pjmedia_transport *tp, const void *pkt, pj_size_t size);
    pj_status_t (*send_rtcp)(pjmedia_transport *tp, const void *pkt, pj_size_t size);
    pj_status_t (*send_rtcp2)(pjmedia_transport *tp, const pj_sockaddr_t *addr, unsigned addr_len,
                              const void *pkt, pj_size_t size);
    pj_status_t (*media_create)(pjmedia_transport *tp, pj_pool_t *sdp_pool, unsigned options,
                                const pjmedia_sdp_session *remote_sdp, unsigned media_index);
    pj_status_t (*encode_sdp)(pjmedia_transport *tp, pj_pool_t *sdp_pool, pjmedia_sdp_session *sdp_local,
                              const pjmedia_sdp_session *rem_sdp, unsigned media_index);
    pj_status_t (*media_start)(pjmedia_transport *tp, pj_pool_t *tmp_pool, const pjmedia_sdp_session *sdp_local,
                               const pjmedia_sdp_session *sdp_remote, unsigned media_index);
    pj_status_t (*media_stop)(pjmedia_transport *tp);
    pj_status_t (*simulate_lost)(pjmedia_transport *tp, pjmedia_dir dir, unsigned pct_lost);
    pj_status_t (*destroy)(pjmedia_transport *tp);
} pjmedia_transport_op;
struct pjmedia_transport {
    char name[PJ_MAX_OBJ_NAME];
    pjmedia_transport_type type;
    pjmedia_transport_op *op;
    void *user_data;
};
}
#endif
