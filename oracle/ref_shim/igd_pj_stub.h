/*
 * igd_pj_stub.h -- the slice of the PJLIB / PJMEDIA / PJSUA API that the reference's
 * TransportAdapter.{h,cpp} touches, restated from the published pjproject 2.x API so that the
 * REFERENCE'S OWN TransportAdapter.cpp compiles where it lies (/root/reference) without
 * pjproject (not vendored, not installable here).
 *
 * TEST INFRASTRUCTURE ONLY: reached through `-I oracle/ref_shim` by oracle/Makefile (-> oracle/_ref/)
 * and by the vtable test of the product shim (tests/host_cpp/vtable_driver.cpp).  The product
 * (igate4xsoftphonedsp_b200/, include/) never includes it: a real build uses the real pjproject
 * headers, whose `pjmedia_transport` / `pjmedia_transport_op` these declarations mirror
 * (12 entries, the pre-`attach2` table of pjproject <= 2.7 that TransportAdapter.cpp:59-73 fills).
 */
#ifndef IGD_PJ_STUB_H
#define IGD_PJ_STUB_H

#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <arpa/inet.h>

#ifdef __cplusplus
#define PJ_BEGIN_DECL extern "C" {
#define PJ_END_DECL }
#else
#define PJ_BEGIN_DECL
#define PJ_END_DECL
#endif
#define PJ_DEF(type) type
#define PJ_DECL(type) extern type
#define PJ_INLINE(type) static inline type
#define PJ_UNUSED_ARG(a) (void)(a)
#define PJ_SUCCESS 0
#define PJ_TRUE 1
#define PJ_FALSE 0
#define PJ_MAX_OBJ_NAME 32
#define PJ_EINVAL 70004
#define PJMEDIA_RTP_EINVER 220121 /* pjmedia/errno.h */
#define PJMEDIA_RTP_EINLEN 220123
#define PJSUA_INVALID_ID (-1)

PJ_BEGIN_DECL

typedef int pj_status_t;
typedef int pj_bool_t;
typedef size_t pj_size_t;
typedef long pj_ssize_t;
typedef uint8_t pj_uint8_t;
typedef uint16_t pj_uint16_t;
typedef uint32_t pj_uint32_t;
typedef int16_t pj_int16_t;
typedef int32_t pj_int32_t;
typedef void pj_sockaddr_t;
typedef int pjsua_call_id;
typedef int pjsua_acc_id;
typedef int pjsua_conf_port_id;

typedef struct pj_str_t { char *ptr; pj_ssize_t slen; } pj_str_t;

/* pj/pool.h -- a pool whose blocks are freed together by pj_pool_release() */
typedef struct igd_pool_block { struct igd_pool_block *next; } igd_pool_block;
typedef struct pj_pool_t {
    char obj_name[PJ_MAX_OBJ_NAME];
    igd_pool_block *blocks;
} pj_pool_t;
void *pj_pool_alloc(pj_pool_t *pool, pj_size_t size);
void *pj_pool_zalloc(pj_pool_t *pool, pj_size_t size);
void pj_pool_release(pj_pool_t *pool);
#define PJ_POOL_ALLOC_T(pool, type) ((type *)pj_pool_alloc(pool, sizeof(type)))
#define PJ_POOL_ZALLOC_T(pool, type) ((type *)pj_pool_zalloc(pool, sizeof(type)))
pj_str_t *pj_strdup2(pj_pool_t *pool, pj_str_t *dst, const char *src);

#define pj_assert(e) ((void)0) /* release build: PJ_ASSERT is compiled out with NDEBUG */
#define pj_memcpy memcpy
#define pj_memset memset
#define pj_ansi_strncpy strncpy
#define pj_ntohs ntohs
#define pj_htons htons
#define pj_ntohl ntohl
#define pj_htonl htonl

/* pjmedia/endpoint.h */
typedef struct pjmedia_endpt pjmedia_endpt;
pj_pool_t *pjmedia_endpt_create_pool(pjmedia_endpt *endpt, const char *name, pj_size_t initial,
                                     pj_size_t increment);

/* pjmedia/sdp.h */
#define PJMEDIA_MAX_SDP_ATTR 68
#define PJMEDIA_MAX_SDP_MEDIA 16
typedef struct pjmedia_sdp_attr { pj_str_t name; pj_str_t value; } pjmedia_sdp_attr;
typedef struct pjmedia_sdp_media {
    unsigned attr_count;
    pjmedia_sdp_attr *attr[PJMEDIA_MAX_SDP_ATTR];
} pjmedia_sdp_media;
typedef struct pjmedia_sdp_session {
    unsigned media_count;
    pjmedia_sdp_media *media[PJMEDIA_MAX_SDP_MEDIA];
} pjmedia_sdp_session;
pj_status_t pjmedia_sdp_attr_add(unsigned *count, pjmedia_sdp_attr *attr_array[], pjmedia_sdp_attr *attr);

/* pjmedia/rtp.h: the 12-byte RTP fixed header (little-endian bit-field order) */
#pragma pack(1)
typedef struct pjmedia_rtp_hdr {
    pj_uint16_t cc : 4;
    pj_uint16_t x : 1;
    pj_uint16_t p : 1;
    pj_uint16_t v : 2;
    pj_uint16_t pt : 7;
    pj_uint16_t m : 1;
    pj_uint16_t seq;
    pj_uint32_t ts;
    pj_uint32_t ssrc;
} pjmedia_rtp_hdr;
#pragma pack()
typedef struct pjmedia_rtp_session pjmedia_rtp_session;
/* RFC 3550 s5.1 / pjmedia rtp.c: version check, CSRC list, header extension, padding */
pj_status_t pjmedia_rtp_decode_rtp(pjmedia_rtp_session *ses, const void *pkt, int pkt_len,
                                   const pjmedia_rtp_hdr **hdr, const void **payload, unsigned *payloadlen);

/* pjmedia/transport.h (pjproject <= 2.7) */
typedef enum pjmedia_dir { PJMEDIA_DIR_NONE = 0, PJMEDIA_DIR_ENCODING = 1, PJMEDIA_DIR_DECODING = 2,
                           PJMEDIA_DIR_ENCODING_DECODING = 3 } pjmedia_dir;
typedef enum pjmedia_transport_type { PJMEDIA_TRANSPORT_TYPE_UDP, PJMEDIA_TRANSPORT_TYPE_ICE,
                                      PJMEDIA_TRANSPORT_TYPE_SRTP, PJMEDIA_TRANSPORT_TYPE_USER } pjmedia_transport_type;
typedef struct pjmedia_transport pjmedia_transport;
typedef struct pjmedia_transport_info { int igd_stub_calls; } pjmedia_transport_info;

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

/* the inline dispatchers of pjmedia/transport.h: what PJSIP (and the adapter, towards its slave
 * transport) calls -- always through tp->op */
PJ_INLINE(pj_status_t) pjmedia_transport_get_info(pjmedia_transport *tp, pjmedia_transport_info *info)
{ return tp && tp->op && tp->op->get_info ? (*tp->op->get_info)(tp, info) : PJ_EINVAL; }
PJ_INLINE(pj_status_t) pjmedia_transport_attach(pjmedia_transport *tp, void *user_data, const pj_sockaddr_t *rem_addr,
                                                const pj_sockaddr_t *rem_rtcp, unsigned addr_len,
                                                void (*rtp_cb)(void *, void *, pj_ssize_t),
                                                void (*rtcp_cb)(void *, void *, pj_ssize_t))
{ return (*tp->op->attach)(tp, user_data, rem_addr, rem_rtcp, addr_len, rtp_cb, rtcp_cb); }
PJ_INLINE(void) pjmedia_transport_detach(pjmedia_transport *tp, void *user_data) { (*tp->op->detach)(tp, user_data); }
PJ_INLINE(pj_status_t) pjmedia_transport_send_rtp(pjmedia_transport *tp, const void *pkt, pj_size_t size)
{ return (*tp->op->send_rtp)(tp, pkt, size); }
PJ_INLINE(pj_status_t) pjmedia_transport_send_rtcp(pjmedia_transport *tp, const void *pkt, pj_size_t size)
{ return (*tp->op->send_rtcp)(tp, pkt, size); }
PJ_INLINE(pj_status_t) pjmedia_transport_send_rtcp2(pjmedia_transport *tp, const pj_sockaddr_t *addr,
                                                    unsigned addr_len, const void *pkt, pj_size_t size)
{ return (*tp->op->send_rtcp2)(tp, addr, addr_len, pkt, size); }
PJ_INLINE(pj_status_t) pjmedia_transport_media_create(pjmedia_transport *tp, pj_pool_t *pool, unsigned options,
                                                      const pjmedia_sdp_session *rem_sdp, unsigned media_index)
{ return (*tp->op->media_create)(tp, pool, options, rem_sdp, media_index); }
PJ_INLINE(pj_status_t) pjmedia_transport_encode_sdp(pjmedia_transport *tp, pj_pool_t *sdp_pool,
                                                    pjmedia_sdp_session *sdp, const pjmedia_sdp_session *rem_sdp,
                                                    unsigned media_index)
{ return (*tp->op->encode_sdp)(tp, sdp_pool, sdp, rem_sdp, media_index); }
PJ_INLINE(pj_status_t) pjmedia_transport_media_start(pjmedia_transport *tp, pj_pool_t *tmp_pool,
                                                     const pjmedia_sdp_session *sdp_local,
                                                     const pjmedia_sdp_session *sdp_remote, unsigned media_index)
{ return (*tp->op->media_start)(tp, tmp_pool, sdp_local, sdp_remote, media_index); }
PJ_INLINE(pj_status_t) pjmedia_transport_media_stop(pjmedia_transport *tp) { return (*tp->op->media_stop)(tp); }
PJ_INLINE(pj_status_t) pjmedia_transport_close(pjmedia_transport *tp)
{ return tp->op->destroy ? (*tp->op->destroy)(tp) : PJ_SUCCESS; }
PJ_INLINE(pj_status_t) pjmedia_transport_simulate_lost(pjmedia_transport *tp, pjmedia_dir dir, unsigned pct_lost)
{ return (*tp->op->simulate_lost)(tp, dir, pct_lost); }

PJ_END_DECL

#endif
