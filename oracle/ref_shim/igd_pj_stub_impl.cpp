/*
 * igd_pj_stub_impl.cpp -- bodies of the few PJLIB / PJMEDIA functions declared in igd_pj_stub.h
 * (pool, pj_strdup2, pjmedia_sdp_attr_add, pjmedia_rtp_decode_rtp), restated from the published
 * pjproject 2.x behaviour.  TEST INFRASTRUCTURE ONLY: linked into oracle/_ref/libigd_ref_ta*.so next
 * to the reference's own TransportAdapter.cpp, and into the PJSIP-role driver that exercises the
 * product shim through its pjmedia_transport vtable (tests/host_cpp/vtable_driver.cpp).
 */
#include "igd_pj_stub.h"

extern "C" {

void *pj_pool_alloc(pj_pool_t *pool, pj_size_t size)
{
    igd_pool_block *b = (igd_pool_block *)malloc(sizeof(igd_pool_block) + 16 + size);
    b->next = pool->blocks;
    pool->blocks = b;
    return (char *)b + 16;
}
void *pj_pool_zalloc(pj_pool_t *pool, pj_size_t size)
{
    void *p = pj_pool_alloc(pool, size);
    memset(p, 0, size);
    return p;
}
void pj_pool_release(pj_pool_t *pool)
{
    igd_pool_block *b = pool->blocks;
    while (b) { igd_pool_block *n = b->next; free(b); b = n; }
    free(pool);
}
pj_pool_t *pjmedia_endpt_create_pool(pjmedia_endpt *, const char *name, pj_size_t, pj_size_t)
{
    pj_pool_t *p = (pj_pool_t *)calloc(1, sizeof(pj_pool_t));
    snprintf(p->obj_name, sizeof p->obj_name, name, (void *)p);   /* "tpad%p" */
    return p;
}
pj_str_t *pj_strdup2(pj_pool_t *pool, pj_str_t *dst, const char *src)
{
    size_t n = src ? strlen(src) : 0;
    dst->ptr = (char *)pj_pool_alloc(pool, n + 1);
    memcpy(dst->ptr, src ? src : "", n + 1);
    dst->slen = (pj_ssize_t)n;
    return dst;
}
pj_status_t pjmedia_sdp_attr_add(unsigned *count, pjmedia_sdp_attr *attr_array[], pjmedia_sdp_attr *attr)
{
    if (*count >= PJMEDIA_MAX_SDP_ATTR) return PJ_EINVAL;
    attr_array[(*count)++] = attr;
    return PJ_SUCCESS;
}
/* RFC 3550 s5.1 / s5.3.1, the checks of pjmedia's rtp.c */
pj_status_t pjmedia_rtp_decode_rtp(pjmedia_rtp_session *, const void *pkt, int pkt_len,
                                   const pjmedia_rtp_hdr **hdr, const void **payload, unsigned *payloadlen)
{
    *hdr = (const pjmedia_rtp_hdr *)pkt;
    if ((*hdr)->v != 2) return PJMEDIA_RTP_EINVER;
    int offset = (int)sizeof(pjmedia_rtp_hdr) + (int)((*hdr)->cc * sizeof(pj_uint32_t));
    if ((*hdr)->x) {
        const uint8_t *ext = (const uint8_t *)pkt + offset;
        unsigned ext_len = ((unsigned)ext[2] << 8) | ext[3];
        offset += (int)((ext_len + 1) * sizeof(pj_uint32_t));
    }
    if (offset > pkt_len) return PJMEDIA_RTP_EINLEN;
    *payload = (const uint8_t *)pkt + offset;
    *payloadlen = (unsigned)(pkt_len - offset);
    if ((*hdr)->p && *payloadlen > 0) {
        unsigned pad = ((const uint8_t *)(*payload))[*payloadlen - 1];
        if (pad <= *payloadlen) *payloadlen -= pad;
    }
    return PJ_SUCCESS;
}
}  /* extern "C" */
