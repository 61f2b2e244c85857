// igd_device.cuh -- device helpers shared by the kernel translation units (igd_fused.cu,
// igd_codec.cu, igd_packet.cu): streaming loads/stores, packed 16x2 ops, the shared-memory G.711
// decode table, the packed compressor, mbarrier / TMA bulk-copy wrappers.  Everything sits in an
// anonymous namespace (one private copy per translation unit; all of it inlines).
#pragma once
#include <cstdlib>

#include "igd_kernels.cuh"
#include "igd_math.cuh"

namespace {


constexpr int kChunks = IGD_FRAME / 16;   // 16-byte chunks per frame = 10
constexpr int kPst = kChunks + 1;         // padded partial stride (odd => conflict-free)
constexpr int kLutBytes = 256 * 64 * 4;   // 64 KB decode table

// ------------------------------------------------------------------ memory ops
__device__ __forceinline__ uint4 ld16_stream(const void *p)
{
    return __ldcs(reinterpret_cast<const uint4 *>(p));
}
__device__ __forceinline__ void st16_stream(void *p, uint4 v)
{
    __stcs(reinterpret_cast<uint4 *>(p), v);
}
// one 256-bit store (sm_100+): 16 PCM samples of one thread
__device__ __forceinline__ void st32_stream(void *p, const uint32_t (&v)[8])
{
    asm volatile("st.global.cs.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(v[0]),
                 "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}
__device__ __forceinline__ void ld32_stream(const void *p, uint32_t (&v)[8])
{
    asm volatile("ld.global.cs.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]),
                   "=r"(v[6]), "=r"(v[7])
                 : "l"(p));
}
// {hi, lo} -> two saturated int16 packed in one word (I2IP.S16.S32.SAT)
__device__ __forceinline__ uint32_t pack_sat16(int hi, int lo)
{
    uint32_t r;
    asm("cvt.pack.sat.s16.s32 %0, %1, %2;" : "=r"(r) : "r"(hi), "r"(lo));
    return r;
}
__device__ __forceinline__ int clamp16(int v) { return min(max(v, -32768), 32767); }

// ------------------------------------------------------------------ decode LUT
// 64 KB: one 32 KB table per law, row = code (128 B), column = lane (4 B), so a
// lookup can never bank-conflict (every lane owns its bank) and its address is ONE
// instruction on the FMA pipe: IDP.4A(word, 0x80 << 8k, lane_base) = code_k*128 +
// lane_base -- the ALU pipe, which bounds this kernel, is not involved.
// Entry = two int16: low half x/4 (every G.711 sample is a multiple of 4, so this is
// exact and sum((x/4)^2) over a 16-sample chunk fits 32 bits), high half clamp16(2x) =
// the sample after the reference's open-gate gain 2.0 (SLOT_VOLUME, Functions.cpp:
// 1682) with pjmedia's per-port clip.  One IDP.2A then selects 4*(x/4) (gain 1.0),
// clamp16(2x) (gain 2.0) or nothing (gate shut) AND accumulates it into the mix.
// Fills a [2 laws][256 codes][32 lane columns] table: every warp takes runs of 32 (law, code) pairs,
// lane l computes the entry of pair l ONCE, and the 32 entries are broadcast row by row
// (shuffle + one conflict-free store per row) -- 1/30th of the instructions of computing every
// replica, which is what a one-tick launch of the fused kernel mostly consisted of.
template <class F>
__device__ __forceinline__ void fill_lut32(uint32_t *lut, int warp, int lane, int nwarps, F entry_of)
{
    for (int base = warp * 32; base < 2 * 256; base += nwarps * 32) {
        const uint32_t idx = (uint32_t)(base + lane);
        const uint32_t mine = entry_of(idx >> 8, idx & 255u);
#pragma unroll 8
        for (int k = 0; k < 32; k++) lut[(base + k) * 32 + lane] = __shfl_sync(0xFFFFFFFFu, mine, k);
    }
}
__device__ __forceinline__ void build_decode_lut(uint32_t *lut, int tid, int nthreads)
{
    fill_lut32(lut, tid >> 5, tid & 31, nthreads >> 5, [](uint32_t law, uint32_t code) {      // nthreads: a multiple of 32
        const int x = law ? igd_ulaw2lin(code) : igd_alaw2lin(code);
        const int y2 = min(max(2 * x, -32768), 32767);
        return ((uint32_t)y2 << 16) | ((uint32_t)(x >> 2) & 0xFFFFu);
    });
}
// shared-window byte address of this lane's column in the table of `law`
__device__ __forceinline__ uint32_t lut_lane_base(uint32_t lut_s, uint32_t lane, uint32_t law)
{
    return lut_s + 4u * lane + ((law & 1u) << 15);
}
template <int K>
__device__ __forceinline__ uint32_t lut_lookup(uint32_t lane_base, uint32_t word)
{
    const uint32_t a = __dp4a(word, 0x80u << (8 * K), lane_base);
    uint32_t v;
    asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t shared_addr(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}
// d = c + a.lo16 * b.byte0 + a.hi16 * b.byte1   (IDP.2A.LO.S16.U8, FMA pipe)
__device__ __forceinline__ int dp2a_lo(uint32_t a, uint32_t b, int c)
{
    int d;
    asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ uint32_t dp2a_lo_u(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t d;
    asm("dp2a.lo.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
// packed 2 x 16-bit ALU ops (VIMNMX[3].S16x2 / .U16x2, VIADD.16x2, VIADDMNMX.S16x2)
__device__ __forceinline__ uint32_t max_s16x2(uint32_t a, uint32_t b)
{
    uint32_t r; asm("max.s16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
}
__device__ __forceinline__ uint32_t min_s16x2(uint32_t a, uint32_t b)
{
    uint32_t r; asm("min.s16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
}
__device__ __forceinline__ uint32_t max_u16x2(uint32_t a, uint32_t b)
{
    uint32_t r; asm("max.u16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
}
__device__ __forceinline__ uint32_t min_u16x2(uint32_t a, uint32_t b)
{
    uint32_t r; asm("min.u16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
}
// PRMT with the full 4-bit selectors (bit 3 = replicate the byte's sign); the
// __byte_perm intrinsic masks that bit away.
template <uint32_t kSel>
__device__ __forceinline__ uint32_t prmt_full(uint32_t a, uint32_t b)
{
    uint32_t r; asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "n"(kSel)); return r;
}
__device__ __forceinline__ uint32_t add_16x2(uint32_t a, uint32_t b)
{
    uint32_t r; asm("add.u16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
}
// -x per 16-bit half (two's complement, 0x8000 stays 0x8000 = 32768 unsigned): ~x + 1
__device__ __forceinline__ uint32_t neg_16x2(uint32_t x) { return add_16x2(~x, 0x00010001u); }

// ------------------------------------------------------------------ partials
// 8-byte per-chunk meter partial: lo = sum((x/4)^2) over the 16 samples (< 2^31);
// hi = max|x|/4 (16 bit) | bytesum<<16 (signed 16 bit)
__device__ __forceinline__ uint2 partial_pack(uint32_t sq16, uint32_t peakq, int bsum)
{
    return make_uint2(sq16, (peakq & 0xFFFFu) | ((uint32_t)bsum << 16));
}
__device__ __forceinline__ void partial_add(uint2 v, unsigned long long &sq16, uint32_t &peakq, int &bsum)
{
    sq16 += v.x;
    peakq = max(peakq, v.y & 0xFFFFu);
    bsum += (int)v.y >> 16;
}
__device__ __forceinline__ igd_meter_rec meter_finish(unsigned long long sq, uint32_t peak, int bsum,
                                                      bool have_bytes)
{
    igd_meter_rec r;
    r.sumsq_lo = (uint32_t)sq;
    const uint32_t bm = have_bytes ? igd_bytemean_from_sum(bsum, IGD_FRAME) : 0u;
    r.hi = ((uint32_t)(sq >> 32) & 0xFFu) | (bm << 8) | (peak << 16);
    r.rms_dbfs = igd_rms_dbfs(sq);
    r.peak_dbfs = igd_peak_dbfs(peak);
    return r;
}

__device__ __forceinline__ int bytesum4(uint32_t w, bool signed_char, int acc)
{
    return signed_char ? __dp4a((int)w, 0x01010101, acc) : (int)__dp4a(w, 0x01010101u, (uint32_t)acc);
}

// The same compressor (igd_math.cuh) on PACKED pairs of int16: the pre-bias, clip
// and segment normalisation run two samples per instruction on the 16x2 ALU ops,
// only the exponent extraction (one PRMT + FFMA + shift) is per sample.
struct enc_pk {
    uint32_t bias_pos, bias_x;   // packed pre-bias for x>=0, and pos^neg
    uint32_t hi_pos, hi_x;       // packed upper clip of t so that t+bias <= 0x7FFF
    uint32_t thr;                // packed 256 (A-law) / 0 (u-law)
    uint32_t mask4;              // output XOR mask for x>=0, replicated per byte
    uint32_t zero2;              // packed (0, 0) in a register the compiler cannot fold: max.s16x2(x, 0) with a
                                 // literal makes it rebuild the constant (PRMT RZ) for every packed pair
    // the law-independent constants of enc_pair; enc_pk_make leaves them literal (folded into the instructions),
    // a kernel that reads them back from shared memory keeps them in (uniform) registers instead of rebuilding
    // them in every loop trip
    uint32_t sel_lo = 0x0001u, sel_hi = 0x0100u, f_magic = 0x4B000000u;
    float f_scale = 0.0078125f;
};
__device__ __forceinline__ enc_pk enc_pk_make(int law)
{
    const igd_enc_law L = igd_enc_law_make(law);
    enc_pk e;
    const uint32_t bp = (uint32_t)L.bpos & 0xFFFFu, bn = (uint32_t)L.bneg & 0xFFFFu;
    const uint32_t hp = (uint32_t)(0x7FFF - max(L.bpos, 0)), hn = (uint32_t)(0x7FFF - max(L.bneg, 0));
    e.bias_pos = bp * 0x10001u; e.bias_x = (bp ^ bn) * 0x10001u;
    e.hi_pos = hp * 0x10001u;   e.hi_x = (hp ^ hn) * 0x10001u;
    e.thr = (uint32_t)L.thr * 0x10001u;
    e.mask4 = L.mpos * 0x01010101u;
    e.zero2 = 0u;
    return e;
}
// two packed samples -> the float whose bits [26:19] are seg<<4|mant, per sample
__device__ __forceinline__ void enc_pair(uint32_t pk, const enc_pk &E, uint32_t &g0, uint32_t &g1)
{
    const uint32_t sgn = prmt_full<0xBB99>(pk, 0u);                 // sign of each half, replicated
    uint32_t t = pk ^ sgn;                                          // |x| or |x|-1
    t = min_u16x2(t, E.hi_pos ^ (sgn & E.hi_x));                    // u-law clip
    const uint32_t p = max_s16x2(add_16x2(t, E.bias_pos ^ (sgn & E.bias_x)), E.zero2);
    const uint32_t P = add_16x2(p, max_u16x2(p, E.thr));            // leading one -> segment
    // 8388608.0f + P per half, built on the FMA pipe (IDP.2A picks the half and adds the magic;
    // the compressor is ALU-bound, a PRMT here measured 4 % slower)
    g0 = __float_as_uint(fmaf(__uint_as_float(dp2a_lo_u(P, E.sel_lo, E.f_magic)), E.f_scale, -65536.0f));
    g1 = __float_as_uint(fmaf(__uint_as_float(dp2a_lo_u(P, E.sel_hi, E.f_magic)), E.f_scale, -65536.0f));
}
// 2 packed words (4 samples) -> 4 code bytes
__device__ __forceinline__ uint32_t encode4_packed(uint32_t pa, uint32_t pb, const enc_pk &E)
{
    uint32_t g0, g1, g2, g3;
    enc_pair(pa, E, g0, g1);
    enc_pair(pb, E, g2, g3);
    // upper halves of the floats hold the code at bits [10:3]: pack two per word, one shift
    // moves both into bytes 1 and 3, one PRMT gathers the four codes
    const uint32_t h01 = __byte_perm(g0, g1, 0x7632) << 5, h23 = __byte_perm(g2, g3, 0x7632) << 5;
    const uint32_t codes = __byte_perm(h01, h23, 0x7531);
    const uint32_t sg = __byte_perm(pa, pb, 0x7531);   // sign bit of each sample in bit 7
    return codes ^ ((sg & 0x80808080u) ^ E.mask4);
}
// 8 packed words (16 samples) -> 16 code bytes
__device__ __forceinline__ uint4 encode16_packed(const uint32_t (&pk)[8], const enc_pk &E)
{
    uint32_t w[4];
#pragma unroll
    for (int j = 0; j < 4; j++) w[j] = encode4_packed(pk[2 * j], pk[2 * j + 1], E);
    return make_uint4(w[0], w[1], w[2], w[3]);
}

// ---------------------------------------------------------------- mbarrier / TMA
// (all on shared-window byte addresses, so that warp-uniform operands stay in uniform registers)
__device__ __forceinline__ void mbar_init(uint32_t bar_s, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_s), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar_s, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_s), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar_s, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@!p bra WAIT_%=;\n\t}"
        ::"r"(bar_s), "r"(parity) : "memory");
}
// orders this thread's earlier generic-proxy shared accesses before later async-proxy (TMA) ones
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// TMA bulk copy global -> shared (UBLKCP): no registers, no LSU issue slots; the
// bytes land asynchronously and complete_tx on the mbarrier.
__device__ __forceinline__ void bulk_g2s(uint32_t dst_s, const void *src, uint32_t bytes, uint32_t bar_s)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_s), "l"(src), "r"(bytes), "r"(bar_s) : "memory");
}

inline int grid_for(const igd_launch_cfg &c, size_t work_items, int threads, int per_sm)
{
    size_t blocks = (work_items + threads - 1) / threads;
    size_t cap = (size_t)c.sm_count * per_sm;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

}  // namespace
