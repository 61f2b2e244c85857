// igd_kernels.cu -- hand-written sm_100a kernels of the iGate4x voice path.
//
// Everything here is streaming integer/byte work bounded by HBM3e bandwidth or,
// for the fused kernel, by instruction issue (see DESIGN.md "Kernels").  No
// tensor cores: nothing on this path is a contraction.
//
// Common design points
//  * 128-bit (LDG.128) loads / 128- and 256-bit (STG.E.ENL2.256) stores with the
//    evict-first hint: every input byte is read once and every output written
//    once, so nothing should stay in L2.
//  * G.711 expansion through a 64 KB shared-memory table: one 32 KB table per law,
//    row = code (128 B), column = lane, so a lookup can never bank-conflict and its
//    byte address is ONE FMA-pipe instruction: IDP.4A(word, 0x80 << 8k, lane_base).
//  * G.711 compression in ALU through the float-exponent trick of igd_math.cuh
//    (no 64 KB encode table, no data-dependent bank conflicts).
//  * Per-frame meters: exact integer partial sums per 16-sample chunk, combined
//    through a padded shared-memory scratch (conflict-free LDS.64), dB via SFU lg2.
//  * Persistent grids: a multiple of the SM count; the fused kernel runs one CTA of
//    24 autonomous warps per SM, every warp with its own TMA slot and mbarrier.
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
__device__ __forceinline__ void build_decode_lut(uint32_t *lut, int tid, int nthreads)
{
    for (int i = tid; i < 2 * 256 * 32; i += nthreads) {
        const uint32_t law = (uint32_t)i >> 13, code = ((uint32_t)i >> 5) & 255u;
        const int x = law ? igd_ulaw2lin(code) : igd_alaw2lin(code);
        const int y2 = min(max(2 * x, -32768), 32767);
        lut[i] = ((uint32_t)y2 << 16) | ((uint32_t)(x >> 2) & 0xFFFFu);
    }
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
    return e;
}
// two packed samples -> the float whose bits [26:19] are seg<<4|mant, per sample
__device__ __forceinline__ void enc_pair(uint32_t pk, const enc_pk &E, uint32_t &g0, uint32_t &g1)
{
    const uint32_t sgn = prmt_full<0xBB99>(pk, 0u);                 // sign of each half, replicated
    uint32_t t = pk ^ sgn;                                          // |x| or |x|-1
    t = min_u16x2(t, E.hi_pos ^ (sgn & E.hi_x));                    // u-law clip
    const uint32_t p = max_s16x2(add_16x2(t, E.bias_pos ^ (sgn & E.bias_x)), 0u);
    const uint32_t P = add_16x2(p, max_u16x2(p, E.thr));            // leading one -> segment
    // 8388608.0f + P per half, built on the FMA pipe (IDP.2A picks the half and adds the magic;
    // the compressor is ALU-bound, a PRMT here measured 4 % slower)
    g0 = __float_as_uint(fmaf(__uint_as_float(dp2a_lo_u(P, 0x0001u, 0x4B000000u)), 0.0078125f, -65536.0f));
    g1 = __float_as_uint(fmaf(__uint_as_float(dp2a_lo_u(P, 0x0100u, 0x4B000000u)), 0.0078125f, -65536.0f));
}
// 8 packed words (16 samples) -> 16 code bytes
__device__ __forceinline__ uint4 encode16_packed(const uint32_t (&pk)[8], const enc_pk &E)
{
    uint32_t w[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        uint32_t g0, g1, g2, g3;
        enc_pair(pk[2 * j], E, g0, g1);
        enc_pair(pk[2 * j + 1], E, g2, g3);
        // upper halves of the floats hold the code at bits [10:3]: pack two per word, one shift
        // moves both into bytes 1 and 3, one PRMT gathers the four codes
        const uint32_t h01 = __byte_perm(g0, g1, 0x7632) << 5, h23 = __byte_perm(g2, g3, 0x7632) << 5;
        const uint32_t codes = __byte_perm(h01, h23, 0x7531);
        const uint32_t sg = __byte_perm(pk[2 * j], pk[2 * j + 1], 0x7531);   // sign bit of each sample in bit 7
        w[j] = codes ^ ((sg & 0x80808080u) ^ E.mask4);
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}

// ==================================================================== fused
struct FusedParams {
    const uint8_t *codes;
    const uint8_t *law;
    const uint16_t *gain;
    const uint8_t *out_law;
    int16_t *mix;
    uint8_t *enc;
    igd_meter_rec *meter;
    igd_bridge_rec *bmeter;
    long long total_bf;   // F*B bridge-frames
    long long num_tiles;
    int B, G;
    unsigned flags;
};

constexpr uint32_t kSelGeneral = 0xFFFFFFFFu;
// gain_q7 -> IDP.2A selector: 0 -> nothing, 128 -> x, 256 -> clamp16(2x); anything else
// takes the general multiply / shift / clip path.
__device__ __forceinline__ uint32_t gain_selector(uint32_t adj)
{
    return adj == 0u ? 0u : adj == 128u ? 0x0004u : adj == 256u ? 0x0100u : kSelGeneral;
}

// decode + meter + gain/accumulate one 16-sample chunk of one leg
// kMode: 0 = gate shut (meter only), 1 = gain 1.0 / 2.0 through the IDP.2A selector,
//        2 = arbitrary Q7 gain (multiply, shift, clip)
template <bool kSigned, int kMode>
__device__ __forceinline__ uint2 leg_chunk(uint32_t lane_base, uint4 w, uint32_t sel, int adj, int (&acc)[16])
{
    const uint32_t wd[4] = {w.x, w.y, w.z, w.w};
    uint32_t sq = 0;               // sum of (x/4)^2: 16 * 8064^2 < 2^31
    uint32_t mx = 0, mn = 0;       // packed running max / min of (x/4, clamp16(2x))
    int bsum = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const uint32_t e0 = lut_lookup<0>(lane_base, wd[j]);
        const uint32_t e1 = lut_lookup<1>(lane_base, wd[j]);
        const uint32_t e2 = lut_lookup<2>(lane_base, wd[j]);
        const uint32_t e3 = lut_lookup<3>(lane_base, wd[j]);
        // x/4 sign-extended from the low half (IDP.2A with a unit selector)
        const int x0 = dp2a_lo(e0, 1u, 0), x1 = dp2a_lo(e1, 1u, 0);
        const int x2 = dp2a_lo(e2, 1u, 0), x3 = dp2a_lo(e3, 1u, 0);
        sq += (uint32_t)(x0 * x0) + (uint32_t)(x1 * x1) + (uint32_t)(x2 * x2) + (uint32_t)(x3 * x3);
        mx = max_s16x2(max_s16x2(mx, e0), e1); mx = max_s16x2(max_s16x2(mx, e2), e3);
        mn = min_s16x2(min_s16x2(mn, e0), e1); mn = min_s16x2(min_s16x2(mn, e2), e3);
        bsum = kSigned ? __dp4a((int)wd[j], 0x01010101, bsum) : (int)__dp4a(wd[j], 0x01010101u, (uint32_t)bsum);
        if (kMode == 1) {
            acc[4 * j + 0] = dp2a_lo(e0, sel, acc[4 * j + 0]);
            acc[4 * j + 1] = dp2a_lo(e1, sel, acc[4 * j + 1]);
            acc[4 * j + 2] = dp2a_lo(e2, sel, acc[4 * j + 2]);
            acc[4 * j + 3] = dp2a_lo(e3, sel, acc[4 * j + 3]);
        } else if (kMode == 2) {
            acc[4 * j + 0] += clamp16((4 * x0 * adj) >> 7);
            acc[4 * j + 1] += clamp16((4 * x1 * adj) >> 7);
            acc[4 * j + 2] += clamp16((4 * x2 * adj) >> 7);
            acc[4 * j + 3] += clamp16((4 * x3 * adj) >> 7);
        }
    }
    const int pmax = (int)(short)(mx & 0xFFFFu), pmin = (int)(short)(mn & 0xFFFFu);
    return partial_pack(sq, (uint32_t)max(pmax, -pmin), bsum);
}

// bridge output of one 16-sample chunk: saturate, store PCM, compress, store codes
template <bool kSigned>
__device__ __forceinline__ uint2 mix_out_chunk(const int (&acc)[16], const enc_pk &E, int16_t *mix_dst,
                                               uint8_t *enc_dst, bool do_store = true)
{
    uint32_t pk[8];
#pragma unroll
    for (int i = 0; i < 8; i++) pk[i] = pack_sat16(acc[2 * i + 1], acc[2 * i]);
    if (do_store) st32_stream(mix_dst, pk);
    uint32_t mx = max_s16x2(max_s16x2(pk[0], pk[1]), pk[2]), mn = min_s16x2(min_s16x2(pk[0], pk[1]), pk[2]);
    mx = max_s16x2(max_s16x2(mx, pk[3]), pk[4]); mn = min_s16x2(min_s16x2(mn, pk[3]), pk[4]);
    mx = max_s16x2(max_s16x2(mx, pk[5]), pk[6]); mn = min_s16x2(min_s16x2(mn, pk[5]), pk[6]);
    mx = max_s16x2(mx, pk[7]); mn = min_s16x2(mn, pk[7]);
    const int hi = max((int)(short)(mx & 0xFFFFu), (int)mx >> 16);
    const int lo = min((int)(short)(mn & 0xFFFFu), (int)mn >> 16);
    const uint4 e = encode16_packed(pk, E);
    if (do_store) st16_stream(enc_dst, e);
    int esum = 0;
    if (kSigned) {
        esum = __dp4a((int)e.x, 0x01010101, esum); esum = __dp4a((int)e.y, 0x01010101, esum);
        esum = __dp4a((int)e.z, 0x01010101, esum); esum = __dp4a((int)e.w, 0x01010101, esum);
    } else {
        uint32_t u = __dp4a(e.x, 0x01010101u, 0u); u = __dp4a(e.y, 0x01010101u, u);
        u = __dp4a(e.z, 0x01010101u, u); u = __dp4a(e.w, 0x01010101u, u);
        esum = (int)u;
    }
    return make_uint2((uint32_t)esum, (uint32_t)max(hi, -lo));
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
// TMA bulk copy global -> shared (UBLKCP): no registers, no LSU issue slots; the
// bytes land asynchronously and complete_tx on the mbarrier.
__device__ __forceinline__ void bulk_g2s(uint32_t dst_s, const void *src, uint32_t bytes, uint32_t bar_s)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_s), "l"(src), "r"(bytes), "r"(bar_s) : "memory");
}


// raw gain bits of one bridge-frame (G u16 values) in two registers
template <int G>
__device__ __forceinline__ uint2 load_gains(const uint16_t *g)
{
    if (G == 4) return *reinterpret_cast<const uint2 *>(g);
    if (G == 2) return make_uint2(*reinterpret_cast<const uint32_t *>(g), 0u);
    return make_uint2(*g, 0u);
}

// ------------------------------------------------------------------------------------
// Warp-autonomous fused kernel (G in {1,2,4}): no cross-warp handshake at all.
// A warp owns "items" of kBfPerItem = 6 consecutive bridge-frames (6*G*160 contiguous
// code bytes) and strides over them on its own:
//   * the item's codes are fetched by the warp's OWN bulk async copies (TMA, one per
//     bridge-frame into a padded, bank-conflict-free slot), completion on the warp's
//     private mbarrier; the copies of item i+1 are issued as soon as every lane holds
//     the last codes of item i, so HBM latency hides behind half an item (~1.5 us) and
//     ~77 KB per SM are in flight;
//   * lane = two 16-sample chunks (c and c+5) of one bridge-frame, all G legs (5 lanes
//     per bridge-frame, 30 of 32 lanes busy): per-bridge-frame setup (gains, laws,
//     selectors, addresses) is paid once per 32 samples;
//   * the per-chunk meter partials meet in the warp's private shared scratch after a
//     __syncwarp; lanes 0..6G-1 finish one leg record each, the next 6 lanes one bridge
//     record each.
// Warps drift freely: nobody spins on a peer (the CTA-cooperative kernel above spends
// ~16 % of its issue slots in mbarrier spin loops).
//
// Decode table of this kernel: low half |x|/4 (unsigned -> the frame peak is ONE packed
// max per two samples, no min), high half clamp16(2x).  The one-instruction accumulate
// covers the reference's gains 0.0 and 2.0; every other gain (sidetone 0.1, 0.5, 1.0)
// takes the multiply/shift/clip path with the sign recovered from the high half.
constexpr int kBfPerItem = 6;
constexpr int kC32 = IGD_FRAME / 32;      // lanes per bridge-frame = 5
// A slot is the item's 6*G*160 code bytes as they lie in HBM (ONE bulk copy).  Lane (bfl, c)
// works on the 16-sample chunks (c + rot(bfl)) % 10 and that + 5; the per-bridge-frame rotation
// (0,5,0,7,3,0) is the brute-forced minimum of LDS.128 bank conflicts for this layout (12
// wavefronts instead of 8 per leg; padding the slot instead would cost six copies per item).
template <int G> struct slot_geom {
    static constexpr int kBfBytes = G * IGD_FRAME;
    static constexpr int kSlotBytes = kBfPerItem * kBfBytes;
};
__device__ __forceinline__ uint32_t chunk_rotation(uint32_t bfl) { return (0x037050u >> (4 * bfl)) & 0xFu; }

__device__ __forceinline__ void build_decode_lut_abs(uint32_t *lut, int tid, int nthreads)
{
    for (int i = tid; i < 2 * 256 * 32; i += nthreads) {
        const uint32_t law = (uint32_t)i >> 13, code = ((uint32_t)i >> 5) & 255u;
        const int x = law ? igd_ulaw2lin(code) : igd_alaw2lin(code);
        const int y2 = min(max(2 * x, -32768), 32767);
        lut[i] = ((uint32_t)y2 << 16) | ((uint32_t)(abs(x) >> 2) & 0xFFFFu);
    }
}

// bit 15 / 31 set for every non-zero 16-bit half of x
__device__ __forceinline__ uint32_t nonzero_halves(uint32_t x)
{
    return (((x & 0x7FFF7FFFu) + 0x7FFF7FFFu) | x) & 0x80008000u;
}

// decode + meter + gain/accumulate one 16-sample chunk of one leg (|x|/4 table)
// kMode: 0 = gate shut for the whole warp (meter only), 1 = gain 2.0 or shut through the
//        IDP.2A selector (0x0100 / 0), 2 = arbitrary Q7 gain (multiply, shift, clip)
template <bool kSigned, int kMode>
__device__ __forceinline__ uint2 leg_chunk_u(uint32_t lane_base, uint4 w, uint32_t sel, int adj, int (&acc)[16])
{
    const uint32_t wd[4] = {w.x, w.y, w.z, w.w};
    uint32_t sq = 0;               // sum of (x/4)^2: 16 * 8064^2 < 2^31
    uint32_t mx = 0;               // packed running max of (|x|/4, junk)
    int bsum = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const uint32_t e0 = lut_lookup<0>(lane_base, wd[j]);
        const uint32_t e1 = lut_lookup<1>(lane_base, wd[j]);
        const uint32_t e2 = lut_lookup<2>(lane_base, wd[j]);
        const uint32_t e3 = lut_lookup<3>(lane_base, wd[j]);
        // |x|/4 out of the low half on the ALU pipe (LOP3): everything else in this loop body is
        // IDP/IMAD on the FMA pipe, which bounds this phase (measured: any IDP.2A here is slower)
        const uint32_t x0 = e0 & 0xFFFFu, x1 = e1 & 0xFFFFu, x2 = e2 & 0xFFFFu, x3 = e3 & 0xFFFFu;
        sq += x0 * x0 + x1 * x1 + x2 * x2 + x3 * x3;
        mx = max_u16x2(max_u16x2(mx, e0), e1); mx = max_u16x2(max_u16x2(mx, e2), e3);
        bsum = kSigned ? __dp4a((int)wd[j], 0x01010101, bsum) : (int)__dp4a(wd[j], 0x01010101u, (uint32_t)bsum);
        if (kMode == 1) {
            acc[4 * j + 0] = dp2a_lo(e0, sel, acc[4 * j + 0]);
            acc[4 * j + 1] = dp2a_lo(e1, sel, acc[4 * j + 1]);
            acc[4 * j + 2] = dp2a_lo(e2, sel, acc[4 * j + 2]);
            acc[4 * j + 3] = dp2a_lo(e3, sel, acc[4 * j + 3]);
        } else if (kMode == 2) {
            const int s0 = (int)e0 < 0 ? -(int)x0 : (int)x0, s1 = (int)e1 < 0 ? -(int)x1 : (int)x1;
            const int s2 = (int)e2 < 0 ? -(int)x2 : (int)x2, s3 = (int)e3 < 0 ? -(int)x3 : (int)x3;
            acc[4 * j + 0] += clamp16((4 * s0 * adj) >> 7);
            acc[4 * j + 1] += clamp16((4 * s1 * adj) >> 7);
            acc[4 * j + 2] += clamp16((4 * s2 * adj) >> 7);
            acc[4 * j + 3] += clamp16((4 * s3 * adj) >> 7);
        }
    }
    return make_uint2(sq, __byte_perm(mx, (uint32_t)bsum, 0x5410));   // {sq, peak/4 | bsum << 16}
}

template <int G, bool kSigned, int kWarps>
__global__ void __launch_bounds__(kWarps * 32, 1) k_fused_w(const FusedParams q)
{
    static_assert(kBfPerItem * G + kBfPerItem <= 32, "finish needs one lane per record");
    using geom = slot_geom<G>;
    constexpr int kLegParts = kBfPerItem * G * kPst, kBrParts = kBfPerItem * kPst;
    __shared__ uint64_t bars[kWarps];
    __shared__ __align__(16) uint32_t enc_tab[2][8];
    extern __shared__ __align__(128) uint8_t smem[];
    uint32_t *lut = reinterpret_cast<uint32_t *>(smem);
    const uint32_t lut_bytes = shared_addr(smem);

    const int t = threadIdx.x;
    const uint32_t lane = t & 31;
    const uint32_t warp = __shfl_sync(0xFFFFFFFFu, (uint32_t)t >> 5, 0);   // warp-uniform for the compiler too
    const uint32_t slot_s = shared_addr(smem + kLutBytes) + warp * geom::kSlotBytes;
    uint2 *part = reinterpret_cast<uint2 *>(smem + kLutBytes + (size_t)kWarps * geom::kSlotBytes) +
                  (size_t)warp * (kLegParts + kBrParts);
    uint2 *bpart = part + kLegParts;
    const uint32_t bar_s = shared_addr(bars) + warp * 8;

    build_decode_lut_abs(lut, t, kWarps * 32);
    if (t < 2) {
        const enc_pk e = enc_pk_make(t);
        enc_tab[t][0] = e.bias_pos; enc_tab[t][1] = e.bias_x; enc_tab[t][2] = e.hi_pos; enc_tab[t][3] = e.hi_x;
        enc_tab[t][4] = e.thr; enc_tab[t][5] = e.mask4;
    }
    if (lane == 0) {
        mbar_init(bar_s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // bridge-frame and item indices fit 32 bits (the launcher checks total_bf < 2^31 - slack)
    const uint32_t total_bf = (uint32_t)q.total_bf;
    const uint32_t items = (total_bf + kBfPerItem - 1) / kBfPerItem;
    const uint32_t nw = gridDim.x * kWarps;
    uint32_t item = warp * gridDim.x + blockIdx.x;      // neighbouring items on different SMs
    // one elected lane posts the byte count and issues the item's bulk async copy; every operand is
    // warp-uniform (uniform datapath, no R2UR shuffles)
    auto fetch = [&](uint32_t it_idx) {
        const uint32_t bf0 = it_idx * kBfPerItem;
        const uint32_t left = total_bf - bf0;
        const uint32_t bytes = (left < (uint32_t)kBfPerItem ? left : (uint32_t)kBfPerItem) * geom::kBfBytes;
        if (lane == 0) {
            mbar_expect_tx(bar_s, bytes);
            bulk_g2s(slot_s, q.codes + (size_t)bf0 * geom::kBfBytes, bytes, bar_s);
        }
    };
    if (item < items) fetch(item);

    const bool worker = lane < kBfPerItem * kC32;
    const uint32_t bfl = worker ? lane / kC32 : 0u;
    const uint32_t c0 = worker ? (lane - bfl * kC32 + chunk_rotation(bfl)) % kChunks : 0u;   // first chunk; second = +-5
    const uint32_t src = slot_s + bfl * geom::kBfBytes;
    const uint32_t lane4 = lut_bytes + 4u * lane;
    const uint32_t b_step = (uint32_t)(((unsigned long long)nw * kBfPerItem) % (uint32_t)q.B);
    uint32_t b = (item * kBfPerItem + bfl) % (uint32_t)q.B;
    auto load_laws = [&](uint32_t bb) -> uint32_t {      // the G leg laws (1 bit each) and the output law (bit 8)
        uint32_t r = 0;
        if (G == 4) {
            const uint32_t lw = __ldg(reinterpret_cast<const uint32_t *>(q.law + (size_t)bb * 4));
            r = (lw & 1u) | ((lw >> 7) & 2u) | ((lw >> 14) & 4u) | ((lw >> 21) & 8u);
        } else {
#pragma unroll
            for (int g = 0; g < G; g++) r |= (uint32_t)(__ldg(q.law + (size_t)bb * G + g) & 1u) << g;
        }
        return r | ((uint32_t)(__ldg(q.out_law + bb) & 1u) << 8);
    };
    uint2 gq = make_uint2(0u, 0u);
    uint32_t lwq = 0u;
    if (worker && item < items && item * kBfPerItem + bfl < total_bf) {
        gq = load_gains<G>(q.gain + (size_t)(item * kBfPerItem + bfl) * G);
        lwq = load_laws(b);
    }

    for (uint32_t it = 0; item < items; item += nw, it++) {
        const uint32_t bf = item * kBfPerItem + bfl;
        const uint32_t next = item + nw;
        mbar_wait(bar_s, it & 1u);                           // this item's codes have landed
        const uint2 gcur = gq;
        const uint32_t lcur = lwq;
        const bool valid = worker && bf < total_bf;
        {
            b += b_step;
            if (b >= (uint32_t)q.B) b -= (uint32_t)q.B;
            const uint32_t bfn = bf + nw * kBfPerItem;
            if (worker && next < items && bfn < total_bf) {      // next item's gains and laws ride in three registers
                gq = load_gains<G>(q.gain + (size_t)bfn * G);
                lwq = load_laws(b);
            }
        }
        // every lane runs the same instruction stream (idle / tail lanes on stale bytes with all
        // gates shut); only the stores are predicated, and the mode branches are warp-uniform
        auto adj_of = [&](int g) -> uint32_t { return g == 0 ? (gcur.x & 0xFFFFu) : g == 1 ? (gcur.x >> 16) : g == 2 ? (gcur.y & 0xFFFFu) : (gcur.y >> 16); };
        // warp-wide OR of the gains (REDUX): anything but 0 / 256 anywhere in the warp -> general
        // path; bit g of open_mask: some lane of the warp has leg g open
        const uint32_t orx = __reduce_or_sync(0xFFFFFFFFu, gcur.x), ory = G > 2 ? __reduce_or_sync(0xFFFFFFFFu, gcur.y) : 0u;
        const bool general = ((orx | ory) & 0xFEFFFEFFu) != 0u;
        const uint32_t open_mask = ((orx & 0xFFFFu) ? 1u : 0u) | ((orx >> 16) ? 2u : 0u) | ((ory & 0xFFFFu) ? 4u : 0u) |
                                   ((ory >> 16) ? 8u : 0u);
        // open legs of this lane's bridge-frame, pre-shifted into the high half of the bridge partial
        // (every one of the ten partials carries it; the finish divides the sum by ten)
        const uint32_t n_open16 = (uint32_t)(__popc(nonzero_halves(gcur.x)) + __popc(nonzero_halves(gcur.y))) << 16;
#pragma unroll 1
        for (int h = 0; h < 2; h++) {
            const uint32_t ch = h == 0 ? c0 : (c0 >= (uint32_t)kC32 ? c0 - kC32 : c0 + kC32);   // this pass's chunk
            uint4 wh[G];
#pragma unroll
            for (int g = 0; g < G; g++)
                asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                             : "=r"(wh[g].x), "=r"(wh[g].y), "=r"(wh[g].z), "=r"(wh[g].w)
                             : "r"(src + g * IGD_FRAME + ch * 16));
            if (h == 1) {        // every lane holds the rest of its codes: refill the slot
                __syncwarp();
                if (next < items) fetch(next);
            }
            uint2 *mypart = part + bfl * (G * kPst) + ch;
            int acc[16];
#pragma unroll
            for (int i = 0; i < 16; i++) acc[i] = 0;
            if (!general) {        // gains in {0, 2.0}: one IDP.2A per sample of an open leg
#pragma unroll
                for (int g = 0; g < G; g++) {
                    const uint32_t lb = lane4 + (((lcur >> g) & 1u) << 15);
                    const uint2 ph = (open_mask >> g) & 1u ? leg_chunk_u<kSigned, 1>(lb, wh[g], adj_of(g), 0, acc)
                                                           : leg_chunk_u<kSigned, 0>(lb, wh[g], 0u, 0, acc);
                    if (valid) mypart[g * kPst] = ph;
                }
            } else {               // arbitrary Q7 gains: multiply, shift, clip
#pragma unroll
                for (int g = 0; g < G; g++) {
                    const uint32_t lb = lane4 + (((lcur >> g) & 1u) << 15);
                    const uint2 ph = leg_chunk_u<kSigned, 2>(lb, wh[g], 0u, (int)adj_of(g), acc);
                    if (valid) mypart[g * kPst] = ph;
                }
            }
            enc_pk E;
            {
                const uint32_t *et = enc_tab[(lcur >> 8) & 1u];
                const uint4 e0 = *reinterpret_cast<const uint4 *>(et);
                const uint2 e1 = *reinterpret_cast<const uint2 *>(et + 4);
                E.bias_pos = e0.x; E.bias_x = e0.y; E.hi_pos = e0.z; E.hi_x = e0.w; E.thr = e1.x; E.mask4 = e1.y;
            }
            const uint32_t o16 = bf * kChunks + ch;     // 16-sample chunk index of the outputs
            const uint2 mo = mix_out_chunk<kSigned>(acc, E, q.mix + (size_t)o16 * 16, q.enc + (size_t)o16 * 16, valid);
            if (valid) bpart[bfl * kPst + ch] = make_uint2(mo.x, mo.y | n_open16);
        }
        __syncwarp();
        // ---- finish: one lane per record (leg records, then bridge records: bpart follows part,
        // and both kinds of partial are {sum, max | sum << 16}, so the ten-partial walk is shared)
        {
            const uint32_t bf0 = item * kBfPerItem;
            const uint2 *src_p = part + lane * kPst;
            unsigned long long sq = 0; uint32_t pk = 0; int bsum = 0;
            if (lane < kBfPerItem * G + kBfPerItem) {
#pragma unroll
                for (int i = 0; i < kChunks; i++) {
                    const uint2 v = src_p[i];
                    sq += v.x; pk = max_u16x2(pk, v.y); bsum = dp2a_lo(v.y, 0x0100u, bsum);
                }
            }
            if (lane < kBfPerItem * G) {
                if (bf0 + lane / G < total_bf) {
                    const igd_meter_rec r = meter_finish(sq << 4, (pk & 0xFFFFu) << 2, bsum, true);
                    st16_stream(q.meter + ((size_t)bf0 * G + lane), *reinterpret_cast<const uint4 *>(&r));
                }
            } else if (lane < kBfPerItem * G + kBfPerItem) {
                const uint32_t j = lane - kBfPerItem * G;
                if (bf0 + j < total_bf) {
                    igd_bridge_rec r;
                    r.bytemean_out = (uint8_t)igd_bytemean_from_sum((int)(uint32_t)sq, IGD_FRAME);
                    r.n_open = (uint8_t)((uint32_t)bsum / kChunks);
                    r.mix_peak = (uint16_t)(pk & 0xFFFFu);
                    q.bmeter[bf0 + j] = r;
                }
            }
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------
// Warp-autonomous fused kernel for ANY leg count (1..IGD_MAX_LEGS), e.g. the 32 inbound calls of a
// CLIENT-mode softphone (roip_ed137.cpp:141-150).  Same building blocks as k_fused_w; what changes:
//   * a warp's item is 3 consecutive bridge-frames, lane = ONE 16-sample chunk (10 lanes per
//     bridge-frame, 30 of 32 lanes busy), so only 16 mix accumulators live across the legs;
//   * the legs are walked in groups of four: per group the warp fetches 3 x (4 legs x 160 B) with
//     three bulk copies (the groups of different bridge-frames are G*160 bytes apart) into a
//     double-buffered, bank-conflict-free slot (bridge-frame stride 672 B == 2 mod 8 sixteenths),
//     the next group is in flight while the current one is processed;
//   * after every group lanes 0..11 finish that group's leg records; after the last group the mix
//     is saturated / stored / compressed and lanes 0..2 finish the bridge records.
constexpr int kGBf = 3;                    // bridge-frames per item
constexpr int kGLegs = 4;                  // legs per group
constexpr int kGStride = kGLegs * IGD_FRAME + 32;          // 672
constexpr int kGSlotBytes = kGBf * kGStride;                // 2016
constexpr int kGParts = (kGBf * kGLegs + kGBf) * kPst;      // leg + bridge partials per warp

template <bool kSigned, int kWarps>
__global__ void __launch_bounds__(kWarps * 32, 1) k_fused_g(const FusedParams q)
{
    __shared__ uint64_t bars[kWarps * 2];
    __shared__ __align__(16) uint32_t enc_tab[2][8];
    extern __shared__ __align__(128) uint8_t smem[];
    uint32_t *lut = reinterpret_cast<uint32_t *>(smem);
    const uint32_t lut_bytes = shared_addr(smem);
    const int t = threadIdx.x;
    const uint32_t lane = t & 31;
    const uint32_t warp = __shfl_sync(0xFFFFFFFFu, (uint32_t)t >> 5, 0);
    const uint32_t slot_s = shared_addr(smem + kLutBytes) + warp * (2 * kGSlotBytes);
    uint2 *part = reinterpret_cast<uint2 *>(smem + kLutBytes + (size_t)kWarps * 2 * kGSlotBytes) + (size_t)warp * kGParts;
    uint2 *bpart = part + kGBf * kGLegs * kPst;
    const uint32_t bar_s = shared_addr(bars) + warp * 16;

    build_decode_lut_abs(lut, t, kWarps * 32);
    if (t < 2) {
        const enc_pk e = enc_pk_make(t);
        enc_tab[t][0] = e.bias_pos; enc_tab[t][1] = e.bias_x; enc_tab[t][2] = e.hi_pos; enc_tab[t][3] = e.hi_x;
        enc_tab[t][4] = e.thr; enc_tab[t][5] = e.mask4;
    }
    if (lane == 0) {
        mbar_init(bar_s, 1); mbar_init(bar_s + 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const uint32_t G = (uint32_t)q.G, ngrp = (G + kGLegs - 1) / kGLegs;
    const uint32_t total_bf = (uint32_t)q.total_bf;
    const uint32_t items = (total_bf + kGBf - 1) / kGBf;
    const uint32_t nw = gridDim.x * kWarps;
    uint32_t item = warp * gridDim.x + blockIdx.x;
    // unit = (item, leg group); the n-th unit of this warp lives in slot n & 1
    auto fetch = [&](uint32_t it_idx, uint32_t grp, uint32_t n) {
        const uint32_t bf0 = it_idx * kGBf;
        const uint32_t left = total_bf - bf0;
        const uint32_t nbf = left < (uint32_t)kGBf ? left : (uint32_t)kGBf;
        const uint32_t legs = min((uint32_t)kGLegs, G - grp * kGLegs), bytes = legs * IGD_FRAME;
        if (lane == 0) {
            const uint32_t bs = bar_s + (n & 1u) * 8, dst = slot_s + (n & 1u) * kGSlotBytes;
            mbar_expect_tx(bs, nbf * bytes);
            const uint8_t *g = q.codes + ((size_t)bf0 * G + (size_t)grp * kGLegs) * IGD_FRAME;
#pragma unroll
            for (int k = 0; k < kGBf; k++)
                if ((uint32_t)k < nbf) bulk_g2s(dst + k * kGStride, g + (size_t)k * G * IGD_FRAME, bytes, bs);
        }
    };
    const bool worker = lane < kGBf * kChunks;
    const uint32_t bfl = worker ? lane / kChunks : 0u, c = worker ? lane - bfl * kChunks : 0u;
    const uint32_t src_off = bfl * kGStride + c * 16;
    const uint32_t lane4 = lut_bytes + 4u * lane;
    const uint32_t b_step = (uint32_t)(((unsigned long long)nw * kGBf) % (uint32_t)q.B);
    uint32_t b = (item * kGBf + bfl) % (uint32_t)q.B;
    // gains (two packed words) and laws (4 bits) of one unit for this lane's bridge-frame
    auto load_unit = [&](uint32_t it_idx, uint32_t grp, uint32_t bb, uint2 &gq, uint32_t &lw) {
        gq = make_uint2(0u, 0u); lw = 0u;
        const uint32_t bf = it_idx * kGBf + bfl;
        if (!worker || it_idx >= items || bf >= total_bf) return;
        const uint32_t legs = min((uint32_t)kGLegs, G - grp * kGLegs);
        const uint16_t *gp = q.gain + (size_t)bf * G + grp * kGLegs;
        const uint8_t *lp = q.law + (size_t)bb * G + grp * kGLegs;
        uint32_t g0 = 0, g1 = 0, g2 = 0, g3 = 0;
        if (legs > 0) { g0 = __ldg(gp + 0); lw |= (uint32_t)(__ldg(lp + 0) & 1u); }
        if (legs > 1) { g1 = __ldg(gp + 1); lw |= (uint32_t)(__ldg(lp + 1) & 1u) << 1; }
        if (legs > 2) { g2 = __ldg(gp + 2); lw |= (uint32_t)(__ldg(lp + 2) & 1u) << 2; }
        if (legs > 3) { g3 = __ldg(gp + 3); lw |= (uint32_t)(__ldg(lp + 3) & 1u) << 3; }
        gq = make_uint2(g0 | (g1 << 16), g2 | (g3 << 16));
    };
    uint32_t n = 0;                       // units fetched so far == index of the unit being fetched next
    if (item < items) fetch(item, 0, n);
    uint2 gq; uint32_t lwq;
    load_unit(item, 0, b, gq, lwq);

    for (; item < items; item += nw) {
        const uint32_t bf = item * kGBf + bfl;
        const bool valid = worker && bf < total_bf;
        const uint32_t b_next = (b + b_step >= (uint32_t)q.B) ? b + b_step - (uint32_t)q.B : b + b_step;
        const uint32_t olaw = valid ? (uint32_t)(__ldg(q.out_law + b) & 1u) : 0u;
        int acc[16];
#pragma unroll
        for (int i = 0; i < 16; i++) acc[i] = 0;
        uint32_t n_open = 0;
#pragma unroll 1
        for (uint32_t grp = 0; grp < ngrp; grp++, n++) {
            const uint32_t legs = min((uint32_t)kGLegs, G - grp * kGLegs);
            const bool last = grp + 1 == ngrp;
            const uint32_t it_n = last ? item + nw : item, grp_n = last ? 0u : grp + 1;
            if (it_n < items) fetch(it_n, grp_n, n + 1);                 // the other slot was drained one unit ago
            mbar_wait(bar_s + (n & 1u) * 8, (n >> 1) & 1u);
            const uint2 gcur = gq;
            const uint32_t lcur = lwq;
            load_unit(it_n, grp_n, last ? b_next : b, gq, lwq);          // next unit's gains / laws ride in registers
            const uint32_t orx = __reduce_or_sync(0xFFFFFFFFu, gcur.x), ory = __reduce_or_sync(0xFFFFFFFFu, gcur.y);
            const bool general = ((orx | ory) & 0xFEFFFEFFu) != 0u;
            const uint32_t open_mask = ((orx & 0xFFFFu) ? 1u : 0u) | ((orx >> 16) ? 2u : 0u) | ((ory & 0xFFFFu) ? 4u : 0u) |
                                       ((ory >> 16) ? 8u : 0u);
            auto adj_of = [&](int g) -> uint32_t { return g == 0 ? (gcur.x & 0xFFFFu) : g == 1 ? (gcur.x >> 16) : g == 2 ? (gcur.y & 0xFFFFu) : (gcur.y >> 16); };
            n_open += (uint32_t)(__popc(nonzero_halves(gcur.x)) + __popc(nonzero_halves(gcur.y)));
            const uint32_t src = slot_s + (n & 1u) * kGSlotBytes + src_off;
            uint4 wh[kGLegs];
#pragma unroll
            for (int g = 0; g < kGLegs; g++)
                asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                             : "=r"(wh[g].x), "=r"(wh[g].y), "=r"(wh[g].z), "=r"(wh[g].w)
                             : "r"(src + g * IGD_FRAME));
#pragma unroll
            for (int g = 0; g < kGLegs; g++) {
                if ((uint32_t)g < legs) {                                 // warp-uniform
                    const uint32_t lb = lane4 + (((lcur >> g) & 1u) << 15);
                    uint2 ph;
                    if (general) ph = leg_chunk_u<kSigned, 2>(lb, wh[g], 0u, (int)adj_of(g), acc);
                    else if ((open_mask >> g) & 1u) ph = leg_chunk_u<kSigned, 1>(lb, wh[g], adj_of(g), 0, acc);
                    else ph = leg_chunk_u<kSigned, 0>(lb, wh[g], 0u, 0, acc);
                    if (valid) part[(bfl * kGLegs + g) * kPst + c] = ph;
                }
            }
            __syncwarp();
            if (lane < kGBf * kGLegs) {                                  // this group's leg records
                const uint32_t fb = lane / kGLegs, g = lane - fb * kGLegs;
                if (g < legs && item * kGBf + fb < total_bf) {
                    const uint2 *src_p = part + lane * kPst;
                    unsigned long long sq = 0; uint32_t pk = 0; int bsum = 0;
#pragma unroll
                    for (int i = 0; i < kChunks; i++) {
                        const uint2 v = src_p[i];
                        sq += v.x; pk = max_u16x2(pk, v.y); bsum = dp2a_lo(v.y, 0x0100u, bsum);
                    }
                    const igd_meter_rec r = meter_finish(sq << 4, (pk & 0xFFFFu) << 2, bsum, true);
                    st16_stream(q.meter + ((size_t)(item * kGBf + fb) * G + grp * kGLegs + g), *reinterpret_cast<const uint4 *>(&r));
                }
            }
            __syncwarp();
        }
        // ---- bridge output of this lane's chunk
        enc_pk E;
        {
            const uint32_t *et = enc_tab[olaw];
            const uint4 e0 = *reinterpret_cast<const uint4 *>(et);
            const uint2 e1 = *reinterpret_cast<const uint2 *>(et + 4);
            E.bias_pos = e0.x; E.bias_x = e0.y; E.hi_pos = e0.z; E.hi_x = e0.w; E.thr = e1.x; E.mask4 = e1.y;
        }
        const size_t o16 = (size_t)bf * kChunks + c;
        const uint2 mo = mix_out_chunk<kSigned>(acc, E, q.mix + o16 * 16, q.enc + o16 * 16, valid);
        if (valid) bpart[bfl * kPst + c] = make_uint2(mo.x, mo.y | (n_open << 16));
        __syncwarp();
        if (lane < kGBf && item * kGBf + lane < total_bf) {
            const uint2 *src_p = bpart + lane * kPst;
            int esum = 0; uint32_t pk = 0;
#pragma unroll
            for (int i = 0; i < kChunks; i++) { esum += (int)src_p[i].x; pk = max(pk, src_p[i].y & 0xFFFFu); }
            igd_bridge_rec r;
            r.bytemean_out = (uint8_t)igd_bytemean_from_sum(esum, IGD_FRAME);
            r.n_open = (uint8_t)(src_p[0].y >> 16);
            r.mix_peak = (uint16_t)pk;
            q.bmeter[item * kGBf + lane] = r;
        }
        __syncwarp();
        b = b_next;
    }
}

// Any number of legs per bridge (1..IGD_MAX_LEGS): same algorithm, legs walked
// in a loop with the partials reduced per leg through shared memory.
template <int BFPC, bool kSigned>
__global__ void __launch_bounds__(BFPC * kChunks, 2) k_fused_anyg(const FusedParams q)
{
    extern __shared__ __align__(128) uint8_t smem[];
    uint32_t *lut = reinterpret_cast<uint32_t *>(smem);
    uint2 *part = reinterpret_cast<uint2 *>(smem + kLutBytes);       // [BFPC][kPst]
    uint2 *bpart = part + BFPC * kPst;                               // [BFPC][kPst]
    const uint32_t lut_bytes = shared_addr(smem);
    const int t = threadIdx.x, G = q.G;
    const uint32_t lane = t & 31;
    build_decode_lut(lut, t, BFPC * kChunks);
    __syncthreads();
    const int bfl = t / kChunks, p = t - bfl * kChunks;
    for (long long tile = blockIdx.x; tile < q.num_tiles; tile += gridDim.x) {
        const long long bf = tile * BFPC + bfl;
        const bool valid = bf < q.total_bf;
        const int b = valid ? (int)(bf % q.B) : 0;
        int acc[16];
#pragma unroll
        for (int i = 0; i < 16; i++) acc[i] = 0;
        for (int g = 0; g < G; g++) {
            if (valid) {
                const uint4 w = ld16_stream(q.codes + ((size_t)bf * G + g) * IGD_FRAME + p * 16);
                const uint32_t a = q.gain[(size_t)bf * G + g];
                const uint32_t sl = gain_selector(a), lb = lut_lane_base(lut_bytes, lane, q.law[(size_t)b * G + g]);
                part[bfl * kPst + p] = sl == kSelGeneral ? leg_chunk<kSigned, 2>(lb, w, 0u, (int)a, acc)
                                       : sl             ? leg_chunk<kSigned, 1>(lb, w, sl, 0, acc)
                                                        : leg_chunk<kSigned, 0>(lb, w, 0u, 0, acc);
            }
            __syncthreads();
            if (t < BFPC) {
                const long long bf2 = tile * BFPC + t;
                if (bf2 < q.total_bf) {
                    unsigned long long sq = 0; uint32_t peak = 0; int bsum = 0;
#pragma unroll
                    for (int i = 0; i < kChunks; i++) partial_add(part[t * kPst + i], sq, peak, bsum);
                    const igd_meter_rec r = meter_finish(sq << 4, peak << 2, bsum, true);
                    st16_stream(q.meter + bf2 * G + g, *reinterpret_cast<const uint4 *>(&r));
                }
            }
            __syncthreads();
        }
        if (valid)
            bpart[bfl * kPst + p] = mix_out_chunk<kSigned>(acc, enc_pk_make(q.out_law[b]),
                                                           q.mix + (size_t)bf * IGD_FRAME + p * 16,
                                                           q.enc + (size_t)bf * IGD_FRAME + p * 16);
        __syncthreads();
        if (t < BFPC) {
            const long long bf2 = tile * BFPC + t;
            if (bf2 < q.total_bf) {
                int esum = 0, mpeak = 0;
#pragma unroll
                for (int i = 0; i < kChunks; i++) { esum += (int)bpart[t * kPst + i].x; mpeak = max(mpeak, (int)bpart[t * kPst + i].y); }
                int n_open = 0;
                for (int g = 0; g < G; g++) n_open += q.gain[(size_t)bf2 * G + g] != 0;
                igd_bridge_rec r;
                r.bytemean_out = (uint8_t)igd_bytemean_from_sum(esum, IGD_FRAME);
                r.n_open = (uint8_t)n_open;
                r.mix_peak = (uint16_t)mpeak;
                q.bmeter[bf2] = r;
            }
        }
        __syncthreads();
    }
}

// ============================================================ stand-alone G.711
// codes -> PCM.  One thread per 16 codes.  law_ch != nullptr: law of chunk i is
// law_ch[(i / 10) % nch] (frames of 160 samples laid out [frame][channel]).
__global__ void __launch_bounds__(512) k_g711_decode(const uint8_t *__restrict__ codes,
                                                     const uint8_t *__restrict__ law_ch, int law,
                                                     int16_t *__restrict__ pcm, size_t n, size_t nch)
{
    extern __shared__ __align__(128) uint8_t smem[];
    build_decode_lut(reinterpret_cast<uint32_t *>(smem), threadIdx.x, blockDim.x);
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t lut_s = shared_addr(smem);
    const size_t nchunk = n / 16;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nchunk;
         i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t lw = law_ch ? law_ch[(i / kChunks) % nch] : (uint32_t)law;
        const uint32_t lb = lut_lane_base(lut_s, lane, lw);
        const uint4 w = ld16_stream(codes + i * 16);
        const uint32_t wd[4] = {w.x, w.y, w.z, w.w};
        uint32_t pk[8];
#pragma unroll
        for (int j = 0; j < 4; j++) {      // low halves of the table entries are the samples / 4
            const uint32_t e0 = lut_lookup<0>(lb, wd[j]), e1 = lut_lookup<1>(lb, wd[j]);
            const uint32_t e2 = lut_lookup<2>(lb, wd[j]), e3 = lut_lookup<3>(lb, wd[j]);
            pk[2 * j] = (__byte_perm(e0, e1, 0x5410) << 2) & 0xFFFCFFFCu;
            pk[2 * j + 1] = (__byte_perm(e2, e3, 0x5410) << 2) & 0xFFFCFFFCu;
        }
        st32_stream(pcm + i * 16, pk);
    }
    // ragged tail (< 16 codes): straight formula, one thread
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        for (size_t i = nchunk * 16; i < n; i++) {
            const uint32_t lw = law_ch ? law_ch[(i / IGD_FRAME) % nch] : (uint32_t)law;
            pcm[i] = (int16_t)(lw ? igd_ulaw2lin(codes[i]) : igd_alaw2lin(codes[i]));
        }
    }
}

// PCM -> codes.  One thread per 16 samples (256-bit load, 128-bit store).
__global__ void __launch_bounds__(256) k_g711_encode(const int16_t *__restrict__ pcm,
                                                     const uint8_t *__restrict__ law_ch, int law,
                                                     uint8_t *__restrict__ codes, size_t n, size_t nch)
{
    const size_t nchunk = n / 16;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nchunk;
         i += (size_t)gridDim.x * blockDim.x) {
        const int lw = law_ch ? law_ch[(i / kChunks) % nch] : law;
        uint32_t pk[8];
        ld32_stream(pcm + i * 16, pk);
        st16_stream(codes + i * 16, encode16_packed(pk, enc_pk_make(lw)));
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        for (size_t i = nchunk * 16; i < n; i++) {
            const int lw = law_ch ? law_ch[(i / IGD_FRAME) % nch] : law;
            codes[i] = (uint8_t)igd_g711_enc1(pcm[i], igd_enc_law_make(lw));
        }
    }
}

// ============================================================ stand-alone meter
// PCM frames -> records.  10 threads per frame, 16 samples (32 B) each.
template <int FPC>
__global__ void __launch_bounds__(FPC * kChunks) k_frame_meter(const int16_t *__restrict__ pcm,
                                                               long long nframes, long long num_tiles,
                                                               igd_meter_rec *__restrict__ out)
{
    __shared__ uint4 part[FPC * kPst];   // {sq_lo, sq_hi, peak, -}: general int16 needs 40 bits
    const int t = threadIdx.x, fl = t / kChunks, p = t - fl * kChunks;
    for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const long long f = tile * FPC + fl;
        if (f < nframes) {
            uint32_t pk[8];
            ld32_stream(pcm + (size_t)f * IGD_FRAME + p * 16, pk);
            unsigned long long sq = 0;
            int peak = 0;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const int a = (int)(short)(pk[j] & 0xFFFFu), b = (int)pk[j] >> 16;
                sq += (unsigned long long)(uint32_t)(a * a) + (unsigned long long)(uint32_t)(b * b);
                peak = max(peak, max(abs(a), abs(b)));
            }
            part[fl * kPst + p] = make_uint4((uint32_t)sq, (uint32_t)(sq >> 32), (uint32_t)peak, 0u);
        }
        __syncthreads();
        if (t < FPC) {
            const long long f2 = tile * FPC + t;
            if (f2 < nframes) {
                unsigned long long sq = 0; uint32_t peak = 0;
#pragma unroll
                for (int i = 0; i < kChunks; i++) {
                    const uint4 v = part[t * kPst + i];
                    sq += (unsigned long long)v.x | ((unsigned long long)v.y << 32);
                    peak = max(peak, v.z);
                }
                const igd_meter_rec r = meter_finish(sq, peak, 0, false);
                st16_stream(out + f2, *reinterpret_cast<const uint4 *>(&r));
            }
        }
        __syncthreads();
    }
}

// reference per-packet level over arbitrary (stride, len): one warp per payload
__global__ void __launch_bounds__(256) k_bytemean(const uint8_t *__restrict__ base, size_t n, size_t len,
                                                  size_t stride, unsigned flags, uint8_t *__restrict__ out)
{
    const uint32_t lane = threadIdx.x & 31;
    const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const size_t nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
    const bool sc = (flags & IGD_F_SIGNED_CHAR) != 0;
    for (size_t i = warp; i < n; i += nwarps) {
        const uint8_t *p = base + i * stride;
        int s = 0;
        for (size_t k = lane; k < len; k += 32) s += sc ? (int)(signed char)p[k] : (int)p[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) out[i] = (uint8_t)igd_bytemean_from_sum(s, (int)len);
    }
}

// audiometer.cpp:30-31: int(float(v*100.0/30000.0))
__global__ void k_level_percent(const int32_t *__restrict__ v, size_t n, int32_t *__restrict__ out)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x)
        out[i] = (int)((float)(((double)v[i] * 100.0) / 30000.0));
}

// ============================================================ stand-alone mix
// PCM legs -> saturating mix.  One thread per 8 samples of one bridge-frame.
__global__ void __launch_bounds__(256) k_mix(const int16_t *__restrict__ pcm, const uint16_t *__restrict__ gain,
                                             long long total_bf, int G, int16_t *__restrict__ mix)
{
    constexpr int kPer = IGD_FRAME / 8;   // 20 threads per bridge-frame
    const long long nthreads_total = total_bf * kPer;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nthreads_total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long bf = i / kPer;
        const int p = (int)(i - bf * kPer);
        int acc[8];
#pragma unroll
        for (int k = 0; k < 8; k++) acc[k] = 0;
        for (int g = 0; g < G; g++) {
            const int a = gain[bf * G + g];
            if (a == 0) continue;
            const uint4 w = ld16_stream(pcm + ((size_t)bf * G + g) * IGD_FRAME + p * 8);
            const uint32_t wd[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int lo = (int)(short)(wd[j] & 0xFFFFu), hi = (int)wd[j] >> 16;
                acc[2 * j] += clamp16((lo * a) >> 7);
                acc[2 * j + 1] += clamp16((hi * a) >> 7);
            }
        }
        uint4 o;
        o.x = pack_sat16(acc[1], acc[0]); o.y = pack_sat16(acc[3], acc[2]);
        o.z = pack_sat16(acc[5], acc[4]); o.w = pack_sat16(acc[7], acc[6]);
        st16_stream(mix + (size_t)bf * IGD_FRAME + p * 8, o);
    }
}

// ============================================================ event summary
// Functions.cpp:2126-2145 over the frames whose gate is open; one warp per 32
// channels x frame slice, slices combined through shared memory.
constexpr int kSumSlices = 32;
__global__ void __launch_bounds__(32 * kSumSlices) k_event_summary(const igd_meter_rec *__restrict__ meter,
                                                                   const uint16_t *__restrict__ gain,
                                                                   long long F, long long C,
                                                                   igd_summary_rec *__restrict__ out,
                                                                   igd_summary_db *__restrict__ db)
{
    __shared__ igd_summary_rec sh[kSumSlices][32];
    const int lane = threadIdx.x & 31, slice = threadIdx.x >> 5;
    const long long c = (long long)blockIdx.x * 32 + lane;
    igd_summary_rec r;
    r.count = 0; r.bm_sum = 0; r.bm_max = 0; r.bm_min = 255;          // Functions.cpp:2159-2167
    r.sum_s = 0; r.max_s = 0; r.min_s = 255ull * IGD_FRAME;
    if (c < C) {
        constexpr int kAhead = 4;                    // four frames of this slice in flight
        for (long long f0 = slice; f0 < F; f0 += (long long)kSumSlices * kAhead) {
            uint32_t g[kAhead];
            uint4 mm[kAhead];
#pragma unroll
            for (int u = 0; u < kAhead; u++) {
                const long long f = f0 + (long long)u * kSumSlices;
                g[u] = f < F ? (uint32_t)gain[(size_t)f * C + c] : 0u;
            }
#pragma unroll
            for (int u = 0; u < kAhead; u++) {
                const long long f = f0 + (long long)u * kSumSlices;
                if (g[u]) mm[u] = __ldcs(reinterpret_cast<const uint4 *>(meter + (size_t)f * C + c));
            }
#pragma unroll
            for (int u = 0; u < kAhead; u++) {
                if (!g[u]) continue;
                const uint4 m = mm[u];
                const uint64_t s = (uint64_t)m.x | ((uint64_t)(m.y & 0xFFu) << 32);
                const uint32_t bm = (m.y >> 8) & 0xFFu;
                r.count += 1;
                r.sum_s += s;
                r.bm_sum = (uint16_t)(r.bm_sum + bm);
                r.max_s = s > r.max_s ? s : r.max_s;
                r.min_s = s < r.min_s ? s : r.min_s;
                r.bm_max = (uint8_t)max((uint32_t)r.bm_max, bm);
                r.bm_min = (uint8_t)min((uint32_t)r.bm_min, bm);
            }
        }
    }
    sh[slice][lane] = r;
    __syncthreads();
    if (slice == 0 && c < C) {
        for (int k = 1; k < kSumSlices; k++) {
            const igd_summary_rec o = sh[k][lane];
            r.count += o.count;
            r.sum_s += o.sum_s;
            r.bm_sum = (uint16_t)(r.bm_sum + o.bm_sum);
            r.max_s = o.max_s > r.max_s ? o.max_s : r.max_s;
            r.min_s = o.min_s < r.min_s ? o.min_s : r.min_s;
            r.bm_max = o.bm_max > r.bm_max ? o.bm_max : r.bm_max;
            r.bm_min = o.bm_min < r.bm_min ? o.bm_min : r.bm_min;
        }
        out[c] = r;
        if (db) {
            igd_summary_db d;                                             // Functions.cpp:2196-2200
            d.level_av_db = (float)(10.0 * log10(((double)r.sum_s / IGD_FRAME) / (double)r.count));
            d.level_max_db = (float)(10.0 * log10((double)r.max_s / IGD_FRAME));
            d.level_min_db = (float)(10.0 * log10((double)r.min_s / IGD_FRAME));
            d.bm_av = r.count ? (uint32_t)(uint8_t)(r.bm_sum / r.count) : 0u;
            db[c] = d;
        }
    }
}

// ============================================================ ED-137 parse
// transport_rtp_cb (TransportAdapter.cpp:240-316) + field getters
// (Functions.cpp:1001-1179).  One warp per packet: lanes 0..4 fetch the five
// header words, every lane copies payload words.
__device__ __forceinline__ uint32_t bswap32(uint32_t v) { return __byte_perm(v, 0, 0x0123); }

// Tile form of the receive parse for the wire stride (180 B): a CTA stages 64 packets
// (11 520 contiguous bytes) in shared memory with 16-byte loads, 64 threads extract the fields,
// all threads write the 64 x 160 payload bytes with 16-byte stores.  Same results as
// k_ed137_parse below (which keeps serving other strides / unaligned buffers).
constexpr int kPktTile = 64, kPktWords = IGD_PKT_MAX / 4;      // 45 words per packet
__global__ void __launch_bounds__(256) k_ed137_parse_tile(const uint8_t *__restrict__ pkts,
                                                          const uint32_t *__restrict__ sizes, size_t npkts,
                                                          igd_ed137_fields *__restrict__ fields,
                                                          uint8_t *__restrict__ payload_out)
{
    __shared__ __align__(128) uint32_t img[kPktTile * kPktWords];
    __shared__ uint32_t plen_s[kPktTile];
    const size_t first = (size_t)blockIdx.x * kPktTile;
    const uint32_t np = (uint32_t)min((size_t)kPktTile, npkts - first);
    const uint4 *src = reinterpret_cast<const uint4 *>(pkts + first * IGD_PKT_MAX);
    const uint32_t tile_bytes = np * IGD_PKT_MAX;
    if ((tile_bytes & 15u) == 0) {
        // full tiles (and any tail whose size is a multiple of 16): ONE bulk async copy (TMA) stages the
        // tile, completion on the CTA's mbarrier -- no registers, no LSU issue slots
        __shared__ uint64_t bar;
        const uint32_t bar_s = shared_addr(&bar);
        if (threadIdx.x == 0) {
            mbar_init(bar_s, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            mbar_expect_tx(bar_s, tile_bytes);
            bulk_g2s(shared_addr(img), src, tile_bytes, bar_s);
        }
        __syncthreads();                 // the barrier is initialised before anyone polls it
        mbar_wait(bar_s, 0);
    } else {
        const uint32_t nvec = (tile_bytes + 15) / 16;
        for (uint32_t j = threadIdx.x; j < nvec; j += blockDim.x) {
            uint4 v;
            if ((j + 1) * 16 <= tile_bytes) v = __ldcs(src + j);
            else {                                               // ragged tail of the whole buffer: word loads
                const uint32_t *w = reinterpret_cast<const uint32_t *>(src + j);
                const uint32_t left = (tile_bytes - j * 16) / 4;
                v.x = left > 0 ? w[0] : 0u; v.y = left > 1 ? w[1] : 0u; v.z = left > 2 ? w[2] : 0u; v.w = 0u;
            }
            reinterpret_cast<uint4 *>(img)[j] = v;
        }
        __syncthreads();
    }
    if (threadIdx.x < np) {
        const uint32_t p = threadIdx.x;
        const size_t i = first + p;
        const uint32_t size = sizes ? sizes[i] : (uint32_t)IGD_PKT_MAX;
        const uint32_t navail = min(size, (uint32_t)IGD_PKT_MAX) / 4;
        const uint32_t *hw = img + p * kPktWords;
        const uint32_t w0 = navail > 0 ? hw[0] : 0u, w3 = navail > 3 ? hw[3] : 0u, w4 = navail > 4 ? hw[4] : 0u;
        const uint32_t pt = (w0 >> 8) & 0x7Fu;
        const bool too_short = size < IGD_PKT_HDR;
        const uint32_t plen_raw = size - IGD_PKT_HDR;
        const bool dropped = too_short || plen_raw >= 1024u;
        const bool accepted = !too_short && (pt == 8 || pt == 0 || pt == 18 || pt == 123);
        const uint32_t plen = dropped ? 0u : min(plen_raw, (uint32_t)IGD_FRAME);
        plen_s[p] = plen;
        igd_ed137_fields f;
        const uint32_t word = accepted ? bswap32(w4) : 0u;
        const igd_edf e = igd_ed137_fields_of(word);
        f.word = word;
        f.length_raw = accepted ? (uint16_t)(w3 >> 16) : (uint16_t)0;
        f.payload_len = (uint16_t)plen;
        f.pt = (uint8_t)pt;
        f.accepted = accepted;
        f.keepalive = (!too_short && pt == 123);
        f.ptt_type = (uint8_t)e.ptt_type;
        f.ptt_id = (uint8_t)e.ptt_id;
        f.squelch = (uint8_t)e.squelch;
        f.bss = (uint8_t)e.bss;
        f.flags = (uint8_t)(e.flags | (dropped ? IGD_EDF_DROPPED : 0u));
        *reinterpret_cast<uint4 *>(fields + i) = *reinterpret_cast<const uint4 *>(&f);
    }
    if (!payload_out) return;
    __syncthreads();
    uint4 *dst = reinterpret_cast<uint4 *>(payload_out + first * IGD_FRAME);
    for (uint32_t j = threadIdx.x; j < np * kChunks; j += blockDim.x) {
        const uint32_t p = j / kChunks, ch = j - p * kChunks;
        const uint32_t plen = plen_s[p], b0 = ch * 16;
        const uint32_t *w = img + p * kPktWords + 5 + ch * 4;
        uint32_t v[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const uint32_t b = b0 + 4 * k;
            v[k] = b < plen ? w[k] : 0u;
            if (b < plen && plen - b < 4) v[k] &= (1u << (8 * (plen - b))) - 1u;
        }
        __stcs(dst + j, make_uint4(v[0], v[1], v[2], v[3]));
    }
}

__global__ void __launch_bounds__(256) k_ed137_parse(const uint8_t *__restrict__ pkts,
                                                     const uint32_t *__restrict__ sizes, size_t npkts,
                                                     size_t stride, igd_ed137_fields *__restrict__ fields,
                                                     uint8_t *__restrict__ payload_out)
{
    const uint32_t lane = threadIdx.x & 31;
    const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const size_t nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
    for (size_t i = warp; i < npkts; i += nwarps) {
        const uint32_t *pw = reinterpret_cast<const uint32_t *>(pkts + i * stride);
        const uint32_t size = sizes ? sizes[i] : (uint32_t)stride;
        const uint32_t navail = min(size, (uint32_t)stride) / 4;
        const uint32_t hw = (lane < 5 && lane < navail) ? __ldcs(pw + lane) : 0u;
        const uint32_t w0 = __shfl_sync(0xffffffffu, hw, 0);
        const uint32_t w3 = __shfl_sync(0xffffffffu, hw, 3);
        const uint32_t w4 = __shfl_sync(0xffffffffu, hw, 4);
        const uint32_t pt = (w0 >> 8) & 0x7Fu;                           // byte 1, low 7 bits
        const bool too_short = size < IGD_PKT_HDR;
        const uint32_t plen_raw = size - IGD_PKT_HDR;                     // unsigned wrap (:279)
        const bool dropped = too_short || plen_raw >= 1024u;              // :286-291
        const bool accepted = !too_short && (pt == 8 || pt == 0 || pt == 18 || pt == 123);   // :252
        const uint32_t plen = dropped ? 0u : min(plen_raw, (uint32_t)IGD_FRAME);
        if (lane == 0) {
            igd_ed137_fields f;
            const uint32_t word = accepted ? bswap32(w4) : 0u;            // ntohl (:342)
            const igd_edf e = igd_ed137_fields_of(word);
            f.word = word;
            f.length_raw = accepted ? (uint16_t)(w3 >> 16) : (uint16_t)0; // bytes 14..15 as stored
            f.payload_len = (uint16_t)plen;
            f.pt = (uint8_t)pt;
            f.accepted = accepted;
            f.keepalive = (!too_short && pt == 123);
            f.ptt_type = (uint8_t)e.ptt_type;
            f.ptt_id = (uint8_t)e.ptt_id;
            f.squelch = (uint8_t)e.squelch;
            f.bss = (uint8_t)e.bss;
            f.flags = (uint8_t)(e.flags | (dropped ? IGD_EDF_DROPPED : 0u));
            *reinterpret_cast<uint4 *>(fields + i) = *reinterpret_cast<const uint4 *>(&f);
        }
        if (payload_out) {
            uint32_t *dst = reinterpret_cast<uint32_t *>(payload_out + i * IGD_FRAME);
            for (uint32_t k = lane; k < IGD_FRAME / 4; k += 32) {
                uint32_t v = 0;
                if (k * 4 < plen) {
                    v = __ldcs(pw + 5 + k);
                    const uint32_t rem = plen - k * 4;
                    if (rem < 4) v &= (1u << (8 * rem)) - 1u;
                }
                dst[k] = v;
            }
        }
    }
}

// ============================================================ ED-137 pack
// ============================================================ RX liveness walk
// One thread per channel walks its frames through the receive-side state of
// transport_rtp_cb (TransportAdapter.cpp:240-316) and the R2S watchdog
// (roip_ed137.cpp:1767-1780).  Consecutive threads read consecutive 16-byte field
// records and write consecutive 8-byte events: coalesced, latency-bound.
__global__ void __launch_bounds__(128) k_rx_track(const igd_rx_track_desc d)
{
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= d.C) return;
    igd_rx_state s = d.state[c];
    constexpr int kAhead = 8;            // the walk is sequential, its inputs are not: fetch 8 frames ahead
    for (int f0 = 0; f0 < d.F; f0 += kAhead) {
        uint4 raw[kAhead];
        uint8_t pres[kAhead];
#pragma unroll
        for (int u = 0; u < kAhead; u++) {
            const int f = f0 + u;
            if (f < d.F) {
                const size_t i = (size_t)f * d.C + c;
                raw[u] = __ldg(reinterpret_cast<const uint4 *>(d.fields + i));
                pres[u] = d.present ? d.present[i] : (uint8_t)1;
            }
        }
#pragma unroll
        for (int u = 0; u < kAhead; u++) {
            const int f = f0 + u;
            if (f >= d.F) break;
            const size_t i = (size_t)f * d.C + c;
            const igd_ed137_fields fl = *reinterpret_cast<const igd_ed137_fields *>(&raw[u]);
            const bool wd = d.wd_ticks > 0 && ((d.frame0 + f) % d.wd_ticks) == d.wd_ticks - 1;
            const uint32_t ev = igd_rx_step(s, fl, pres[u] != 0, wd, d.now_ms0 + (long long)f * d.tick_ms, d.r2s_period_ms);
            igd_rx_event e;
            e.word = s.ed137_value;
            e.flags = (uint8_t)ev;
            e.r2sCount = s.r2sCount;
            e.reserved = 0;
            *reinterpret_cast<uint2 *>(d.events + i) = *reinterpret_cast<const uint2 *>(&e);
        }
    }
    d.state[c] = s;
}

// ============================================================ gate arbitration
// One thread per bridge walks its frames through checkEvents()'s gate decisions
// (igd_math.cuh: igd_arb_client_tick / igd_arb_server_best_tick); the leg state of
// the block's bridges lives in shared memory for the walk.
constexpr int kArbThreads = 64;
constexpr int kArbStageWords = 4096;     // words (and gains) of a run of ticks staged per block
template <int kG> struct arb_legs {      // compile-time leg count: the leg state lives in registers
    igd_arb_leg v[kG > 0 ? kG : 1];
};
template <int kG>
__global__ void __launch_bounds__(kArbThreads) k_gate_arbitrate(const igd_arb_desc d)
{
    extern __shared__ __align__(128) uint8_t smem[];
    const int G = kG > 0 ? kG : d.G;
    const int bpb = kArbThreads;                                        // bridges per block
    uint32_t *words_s = reinterpret_cast<uint32_t *>(smem);             // [T][bpb*G]
    uint16_t *gain_s = reinterpret_cast<uint16_t *>(words_s + kArbStageWords);   // [T][bpb*G]
    igd_arb_leg *legs_s = reinterpret_cast<igd_arb_leg *>(gain_s + kArbStageWords);   // [bpb][G] (runtime G only)
    const int b0 = blockIdx.x * bpb;
    const int b = b0 + threadIdx.x;
    const int nb = min(bpb, d.B - b0);
    const int row = nb * G;                                             // words of this block per tick
    const int T = max(1, kArbStageWords / (bpb * G));                   // ticks per staged run
    const size_t Cn = (size_t)d.B * G;
    arb_legs<kG> lr;
    igd_arb_leg *legs;
    if (kG > 0) {
        legs = lr.v;
        if (b < d.B) {
#pragma unroll
            for (int g = 0; g < (kG > 0 ? kG : 1); g++) lr.v[g] = d.legs[(size_t)b * G + g];
        }
    } else {
        for (int k = threadIdx.x; k < row; k += kArbThreads) legs_s[k] = d.legs[(size_t)b0 * G + k];
        legs = legs_s + threadIdx.x * G;
    }
    igd_arb_bridge br;
    if (b < d.B) br = d.bridges[b];
    const uint8_t *act = d.active ? d.active + (size_t)b * G : nullptr;
    uint32_t act_mask = 0xFFFFFFFFu;                                    // G <= 32 legs
    if (act && b < d.B) {
        act_mask = 0;
        for (int g = 0; g < G; g++) act_mask |= (act[g] != 0 ? 1u : 0u) << g;
    }
    for (int f0 = 0; f0 < d.F; f0 += T) {
        const int nt = min(T, d.F - f0);
        __syncthreads();
        for (int t = 0; t < nt; t++) {                                  // coalesced: a tick's words are contiguous
            const uint8_t *wrow = reinterpret_cast<const uint8_t *>(d.words) + ((size_t)(f0 + t) * Cn + (size_t)b0 * G) * d.word_stride;
            for (int j = threadIdx.x; j < row; j += kArbThreads)
                words_s[t * row + j] = *reinterpret_cast<const uint32_t *>(wrow + (size_t)j * d.word_stride);
        }
        __syncthreads();
        if (b < d.B) {
            for (int t = 0; t < nt; t++) {
                const uint32_t *wt = words_s + t * row + threadIdx.x * G;
                auto word = [&](int g) { return wt[g]; };
                auto active = [&](int g) { return ((act_mask >> g) & 1u) != 0u; };
                uint16_t *gt = gain_s + t * row + threadIdx.x * G;
                if (kG > 0) {
                    const igd_const_int<(kG > 0 ? kG : 1)> Gc;
                    if (d.mode == IGD_ARB_CLIENT_PTT) igd_arb_client_tick(br, legs, Gc, word, active);
                    else igd_arb_server_best_tick(br, legs, Gc, word, active);
#pragma unroll
                    for (int g = 0; g < (kG > 0 ? kG : 1); g++) gt[g] = legs[g].gain_q7;
                } else {
                    if (d.mode == IGD_ARB_CLIENT_PTT) igd_arb_client_tick(br, legs, G, word, active);
                    else igd_arb_server_best_tick(br, legs, G, word, active);
                    for (int g = 0; g < G; g++) gt[g] = legs[g].gain_q7;
                }
            }
        }
        __syncthreads();
        for (int t = 0; t < nt; t++) {
            uint16_t *grow = d.gain_q7 + (size_t)(f0 + t) * Cn + (size_t)b0 * G;
            for (int j = threadIdx.x; j < row; j += kArbThreads) grow[j] = gain_s[t * row + j];
        }
    }
    __syncthreads();
    if (b < d.B) d.bridges[b] = br;
    if (kG > 0) {
        if (b < d.B) {
#pragma unroll
            for (int g = 0; g < (kG > 0 ? kG : 1); g++) d.legs[(size_t)b * G + g] = lr.v[g];
        }
    } else {
        for (int k = threadIdx.x; k < row; k += kArbThreads) d.legs[(size_t)b0 * G + k] = legs_s[k];
    }
}

// Phase 1: one thread per channel walks its frames through the sender state
// machine (transport_send_rtp, TransportAdapter.cpp:635-874) and writes a plan
// record per packet.  The state is tiny and strictly sequential per channel.
__global__ void __launch_bounds__(128) k_ed137_plan(const igd_ed137_pack_desc d, igd_tx_plan_rec *__restrict__ plan,
                                                    int32_t *__restrict__ last_src)
{
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= d.C) return;
    igd_ed137_state s = d.state[c];
    int32_t src = -1;
    const bool stuck = 12u + d.payload_len > 60u;
    constexpr int kAhead = 8;            // the walk is sequential, its inputs are not: fetch 8 frames ahead
    for (int f0 = 0; f0 < d.F; f0 += kAhead) {
        igd_ed137_ctl kk[kAhead];
        uint32_t a40[kAhead], a50[kAhead], a60[kAhead];
#pragma unroll
        for (int u = 0; u < kAhead; u++) {
            const int f = f0 + u;
            if (f < d.F) {
                const size_t i = (size_t)f * d.C + c;
                if (d.ctl) kk[u] = d.ctl[i];
                if (stuck) {
                    const uint8_t *pl = d.payload + i * IGD_FRAME;
                    a40[u] = pl[40 - 12]; a50[u] = pl[50 - 12]; a60[u] = pl[60 - 12];
                }
            }
        }
#pragma unroll
        for (int u = 0; u < kAhead; u++) {
            const int f = f0 + u;
            if (f >= d.F) break;
            const size_t i = (size_t)f * d.C + c;
            if (d.ctl) {                                                     // the setters, :135-213
                const igd_ed137_ctl k = kk[u];
                s.pttstatus = k.pttstatus; s.pttpriority = k.pttpriority; s.callRecorder = k.callRecorder;
                s.sqlstatus = k.sqlstatus; s.ed137_bssi = k.ed137_bssi; s.pttid = k.pttid;
            }
            if (s.radiostatus && stuck) {                                    // stuck-audio detector :657-673
                if (a40[u] == a50[u] && a40[u] == a60[u] && a40[u] == 0xd5) s.rtpFalse += 1; else s.rtpFalse = 0;
            }
            const igd_tx_plan t = igd_ed137_tx_step(s, d.payload_len, d.now_ms0 + (long long)f * d.tick_ms);
            if (t.copy_payload) src = f;
            igd_tx_plan_rec r;
            r.word = t.word;
            r.size = (uint16_t)t.size;
            r.flags = (uint8_t)(t.pt123 | (t.marker << 1) | (t.copy_payload << 2));
            r.reserved = 0;
            r.src_frame = (d.flags & IGD_F_REF_QUIRKS) ? src : f;            // quirk Q2
            plan[i] = r;
        }
    }
    d.state[c] = s;
    last_src[c] = src;
}

// sendR2SStatus (TransportAdapter.cpp:422-633): thread per channel, one timer tick
__global__ void __launch_bounds__(128) k_ed137_keepalive(uint8_t *__restrict__ hdr20, igd_ed137_state *__restrict__ state,
                                                         size_t C, long long now, uint32_t *__restrict__ sizes)
{
    const size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    igd_ed137_state s = state[c];
    uint32_t *h = reinterpret_cast<uint32_t *>(hdr20 + c * IGD_PKT_HDR);
    uint32_t w0 = h[0];
    const igd_tx_plan t = igd_ed137_r2s_step(s, now, (w0 >> 8) & 0x7Fu);
    if (t.copy_payload) {                                   // the header fields are stamped into the send buffer
        w0 |= 0x10u;                                                           // x = 1 (:495)
        w0 = (w0 & ~0x8000u) | (t.marker ? 0x8000u : 0u);                      // m (:485-493)
        if (t.pt123) w0 = (w0 & ~0x7F00u) | (123u << 8);
        h[0] = w0;
        h[3] = 0x01006701u;                                                    // 0x0167, 0x0001 big-endian
        h[4] = bswap32(t.word);
    }
    sizes[c] = t.size;
    state[c] = s;
}

// after the packets are assembled: remember the payload each adapter's send buffer ends up holding
__global__ void __launch_bounds__(256) k_ed137_stale_update(const igd_ed137_pack_desc d,
                                                            const int32_t *__restrict__ last_src)
{
    const uint32_t lane = threadIdx.x & 31;
    const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const size_t nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
    for (size_t c = warp; c < (size_t)d.C; c += nwarps) {
        const int32_t f = last_src[c];
        if (f < 0) continue;
        const uint32_t *src = reinterpret_cast<const uint32_t *>(d.payload + ((size_t)f * d.C + c) * IGD_FRAME);
        uint32_t *dst = reinterpret_cast<uint32_t *>(d.stale_payload + c * IGD_FRAME);
        for (uint32_t k = lane; k < IGD_FRAME / 4; k += 32) dst[k] = src[k];
    }
}

// Phase 2: one warp per packet assembles header + payload (4-byte words; the
// payload sits at byte 20 of a 180-byte packet, so 16-byte accesses do not apply)
__global__ void __launch_bounds__(256) k_ed137_assemble(const igd_ed137_pack_desc d,
                                                        const igd_tx_plan_rec *__restrict__ plan)
{
    const uint32_t lane = threadIdx.x & 31;
    const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const size_t nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
    const size_t npkts = (size_t)d.F * d.C;
    const bool sc = (d.flags & IGD_F_SIGNED_CHAR) != 0, quirks = (d.flags & IGD_F_REF_QUIRKS) != 0;
    const uint32_t nwords = d.payload_len / 4;
    for (size_t i = warp; i < npkts; i += nwarps) {
        const igd_tx_plan_rec r = plan[i];
        uint32_t *out = reinterpret_cast<uint32_t *>(d.pkts + i * d.out_stride);
        const uint32_t *hdr = reinterpret_cast<const uint32_t *>(d.rtp12 + i * 12);
        const size_t c = i % (size_t)d.C;
        if (lane == 0) d.sizes[i] = r.size;
        int bsum = 0;
        if (r.size == 0) {
            if (lane == 0) d.bytemean_out[i] = 0;
            continue;
        }
        if (lane < 5) {
            uint32_t v;
            if (lane < 3) {
                v = hdr[lane];
                if (lane == 0) {
                    v |= 0x10u;                                            // x = 1 (:725)
                    v = (v & ~0x8000u) | ((r.flags & 2u) ? 0x8000u : 0u);  // m (:715-723)
                    if (r.flags & 1u) v = (v & ~0x7F00u) | (123u << 8);    // pt = 123
                }
            } else if (lane == 3) {
                v = 0x01006701u;                                           // 0x0167, 0x0001 big-endian
            } else {
                v = bswap32(r.word);                                       // htonl (:800)
            }
            out[lane] = v;
        }
        const bool audio = !(r.flags & 1u);
        if (r.size > IGD_PKT_HDR) {
            const uint32_t *src = r.src_frame >= 0
                ? reinterpret_cast<const uint32_t *>(d.payload + ((size_t)r.src_frame * d.C + c) * IGD_FRAME)
                : (d.stale_payload ? reinterpret_cast<const uint32_t *>(d.stale_payload + c * IGD_FRAME) : nullptr);
            for (uint32_t k = lane; k < nwords; k += 32) out[5 + k] = src ? src[k] : 0u;
        }
        if (audio) {
            // setOutgoingRTP (roip_ed137.cpp:6500-6536).  Clean: mean of the payload
            // bytes.  Quirk Q3: mean of the first payload_len bytes of the ORIGINAL
            // packet (12 header bytes + payload[0 .. len-12)), TransportAdapter.cpp:654.
            const uint32_t *cur = reinterpret_cast<const uint32_t *>(d.payload + i * IGD_FRAME);
            if (!quirks) {
                for (uint32_t k = lane; k < nwords; k += 32) bsum = bytesum4(cur[k], sc, bsum);
            } else {
                for (uint32_t k = lane; k < nwords; k += 32) {
                    const uint32_t v = k < 3 ? hdr[k] : cur[k - 3];
                    bsum = bytesum4(v, sc, bsum);
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) bsum += __shfl_xor_sync(0xffffffffu, bsum, o);
        }
        if (lane == 0) d.bytemean_out[i] = audio ? (uint8_t)igd_bytemean_from_sum(bsum, (int)d.payload_len) : 0;
    }
}

// Tile form of the assembly for the wire shape (out_stride 180, 160-byte payloads): a CTA builds
// 64 packets in shared memory (headers by 64 threads, payload chunks by all threads with 16-byte
// loads) and writes the image out with coalesced word stores, only the bytes each packet's size
// covers -- same bytes, sizes and levels as k_ed137_assemble above.
__global__ void __launch_bounds__(256) k_ed137_assemble_tile(const igd_ed137_pack_desc d,
                                                             const igd_tx_plan_rec *__restrict__ plan)
{
    __shared__ __align__(128) uint32_t img[kPktTile * kPktWords];           // the 64 packets being built
    __shared__ __align__(128) uint32_t pay[kPktTile * (IGD_FRAME / 4)];      // this tick's payloads of the tile
    __shared__ uint64_t bar;
    __shared__ uint32_t size_s[kPktTile];
    __shared__ int32_t srcf_s[kPktTile];
    __shared__ uint32_t flag_s[kPktTile];
    __shared__ int bsum_s[kPktTile];
    __shared__ uint32_t chan_s[kPktTile];
    __shared__ int32_t frame_s[kPktTile];
    __shared__ uint32_t not_full;                                            // some packet of the tile is not 180 bytes
    const size_t npkts = (size_t)d.F * d.C;
    const size_t first = (size_t)blockIdx.x * kPktTile;
    const uint32_t np = (uint32_t)min((size_t)kPktTile, npkts - first);
    const bool sc = (d.flags & IGD_F_SIGNED_CHAR) != 0, quirks = (d.flags & IGD_F_REF_QUIRKS) != 0;
    const uint32_t bar_s = shared_addr(&bar);
    if (threadIdx.x == 0) {          // ONE bulk async copy (TMA) stages the tile's current payloads (np*160 contiguous bytes)
        not_full = 0u;
        mbar_init(bar_s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect_tx(bar_s, np * IGD_FRAME);
        bulk_g2s(shared_addr(pay), d.payload + first * IGD_FRAME, np * IGD_FRAME, bar_s);
    }
    if (threadIdx.x < np) {
        const uint32_t p = threadIdx.x;
        const size_t i = first + p;
        const igd_tx_plan_rec r = plan[i];
        const uint32_t *hdr = reinterpret_cast<const uint32_t *>(d.rtp12 + i * 12);
        const uint32_t h0 = hdr[0], h1 = hdr[1], h2 = hdr[2];
        uint32_t v0 = h0 | 0x10u;                                              // x = 1 (:725)
        v0 = (v0 & ~0x8000u) | ((r.flags & 2u) ? 0x8000u : 0u);                // m (:715-723)
        if (r.flags & 1u) v0 = (v0 & ~0x7F00u) | (123u << 8);                  // pt = 123
        uint32_t *o = img + p * kPktWords;
        o[0] = v0; o[1] = h1; o[2] = h2; o[3] = 0x01006701u; o[4] = bswap32(r.word);
        size_s[p] = r.size;
        srcf_s[p] = r.src_frame;
        flag_s[p] = r.flags;
        frame_s[p] = (int32_t)(i / (size_t)d.C);
        chan_s[p] = (uint32_t)(i - (size_t)frame_s[p] * d.C);
        d.sizes[i] = r.size;
        // quirk Q3: the level sums the first 160 bytes of the ORIGINAL packet = these 12 header bytes ...
        int hs = 0;
        if (quirks) { hs = bytesum4(h0, sc, hs); hs = bytesum4(h1, sc, hs); hs = bytesum4(h2, sc, hs); }
        bsum_s[p] = hs;
    }
    __syncthreads();                 // barrier initialised, per-packet plan visible
    if (threadIdx.x < np && size_s[threadIdx.x] != IGD_PKT_MAX) not_full = 1u;
    mbar_wait(bar_s, 0);
    for (uint32_t j = threadIdx.x; j < np * kChunks; j += blockDim.x) {
        const uint32_t p = j / kChunks, ch = j - p * kChunks;
        const uint32_t size = size_s[p];
        if (size == 0) continue;
        const bool audio = !(flag_s[p] & 1u);
        const size_t c = chan_s[p];
        const int32_t f = frame_s[p], sf = srcf_s[p];
        const uint4 cur = *reinterpret_cast<const uint4 *>(pay + p * (IGD_FRAME / 4) + ch * 4);
        if (size > IGD_PKT_HDR) {
            uint4 v = cur;
            if (sf != f) {           // quirk Q2: the send buffer still holds an older frame's payload
                if (sf >= 0) v = __ldg(reinterpret_cast<const uint4 *>(d.payload + ((size_t)sf * d.C + c) * IGD_FRAME) + ch);
                else if (d.stale_payload) v = __ldg(reinterpret_cast<const uint4 *>(d.stale_payload + c * IGD_FRAME) + ch);
                else v = make_uint4(0u, 0u, 0u, 0u);
            }
            uint32_t *o = img + p * kPktWords + 5 + ch * 4;
            o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
        }
        if (audio) {        // setOutgoingRTP (roip_ed137.cpp:6500-6536): clean = the payload bytes;
            int bs = 0;     // Q3 = ... + payload[0 .. 148) (TransportAdapter.cpp:654)
            if (!quirks || ch < 9) {
                bs = bytesum4(cur.x, sc, bs); bs = bytesum4(cur.y, sc, bs); bs = bytesum4(cur.z, sc, bs); bs = bytesum4(cur.w, sc, bs);
            } else {
                bs = bytesum4(cur.x, sc, bs);                                  // bytes 144..147
            }
            atomicAdd(&bsum_s[p], bs);
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // the image is read by the async proxy below
    __syncthreads();
    const uint32_t tile_bytes = np * IGD_PKT_MAX;
    if (!not_full && (tile_bytes & 15u) == 0 && (reinterpret_cast<uintptr_t>(d.pkts) & 15u) == 0) {
        // every packet of the tile is a full 180-byte audio packet: ONE bulk async store (TMA) writes the image
        if (threadIdx.x == 0) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                         ::"l"(d.pkts + first * IGD_PKT_MAX), "r"(shared_addr(img)), "r"(tile_bytes) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");     // smem must outlive the read
        }
    } else {
        uint32_t *out = reinterpret_cast<uint32_t *>(d.pkts + first * IGD_PKT_MAX);
        for (uint32_t w = threadIdx.x; w < np * kPktWords; w += blockDim.x) {
            const uint32_t p = w / kPktWords, k = w - p * kPktWords;
            if (4 * k < size_s[p]) out[w] = img[w];
        }
    }
    if (threadIdx.x < np) {
        const uint32_t p = threadIdx.x;
        const bool audio = size_s[p] != 0 && !(flag_s[p] & 1u);
        d.bytemean_out[first + p] = audio ? (uint8_t)igd_bytemean_from_sum(bsum_s[p], IGD_FRAME) : (uint8_t)0;
    }
}

// ============================================================ recorder sink
__global__ void k_wav_image(const uint8_t *__restrict__ payload, size_t n, int rate, int law, int ref_quirks,
                            uint8_t *__restrict__ out)
{
    const size_t body = ref_quirks ? 2 * n : n, total = 44 + body;
    if (blockIdx.x == 0 && threadIdx.x < 11) {
        const uint32_t channels = ref_quirks ? 2 : 1, bits = ref_quirks ? 16 : 8;
        const uint32_t fmt = ref_quirks ? 7u : (law == IGD_LAW_ALAW ? 6u : 7u);
        const uint32_t align = bits / 8 * channels;
        uint32_t h[11];
        h[0] = 0x46464952u;                      // "RIFF"
        h[1] = (uint32_t)(total - 8);
        h[2] = 0x45564157u;                      // "WAVE"
        h[3] = 0x20746d66u;                      // "fmt "
        h[4] = 16u;
        h[5] = fmt | (channels << 16);
        h[6] = (uint32_t)rate;
        h[7] = (uint32_t)rate * align;
        h[8] = align | (bits << 16);
        h[9] = 0x61746164u;                      // "data"
        h[10] = (uint32_t)body;
        reinterpret_cast<uint32_t *>(out)[threadIdx.x] = h[threadIdx.x];
    }
    uint8_t *dst = out + 44;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x) {
        if (ref_quirks) reinterpret_cast<uint16_t *>(dst)[i] = payload[i];   // {b, 0x00}
        else dst[i] = payload[i];
    }
}

// Recorder sink for many calls at once: file image k = the WAV of channel chans[k], gathered from the
// frame-major batch layout codes [F][C][160] (a channel's frames are C*160 bytes apart), written to
// out + k*image_stride.  One warp moves one 160-byte frame (40 words in, 40 or 80 words out).
__global__ void __launch_bounds__(256) k_wav_images(const uint8_t *__restrict__ codes, size_t F, size_t C,
                                                    const uint32_t *__restrict__ chans, size_t nchan,
                                                    const uint8_t *__restrict__ law_ch, int rate, int ref_quirks,
                                                    uint8_t *__restrict__ out, size_t image_stride)
{
    const size_t n = F * IGD_FRAME;
    const size_t body = ref_quirks ? 2 * n : n, total = 44 + body;
    const uint32_t lane = threadIdx.x & 31;
    const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const size_t nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
    for (size_t u = warp; u < nchan * (F + 1); u += nwarps) {       // unit F of a channel = its header
        const size_t k = u / (F + 1), f = u - k * (F + 1);
        const uint32_t ch = chans ? chans[k] : (uint32_t)k;
        uint8_t *img = out + k * image_stride;
        if (f == F) {
            if (lane < 11) {
                const uint32_t law = law_ch ? law_ch[ch] : (uint32_t)IGD_LAW_ULAW;
                const uint32_t channels = ref_quirks ? 2 : 1, bits = ref_quirks ? 16 : 8;
                const uint32_t fmt = ref_quirks ? 7u : (law == IGD_LAW_ALAW ? 6u : 7u);
                const uint32_t align = bits / 8 * channels;
                const uint32_t h[11] = {0x46464952u, (uint32_t)(total - 8), 0x45564157u, 0x20746d66u, 16u,
                                        fmt | (channels << 16), (uint32_t)rate, (uint32_t)rate * align,
                                        align | (bits << 16), 0x61746164u, (uint32_t)body};
                reinterpret_cast<uint32_t *>(img)[lane] = h[lane];
            }
            continue;
        }
        const uint32_t *src = reinterpret_cast<const uint32_t *>(codes + (f * C + ch) * IGD_FRAME);
        uint32_t *dst = reinterpret_cast<uint32_t *>(img + 44 + (ref_quirks ? 2 : 1) * f * IGD_FRAME);
        for (uint32_t w = lane; w < IGD_FRAME / 4; w += 32) {
            const uint32_t v = __ldcs(src + w);
            if (ref_quirks) {                                            // every byte b -> {b, 0x00}
                dst[2 * w] = __byte_perm(v, 0u, 0x4140);
                dst[2 * w + 1] = __byte_perm(v, 0u, 0x4342);
            } else {
                dst[w] = v;
            }
        }
    }
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

// ============================================================ launchers
cudaError_t igd_k_g711_decode(const igd_launch_cfg &c, const uint8_t *codes, const uint8_t *law_ch,
                              int law, int16_t *pcm, size_t n, size_t nch)
{
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(k_g711_decode, cudaFuncAttributeMaxDynamicSharedMemorySize, kLutBytes);
        if (e != cudaSuccess) return e;
        attr = true;
    }
    const int threads = 512;
    k_g711_decode<<<grid_for(c, n / 16 + 1, threads, 3), threads, kLutBytes, c.stream>>>(codes, law_ch, law, pcm, n,
                                                                                  nch ? nch : 1);
    return cudaGetLastError();
}

cudaError_t igd_k_g711_encode(const igd_launch_cfg &c, const int16_t *pcm, const uint8_t *law_ch,
                              int law, uint8_t *codes, size_t n, size_t nch)
{
    const int threads = 256;
    k_g711_encode<<<grid_for(c, n / 16 + 1, threads, 8), threads, 0, c.stream>>>(pcm, law_ch, law, codes, n,
                                                                          nch ? nch : 1);
    return cudaGetLastError();
}

cudaError_t igd_k_frame_meter(const igd_launch_cfg &c, const int16_t *pcm, size_t nframes,
                              igd_meter_rec *out)
{
    constexpr int FPC = 32;
    const long long tiles = ((long long)nframes + FPC - 1) / FPC;
    k_frame_meter<FPC><<<grid_for(c, (size_t)tiles, 1, 6), FPC * kChunks, 0, c.stream>>>(pcm, (long long)nframes,
                                                                                  tiles, out);
    return cudaGetLastError();
}

cudaError_t igd_k_bytemean(const igd_launch_cfg &c, const uint8_t *payloads, size_t n, size_t len,
                           size_t stride, unsigned flags, uint8_t *out)
{
    k_bytemean<<<grid_for(c, n * 32, 256, 8), 256, 0, c.stream>>>(payloads, n, len, stride, flags, out);
    return cudaGetLastError();
}

cudaError_t igd_k_level_percent(const igd_launch_cfg &c, const int32_t *v, size_t n, int32_t *out)
{
    k_level_percent<<<grid_for(c, n, 256, 8), 256, 0, c.stream>>>(v, n, out);
    return cudaGetLastError();
}

cudaError_t igd_k_mix(const igd_launch_cfg &c, const int16_t *pcm, const uint16_t *gain,
                      size_t nframes, size_t nbridges, int legs, int16_t *mix)
{
    const long long total_bf = (long long)nframes * (long long)nbridges;
    k_mix<<<grid_for(c, (size_t)total_bf * 20, 256, 8), 256, 0, c.stream>>>(pcm, gain, total_bf, legs, mix);
    return cudaGetLastError();
}

namespace {
template <int G, bool kSigned, int kWarps>
cudaError_t launch_fused_w(const igd_launch_cfg &c, const FusedParams &q)
{
    auto kern = k_fused_w<G, kSigned, kWarps>;
    const size_t smem = kLutBytes + (size_t)kWarps * slot_geom<G>::kSlotBytes +
                        (size_t)kWarps * (kBfPerItem * G * kPst + kBfPerItem * kPst) * 8;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const long long items = (q.total_bf + kBfPerItem - 1) / kBfPerItem;
    long long grid = c.sm_count;
    if (grid > items) grid = items;      // small ticks: one item per SM before a second warp gets one
    kern<<<(int)grid, kWarps * 32, smem, c.stream>>>(q);
    return cudaGetLastError();
}

template <bool kSigned, int kWarps>
cudaError_t launch_fused_g(const igd_launch_cfg &c, const FusedParams &q)
{
    auto kern = k_fused_g<kSigned, kWarps>;
    const size_t smem = kLutBytes + (size_t)kWarps * 2 * kGSlotBytes + (size_t)kWarps * kGParts * 8;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const long long items = (q.total_bf + kGBf - 1) / kGBf;
    long long grid = c.sm_count;
    if (grid > items) grid = items;
    kern<<<(int)grid, kWarps * 32, smem, c.stream>>>(q);
    return cudaGetLastError();
}

template <int BFPC, bool kSigned>
cudaError_t launch_fused_anyg(const igd_launch_cfg &c, const FusedParams &q)
{
    const size_t smem = kLutBytes + (size_t)2 * BFPC * kPst * 8;
    cudaError_t e = cudaFuncSetAttribute(k_fused_anyg<BFPC, kSigned>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_fused_anyg<BFPC, kSigned>, BFPC * kChunks, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    FusedParams p = q;
    p.num_tiles = (q.total_bf + BFPC - 1) / BFPC;
    long long grid = (long long)c.sm_count * per_sm;
    if (grid > p.num_tiles) grid = p.num_tiles;
    k_fused_anyg<BFPC, kSigned><<<(int)grid, BFPC * kChunks, smem, c.stream>>>(p);
    return cudaGetLastError();
}
}  // namespace

cudaError_t igd_k_fused(const igd_launch_cfg &c, const igd_batch_desc &d)
{
    FusedParams q;
    q.codes = d.codes; q.law = d.law; q.gain = d.gain_q7; q.out_law = d.out_law;
    q.mix = d.mix; q.enc = d.enc; q.meter = d.meter; q.bmeter = d.bmeter;
    q.total_bf = (long long)d.F * d.B;
    q.num_tiles = 0;
    q.B = d.B; q.G = d.G; q.flags = d.flags;
    const bool sc = (d.flags & IGD_F_SIGNED_CHAR) != 0;
    // the warp-autonomous kernels index bridge-frames in 32 bits; anything larger (2^28 bridge-frames =
    // 318 GB of traffic at G = 4) cannot be resident on one GPU anyway and takes the generic kernel
    // ... and reads a bridge-frame's G gains / a bridge's G laws as one 2G- / G-byte word
    const bool fits32 = q.total_bf < (1ll << 28) &&          // 16-sample chunk indices (10 per bridge-frame) stay below 2^32
                        (reinterpret_cast<uintptr_t>(d.gain_q7) & (size_t)(2 * d.G - 1) & 7u) == 0 &&
                        (d.G != 4 || (reinterpret_cast<uintptr_t>(d.law) & 3u) == 0);
    if (fits32 && d.G == 4) return sc ? launch_fused_w<4, true, 24>(c, q) : launch_fused_w<4, false, 24>(c, q);
    if (fits32 && d.G == 2) return sc ? launch_fused_w<2, true, 24>(c, q) : launch_fused_w<2, false, 24>(c, q);
    if (fits32 && d.G == 1) return sc ? launch_fused_w<1, true, 24>(c, q) : launch_fused_w<1, false, 24>(c, q);
    // any other leg count: the warp-autonomous group walk (needs 16-byte aligned codes, which the C ABI
    // checks, and 32-bit bridge-frame indices); the block-cooperative kernel is the last resort
    if (q.total_bf < (1ll << 28) && (long long)q.total_bf * d.G < (1ll << 32) &&
        !getenv("IGD_FUSED_ANYG"))
        return sc ? launch_fused_g<true, 24>(c, q) : launch_fused_g<false, 24>(c, q);
    return sc ? launch_fused_anyg<32, true>(c, q) : launch_fused_anyg<32, false>(c, q);
}

cudaError_t igd_k_event_summary(const igd_launch_cfg &c, const igd_meter_rec *meter,
                                const uint16_t *gain, size_t F, size_t C, igd_summary_rec *out,
                                igd_summary_db *db)
{
    const int blocks = (int)((C + 31) / 32);
    k_event_summary<<<blocks, 32 * kSumSlices, 0, c.stream>>>(meter, gain, (long long)F, (long long)C, out, db);
    return cudaGetLastError();
}

cudaError_t igd_k_ed137_parse(const igd_launch_cfg &c, const uint8_t *pkts, const uint32_t *sizes,
                              size_t npkts, size_t stride, igd_ed137_fields *fields,
                              uint8_t *payload_out)
{
    const bool al16 = ((reinterpret_cast<uintptr_t>(pkts) | reinterpret_cast<uintptr_t>(payload_out)) & 15) == 0;
    if (stride == IGD_PKT_MAX && al16 && npkts > 0) {
        const size_t tiles = (npkts + kPktTile - 1) / kPktTile;
        k_ed137_parse_tile<<<(unsigned)tiles, 256, 0, c.stream>>>(pkts, sizes, npkts, fields, payload_out);
        return cudaGetLastError();
    }
    k_ed137_parse<<<grid_for(c, npkts * 32, 256, 8), 256, 0, c.stream>>>(pkts, sizes, npkts, stride, fields,
                                                                  payload_out);
    return cudaGetLastError();
}

cudaError_t igd_k_rx_track(const igd_launch_cfg &c, const igd_rx_track_desc &d)
{
    k_rx_track<<<(d.C + 127) / 128, 128, 0, c.stream>>>(d);
    return cudaGetLastError();
}

cudaError_t igd_k_gate_arbitrate(const igd_launch_cfg &c, const igd_arb_desc &d)
{
    const unsigned blocks = (unsigned)((d.B + kArbThreads - 1) / kArbThreads);
    const size_t stage = (size_t)kArbStageWords * 6;
    switch (d.G) {
    case 1: k_gate_arbitrate<1><<<blocks, kArbThreads, stage, c.stream>>>(d); break;
    case 2: k_gate_arbitrate<2><<<blocks, kArbThreads, stage, c.stream>>>(d); break;
    case 4: k_gate_arbitrate<4><<<blocks, kArbThreads, stage, c.stream>>>(d); break;
    default:   // runtime leg count: leg state in shared memory (<= 64 * 32 * 8 B = 16 KB, 40 KB in all)
        k_gate_arbitrate<0><<<blocks, kArbThreads, stage + (size_t)kArbThreads * d.G * sizeof(igd_arb_leg), c.stream>>>(d);
    }
    return cudaGetLastError();
}

cudaError_t igd_k_ed137_keepalive(const igd_launch_cfg &c, uint8_t *hdr20, igd_ed137_state *state, size_t C,
                                  long long now, uint32_t *sizes)
{
    k_ed137_keepalive<<<(unsigned)((C + 127) / 128), 128, 0, c.stream>>>(hdr20, state, C, now, sizes);
    return cudaGetLastError();
}

int igd_k_launches_ed137_pack() { return 3; }

cudaError_t igd_k_ed137_pack(const igd_launch_cfg &c, const igd_ed137_pack_desc &d,
                             igd_tx_plan_rec *plan, int32_t *last_src)
{
    k_ed137_plan<<<(d.C + 127) / 128, 128, 0, c.stream>>>(d, plan, last_src);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    const bool al16 = ((reinterpret_cast<uintptr_t>(d.payload) | reinterpret_cast<uintptr_t>(d.stale_payload)) & 15) == 0 &&
                      (reinterpret_cast<uintptr_t>(d.pkts) & 3) == 0;
    if (d.out_stride == IGD_PKT_MAX && d.payload_len == IGD_FRAME && al16) {
        const size_t tiles = ((size_t)d.F * d.C + kPktTile - 1) / kPktTile;
        k_ed137_assemble_tile<<<(unsigned)tiles, 256, 0, c.stream>>>(d, plan);
    } else {
        k_ed137_assemble<<<grid_for(c, (size_t)d.F * d.C * 32, 256, 8), 256, 0, c.stream>>>(d, plan);
    }
    e = cudaGetLastError();
    if (e != cudaSuccess || !d.stale_payload) return e;
    k_ed137_stale_update<<<grid_for(c, (size_t)d.C * 32, 256, 8), 256, 0, c.stream>>>(d, last_src);
    return cudaGetLastError();
}

cudaError_t igd_k_wav_images(const igd_launch_cfg &c, const uint8_t *codes, size_t F, size_t C, const uint32_t *chans,
                             size_t nchan, const uint8_t *law_ch, int rate, int ref_quirks, uint8_t *out,
                             size_t image_stride)
{
    k_wav_images<<<grid_for(c, nchan * (F + 1) * 32, 256, 8), 256, 0, c.stream>>>(codes, F, C, chans, nchan, law_ch, rate,
                                                                             ref_quirks, out, image_stride);
    return cudaGetLastError();
}

cudaError_t igd_k_wav_image(const igd_launch_cfg &c, const uint8_t *payload, size_t n, int rate,
                            int law, int ref_quirks, uint8_t *out)
{
    k_wav_image<<<grid_for(c, n + 1, 256, 8), 256, 0, c.stream>>>(payload, n, rate, law, ref_quirks, out);
    return cudaGetLastError();
}
