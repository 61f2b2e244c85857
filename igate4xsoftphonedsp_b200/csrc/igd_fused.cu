// igd_fused.cu -- the fused decode -> meter -> gate/gain -> mix -> encode kernels (sm_100a) of the iGate4x voice path.
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
#include "igd_device.cuh"

namespace {

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
    const igd_ed137_fields *fields;   // packet form only: [F][C]; NULL = the gains already carry IGD_GAIN_NO_AUDIO
    // packets OUT (gateway form): the bridge output leaves as finished ED-137 packets, one sender per bridge
    const igd_tx_plan_rec *plan;      // [F][B] from the sender walk (k_ed137_plan)
    const uint8_t *rtp12;             // [F][B][12] the PJSIP-built RTP headers
    uint8_t *tx_pkts;                 // [F][B][180]
    uint32_t *tx_sizes;               // [F][B]
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
                                               uint8_t *enc_dst, bool st_mix, bool st_enc, uint4 *enc_out = nullptr)
{
    uint32_t pk[8];
#pragma unroll
    for (int i = 0; i < 8; i++) pk[i] = pack_sat16(acc[2 * i + 1], acc[2 * i]);
    if (st_mix) st32_stream(mix_dst, pk);
#ifdef IGD_X_NOPEAK      // measurement only: what the bridge record's mix_peak costs
    const uint4 e_np = encode16_packed(pk, E);
    if (st_enc) st16_stream(enc_dst, e_np);
    uint32_t u_np = __dp4a(e_np.x, 0x01010101u, 0u); u_np = __dp4a(e_np.y, 0x01010101u, u_np);
    u_np = __dp4a(e_np.z, 0x01010101u, u_np); u_np = __dp4a(e_np.w, 0x01010101u, u_np);
    return make_uint2(u_np, 0u);
#endif
    // max|v| of the 16 saturated samples: packed running max / min that start at zero (a free third input of the
    // 3-input min/max), so max >= 0 >= min and |.| is max_u16(max, 0 - min) per half -- 32768 for -32768 included
    uint32_t mx = max_s16x2(max_s16x2(E.zero2, pk[0]), pk[1]), mn = min_s16x2(min_s16x2(E.zero2, pk[0]), pk[1]);
    mx = max_s16x2(max_s16x2(mx, pk[2]), pk[3]); mn = min_s16x2(min_s16x2(mn, pk[2]), pk[3]);
    mx = max_s16x2(max_s16x2(mx, pk[4]), pk[5]); mn = min_s16x2(min_s16x2(mn, pk[4]), pk[5]);
    mx = max_s16x2(max_s16x2(mx, pk[6]), pk[7]); mn = min_s16x2(min_s16x2(mn, pk[6]), pk[7]);
    const uint32_t pk2 = max_u16x2(mx, neg_16x2(mn));
    const uint32_t peak = max(pk2 & 0xFFFFu, pk2 >> 16);
    const uint4 e = encode16_packed(pk, E);
    if (st_enc) st16_stream(enc_dst, e);
    if (enc_out) *enc_out = e;
    int esum = 0;
    if (kSigned) {
        esum = __dp4a((int)e.x, 0x01010101, esum); esum = __dp4a((int)e.y, 0x01010101, esum);
        esum = __dp4a((int)e.z, 0x01010101, esum); esum = __dp4a((int)e.w, 0x01010101, esum);
    } else {
        uint32_t u = __dp4a(e.x, 0x01010101u, 0u); u = __dp4a(e.y, 0x01010101u, u);
        u = __dp4a(e.z, 0x01010101u, u); u = __dp4a(e.w, 0x01010101u, u);
        esum = (int)u;
    }
    return make_uint2((uint32_t)esum, peak);
}

// raw gain bits of one bridge-frame (G u16 values) in two registers
template <int G>
__device__ __forceinline__ uint2 load_gains(const uint16_t *g)
{
    if (G == 4) return *reinterpret_cast<const uint2 *>(g);
    if (G == 2) return make_uint2(*reinterpret_cast<const uint32_t *>(g), 0u);
    if (G == 3) return make_uint2((uint32_t)g[0] | ((uint32_t)g[1] << 16), (uint32_t)g[2]);   // 6-byte rows: halfword loads
    return make_uint2(*g, 0u);
}

// ------------------------------------------------------------------------------------
// Warp-autonomous fused kernel (G in {1,2,3,4}): no cross-warp handshake at all.
// A warp owns "items" of kBfPerItem = 6 consecutive bridge-frames (6*G*160 contiguous
// code bytes) and strides over them on its own:
//   * the item's codes are fetched by the warp's OWN bulk async copy (TMA, one copy of
//     the contiguous item into the warp's private slot), completion on the warp's
//     private mbarrier; the copy of item i+1 is issued once every lane has consumed
//     the last codes of item i (proxy fence, then the copy), so HBM latency hides behind
//     the rest of the item and ~90 KB per SM are in flight;
//   * lane = two 16-sample chunks (c' and c'+5, c' rotated per bridge-frame, see
//     chunk_rotation) of one bridge-frame, all G legs (5 lanes
//     per bridge-frame, 30 of 32 lanes busy): per-bridge-frame setup (gains, laws,
//     selectors, addresses) is paid once per 32 samples;
//   * the per-chunk meter partials meet in the warp's private shared scratch after a
//     __syncwarp; lanes 0..6G-1 finish one leg record each, the next 6 lanes one bridge
//     record each.
// Warps drift freely: nobody spins on a peer (the CTA-cooperative first version of this kernel
// spent ~16 % of its issue slots in mbarrier spin loops).
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
template <int G, bool kPkt = false> struct slot_geom {
    static constexpr int kLegBytes = kPkt ? IGD_PKT_MAX : IGD_FRAME;   // packet form: the raw 180-byte packets
    static constexpr int kBfBytes = G * kLegBytes;
    static constexpr int kSlotBytes = kBfPerItem * kBfBytes;
};
__device__ __forceinline__ uint32_t chunk_rotation(uint32_t bfl) { return (0x037050u >> (4 * bfl)) & 0xFu; }

__device__ __forceinline__ void build_decode_lut_abs(uint32_t *lut, int warp, int lane, int nwarps)
{
    fill_lut32(lut, warp, lane, nwarps, [](uint32_t law, uint32_t code) {
        const int x = law ? igd_ulaw2lin(code) : igd_alaw2lin(code);
        const int y2 = min(max(2 * x, -32768), 32767);
        return ((uint32_t)y2 << 16) | ((uint32_t)(abs(x) >> 2) & 0xFFFFu);
    });
}
// Wide form (k_fused_g): 8-byte entries {(|x|/4)^2, |x|/4 | clamp16(2x) << 16} -- the square comes
// out of the table with the sample (one LDS.64), so the meter needs no extract and no multiply per
// sample.  Row = code, 16 columns of 8 bytes (128 B, the same 64 KB as 32 columns of 4 bytes): a
// 64-bit shared load is served per half-warp, lanes l and l + 16 share a column, no bank conflicts.
// Twice the shared-memory bytes per lookup: +5..7 % for k_fused_g (one chunk per lane, issue-bound),
// -26 % for k_fused_w at G = 4, which then sits on the shared-memory bandwidth at every gate density.
__device__ __forceinline__ void build_decode_lut_wide(uint32_t *lut, int tid, int nthreads)
{
    uint2 *lut2 = reinterpret_cast<uint2 *>(lut);
    for (int i = tid; i < 2 * 256 * 16; i += nthreads) {
        const uint32_t law = (uint32_t)i >> 12, code = ((uint32_t)i >> 4) & 255u;
        const int x = law ? igd_ulaw2lin(code) : igd_alaw2lin(code);
        const int y2 = min(max(2 * x, -32768), 32767);
        const uint32_t q = (uint32_t)(abs(x) >> 2);
        lut2[i] = make_uint2(q * q, ((uint32_t)y2 << 16) | q);
    }
}
template <int K>
__device__ __forceinline__ uint2 lut_lookup2(uint32_t lane_base, uint32_t word)
{
    const uint32_t a = __dp4a(word, 0x80u << (8 * K), lane_base);
    uint2 v;
    asm("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
    return v;
}

// IGD_GAIN_NO_AUDIO (bit 15 of a gain): the leg-frame is silent.  Clears every flagged half (the leg then
// walks the shut path like any closed gate) and returns the flags as one bit per leg in `sil`.
__device__ __forceinline__ uint2 split_no_audio(uint2 raw, uint32_t &sil)
{
    const uint32_t fx = raw.x & 0x80008000u, fy = raw.y & 0x80008000u;
    sil = ((fx >> 15) & 1u) | ((fx >> 30) & 2u) | (((fy >> 15) & 1u) << 2) | (((fy >> 30) & 2u) << 2);
    // a flagged half 0x8000 becomes a 0xFFFF mask: (f >> 15) * 0xFFFF
    return make_uint2(raw.x & ~((fx >> 15) * 0xFFFFu), raw.y & ~((fy >> 15) * 0xFFFFu));
}

// bit 15 / 31 set for every non-zero 16-bit half of x
__device__ __forceinline__ uint32_t nonzero_halves(uint32_t x)
{
    return (((x & 0x7FFF7FFFu) + 0x7FFF7FFFu) | x) & 0x80008000u;
}

// decode + meter + gain/accumulate one 16-sample chunk of one leg (|x|/4 table)
// kMode: 0 = gate shut for the whole warp (meter only), 1 = gain 2.0 or shut through the
//        IDP.2A selector (0x0100 / 0), 2 = arbitrary Q7 gain (multiply, shift, clip)
template <bool kSigned, int kMode, bool kWide = false>
__device__ __forceinline__ uint2 leg_chunk_u(uint32_t lane_base, uint4 w, uint32_t sel, int adj, int (&acc)[16])
{
    const uint32_t wd[4] = {w.x, w.y, w.z, w.w};
    uint32_t sq = 0;               // sum of (x/4)^2: 16 * 8064^2 < 2^31
    uint32_t mx = 0;               // packed running max of (|x|/4, junk)
    int bsum = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        uint32_t e0, e1, e2, e3;
        if (kWide) {
            const uint2 t0 = lut_lookup2<0>(lane_base, wd[j]), t1 = lut_lookup2<1>(lane_base, wd[j]);
            const uint2 t2 = lut_lookup2<2>(lane_base, wd[j]), t3 = lut_lookup2<3>(lane_base, wd[j]);
            e0 = t0.y; e1 = t1.y; e2 = t2.y; e3 = t3.y;
            sq += (t0.x + t1.x) + (t2.x + t3.x);      // the squares ride in the table: two IADD3 per four samples
        } else {
            e0 = lut_lookup<0>(lane_base, wd[j]); e1 = lut_lookup<1>(lane_base, wd[j]);
            e2 = lut_lookup<2>(lane_base, wd[j]); e3 = lut_lookup<3>(lane_base, wd[j]);
            // |x|/4 out of the low half on the ALU pipe (LOP3): everything else in this loop body is
            // IDP/IMAD on the FMA pipe, which bounds this phase (measured: any IDP.2A here is slower)
            const uint32_t x0 = e0 & 0xFFFFu, x1 = e1 & 0xFFFFu, x2 = e2 & 0xFFFFu, x3 = e3 & 0xFFFFu;
            sq += x0 * x0 + x1 * x1 + x2 * x2 + x3 * x3;
        }
        mx = max_u16x2(max_u16x2(mx, e0), e1); mx = max_u16x2(max_u16x2(mx, e2), e3);
        bsum = kSigned ? __dp4a((int)wd[j], 0x01010101, bsum) : (int)__dp4a(wd[j], 0x01010101u, (uint32_t)bsum);
        if (kMode == 1) {
            acc[4 * j + 0] = dp2a_lo(e0, sel, acc[4 * j + 0]);
            acc[4 * j + 1] = dp2a_lo(e1, sel, acc[4 * j + 1]);
            acc[4 * j + 2] = dp2a_lo(e2, sel, acc[4 * j + 2]);
            acc[4 * j + 3] = dp2a_lo(e3, sel, acc[4 * j + 3]);
        } else if (kMode == 2) {
            const uint32_t x0 = e0 & 0xFFFFu, x1 = e1 & 0xFFFFu, x2 = e2 & 0xFFFFu, x3 = e3 & 0xFFFFu;
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

// kPkt: the codes are read straight out of the raw ED-137 packets ([F][C][180], payload at byte 20) -- the
// item is still ONE contiguous bulk copy (6 * G * 180 B), the payload array between igd_ed137_parse and
// this kernel is never materialised; a packet that is not a whole audio frame (parsed fields) is silent.
// kTx: the bridge output leaves as finished 180-byte ED-137 packets -- header from the sender plan (k_ed137_plan)
// and the PJSIP RTP header, payload = this tick's encoded mix, every byte of a slot past the packet's size zero
// (the bytes k_ed137_assemble_tile writes without IGD_F_REF_QUIRKS): no enc[] round trip, no assembly kernel.
// kOpt: some of mix / enc / meter / bmeter may be NULL (checked per store); false = all four are there, no checks.
template <int G, bool kSigned, int kWarps, bool kPkt = false, bool kTx = false, bool kOpt = false>
__global__ void __launch_bounds__(kWarps * 32, 1) k_fused_w(const FusedParams q)
{
    static_assert(kBfPerItem * G + kBfPerItem <= 32, "finish needs one lane per record");
    static_assert(!kPkt || G == 4, "packet form: the aligned-window offsets assume 720-byte bridge-frames");
    using geom = slot_geom<G, kPkt>;
    constexpr int kP = kPkt ? kChunks : kPst;      // partial stride: the bigger packet slots leave no room for the pad column
    constexpr int kLegParts = kBfPerItem * G * kP, kBrParts = kBfPerItem * kP;
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

    build_decode_lut_abs(lut, (int)warp, (int)lane, kWarps);
    if (t < 2) {
        const enc_pk e = enc_pk_make(t);
        enc_tab[t][0] = e.bias_pos; enc_tab[t][1] = e.bias_x; enc_tab[t][2] = e.hi_pos; enc_tab[t][3] = e.hi_x;
        enc_tab[t][4] = e.thr; enc_tab[t][5] = e.mask4; enc_tab[t][6] = 0u; enc_tab[t][7] = 0u;
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
    // first chunk; second = +-5.  The per-bridge-frame rotation is the conflict minimum for the 640-byte bridge-frames of
    // the codes form; the 720-byte bridge-frames of the packet form (45 sixteen-byte units, odd) are best left unrotated
    // (brute force over all rotations: 10 instead of 20 LDS.128 wavefronts per pair of passes)
    const uint32_t c0 = worker ? (lane - bfl * kC32 + (kPkt ? 0u : chunk_rotation(bfl))) % kChunks : 0u;
    const uint32_t src = slot_s + bfl * geom::kBfBytes;
    const uint32_t lane4 = lut_bytes + 4u * lane;
    const uint32_t b_step = (uint32_t)(((unsigned long long)nw * kBfPerItem) % (uint32_t)q.B);
    uint32_t b = (item * kBfPerItem + bfl) % (uint32_t)q.B;
    // the G leg laws and the output law of bridge bb as loaded (the bits are picked out when the item is worked on:
    // unpacking them here would make the prefetch wait for its own loads)
    auto load_laws = [&](uint32_t bb, uint32_t &olaw) -> uint32_t {
        uint32_t r = 0;
        if (G == 4) r = __ldg(reinterpret_cast<const uint32_t *>(q.law + (size_t)bb * 4));
        else {
#pragma unroll
            for (int g = 0; g < G; g++) r |= (uint32_t)__ldg(q.law + (size_t)bb * G + g) << (8 * g);
        }
        olaw = __ldg(q.out_law + bb);
        return r;
    };
    // packet form: a leg whose packet is not a whole audio frame (keep-alive, truncated, dropped, absent) is
    // silent -- its gain gets IGD_GAIN_NO_AUDIO, the bytes of its slot are never interpreted
    auto mark_no_audio = [&](uint2 g, uint32_t bfi) -> uint2 {
        const uint32_t *fw = reinterpret_cast<const uint32_t *>(q.fields + (size_t)bfi * G);
#pragma unroll
        for (int l = 0; l < G; l++) {
            if (igd_fields_no_audio(__ldg(fw + 4 * l + 1), __ldg(fw + 4 * l + 2), __ldg(fw + 4 * l + 3))) {
                if (l < 2) g.x |= 0x8000u << (16 * l); else g.y |= 0x8000u << (16 * (l - 2));
            }
        }
        return g;
    };
    uint2 gq = make_uint2(0u, 0u);
    uint32_t lwq = 0u, owq = 0u, pszq = 0u;
    if (worker && item < items && item * kBfPerItem + bfl < total_bf) {
        gq = load_gains<G>(q.gain + (size_t)(item * kBfPerItem + bfl) * G);
        lwq = load_laws(b, owq);
        if (kPkt && q.fields) gq = mark_no_audio(gq, item * kBfPerItem + bfl);
        if (kTx) pszq = __ldg(&q.plan[item * kBfPerItem + bfl].size);
    }

    const bool want_mix = !kOpt || q.mix != nullptr, want_enc = !kOpt || q.enc != nullptr,
               want_meter = !kOpt || q.meter != nullptr, want_bmeter = !kOpt || q.bmeter != nullptr;
    for (uint32_t it = 0; item < items; item += nw, it++) {
        const uint32_t bf = item * kBfPerItem + bfl;
        const uint32_t next = item + nw;
        mbar_wait(bar_s, it & 1u);                           // this item's codes have landed
        const uint2 gcur = gq;
        // bit g = law of leg g, bit 8 = output law, bit 9 (kTx) = the outgoing packet carries the payload
        const uint32_t lcur = (lwq & 1u) | ((lwq >> 7) & 2u) | ((lwq >> 14) & 4u) | ((lwq >> 21) & 8u) | ((owq & 1u) << 8) |
                              (kTx ? (uint32_t)(pszq > IGD_PKT_HDR) << 9 : 0u);
        const bool valid = worker && bf < total_bf;
        {
            b += b_step;
            if (b >= (uint32_t)q.B) b -= (uint32_t)q.B;
            const uint32_t bfn = bf + nw * kBfPerItem;
            if (worker && next < items && bfn < total_bf) {      // next item's gains and laws ride in three registers
                gq = load_gains<G>(q.gain + (size_t)bfn * G);
                lwq = load_laws(b, owq);
                if (kPkt && q.fields) gq = mark_no_audio(gq, bfn);
                if (kTx) pszq = __ldg(&q.plan[bfn].size);
            }
        }
        // every lane runs the same instruction stream (idle / tail lanes on stale bytes with all
        // gates shut); only the stores are predicated, and the mode branches are warp-uniform
        auto adj_of = [&](int g) -> uint32_t { return g == 0 ? (gcur.x & 0xFFFFu) : g == 1 ? (gcur.x >> 16) : g == 2 ? (gcur.y & 0xFFFFu) : (gcur.y >> 16); };
        // warp-wide OR of the gains (REDUX): anything but 0 / 256 anywhere in the warp -> general
        // path; bit g of open_mask: some lane of the warp has leg g open
        // IGD_GAIN_NO_AUDIO (bit 15 of a gain) anywhere in the warp makes `general` true: the rare multiply path
        // below treats a flagged leg as gain 0 and the finish writes its record as digital silence.  The common
        // path (gains 0 / 2.0 only) is untouched by the flag handling.
        const uint32_t orx = __reduce_or_sync(0xFFFFFFFFu, gcur.x), ory = G > 2 ? __reduce_or_sync(0xFFFFFFFFu, gcur.y) : 0u;
        const bool general = ((orx | ory) & 0xFEFFFEFFu) != 0u;
        const uint32_t open_mask = ((orx & 0xFFFFu) ? 1u : 0u) | ((orx >> 16) ? 2u : 0u) | ((ory & 0xFFFFu) ? 4u : 0u) |
                                   ((ory >> 16) ? 8u : 0u);
        // open legs of this lane's bridge-frame, pre-shifted into the high half of the bridge partial
        // (every one of the ten partials carries it; the finish divides the sum by ten)
        uint32_t n_open16 = (uint32_t)(__popc(nonzero_halves(gcur.x)) + __popc(nonzero_halves(gcur.y))) << 16;
        if (general)             // legs without audio do not count as open (their halves are non-zero: the flag)
            n_open16 = (uint32_t)(__popc(nonzero_halves(gcur.x & ~(((gcur.x & 0x80008000u) >> 15) * 0xFFFFu))) +
                                  __popc(nonzero_halves(gcur.y & ~(((gcur.y & 0x80008000u) >> 15) * 0xFFFFu)))) << 16;
#pragma unroll 1
        for (int h = 0; h < 2; h++) {
            const uint32_t ch = h == 0 ? c0 : (c0 >= (uint32_t)kC32 ? c0 - kC32 : c0 + kC32);   // this pass's chunk
            uint4 wh[G];
#pragma unroll
            for (int g = 0; g < G; g++) {
                if (!kPkt) {
                    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                                 : "=r"(wh[g].x), "=r"(wh[g].y), "=r"(wh[g].z), "=r"(wh[g].w)
                                 : "r"(src + g * IGD_FRAME + ch * 16));
                } else {
                    // payload at byte 20 of a 180-byte packet.  With G = 4 a bridge-frame is 720 = 45 * 16
                    // bytes, so the chunk of leg g starts (4 g + 4) % 16 bytes past a 16-byte boundary -- a
                    // compile-time offset: read the aligned 32-byte window with two LDS.128 (one for
                    // g = 3) and pick the four words by register name (four 4-byte loads per chunk
                    // would be 4-way bank conflicts: the lanes' chunks are 16 bytes apart).
                    const int kOff = ((4 * g + 4) % 16) / 4;                   // words into the window (constant once unrolled)
                    const uint32_t a = src + g * IGD_PKT_MAX + IGD_PKT_HDR + ch * 16 - 4 * kOff;
                    uint32_t v[8];
                    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(a));
                    if (kOff != 0)
                        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4+16];"
                                     : "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(a));
                    wh[g] = make_uint4(v[kOff], v[kOff + 1], v[kOff + 2], v[kOff + 3]);
                }
            }
            if (h == 1) {
                // Every lane has read the rest of its codes: refill the slot now, a whole pass of work ahead of the
                // next item.  The fence completes the slot reads above (generic proxy) and orders them before the
                // bulk copy's writes (async proxy).  (Issuing the copy right behind the LDS UNFENCED lost the race at
                // G = 1: a 960-byte copy that hits L2 overtook loads still queued behind the other warps' table
                // lookups -- profiles/tools/determinism_soak.py.)
                fence_proxy_async();
                __syncwarp();
                if (next < items) fetch(next);
            }
            uint2 *mypart = part + bfl * (G * kP) + ch;
            int acc[16];
#pragma unroll
            for (int i = 0; i < 16; i++) acc[i] = 0;
            if (!general) {        // gains in {0, 2.0}: one IDP.2A per sample of an open leg
#pragma unroll
                for (int g = 0; g < G; g++) {
                    const uint32_t lb = lane4 + (((lcur >> g) & 1u) << 15);
                    const uint2 ph = (open_mask >> g) & 1u ? leg_chunk_u<kSigned, 1>(lb, wh[g], adj_of(g), 0, acc)
                                                           : leg_chunk_u<kSigned, 0>(lb, wh[g], 0u, 0, acc);
                    if (valid) mypart[g * kP] = ph;
                }
            } else {               // arbitrary Q7 gains: multiply, shift, clip
#pragma unroll
                for (int g = 0; g < G; g++) {
                    const uint32_t lb = lane4 + (((lcur >> g) & 1u) << 15);
                    const uint32_t a = adj_of(g);
                    const uint2 ph = leg_chunk_u<kSigned, 2>(lb, wh[g], 0u, (a & IGD_GAIN_NO_AUDIO) ? 0 : (int)a, acc);
                    if (valid) mypart[g * kP] = ph;
                }
            }
            enc_pk E;
            {
                const uint32_t *et = enc_tab[(lcur >> 8) & 1u];
                const uint4 e0 = *reinterpret_cast<const uint4 *>(et);
                const uint2 e1 = *reinterpret_cast<const uint2 *>(et + 4);
                E.bias_pos = e0.x; E.bias_x = e0.y; E.hi_pos = e0.z; E.hi_x = e0.w; E.thr = e1.x; E.mask4 = e1.y;
                E.zero2 = et[6];      // a zero the compiler cannot fold (see enc_pk)
            }
            const uint32_t o16 = bf * kChunks + ch;     // 16-sample chunk index of the outputs
            uint4 ev;
            const uint2 mo = mix_out_chunk<kSigned>(acc, E, q.mix + (size_t)o16 * 16, q.enc + (size_t)o16 * 16,
                                                    valid && want_mix, valid && want_enc, kTx ? &ev : nullptr);
            if (kTx && valid) {      // payload bytes [20 + 16 ch, +16) of the packet slot (4-byte aligned): the codes, or zeros
                uint32_t *tp = reinterpret_cast<uint32_t *>(q.tx_pkts + (size_t)bf * IGD_PKT_MAX + IGD_PKT_HDR + ch * 16);
                const bool pay = (lcur >> 9) & 1u;
                __stcs(tp, pay ? ev.x : 0u); __stcs(tp + 1, pay ? ev.y : 0u);
                __stcs(tp + 2, pay ? ev.z : 0u); __stcs(tp + 3, pay ? ev.w : 0u);
            }
            if (valid) bpart[bfl * kP + ch] = make_uint2(mo.x, mo.y | n_open16);
        }
        __syncwarp();
        // ---- finish: one lane per record (leg records, then bridge records: bpart follows part,
        // and both kinds of partial are {sum, max | sum << 16}, so the ten-partial walk is shared)
        {
            const uint32_t bf0 = item * kBfPerItem;
            const uint2 *src_p = part + lane * kP;
            unsigned long long sq = 0; uint32_t pk = 0; int bsum = 0;
            // the record of leg (lane / G, lane % G): the lanes of that bridge-frame hold its no-audio flags
            uint32_t sil_rec = 0u;
            if (general) {       // rare: the record of leg (lane / G, lane % G) is silent when its gain carries the flag
                uint32_t silent;
                (void)split_no_audio(gcur, silent);
                sil_rec = (__shfl_sync(0xFFFFFFFFu, silent, (int)((lane / G) * kC32) & 31) >> (lane % G)) & 1u;
            }
            if (lane < kBfPerItem * G + kBfPerItem && !(lane < kBfPerItem * G && sil_rec)) {
#pragma unroll
                for (int i = 0; i < kChunks; i++) {
                    const uint2 v = src_p[i];
                    sq += v.x; pk = max_u16x2(pk, v.y); bsum = dp2a_lo(v.y, 0x0100u, bsum);
                }
            }
            if (lane < kBfPerItem * G) {
                if (bf0 + lane / G < total_bf && want_meter) {
                    const igd_meter_rec r = meter_finish(sq << 4, (pk & 0xFFFFu) << 2, bsum, true);
                    st16_stream(q.meter + ((size_t)bf0 * G + lane), *reinterpret_cast<const uint4 *>(&r));
                }
            } else if (lane < kBfPerItem * G + kBfPerItem) {
                const uint32_t j = lane - kBfPerItem * G;
                if (kTx && bf0 + j < total_bf) {      // the packet's 20-byte header (TransportAdapter.cpp:715-800) and its size
                    const uint32_t *pr = reinterpret_cast<const uint32_t *>(q.plan + (bf0 + j));
                    const uint32_t pw = __ldg(pr), psf = __ldg(pr + 1);      // word; size | flags << 16
                    const uint32_t psize = psf & 0xFFFFu, pfl = psf >> 16;
                    const uint32_t *hdr = reinterpret_cast<const uint32_t *>(q.rtp12 + (size_t)(bf0 + j) * 12);
                    uint32_t v0 = __ldg(hdr) | 0x10u;                                          // x = 1 (:725)
                    v0 = (v0 & ~0x8000u) | ((pfl & 2u) ? 0x8000u : 0u);                        // m (:715-723)
                    if (pfl & 1u) v0 = (v0 & ~0x7F00u) | (123u << 8);                          // pt = 123
                    uint32_t *tp = reinterpret_cast<uint32_t *>(q.tx_pkts + (size_t)(bf0 + j) * IGD_PKT_MAX);
                    const bool any = psize != 0u;
                    tp[0] = any ? v0 : 0u; tp[1] = any ? __ldg(hdr + 1) : 0u; tp[2] = any ? __ldg(hdr + 2) : 0u;
                    tp[3] = any ? 0x01006701u : 0u;                                            // 0x0167, 0x0001 big-endian
                    tp[4] = any ? __byte_perm(pw, 0u, 0x0123) : 0u;                            // htonl (:800)
                    q.tx_sizes[bf0 + j] = psize;
                }
                if (bf0 + j < total_bf && want_bmeter) {
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
// Quarter-frame lanes (codes form, G in {1,2,3,4}): the same warp-autonomous design as k_fused_w with ALL 32 lanes
// at work.  An item is 8 consecutive bridge-frames (8*G*160 contiguous code bytes, one bulk copy); lane = (bridge-frame
// bfl = lane / 4, quarter qd = lane % 4) and owns the five 8-sample chunks qd, qd + 4, ... qd + 16 of that bridge-frame,
// all G legs: per pass the four lanes of a bridge-frame read 32 contiguous code bytes per leg and write 64 contiguous
// bytes of mix and 32 of codes.  What this buys over 5 lanes per bridge-frame (k_fused_w):
//   * 32 of 32 lanes busy instead of 30, the per-item setup paid once per 8 bridge-frames instead of 6;
//   * a lane's chunks belong to ONE bridge-frame, so the per-leg meter sums run in registers across the five passes:
//     4 partials per record instead of 10, written once per item; the finish gives every lane one leg record (G = 4)
//     instead of running a 24-lane leg branch and a 6-lane bridge branch one after the other.
// The pass order is rotated per bridge-frame (G = 4: by bfl % 4, the brute-forced conflict minimum of the LDS.64 slot
// reads: 8 instead of 20 wavefronts per five passes and half-warp; an 8-byte read cannot be conflict-free here).
constexpr int kQBf = 8;        // bridge-frames per item
constexpr int kQLanes = 4;     // lanes per bridge-frame
constexpr int kQPass = 5;      // 8-sample chunks per lane

// decode + meter + gain/accumulate one 8-sample chunk of one leg; the leg's meter sums run in the caller's registers
// selector constants of the lookups and the byte sum, held in registers the compiler cannot rematerialise (it otherwise
// rebuilds them with a UMOV / IMAD.MOV in every leg block of every pass: 7 of 220 warp-instructions per bridge-frame)
struct q_consts { uint32_t k0, k1, k2, k3, ones; };
__device__ __forceinline__ uint32_t lut_lookup_r(uint32_t lane_base, uint32_t word, uint32_t ksel)
{
    uint32_t v;
    asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(__dp4a(word, ksel, lane_base)));
    return v;
}
__device__ __forceinline__ void sq_acc(uint32_t &sq, uint32_t x)      // sq += x * x as ONE dependent IMAD (no IADD tree)
{
    asm("mad.lo.u32 %0, %1, %1, %0;" : "+r"(sq) : "r"(x));
}
template <bool kSigned, int kMode>
__device__ __forceinline__ void leg_chunk8(uint32_t lane_base, uint2 w, uint32_t sel, int adj, int (&acc)[8], uint32_t &sq,
                                           uint32_t &mx, int &bsum, const q_consts &K)
{
    const uint32_t wd[2] = {w.x, w.y};
#pragma unroll
    for (int j = 0; j < 2; j++) {
        const uint32_t e0 = lut_lookup_r(lane_base, wd[j], K.k0), e1 = lut_lookup_r(lane_base, wd[j], K.k1);
        const uint32_t e2 = lut_lookup_r(lane_base, wd[j], K.k2), e3 = lut_lookup_r(lane_base, wd[j], K.k3);
        const uint32_t x0 = e0 & 0xFFFFu, x1 = e1 & 0xFFFFu, x2 = e2 & 0xFFFFu, x3 = e3 & 0xFFFFu;
        sq_acc(sq, x0); sq_acc(sq, x1); sq_acc(sq, x2); sq_acc(sq, x3);       // 40 samples * 8064^2 < 2^32
        mx = max_u16x2(max_u16x2(mx, e0), e1); mx = max_u16x2(max_u16x2(mx, e2), e3);
        bsum = kSigned ? __dp4a((int)wd[j], (int)K.ones, bsum) : (int)__dp4a(wd[j], K.ones, (uint32_t)bsum);
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
}

// kPkt / kTx as in k_fused_w: the codes are read straight out of the raw 180-byte packets (G = 4: an item is 8 * 720
// contiguous bytes, still ONE bulk copy; 23 warps fit the shared memory), the bridge output leaves as finished packets.
template <int G, bool kSigned, int kWarps, bool kOpt, bool kPkt = false, bool kTx = false>
__global__ void __launch_bounds__((kWarps > 23 ? kWarps : 23) * 32, 1) k_fused_q(const FusedParams q)   // (fewer warps keep the 80-register budget)
{
    constexpr int kLegBytes = kPkt ? IGD_PKT_MAX : IGD_FRAME, kLegOff = kPkt ? IGD_PKT_HDR : 0;
    constexpr int kBfBytes = G * kLegBytes, kSlotBytes = kQBf * kBfBytes;
    static_assert(!kPkt || (kBfBytes % 16) == 0, "packet form: a ragged last item must still be a multiple of 16 bytes");
    constexpr int kLegRecs = kQBf * G;             // leg records per item (<= 32)
    constexpr int kLegStride = kLegRecs + 1;       // partial layout [quarter][record], odd stride: conflict-free both ways
    constexpr int kBrStride = kQBf + 1;
    constexpr int kParts = kQLanes * kLegStride + kQLanes * kBrStride;
    __shared__ uint64_t bars[kWarps];
    __shared__ __align__(16) uint32_t enc_tab[2][8];
    __shared__ __align__(16) uint32_t k_tab[12];
    extern __shared__ __align__(128) uint8_t smem[];
    uint32_t *lut = reinterpret_cast<uint32_t *>(smem);
    const uint32_t lut_bytes = shared_addr(smem);

    const int t = threadIdx.x;
    const uint32_t lane = t & 31;
    const uint32_t warp = __shfl_sync(0xFFFFFFFFu, (uint32_t)t >> 5, 0);
    const uint32_t slot_s = shared_addr(smem + kLutBytes) + warp * kSlotBytes;
    uint2 *part = reinterpret_cast<uint2 *>(smem + kLutBytes + (size_t)kWarps * kSlotBytes) + (size_t)warp * kParts;
    uint2 *bpart = part + kQLanes * kLegStride;
    const uint32_t bar_s = shared_addr(bars) + warp * 8;

    build_decode_lut_abs(lut, (int)warp, (int)lane, kWarps);
    if (t < 2) {
        const enc_pk e = enc_pk_make(t);
        enc_tab[t][0] = e.bias_pos; enc_tab[t][1] = e.bias_x; enc_tab[t][2] = e.hi_pos; enc_tab[t][3] = e.hi_x;
        enc_tab[t][4] = e.thr; enc_tab[t][5] = e.mask4; enc_tab[t][6] = 0u; enc_tab[t][7] = 0u;
    }
    if (t < 9) k_tab[t] = t < 4 ? 0x80u << (8 * t) : t == 4 ? 0x01010101u : t == 5 ? 0x0001u : t == 6 ? 0x0100u : t == 7 ? 0x4B000000u
                                                                                                                  : 0x3C000000u;
    if (lane == 0) {
        mbar_init(bar_s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const uint32_t total_bf = (uint32_t)q.total_bf;
    const uint32_t items = (total_bf + kQBf - 1) / kQBf;
    const uint32_t nw = gridDim.x * kWarps;
    uint32_t item = warp * gridDim.x + blockIdx.x;
    auto fetch = [&](uint32_t it_idx) {
        const uint32_t bf0 = it_idx * kQBf;
        const uint32_t left = total_bf - bf0;
        const uint32_t bytes = (left < (uint32_t)kQBf ? left : (uint32_t)kQBf) * kBfBytes;
        if (lane == 0) {
            mbar_expect_tx(bar_s, bytes);
            bulk_g2s(slot_s, q.codes + (size_t)bf0 * kBfBytes, bytes, bar_s);
        }
    };
    if (item < items) fetch(item);

    const uint32_t bfl = lane >> 2, qd = lane & 3u;
    // first pass of this lane: the four bridge-frames of a half-warp must start in four different 8-bank groups;
    // slot word address = bfl * 40 G + 8 pass + 2 qd, so the group is (5 G bfl + pass) % 4 -- distinct by itself for odd G
    // (packet form, 720-byte bridge-frames: brute force finds nothing better than no rotation)
    const uint32_t rot = kPkt ? 0u : G == 4 ? (bfl & 3u) : G == 2 ? ((bfl >> 1) & 1u) : 0u;
    const uint32_t src = slot_s + bfl * kBfBytes + kLegOff + qd * 8;   // + g * leg bytes + pass * 32
    const uint32_t lane4 = lut_bytes + 4u * lane;
    const uint32_t b_step = (uint32_t)(((unsigned long long)nw * kQBf) % (uint32_t)q.B);
    uint32_t b = (item * kQBf + bfl) % (uint32_t)q.B;
    // laws of bridge bb as loaded (G = 4: one word; else one byte per leg in a word) and the bridge's output law; the
    // bits are picked out when the item is worked on, so that the prefetch below does not wait for its own loads
    auto load_laws = [&](uint32_t bb, uint32_t &olaw) -> uint32_t {
        uint32_t r = 0;
        if (G == 4) r = __ldg(reinterpret_cast<const uint32_t *>(q.law + (size_t)bb * 4));
        else {
#pragma unroll
            for (int g = 0; g < G; g++) r |= (uint32_t)__ldg(q.law + (size_t)bb * G + g) << (8 * g);
        }
        olaw = __ldg(q.out_law + bb);
        return r;
    };
    // packet form: a leg whose packet is not a whole audio frame is silent (see k_fused_w)
    auto mark_no_audio = [&](uint2 g, uint32_t bfi) -> uint2 {
        const uint32_t *fw = reinterpret_cast<const uint32_t *>(q.fields + (size_t)bfi * G);
#pragma unroll
        for (int l = 0; l < G; l++) {
            if (igd_fields_no_audio(__ldg(fw + 4 * l + 1), __ldg(fw + 4 * l + 2), __ldg(fw + 4 * l + 3))) {
                if (l < 2) g.x |= 0x8000u << (16 * l); else g.y |= 0x8000u << (16 * (l - 2));
            }
        }
        return g;
    };
    uint2 gq = make_uint2(0u, 0u);
    uint32_t lwq = 0u, owq = 0u, pszq = 0u;
    if (item < items && item * kQBf + bfl < total_bf) {
        gq = load_gains<G>(q.gain + (size_t)(item * kQBf + bfl) * G);
        lwq = load_laws(b, owq);
        if (kPkt && q.fields) gq = mark_no_audio(gq, item * kQBf + bfl);
        if (kTx) pszq = __ldg(&q.plan[item * kQBf + bfl].size);
    }
    const bool want_mix = !kOpt || q.mix != nullptr, want_enc = !kOpt || q.enc != nullptr,
               want_meter = !kOpt || q.meter != nullptr, want_bmeter = !kOpt || q.bmeter != nullptr;
    q_consts K;      // read back from shared memory: ptxas folds a constant it can see, however it is written
    {
        const volatile uint32_t *kt = k_tab;
        K.k0 = kt[0]; K.k1 = kt[1]; K.k2 = kt[2]; K.k3 = kt[3]; K.ones = kt[4];
    }
    const uint32_t ksel_lo = *(const volatile uint32_t *)(k_tab + 5), ksel_hi = *(const volatile uint32_t *)(k_tab + 6),
                   kf_magic = *(const volatile uint32_t *)(k_tab + 7), kf_scale = *(const volatile uint32_t *)(k_tab + 8);
    // this lane's pass order as byte offsets into the bridge-frame's codes (32 * chunk row), one byte per pass
    uint32_t seq_lo = (32u * (rot % 5u)) | (32u * ((rot + 1u) % 5u)) << 8 | (32u * ((rot + 2u) % 5u)) << 16 |
                            (32u * ((rot + 3u) % 5u)) << 24,
             seq_hi = 32u * ((rot + 4u) % 5u);
    asm volatile("" : "+r"(seq_lo), "+r"(seq_hi));      // computed here, once: not sunk into the pass loop

    for (uint32_t it = 0; item < items; item += nw, it++) {
        const uint32_t bf = item * kQBf + bfl;
        const uint32_t next = item + nw;
        mbar_wait(bar_s, it & 1u);                           // this item's codes have landed
        const uint2 gcur = gq;
        // bit g = law of leg g, bit 8 = output law
        const uint32_t lcur = (lwq & 1u) | ((lwq >> 7) & 2u) | ((lwq >> 14) & 4u) | ((lwq >> 21) & 8u) | ((owq & 1u) << 8);
        const bool pay = kTx && pszq > IGD_PKT_HDR;               // the outgoing packet carries this tick's payload
        const bool valid = bf < total_bf;
        {
            b += b_step;
            if (b >= (uint32_t)q.B) b -= (uint32_t)q.B;
            const uint32_t bfn = bf + nw * kQBf;
            if (next < items && bfn < total_bf) {                // next item's gains and laws ride in three registers
                gq = load_gains<G>(q.gain + (size_t)bfn * G);
                lwq = load_laws(b, owq);
                if (kPkt && q.fields) gq = mark_no_audio(gq, bfn);
                if (kTx) pszq = __ldg(&q.plan[bfn].size);
            }
        }
        auto adj_of = [&](int g) -> uint32_t { return g == 0 ? (gcur.x & 0xFFFFu) : g == 1 ? (gcur.x >> 16) : g == 2 ? (gcur.y & 0xFFFFu) : (gcur.y >> 16); };
        const uint32_t orx = __reduce_or_sync(0xFFFFFFFFu, gcur.x), ory = G > 2 ? __reduce_or_sync(0xFFFFFFFFu, gcur.y) : 0u;
        const bool general = ((orx | ory) & 0xFEFFFEFFu) != 0u;      // also when IGD_GAIN_NO_AUDIO is set anywhere in the warp
        // one warp-uniform flag per leg (a packed mask costs an extract + compare per leg and pass)
        const bool open_leg[4] = {__any_sync(0xFFFFFFFFu, (gcur.x & 0xFFFFu) != 0u) != 0, (orx >> 16) != 0u,
                                  G > 2 && __any_sync(0xFFFFFFFFu, (gcur.y & 0xFFFFu) != 0u) != 0, (ory >> 16) != 0u};
        uint32_t n_open16 = (uint32_t)(__popc(nonzero_halves(gcur.x)) + __popc(nonzero_halves(gcur.y))) << 16;
        if (general)
            n_open16 = (uint32_t)(__popc(nonzero_halves(gcur.x & ~(((gcur.x & 0x80008000u) >> 15) * 0xFFFFu))) +
                                  __popc(nonzero_halves(gcur.y & ~(((gcur.y & 0x80008000u) >> 15) * 0xFFFFu)))) << 16;
        const uint32_t o8_0 = bf * (IGD_FRAME / 8) + qd;
        // the G legs' meter sums of this lane's quarter, and the bridge's
        uint32_t msq[G], mmx[G];
        int mbs[G];
#pragma unroll
        for (int g = 0; g < G; g++) { msq[g] = 0u; mmx[g] = 0u; mbs[g] = 0; }
        uint32_t bmx = 0u, bmn = 0u;
        int esum = 0;
        enc_pk E;
        {
            const uint32_t *et = enc_tab[(lcur >> 8) & 1u];
            const uint4 e0 = *reinterpret_cast<const uint4 *>(et);
            const uint4 e1 = *reinterpret_cast<const uint4 *>(et + 4);
            E.bias_pos = e0.x; E.bias_x = e0.y; E.hi_pos = e0.z; E.hi_x = e0.w; E.thr = e1.x; E.mask4 = e1.y; E.zero2 = e1.z;
            E.sel_lo = ksel_lo; E.sel_hi = ksel_hi; E.f_magic = kf_magic; E.f_scale = __uint_as_float(kf_scale);
        }
        uint8_t *mixp = reinterpret_cast<uint8_t *>(q.mix) + (size_t)o8_0 * 16;
        uint8_t *encp = q.enc + (size_t)o8_0 * 8;
        uint8_t *txp = kTx ? q.tx_pkts + (size_t)bf * IGD_PKT_MAX + IGD_PKT_HDR + qd * 8 : nullptr;   // 4-byte aligned
        asm volatile("" : "+l"(mixp), "+l"(encp), "+l"(txp));          // per item, not per pass
#ifdef IGD_X_QUNROLL
#pragma unroll
#else
#pragma unroll 1
#endif
        for (int p = 0; p < kQPass; p++) {
            const uint32_t pp32 = __byte_perm(seq_lo, seq_hi, 0x7650u + (uint32_t)p);     // byte p, upper bytes zero
            uint2 wh[G];
#pragma unroll
            for (int g = 0; g < G; g++) {
                // packet form: payload at byte 20 of a 180-byte packet -- 8-byte aligned for odd g only
                if ((kBfBytes % 8) == 0 && ((g * kLegBytes + kLegOff) % 8) == 0)
                    asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(wh[g].x), "=r"(wh[g].y) : "r"(src + pp32 + g * kLegBytes));
                else {
                    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(wh[g].x) : "r"(src + pp32 + g * kLegBytes));
                    asm volatile("ld.shared.u32 %0, [%1+4];" : "=r"(wh[g].y) : "r"(src + pp32 + g * kLegBytes));
                }
            }
            if (p == kQPass - 1) {
                // every lane has read the last of its codes: refill the slot now, a whole pass of work ahead of the
                // next item (the fence completes the reads above and orders them before the bulk copy's
                // async-proxy writes, see k_fused_w)
                fence_proxy_async();
                __syncwarp();
                if (next < items) fetch(next);
            }
            int acc[8];
#pragma unroll
            for (int i = 0; i < 8; i++) acc[i] = 0;
            if (!general) {
#pragma unroll
                for (int g = 0; g < G; g++) {
                    const uint32_t lb = lane4 + (((lcur >> g) & 1u) << 15);
                    if (open_leg[g]) leg_chunk8<kSigned, 1>(lb, wh[g], adj_of(g), 0, acc, msq[g], mmx[g], mbs[g], K);
                    else leg_chunk8<kSigned, 0>(lb, wh[g], 0u, 0, acc, msq[g], mmx[g], mbs[g], K);
                }
            } else {
#pragma unroll
                for (int g = 0; g < G; g++) {
                    const uint32_t lb = lane4 + (((lcur >> g) & 1u) << 15);
                    const uint32_t a = adj_of(g);
                    leg_chunk8<kSigned, 2>(lb, wh[g], 0u, (a & IGD_GAIN_NO_AUDIO) ? 0 : (int)a, acc, msq[g], mmx[g], mbs[g], K);
                }
            }
            // bridge output of this 8-sample chunk
            uint32_t pk[4];
#pragma unroll
            for (int i = 0; i < 4; i++) pk[i] = pack_sat16(acc[2 * i + 1], acc[2 * i]);
            if (valid && want_mix) st16_stream(mixp + 2u * pp32, make_uint4(pk[0], pk[1], pk[2], pk[3]));
            bmx = max_s16x2(max_s16x2(bmx, pk[0]), pk[1]); bmx = max_s16x2(max_s16x2(bmx, pk[2]), pk[3]);
            bmn = min_s16x2(min_s16x2(bmn, pk[0]), pk[1]); bmn = min_s16x2(min_s16x2(bmn, pk[2]), pk[3]);
            const uint32_t c0 = encode4_packed(pk[0], pk[1], E), c1 = encode4_packed(pk[2], pk[3], E);
            if (valid && want_enc) __stcs(reinterpret_cast<uint2 *>(encp + pp32), make_uint2(c0, c1));
            if (kTx && valid) {      // payload bytes of the outgoing packet: the codes, or zeros (see k_fused_w)
                uint32_t *tp = reinterpret_cast<uint32_t *>(txp + pp32);
                __stcs(tp, pay ? c0 : 0u); __stcs(tp + 1, pay ? c1 : 0u);
            }
            if (kSigned) { esum = __dp4a((int)c0, (int)K.ones, esum); esum = __dp4a((int)c1, (int)K.ones, esum); }
            else esum = (int)__dp4a(c1, K.ones, __dp4a(c0, K.ones, (uint32_t)esum));
        }
        // ---- partials of this lane's quarter: {sum (x/4)^2, peak/4 | bytesum << 16} per leg, {codesum, peak | open << 16}
#pragma unroll
        for (int g = 0; g < G; g++) part[qd * kLegStride + bfl * G + g] = make_uint2(msq[g], __byte_perm(mmx[g], (uint32_t)mbs[g], 0x5410));
        {
            const uint32_t pk2 = max_u16x2(bmx, neg_16x2(bmn));
            bpart[qd * kBrStride + bfl] = make_uint2((uint32_t)esum, max(pk2 & 0xFFFFu, pk2 >> 16) | n_open16);
        }
        __syncwarp();
        // ---- finish: one leg record per lane (all 32 at G = 4), then one bridge record per lane 0..7
        {
            const uint32_t bf0 = item * kQBf;
            uint32_t sil_rec = 0u;
            if (general) {
                uint32_t silent;
                (void)split_no_audio(gcur, silent);
                sil_rec = (__shfl_sync(0xFFFFFFFFu, silent, (int)((lane / G) * kQLanes) & 31) >> (lane % G)) & 1u;
            }
            if (lane < (uint32_t)kLegRecs && bf0 + lane / G < total_bf && want_meter) {
                unsigned long long sq = 0; uint32_t pk = 0; int bsum = 0;
                if (!sil_rec) {
#pragma unroll
                    for (int j = 0; j < kQLanes; j++) {
                        const uint2 v = part[j * kLegStride + lane];
                        sq += v.x; pk = max_u16x2(pk, v.y); bsum = dp2a_lo(v.y, 0x0100u, bsum);
                    }
                }
                const igd_meter_rec r = meter_finish(sq << 4, (pk & 0xFFFFu) << 2, bsum, true);
                st16_stream(q.meter + ((size_t)bf0 * G + lane), *reinterpret_cast<const uint4 *>(&r));
            }
            if (kTx && lane < (uint32_t)kQBf && bf0 + lane < total_bf) {
                // the packet's 20-byte header (TransportAdapter.cpp:715-800) and its size, as in k_fused_w
                const uint32_t j = lane;
                const uint32_t *pr = reinterpret_cast<const uint32_t *>(q.plan + (bf0 + j));
                const uint32_t pw = __ldg(pr), psf = __ldg(pr + 1);      // word; size | flags << 16
                const uint32_t psize = psf & 0xFFFFu, pfl = psf >> 16;
                const uint32_t *hdr = reinterpret_cast<const uint32_t *>(q.rtp12 + (size_t)(bf0 + j) * 12);
                uint32_t v0 = __ldg(hdr) | 0x10u;                                          // x = 1 (:725)
                v0 = (v0 & ~0x8000u) | ((pfl & 2u) ? 0x8000u : 0u);                        // m (:715-723)
                if (pfl & 1u) v0 = (v0 & ~0x7F00u) | (123u << 8);                          // pt = 123
                uint32_t *tp = reinterpret_cast<uint32_t *>(q.tx_pkts + (size_t)(bf0 + j) * IGD_PKT_MAX);
                const bool any = psize != 0u;
                tp[0] = any ? v0 : 0u; tp[1] = any ? __ldg(hdr + 1) : 0u; tp[2] = any ? __ldg(hdr + 2) : 0u;
                tp[3] = any ? 0x01006701u : 0u;                                            // 0x0167, 0x0001 big-endian
                tp[4] = any ? __byte_perm(pw, 0u, 0x0123) : 0u;                            // htonl (:800)
                q.tx_sizes[bf0 + j] = psize;
            }
            if (lane < (uint32_t)kQBf && bf0 + lane < total_bf && want_bmeter) {
                int es = 0, hi = 0; uint32_t pk = 0;
#pragma unroll
                for (int j = 0; j < kQLanes; j++) {
                    const uint2 v = bpart[j * kBrStride + lane];
                    es += (int)v.x; pk = max_u16x2(pk, v.y); hi = dp2a_lo(v.y, 0x0100u, hi);
                }
                igd_bridge_rec r;
                r.bytemean_out = (uint8_t)igd_bytemean_from_sum(es, IGD_FRAME);
                r.n_open = (uint8_t)((uint32_t)hi / kQLanes);
                r.mix_peak = (uint16_t)(pk & 0xFFFFu);
                q.bmeter[bf0 + lane] = r;
            }
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------
// Eight legs (k_fused_h): the quarter-lane idea turned sideways.  An item is 4 bridge-frames with all 8 legs (8 * 160
// contiguous code bytes per bridge-frame, one 5120-byte bulk copy); the 8 lanes of a bridge-frame are its 8 sample
// columns: a lane walks five 4-sample chunks (column + 8 * pass) of ALL 8 legs, so a leg's meter sums run in three
// registers across the passes like in k_fused_q, the mix of a chunk never leaves the lane, and every lane finishes
// exactly one leg record per item (8 partials, one per column).  20 warps (96 registers: 24 of them the meters).
// Measured at 512 bridges x 8 legs x 1640 frames: 0.417 ms against k_fused_g's 0.468 (58 % / 52 % of the HBM peak).
// The template keeps the leg-group generalisation (GP = 16 / 32: lanes = leg group x column, the chunk's mix summed
// over the leg groups by shuffle-adds) but only GP = 8 is instantiated: with more than one leg group the saturate /
// compress / store of a chunk is done by a quarter or half of the lanes at the whole warp's issue cost, and the
// variant measured no faster than k_fused_g.
constexpr int kHLegs = 8;      // legs per lane
constexpr int kHCols = 8;      // sample columns per bridge-frame
constexpr int kHPass = 5;      // 4-sample chunks per lane and leg
constexpr int kHLegStride = 33;   // partial layout [column][record], odd stride

template <bool kSigned, int kMode>
__device__ __forceinline__ void leg_chunk4(uint32_t lane_base, uint32_t w, uint32_t sel, int adj, int (&acc)[4], uint32_t &sq,
                                           uint32_t &mx, int &bsum, const q_consts &K)
{
    const uint32_t e0 = lut_lookup_r(lane_base, w, K.k0), e1 = lut_lookup_r(lane_base, w, K.k1);
    const uint32_t e2 = lut_lookup_r(lane_base, w, K.k2), e3 = lut_lookup_r(lane_base, w, K.k3);
    const uint32_t x0 = e0 & 0xFFFFu, x1 = e1 & 0xFFFFu, x2 = e2 & 0xFFFFu, x3 = e3 & 0xFFFFu;
    sq_acc(sq, x0); sq_acc(sq, x1); sq_acc(sq, x2); sq_acc(sq, x3);       // 20 samples * 8064^2 < 2^32
    mx = max_u16x2(max_u16x2(mx, e0), e1); mx = max_u16x2(max_u16x2(mx, e2), e3);
    bsum = kSigned ? __dp4a((int)w, (int)K.ones, bsum) : (int)__dp4a(w, K.ones, (uint32_t)bsum);
    if (kMode == 1) {
        acc[0] = dp2a_lo(e0, sel, acc[0]); acc[1] = dp2a_lo(e1, sel, acc[1]);
        acc[2] = dp2a_lo(e2, sel, acc[2]); acc[3] = dp2a_lo(e3, sel, acc[3]);
    } else if (kMode == 2) {
        const int s0 = (int)e0 < 0 ? -(int)x0 : (int)x0, s1 = (int)e1 < 0 ? -(int)x1 : (int)x1;
        const int s2 = (int)e2 < 0 ? -(int)x2 : (int)x2, s3 = (int)e3 < 0 ? -(int)x3 : (int)x3;
        acc[0] += clamp16((4 * s0 * adj) >> 7); acc[1] += clamp16((4 * s1 * adj) >> 7);
        acc[2] += clamp16((4 * s2 * adj) >> 7); acc[3] += clamp16((4 * s3 * adj) >> 7);
    }
}

template <int GP, bool kSigned, int kWarps, bool kOpt>
__global__ void __launch_bounds__(kWarps * 32, 1) k_fused_h(const FusedParams q)
{
    static_assert(GP == 8, "leg groups (GP = 16 / 32) need the pass order shared by the lanes that are summed: not finished");
    constexpr int kBf = 32 / GP;                        // bridge-frames per item
    constexpr int kLG = GP / kHLegs;                    // leg groups = lanes that share a sample column
    constexpr int kSlotBytes = kBf * GP * IGD_FRAME;    // 5120
    constexpr int kParts = kHCols * kHLegStride;
    __shared__ uint64_t bars[kWarps];
    __shared__ __align__(16) uint32_t enc_tab[2][8];
    __shared__ __align__(16) uint32_t k_tab[12];
    extern __shared__ __align__(128) uint8_t smem[];
    uint32_t *lut = reinterpret_cast<uint32_t *>(smem);
    const uint32_t lut_bytes = shared_addr(smem);

    const int t = threadIdx.x;
    const uint32_t lane = t & 31;
    const uint32_t warp = __shfl_sync(0xFFFFFFFFu, (uint32_t)t >> 5, 0);
    const uint32_t slot_s = shared_addr(smem + kLutBytes) + warp * kSlotBytes;
    uint2 *part = reinterpret_cast<uint2 *>(smem + kLutBytes + (size_t)kWarps * kSlotBytes) + (size_t)warp * kParts;
    const uint32_t bar_s = shared_addr(bars) + warp * 8;

    build_decode_lut_abs(lut, (int)warp, (int)lane, kWarps);
    if (t < 2) {
        const enc_pk e = enc_pk_make(t);
        enc_tab[t][0] = e.bias_pos; enc_tab[t][1] = e.bias_x; enc_tab[t][2] = e.hi_pos; enc_tab[t][3] = e.hi_x;
        enc_tab[t][4] = e.thr; enc_tab[t][5] = e.mask4; enc_tab[t][6] = 0u; enc_tab[t][7] = 0u;
    }
    if (t < 9) k_tab[t] = t < 4 ? 0x80u << (8 * t) : t == 4 ? 0x01010101u : t == 5 ? 0x0001u : t == 6 ? 0x0100u : t == 7 ? 0x4B000000u
                                                                                                                  : 0x3C000000u;
    if (lane == 0) {
        mbar_init(bar_s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const uint32_t G = (uint32_t)q.G;                   // a multiple of 8, <= GP
    const uint32_t bf_bytes = G * IGD_FRAME;
    const uint32_t total_bf = (uint32_t)q.total_bf;
    const uint32_t items = (total_bf + kBf - 1) / kBf;
    const uint32_t nw = gridDim.x * kWarps;
    uint32_t item = warp * gridDim.x + blockIdx.x;
    auto fetch = [&](uint32_t it_idx) {
        const uint32_t bf0 = it_idx * kBf;
        const uint32_t left = total_bf - bf0;
        const uint32_t bytes = (left < (uint32_t)kBf ? left : (uint32_t)kBf) * bf_bytes;
        if (lane == 0) {
            mbar_expect_tx(bar_s, bytes);
            bulk_g2s(slot_s, q.codes + (size_t)bf0 * bf_bytes, bytes, bar_s);
        }
    };
    if (item < items) fetch(item);

    const uint32_t bfl = lane / GP, r = lane % GP, sc = r & 7u, lg = r >> 3;
    const bool active = lg * kHLegs < G;               // G = 24 under GP = 32: the lanes of leg group 3 idle
    // pass order rotated per group of 8 lanes: the groups of a warp read 8 consecutive words each, all at a multiple of
    // 32 words from each other -- four different rows of the slot keep them on different banks
    const uint32_t rot = (lane >> 3) & 3u;
    uint32_t seq_lo = (32u * (rot % 5u)) | (32u * ((rot + 1u) % 5u)) << 8 | (32u * ((rot + 2u) % 5u)) << 16 |
                      (32u * ((rot + 3u) % 5u)) << 24,
             seq_hi = 32u * ((rot + 4u) % 5u);
    asm volatile("" : "+r"(seq_lo), "+r"(seq_hi));
    const uint32_t src = slot_s + bfl * bf_bytes + lg * (kHLegs * IGD_FRAME) + sc * 4;    // + leg * 160 + pass * 32
    const uint32_t lane4 = lut_bytes + 4u * lane;
    const uint32_t b_step = (uint32_t)(((unsigned long long)nw * kBf) % (uint32_t)q.B);
    uint32_t b = (item * kBf + bfl) % (uint32_t)q.B;
    // record of this lane in the finish: (bridge-frame, leg) = (lane / G, lane % G); its partials come from the lanes
    // of that bridge-frame's leg group, its no-audio flag from any of them
    const uint32_t rec_bf = lane / G, rec_leg = lane - rec_bf * G;
    const uint32_t rec_src = rec_bf * GP + (rec_leg >> 3) * 8u;

    uint4 gq = make_uint4(0u, 0u, 0u, 0u);
    uint2 lwq = make_uint2(0u, 0u);
    uint32_t owq = 0u;
    auto prefetch = [&](uint32_t bfi, uint32_t bb) {
        gq = __ldg(reinterpret_cast<const uint4 *>(q.gain + (size_t)bfi * G + lg * kHLegs));
        lwq = __ldg(reinterpret_cast<const uint2 *>(q.law + (size_t)bb * G + lg * kHLegs));
        owq = __ldg(q.out_law + bb);
    };
    if (active && item < items && item * kBf + bfl < total_bf) prefetch(item * kBf + bfl, b);

    const bool want_mix = !kOpt || q.mix != nullptr, want_enc = !kOpt || q.enc != nullptr,
               want_meter = !kOpt || q.meter != nullptr, want_bmeter = !kOpt || q.bmeter != nullptr;
    q_consts K;
    {
        const volatile uint32_t *kt = k_tab;
        K.k0 = kt[0]; K.k1 = kt[1]; K.k2 = kt[2]; K.k3 = kt[3]; K.ones = kt[4];
    }
    const uint32_t ksel_lo = *(const volatile uint32_t *)(k_tab + 5), ksel_hi = *(const volatile uint32_t *)(k_tab + 6),
                   kf_magic = *(const volatile uint32_t *)(k_tab + 7), kf_scale = *(const volatile uint32_t *)(k_tab + 8);

    for (uint32_t it = 0; item < items; item += nw, it++) {
        const uint32_t bf = item * kBf + bfl;
        const uint32_t next = item + nw;
        mbar_wait(bar_s, it & 1u);
        const uint4 gcur = gq;
        const uint32_t gw[4] = {gcur.x, gcur.y, gcur.z, gcur.w};
        // bit j = law of this lane's leg j
        const uint32_t lbits = (lwq.x & 1u) | ((lwq.x >> 7) & 2u) | ((lwq.x >> 14) & 4u) | ((lwq.x >> 21) & 8u) |
                               ((lwq.y & 1u) << 4) | ((lwq.y >> 3) & 0x20u) | ((lwq.y >> 10) & 0x40u) | ((lwq.y >> 17) & 0x80u);
        const uint32_t olaw = owq & 1u;
        const bool valid = bf < total_bf;
        {
            b += b_step;
            if (b >= (uint32_t)q.B) b -= (uint32_t)q.B;
            const uint32_t bfn = bf + nw * kBf;
            gq = make_uint4(0u, 0u, 0u, 0u);
            if (active && next < items && bfn < total_bf) prefetch(bfn, b);
        }
        auto adj_of = [&](int j) -> uint32_t { return (j & 1) ? (gw[j >> 1] >> 16) : (gw[j >> 1] & 0xFFFFu); };
        uint32_t orw[4];
#pragma unroll
        for (int k = 0; k < 4; k++) orw[k] = __reduce_or_sync(0xFFFFFFFFu, gw[k]);
        const bool general = ((orw[0] | orw[1] | orw[2] | orw[3]) & 0xFEFFFEFFu) != 0u;
        bool open_leg[kHLegs];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            open_leg[2 * k] = __any_sync(0xFFFFFFFFu, (gw[k] & 0xFFFFu) != 0u) != 0;
            open_leg[2 * k + 1] = (orw[k] >> 16) != 0u;
        }
        // open legs among this lane's eight; one lane per leg group (column 0) carries the count into the bridge record
        uint32_t n_open = 0, silent8 = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            uint32_t w = gw[k];
            if (general) {
                const uint32_t f = w & 0x80008000u;
                silent8 |= (((f >> 15) & 1u) | ((f >> 30) & 2u)) << (2 * k);
                w &= ~((f >> 15) * 0xFFFFu);
            }
            n_open += (uint32_t)__popc(nonzero_halves(w));
        }
        if (sc != 0u) n_open = 0u;

        uint32_t msq[kHLegs], mmx[kHLegs];
        int mbs[kHLegs];
#pragma unroll
        for (int j = 0; j < kHLegs; j++) { msq[j] = 0u; mmx[j] = 0u; mbs[j] = 0; }
        uint32_t bmx = 0u, bmn = 0u;
        int esum = 0;
        enc_pk E;
        {
            const uint32_t *et = enc_tab[olaw];
            const uint4 e0 = *reinterpret_cast<const uint4 *>(et);
            const uint4 e1 = *reinterpret_cast<const uint4 *>(et + 4);
            E.bias_pos = e0.x; E.bias_x = e0.y; E.hi_pos = e0.z; E.hi_x = e0.w; E.thr = e1.x; E.mask4 = e1.y; E.zero2 = e1.z;
            E.sel_lo = ksel_lo; E.sel_hi = ksel_hi; E.f_magic = kf_magic; E.f_scale = __uint_as_float(kf_scale);
        }
        uint8_t *mixp = reinterpret_cast<uint8_t *>(q.mix) + ((size_t)bf * IGD_FRAME + sc * 4) * 2;
        uint8_t *encp = q.enc + (size_t)bf * IGD_FRAME + sc * 4;
        asm volatile("" : "+l"(mixp), "+l"(encp));
#pragma unroll 1
        for (int p = 0; p < kHPass; p++) {
            const uint32_t pp32 = __byte_perm(seq_lo, seq_hi, 0x7650u + (uint32_t)p);
            uint32_t wh[kHLegs];
#pragma unroll
            for (int j = 0; j < kHLegs; j++)
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(wh[j]) : "r"(src + pp32 + j * IGD_FRAME));
            if (p == kHPass - 1) {      // refill right behind the last reads (see k_fused_q)
                fence_proxy_async();
                __syncwarp();
                if (next < items) fetch(next);
            }
            int acc[4] = {0, 0, 0, 0};
            if (!general) {
#pragma unroll
                for (int j = 0; j < kHLegs; j++) {
                    const uint32_t lb = lane4 + (((lbits >> j) & 1u) << 15);
                    if (open_leg[j]) leg_chunk4<kSigned, 1>(lb, wh[j], adj_of(j), 0, acc, msq[j], mmx[j], mbs[j], K);
                    else leg_chunk4<kSigned, 0>(lb, wh[j], 0u, 0, acc, msq[j], mmx[j], mbs[j], K);
                }
            } else {
#pragma unroll
                for (int j = 0; j < kHLegs; j++) {
                    const uint32_t lb = lane4 + (((lbits >> j) & 1u) << 15);
                    const uint32_t a = adj_of(j);
                    leg_chunk4<kSigned, 2>(lb, wh[j], 0u, (a & IGD_GAIN_NO_AUDIO) ? 0 : (int)a, acc, msq[j], mmx[j], mbs[j], K);
                }
            }
            // the chunk's mix is the sum over the leg groups of this bridge-frame
            if (kLG >= 2) {
#pragma unroll
                for (int i = 0; i < 4; i++) acc[i] += __shfl_xor_sync(0xFFFFFFFFu, acc[i], 8);
            }
            if (kLG >= 4) {
#pragma unroll
                for (int i = 0; i < 4; i++) acc[i] += __shfl_xor_sync(0xFFFFFFFFu, acc[i], 16);
            }
            if (kLG == 1 || lg == 0u) {
                const uint32_t pk0 = pack_sat16(acc[1], acc[0]), pk1 = pack_sat16(acc[3], acc[2]);
                if (valid && want_mix) __stcs(reinterpret_cast<uint2 *>(mixp + 2u * pp32), make_uint2(pk0, pk1));
                bmx = max_s16x2(max_s16x2(bmx, pk0), pk1);
                bmn = min_s16x2(min_s16x2(bmn, pk0), pk1);
                const uint32_t c0 = encode4_packed(pk0, pk1, E);
                if (valid && want_enc) __stcs(reinterpret_cast<uint32_t *>(encp + pp32), c0);
                esum = kSigned ? __dp4a((int)c0, (int)K.ones, esum) : (int)__dp4a(c0, K.ones, (uint32_t)esum);
            }
        }
        // ---- partials of this lane's column, one per leg
        if (active) {
#pragma unroll
            for (int j = 0; j < kHLegs; j++)
                part[sc * kHLegStride + bfl * G + lg * kHLegs + j] = make_uint2(msq[j], __byte_perm(mmx[j], (uint32_t)mbs[j], 0x5410));
        }
        // ---- bridge record: code sum and open count (packed: |code sum| < 2^19), peak -- reduced over the bridge-frame's lanes
        {
            const uint32_t pk2 = max_u16x2(bmx, neg_16x2(bmn));
            uint32_t peak = max(pk2 & 0xFFFFu, pk2 >> 16);
            int es = esum + (int)(n_open << 20);
#pragma unroll
            for (int m = 1; m < GP; m <<= 1) {
                es += __shfl_xor_sync(0xFFFFFFFFu, es, m);
                peak = max(peak, __shfl_xor_sync(0xFFFFFFFFu, peak, m));
            }
            if (r == 0u && valid && want_bmeter) {
                const int n = (es + (1 << 19)) >> 20;
                igd_bridge_rec br;
                br.bytemean_out = (uint8_t)igd_bytemean_from_sum(es - (n << 20), IGD_FRAME);
                br.n_open = (uint8_t)n;
                br.mix_peak = (uint16_t)peak;
                q.bmeter[bf] = br;
            }
        }
        __syncwarp();
        // ---- finish: one leg record per lane
        {
            const uint32_t bf0 = item * kBf;
            uint32_t sil_rec = 0u;
            if (general) sil_rec = (__shfl_sync(0xFFFFFFFFu, silent8, (int)(rec_src & 31u)) >> (rec_leg & 7u)) & 1u;
            if (lane < (uint32_t)kBf * G && bf0 + rec_bf < total_bf && want_meter) {
                unsigned long long sq = 0; uint32_t pk = 0; int bsum = 0;
                if (!sil_rec) {
#pragma unroll
                    for (int j = 0; j < kHCols; j++) {
                        const uint2 v = part[j * kHLegStride + lane];
                        sq += v.x; pk = max_u16x2(pk, v.y); bsum = dp2a_lo(v.y, 0x0100u, bsum);
                    }
                }
                const igd_meter_rec rr = meter_finish(sq << 4, (pk & 0xFFFFu) << 2, bsum, true);
                st16_stream(q.meter + ((size_t)bf0 * G + lane), *reinterpret_cast<const uint4 *>(&rr));
            }
        }
        __syncwarp();
    }
}

template <int GP, bool kSigned, int kWarps>
cudaError_t launch_fused_h(const igd_launch_cfg &c, const FusedParams &q)
{
    const bool all_out = q.mix && q.enc && q.meter && q.bmeter;
    auto kern = all_out ? k_fused_h<GP, kSigned, kWarps, false> : k_fused_h<GP, kSigned, kWarps, true>;
    const size_t smem = kLutBytes + (size_t)kWarps * (32 * IGD_FRAME + kHCols * kHLegStride * 8);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const long long items = (q.total_bf + (32 / GP) - 1) / (32 / GP);
    long long grid = c.sm_count;
    if (grid > items) grid = items;
    kern<<<(int)grid, kWarps * 32, smem, c.stream>>>(q);
    return cudaGetLastError();
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

// Meter-only table of k_fused_g for legs shut in the whole warp: clean 32-bit |x|/4 entries (no
// extract before the square), indexed by the 7 magnitude bits of the code (the sign of both laws
// is bit 7), rows of 256 B = [law][lane column] so that ONE PRMT builds the lookup offset
// (code << 8 | law << 7 | lane << 2): the address moves from the FMA pipe (IDP.4A) to the ALU.
constexpr int kMlutBytes = 128 * 256;
__device__ __forceinline__ void build_meter_lut(uint32_t *mlut, int tid, int nthreads)
{
    for (int i = tid; i < kMlutBytes / 4; i += nthreads) {
        const uint32_t code = (uint32_t)i >> 6, law = ((uint32_t)i >> 5) & 1u;
        const int x = law ? igd_ulaw2lin(code) : igd_alaw2lin(code);
        mlut[i] = (uint32_t)(abs(x) >> 2);
    }
}
template <bool kSigned>
__device__ __forceinline__ uint2 leg_chunk_shut(const uint8_t *mlut, uint32_t lanereg, uint4 w)
{
    const uint32_t wd[4] = {w.x, w.y, w.z, w.w};
    uint32_t sq = 0, mx = 0;
    int bsum = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const uint32_t wm = wd[j] & 0x7F7F7F7Fu;
        const uint32_t e0 = *reinterpret_cast<const uint32_t *>(mlut + __byte_perm(wm, lanereg, 0x7604));
        const uint32_t e1 = *reinterpret_cast<const uint32_t *>(mlut + __byte_perm(wm, lanereg, 0x7614));
        const uint32_t e2 = *reinterpret_cast<const uint32_t *>(mlut + __byte_perm(wm, lanereg, 0x7624));
        const uint32_t e3 = *reinterpret_cast<const uint32_t *>(mlut + __byte_perm(wm, lanereg, 0x7634));
        sq += e0 * e0 + e1 * e1 + e2 * e2 + e3 * e3;
        mx = max_u16x2(max_u16x2(mx, e0), e1); mx = max_u16x2(max_u16x2(mx, e2), e3);
        bsum = kSigned ? __dp4a((int)wd[j], 0x01010101, bsum) : (int)__dp4a(wd[j], 0x01010101u, (uint32_t)bsum);
    }
    return make_uint2(sq, __byte_perm(mx, (uint32_t)bsum, 0x5410));
}

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
    const uint8_t *mlut = smem + kLutBytes;
    const uint32_t slot_s = shared_addr(smem + kLutBytes + kMlutBytes) + warp * (2 * kGSlotBytes);
    uint2 *part = reinterpret_cast<uint2 *>(smem + kLutBytes + kMlutBytes + (size_t)kWarps * 2 * kGSlotBytes) + (size_t)warp * kGParts;
    uint2 *bpart = part + kGBf * kGLegs * kPst;
    const uint32_t bar_s = shared_addr(bars) + warp * 16;

    build_decode_lut_wide(lut, t, kWarps * 32);
    build_meter_lut(reinterpret_cast<uint32_t *>(smem + kLutBytes), t, kWarps * 32);
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
    const size_t row_bytes = (size_t)G * IGD_FRAME;       // a bridge-frame's legs
    // unit = (item, leg group); the n-th unit of this warp lives in slot n & 1
    auto fetch = [&](uint32_t it_idx, uint32_t grp, uint32_t n) {
        const uint32_t bf0 = it_idx * kGBf;
        const uint32_t left = total_bf - bf0;
        const uint32_t nbf = left < (uint32_t)kGBf ? left : (uint32_t)kGBf;
        const uint32_t legs = min((uint32_t)kGLegs, G - grp * kGLegs), bytes = legs * IGD_FRAME;
        if (lane == 0) {
            const uint32_t bs = bar_s + (n & 1u) * 8, dst = slot_s + (n & 1u) * kGSlotBytes;
            mbar_expect_tx(bs, nbf * bytes);
            const uint8_t *g0 = q.codes + ((size_t)bf0 * G + (size_t)grp * kGLegs) * IGD_FRAME;
            if (nbf == (uint32_t)kGBf) {          // whole item: three copies, row pointers by addition
                const uint8_t *g1 = g0 + row_bytes, *g2 = g1 + row_bytes;
                bulk_g2s(dst, g0, bytes, bs);
                bulk_g2s(dst + kGStride, g1, bytes, bs);
                bulk_g2s(dst + 2 * kGStride, g2, bytes, bs);
            } else {
                bulk_g2s(dst, g0, bytes, bs);
                if (nbf > 1) bulk_g2s(dst + kGStride, g0 + row_bytes, bytes, bs);
            }
        }
    };
    const bool worker = lane < kGBf * kChunks;
    const uint32_t bfl = worker ? lane / kChunks : 0u, c = worker ? lane - bfl * kChunks : 0u;
    const uint32_t src_off = bfl * kGStride + c * 16;
    const uint32_t lane4 = lut_bytes + 8u * (lane & 15u);
    const uint32_t b_step = (uint32_t)(((unsigned long long)nw * kGBf) % (uint32_t)q.B);
    uint32_t b = (item * kGBf + bfl) % (uint32_t)q.B;
    const bool wide = (G & 3u) == 0u && (reinterpret_cast<uintptr_t>(q.gain) & 7u) == 0 && (reinterpret_cast<uintptr_t>(q.law) & 3u) == 0;
    // gains (two packed words) and laws (4 bits) of one unit for this lane's bridge-frame
    auto load_unit = [&](uint32_t it_idx, uint32_t grp, uint32_t bb, uint2 &gq, uint32_t &lw) {
        gq = make_uint2(0u, 0u); lw = 0u;
        const uint32_t bf = it_idx * kGBf + bfl;
        if (!worker || it_idx >= items || bf >= total_bf) return;
        const uint32_t legs = min((uint32_t)kGLegs, G - grp * kGLegs);
        const uint16_t *gp = q.gain + (size_t)bf * G + grp * kGLegs;
        const uint8_t *lp = q.law + (size_t)bb * G + grp * kGLegs;
        if (wide) {                          // G % 4 == 0 and aligned arrays: one 8-byte and one 4-byte load
            gq = __ldg(reinterpret_cast<const uint2 *>(gp));
            const uint32_t l4 = __ldg(reinterpret_cast<const uint32_t *>(lp));
            lw = (l4 & 1u) | ((l4 >> 7) & 2u) | ((l4 >> 14) & 4u) | ((l4 >> 21) & 8u);
            return;
        }
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
            // the other slot was drained one unit ago (its loads were consumed by that unit's lookups);
            // the fence makes that generic-before-async order formal
            fence_proxy_async();
            if (it_n < items) fetch(it_n, grp_n, n + 1);
            mbar_wait(bar_s + (n & 1u) * 8, (n >> 1) & 1u);
            const uint2 gcur = gq;
            const uint32_t lcur = lwq;
            load_unit(it_n, grp_n, last ? b_next : b, gq, lwq);          // next unit's gains / laws ride in registers
            const uint32_t orx = __reduce_or_sync(0xFFFFFFFFu, gcur.x), ory = __reduce_or_sync(0xFFFFFFFFu, gcur.y);
            const bool general = ((orx | ory) & 0xFEFFFEFFu) != 0u;
            const uint32_t open_mask = ((orx & 0xFFFFu) ? 1u : 0u) | ((orx >> 16) ? 2u : 0u) | ((ory & 0xFFFFu) ? 4u : 0u) |
                                       ((ory >> 16) ? 8u : 0u);
            auto adj_of = [&](int g) -> uint32_t { return g == 0 ? (gcur.x & 0xFFFFu) : g == 1 ? (gcur.x >> 16) : g == 2 ? (gcur.y & 0xFFFFu) : (gcur.y >> 16); };
            if (!general) n_open += (uint32_t)(__popc(nonzero_halves(gcur.x)) + __popc(nonzero_halves(gcur.y)));
            else          // IGD_GAIN_NO_AUDIO makes `general` true: flagged legs do not count as open
                n_open += (uint32_t)(__popc(nonzero_halves(gcur.x & ~(((gcur.x & 0x80008000u) >> 15) * 0xFFFFu))) +
                                     __popc(nonzero_halves(gcur.y & ~(((gcur.y & 0x80008000u) >> 15) * 0xFFFFu))));
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
                    if (general) ph = leg_chunk_u<kSigned, 2, true>(lb, wh[g], 0u, (adj_of(g) & IGD_GAIN_NO_AUDIO) ? 0 : (int)adj_of(g), acc);
                    else if ((open_mask >> g) & 1u) ph = leg_chunk_u<kSigned, 1, true>(lb, wh[g], adj_of(g), 0, acc);
                    else ph = leg_chunk_shut<kSigned>(mlut, 4u * lane + (((lcur >> g) & 1u) << 7), wh[g]);
                    if (valid) part[(bfl * kGLegs + g) * kPst + c] = ph;
                }
            }
            __syncwarp();
            // the record of leg (lane / 4, lane % 4): the lanes of that bridge-frame hold its no-audio flags
            uint32_t sil_rec = 0u;
            if (general) {
                uint32_t silent;
                (void)split_no_audio(gcur, silent);
                sil_rec = (__shfl_sync(0xFFFFFFFFu, silent, (int)((lane / kGLegs) * kChunks) & 31) >> (lane % kGLegs)) & 1u;
            }
            if (lane < kGBf * kGLegs) {                                  // this group's leg records
                const uint32_t fb = lane / kGLegs, g = lane - fb * kGLegs;
                if (g < legs && item * kGBf + fb < total_bf && q.meter != nullptr) {
                    const uint2 *src_p = part + lane * kPst;
                    unsigned long long sq = 0; uint32_t pk = 0; int bsum = 0;
#pragma unroll
                    for (int i = 0; i < kChunks; i++) {
                        const uint2 v = sil_rec ? make_uint2(0u, 0u) : src_p[i];
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
            E.zero2 = 0u;
        }
        const size_t o16 = (size_t)bf * kChunks + c;
        const uint2 mo = mix_out_chunk<kSigned>(acc, E, q.mix + o16 * 16, q.enc + o16 * 16, valid && q.mix != nullptr,
                                                valid && q.enc != nullptr);
        if (valid) bpart[bfl * kPst + c] = make_uint2(mo.x, mo.y | (n_open << 16));
        __syncwarp();
        if (lane < kGBf && item * kGBf + lane < total_bf && q.bmeter != nullptr) {
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
                part[bfl * kPst + p] = (a & IGD_GAIN_NO_AUDIO) ? make_uint2(0u, 0u)                 // silent leg-frame
                                       : sl == kSelGeneral ? leg_chunk<kSigned, 2>(lb, w, 0u, (int)a, acc)
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
                    if (q.meter) st16_stream(q.meter + bf2 * G + g, *reinterpret_cast<const uint4 *>(&r));
                }
            }
            __syncthreads();
        }
        if (valid)
            bpart[bfl * kPst + p] = mix_out_chunk<kSigned>(acc, enc_pk_make(q.out_law[b]),
                                                           q.mix + (size_t)bf * IGD_FRAME + p * 16,
                                                           q.enc + (size_t)bf * IGD_FRAME + p * 16, q.mix != nullptr,
                                                           q.enc != nullptr);
        __syncthreads();
        if (t < BFPC) {
            const long long bf2 = tile * BFPC + t;
            if (bf2 < q.total_bf && q.bmeter != nullptr) {
                int esum = 0, mpeak = 0;
#pragma unroll
                for (int i = 0; i < kChunks; i++) { esum += (int)bpart[t * kPst + i].x; mpeak = max(mpeak, (int)bpart[t * kPst + i].y); }
                int n_open = 0;
                for (int g = 0; g < G; g++) {
                    const uint32_t a = q.gain[(size_t)bf2 * G + g];
                    n_open += a != 0 && !(a & IGD_GAIN_NO_AUDIO);
                }
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

}  // namespace

// ============================================================ launchers
namespace {
template <int G, bool kSigned, int kWarps, bool kPkt = false, bool kTx = false>
cudaError_t launch_fused_w(const igd_launch_cfg &c, const FusedParams &q)
{
    const bool all_out = q.mix && q.enc && q.meter && q.bmeter;
    auto kern = all_out ? k_fused_w<G, kSigned, kWarps, kPkt, kTx, false> : k_fused_w<G, kSigned, kWarps, kPkt, kTx, true>;
    constexpr int kP = kPkt ? kChunks : kPst;
    const size_t smem = kLutBytes + (size_t)kWarps * slot_geom<G, kPkt>::kSlotBytes +
                        (size_t)kWarps * (kBfPerItem * G * kP + kBfPerItem * kP) * 8;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const long long items = (q.total_bf + kBfPerItem - 1) / kBfPerItem;
    long long grid = c.sm_count;
    if (grid > items) grid = items;      // small ticks: one item per SM before a second warp gets one
    kern<<<(int)grid, kWarps * 32, smem, c.stream>>>(q);
    return cudaGetLastError();
}

template <int G, bool kSigned, int kWarps, bool kPkt = false, bool kTx = false>
cudaError_t launch_fused_q(const igd_launch_cfg &c, const FusedParams &q)
{
    const bool all_out = q.mix && q.enc && q.meter && q.bmeter;
    auto kern = all_out ? k_fused_q<G, kSigned, kWarps, false, kPkt, kTx> : k_fused_q<G, kSigned, kWarps, true, kPkt, kTx>;
    const size_t smem = kLutBytes + (size_t)kWarps * kQBf * G * (kPkt ? IGD_PKT_MAX : IGD_FRAME) +
                        (size_t)kWarps * (kQLanes * (kQBf * G + 1) + kQLanes * (kQBf + 1)) * 8;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const long long items = (q.total_bf + kQBf - 1) / kQBf;
    long long grid = c.sm_count;
    if (grid > items) grid = items;
    kern<<<(int)grid, kWarps * 32, smem, c.stream>>>(q);
    return cudaGetLastError();
}

template <bool kSigned, int kWarps>
cudaError_t launch_fused_g(const igd_launch_cfg &c, const FusedParams &q)
{
    auto kern = k_fused_g<kSigned, kWarps>;
    const size_t smem = kLutBytes + kMlutBytes + (size_t)kWarps * 2 * kGSlotBytes + (size_t)kWarps * kGParts * 8;
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
    q.fields = nullptr;
    q.plan = nullptr; q.rtp12 = nullptr; q.tx_pkts = nullptr; q.tx_sizes = nullptr;
    q.total_bf = (long long)d.F * d.B;
    q.num_tiles = 0;
    q.B = d.B; q.G = d.G; q.flags = d.flags;
    const bool sc = (d.flags & IGD_F_SIGNED_CHAR) != 0;
    // the warp-autonomous kernels index bridge-frames in 32 bits; anything larger (2^28 bridge-frames =
    // 318 GB of traffic at G = 4) cannot be resident on one GPU anyway and takes the generic kernel
    // ... and reads a bridge-frame's G gains / a bridge's G laws as one 2G- / G-byte word
    if (d.flags & IGD_F_GENERIC_KERNEL) return sc ? launch_fused_anyg<32, true>(c, q) : launch_fused_anyg<32, false>(c, q);
    const bool fits32 = q.total_bf < (1ll << 28) &&          // 16-sample chunk indices (10 per bridge-frame) stay below 2^32
                        (reinterpret_cast<uintptr_t>(d.gain_q7) & (size_t)(2 * d.G - 1) & 7u) == 0 &&
                        (d.G != 4 || (reinterpret_cast<uintptr_t>(d.law) & 3u) == 0);
    if (fits32 && !(d.flags & IGD_F_KERNEL_W) && (reinterpret_cast<uintptr_t>(d.mix) & 15u) == 0 && (reinterpret_cast<uintptr_t>(d.enc) & 7u) == 0) {
        if (d.G == 4) return sc ? launch_fused_q<4, true, 24>(c, q) : launch_fused_q<4, false, 24>(c, q);
        if (d.G == 3) return sc ? launch_fused_q<3, true, 24>(c, q) : launch_fused_q<3, false, 24>(c, q);
        if (d.G == 2) return sc ? launch_fused_q<2, true, 24>(c, q) : launch_fused_q<2, false, 24>(c, q);
        if (d.G == 1) return sc ? launch_fused_q<1, true, 24>(c, q) : launch_fused_q<1, false, 24>(c, q);
    }
    if (fits32 && d.G == 4) return sc ? launch_fused_w<4, true, 24>(c, q) : launch_fused_w<4, false, 24>(c, q);
    if (fits32 && d.G == 2) return sc ? launch_fused_w<2, true, 24>(c, q) : launch_fused_w<2, false, 24>(c, q);
    if (fits32 && d.G == 1) return sc ? launch_fused_w<1, true, 24>(c, q) : launch_fused_w<1, false, 24>(c, q);
    if (q.total_bf < (1ll << 28) && d.G == 3) return sc ? launch_fused_w<3, true, 24>(c, q) : launch_fused_w<3, false, 24>(c, q);
    // eight legs: a lane walks all 8 legs of one sample column (k_fused_h)
    if (!(d.flags & IGD_F_KERNEL_W) && d.G == 8 && q.total_bf < (1ll << 26) &&
        ((reinterpret_cast<uintptr_t>(d.gain_q7) & 15u) | (reinterpret_cast<uintptr_t>(d.law) & 7u) |
         (reinterpret_cast<uintptr_t>(d.mix) & 7u) | (reinterpret_cast<uintptr_t>(d.enc) & 3u)) == 0)
        return sc ? launch_fused_h<8, true, 20>(c, q) : launch_fused_h<8, false, 20>(c, q);
    // any other leg count: the warp-autonomous group walk (needs 16-byte aligned codes, which the C ABI
    // checks, and 32-bit bridge-frame indices); the block-cooperative kernel is the last resort
    if (q.total_bf < (1ll << 28) && (long long)q.total_bf * d.G < (1ll << 32))
        return sc ? launch_fused_g<true, 24>(c, q) : launch_fused_g<false, 24>(c, q);
    return sc ? launch_fused_anyg<32, true>(c, q) : launch_fused_anyg<32, false>(c, q);
}

// packets in: [F][B*4][180] raw ED-137 packets + their parsed fields (payload_len) instead of codes
cudaError_t igd_k_fused_packets(const igd_launch_cfg &c, const igd_packets_desc &d)
{
    FusedParams q;
    q.codes = d.pkts; q.fields = d.fields; q.law = d.law; q.gain = d.gain_q7; q.out_law = d.out_law;
    q.mix = d.mix; q.enc = d.enc; q.meter = d.meter; q.bmeter = d.bmeter;
    q.plan = nullptr; q.rtp12 = nullptr; q.tx_pkts = nullptr; q.tx_sizes = nullptr;
    q.total_bf = (long long)d.F * d.B;
    q.num_tiles = 0;
    q.B = d.B; q.G = d.G; q.flags = d.flags;
    if (d.G != 4 || q.total_bf >= (1ll << 28) || (reinterpret_cast<uintptr_t>(d.gain_q7) & 7u) ||
        (reinterpret_cast<uintptr_t>(d.law) & 3u))
        return cudaErrorInvalidValue;
    const bool sc = (d.flags & IGD_F_SIGNED_CHAR) != 0;
    // quarter-lane kernel (23 warps: 8 x 720-byte bridge-frames per slot) unless asked for the older one
    if (!(d.flags & IGD_F_KERNEL_W) && (reinterpret_cast<uintptr_t>(d.mix) & 15u) == 0 && (reinterpret_cast<uintptr_t>(d.enc) & 7u) == 0)
        return sc ? launch_fused_q<4, true, 23, true>(c, q) : launch_fused_q<4, false, 23, true>(c, q);
    return sc ? launch_fused_w<4, true, 24, true>(c, q) : launch_fused_w<4, false, 24, true>(c, q);
}

#ifndef IGD_GW_FUSED_WARPS
#define IGD_GW_FUSED_WARPS 23
#endif
// gateway form: packets in (gains carry IGD_GAIN_NO_AUDIO, no field records needed) -> packets out
cudaError_t igd_k_fused_gateway(const igd_launch_cfg &c, const igd_packets_desc &d, const igd_tx_plan_rec *plan,
                                const uint8_t *tx_rtp12, uint8_t *tx_pkts, uint32_t *tx_sizes)
{
    FusedParams q;
    q.codes = d.pkts; q.fields = d.fields; q.law = d.law; q.gain = d.gain_q7; q.out_law = d.out_law;
    q.mix = d.mix; q.enc = d.enc; q.meter = d.meter; q.bmeter = d.bmeter;
    q.plan = plan; q.rtp12 = tx_rtp12; q.tx_pkts = tx_pkts; q.tx_sizes = tx_sizes;
    q.total_bf = (long long)d.F * d.B;
    q.num_tiles = 0;
    q.B = d.B; q.G = d.G; q.flags = d.flags;
    if (d.G != 4 || q.total_bf >= (1ll << 28) || (reinterpret_cast<uintptr_t>(d.gain_q7) & 7u) ||
        (reinterpret_cast<uintptr_t>(d.law) & 3u) || (reinterpret_cast<uintptr_t>(tx_pkts) & 3u) ||
        (reinterpret_cast<uintptr_t>(tx_rtp12) & 3u))
        return cudaErrorInvalidValue;
    const bool sc = (d.flags & IGD_F_SIGNED_CHAR) != 0;
    if (!(d.flags & IGD_F_KERNEL_W) && (reinterpret_cast<uintptr_t>(d.mix) & 15u) == 0 && (reinterpret_cast<uintptr_t>(d.enc) & 7u) == 0)
        return sc ? launch_fused_q<4, true, IGD_GW_FUSED_WARPS, true, true>(c, q) : launch_fused_q<4, false, IGD_GW_FUSED_WARPS, true, true>(c, q);
    return sc ? launch_fused_w<4, true, 24, true, true>(c, q) : launch_fused_w<4, false, 24, true, true>(c, q);
}
