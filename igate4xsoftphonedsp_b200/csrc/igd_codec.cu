// igd_codec.cu -- stand-alone G.711 decode / encode, frame meter, byte-mean, percent scale, mix, event summary
// (hand-written sm_100a kernels of the iGate4x voice path; design notes in igd_fused.cu and DESIGN.md)
#include "igd_device.cuh"

namespace {

// ============================================================ stand-alone G.711
// law of 16-sample chunk i when frames are laid out [frame][channel]: law_ch[(i / 10) % nch].  The chunk
// index fits 32 bits for anything below 64 G samples; 64-bit division would cost ~80 instructions per chunk.
__device__ __forceinline__ uint32_t chunk_law(const uint8_t *law_ch, int law, size_t i, size_t nchunk, size_t nch)
{
    if (!law_ch) return (uint32_t)law;
    if (nchunk <= 0xFFFFFFFFull) return law_ch[((uint32_t)i / (uint32_t)kChunks) % (uint32_t)nch];
    return law_ch[(i / kChunks) % nch];
}

// codes -> PCM.  One thread per 16 codes, shared-memory table, one 256-bit store.
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
        const uint32_t lw = chunk_law(law_ch, law, i, nchunk, nch);
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
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + stride < nchunk; i += 2 * stride) {       // two chunks per trip: two 32-byte loads in flight
        uint32_t pa[8], pb[8];
        ld32_stream(pcm + i * 16, pa);
        ld32_stream(pcm + (i + stride) * 16, pb);
        st16_stream(codes + i * 16, encode16_packed(pa, enc_pk_make((int)chunk_law(law_ch, law, i, nchunk, nch))));
        st16_stream(codes + (i + stride) * 16,
                    encode16_packed(pb, enc_pk_make((int)chunk_law(law_ch, law, i + stride, nchunk, nch))));
    }
    if (i < nchunk) {
        uint32_t pk[8];
        ld32_stream(pcm + i * 16, pk);
        st16_stream(codes + i * 16, encode16_packed(pk, enc_pk_make((int)chunk_law(law_ch, law, i, nchunk, nch))));
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
            if (a == 0 || (a & (int)IGD_GAIN_NO_AUDIO)) continue;
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
                if (g[u] & IGD_GAIN_NO_AUDIO) g[u] = 0u;      // no audio frame on this tick: no level either
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

}  // namespace

// ============================================================ launchers
cudaError_t igd_k_g711_decode(const igd_launch_cfg &c, const uint8_t *codes, const uint8_t *law_ch,
                              int law, int16_t *pcm, size_t n, size_t nch)
{
    // per launch, like the fused launchers: the attribute is per device, a process may hold contexts on several
    cudaError_t e = cudaFuncSetAttribute(k_g711_decode, cudaFuncAttributeMaxDynamicSharedMemorySize, kLutBytes);
    if (e != cudaSuccess) return e;
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

cudaError_t igd_k_event_summary(const igd_launch_cfg &c, const igd_meter_rec *meter,
                                const uint16_t *gain, size_t F, size_t C, igd_summary_rec *out,
                                igd_summary_db *db)
{
    const int blocks = (int)((C + 31) / 32);
    k_event_summary<<<blocks, 32 * kSumSlices, 0, c.stream>>>(meter, gain, (long long)F, (long long)C, out, db);
    return cudaGetLastError();
}
