// ubench.cu -- integer-pipe issue rates on sm_100a for the ops the fused voice-path
// kernel is made of (it is issue-bound before it is HBM-bound).
// Prints lane-ops per clock per SM for each op.  nvcc -arch=sm_100a -O3 ubench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 4096
#define NCH 8

#define BENCH_KERNEL(name, INIT, BODY)                                             \
__global__ void __launch_bounds__(1024) k_##name(unsigned *out, unsigned seed, long long *cyc) { \
    unsigned a[NCH]; unsigned b = seed | 1u, c = seed * 3u + 7u;                   \
    __shared__ unsigned sh[1024];                                                  \
    sh[threadIdx.x] = (threadIdx.x * 4 + seed) & 0xFFCu;                                             \
    for (int i = 0; i < NCH; i++) a[i] = seed + threadIdx.x * 17u + i * 101u;      \
    INIT;                                                                          \
    __syncthreads();                                                               \
    long long t0 = clock64();                                                      \
    for (int it = 0; it < ITERS; it++) {                                           \
        _Pragma("unroll") for (int i = 0; i < NCH; i++) { BODY; }                  \
    }                                                                              \
    long long t1 = clock64();                                                      \
    unsigned s = 0; for (int i = 0; i < NCH; i++) s += a[i];                       \
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + b + c;                        \
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;                               \
}

BENCH_KERNEL(imad,    , asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c)))
BENCH_KERNEL(imadhi,  , asm volatile("mad.hi.s32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c)))
BENCH_KERNEL(imadwide, unsigned long long w[NCH]; for (int i = 0; i < NCH; i++) w[i] = a[i];,
             asm volatile("mad.wide.s32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"((unsigned)(w[i] >> 7)), "r"(c)); a[i] = (unsigned)w[i])
BENCH_KERNEL(vimnmx,  , asm volatile("max.s32 %0, %0, %1;" : "+r"(a[i]) : "r"(b)); asm volatile("min.s32 %0, %0, %1;" : "+r"(a[i]) : "r"(c)))
BENCH_KERNEL(vimnmx3, , a[i] = (unsigned)max(max((int)a[i], (int)b + i), (int)c + it))
BENCH_KERNEL(prmt,    , asm volatile("prmt.b32 %0, %0, %1, 0x7614;" : "+r"(a[i]) : "r"(b)))
BENCH_KERNEL(lop3,    , asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b), "r"(c)))
BENCH_KERNEL(iadd3,   , asm volatile("add.s32 %0, %0, %1;" : "+r"(a[i]) : "r"(b)))
BENCH_KERNEL(shf,     , asm volatile("shr.s32 %0, %0, 1;" : "+r"(a[i])); asm volatile("add.s32 %0, %0, %1;" : "+r"(a[i]) : "r"(b)))
BENCH_KERNEL(i2ip,    , asm volatile("cvt.pack.sat.s16.s32 %0, %0, %1;" : "+r"(a[i]) : "r"(b)))
BENCH_KERNEL(i2isat,  , { short h; asm volatile("cvt.sat.s16.s32 %0, %1;" : "=h"(h) : "r"(a[i])); a[i] = (unsigned)(int)h + b; })
BENCH_KERNEL(idp4a,   , asm volatile("dp4a.u32.u32 %0, %1, %2, %0;" : "+r"(a[i]) : "r"(b), "r"(c)))
BENCH_KERNEL(idp2a,   , asm volatile("dp2a.lo.s32.u32 %0, %1, %2, %0;" : "+r"(a[i]) : "r"(b), "r"(c)))
BENCH_KERNEL(ffma,    , asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c)))
BENCH_KERNEL(iabs,    , asm volatile("abs.s32 %0, %0;" : "+r"(a[i])); asm volatile("sub.s32 %0, %0, %1;" : "+r"(a[i]) : "r"(b)))
BENCH_KERNEL(flo,     , asm volatile("clz.b32 %0, %0;" : "+r"(a[i])); asm volatile("add.s32 %0, %0, %1;" : "+r"(a[i]) : "r"(b)))
BENCH_KERNEL(lds,     , a[i] = ((volatile unsigned *)sh)[(a[i] >> 2) & 1023u])
BENCH_KERNEL(lds_prmt_imad, , { unsigned ad = __byte_perm(a[i], b, 0x7614) & 0xFFCu; unsigned v = ((volatile unsigned *)sh)[ad >> 2]; a[i] = v * v + a[i]; })
BENCH_KERNEL(vimnmx16x2, , asm volatile("max.s16x2 %0, %0, %1;" : "+r"(a[i]) : "r"(b)))
BENCH_KERNEL(mix_imad_alu, , asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c)); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b), "r"(c)))

template <class K>
void run(const char *name, K kern, int ops_per_body, int threads)
{
    int dev = 0, sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    unsigned *out; long long *cyc;
    cudaMalloc(&out, sizeof(unsigned) * sms * threads);
    cudaMalloc(&cyc, sizeof(long long) * sms);
    kern<<<sms, threads>>>(out, 12345u, cyc);
    kern<<<sms, threads>>>(out, 12345u, cyc);
    cudaDeviceSynchronize();
    long long h[512];
    cudaMemcpy(h, cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < sms; i++) avg += (double)h[i]; avg /= sms;
    double ops = (double)threads * ITERS * NCH * ops_per_body;
    printf("%-14s threads/SM=%4d  lane-ops/clk/SM=%7.2f  (warp-instr/clk/SM=%5.2f)  err=%s\n", name, threads,
           ops / avg, ops / avg / 32.0, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out); cudaFree(cyc);
}

int main()
{
    for (int threads : {256, 1024}) {
        run("imad", k_imad, 1, threads);
        run("imad.hi", k_imadhi, 1, threads);
        run("imad.wide", k_imadwide, 1, threads);
        run("vimnmx(x2)", k_vimnmx, 2, threads);
        run("vimnmx3", k_vimnmx3, 1, threads);
        run("prmt", k_prmt, 1, threads);
        run("lop3", k_lop3, 1, threads);
        run("iadd", k_iadd3, 1, threads);
        run("shf+add", k_shf, 2, threads);
        run("i2ip.sat", k_i2ip, 1, threads);
        run("i2i.sat+add", k_i2isat, 2, threads);
        run("idp4a", k_idp4a, 1, threads);
        run("idp2a", k_idp2a, 1, threads);
        run("ffma", k_ffma, 1, threads);
        run("iabs+sub", k_iabs, 2, threads);
        run("flo+add", k_flo, 2, threads);
        run("lds", k_lds, 1, threads);
        run("lds+prmt+imad", k_lds_prmt_imad, 3, threads);
        run("vimnmx.s16x2", k_vimnmx16x2, 1, threads);
        run("imad+lop3", k_mix_imad_alu, 2, threads);
    }
    return 0;
}
