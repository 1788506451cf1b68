// Pipe-throughput microbenchmark for sm_100a integer instructions (development aid).
// Each kernel runs N_ITER iterations of UNROLL independent chains per thread; reports
// warp-instructions per clock per SM sub-partition (SMSP).  Build: nvcc -arch=sm_100a -O3 pipes.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

#define N_ITER 4096
#define CHAINS 8

enum Op { IMAD, IMADWIDE, IMADHI, LOP3, PRMT, IADD3, SHF, IDP4A, ISETP_SEL, VIADD, IMAD_LOP3, WIDE_LOP3, IDP_LOP3, IMAD_PRMT,
          PHILOX_LIKE, IDP_IMAD, LDS16, LDS16_LOP3, NOPS };
const char* names[] = {"IMAD", "IMAD.WIDE", "IMAD.HI", "LOP3", "PRMT", "IADD3", "SHF", "IDP.4A", "ISETP+SEL", "VIADD",
                       "IMAD+LOP3 (1:1)", "IMAD.WIDE+LOP3 (1:1)", "IDP.4A+LOP3 (1:1)", "IMAD+PRMT (1:1)", "2xWIDE+2xLOP3 (philox round)",
                       "IDP.4A+IMAD (1:1)", "LDS.U16", "LDS.U16+LOP3 (1:1)"};
const int instr_per_step[] = {1, 1, 1, 1, 1, 1, 1, 1, 2, 1, 2, 2, 2, 2, 4, 2, 1, 2};

template <int OP>
__global__ void k(uint32_t* out, uint32_t seed) {
    __shared__ uint16_t tab[64];
    if (threadIdx.x < 64) tab[threadIdx.x] = threadIdx.x * 3;
    __syncthreads();
    uint32_t a[CHAINS], b[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) { a[i] = seed + threadIdx.x * 7 + i; b[i] = seed * 3 + i; }
    const uint32_t m = seed | 0x10001u, c = seed ^ 0x12345u;
#pragma unroll 1
    for (int it = 0; it < N_ITER; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) {
            if (OP == IMAD) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(m), "r"(c));
            if (OP == IMADWIDE) { uint64_t w; asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w) : "r"(a[i]), "r"(m)); a[i] = (uint32_t)w; b[i] ^= (uint32_t)(w >> 32); }
            if (OP == IMADHI) asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(m));
            if (OP == LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(m), "r"(c));
            if (OP == PRMT) asm volatile("prmt.b32 %0, %0, %1, 0x2143;" : "+r"(a[i]) : "r"(m));
            if (OP == IADD3) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(m));
            if (OP == SHF) asm volatile("shf.l.wrap.b32 %0, %0, %1, 5;" : "+r"(a[i]) : "r"(m));
            if (OP == IDP4A) asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(m), "r"(c));
            if (OP == ISETP_SEL) asm volatile("{.reg .pred p; setp.lt.u32 p, %0, %1; @p add.u32 %0, %0, 0x100;}" : "+r"(a[i]) : "r"(m));
            if (OP == VIADD) asm volatile("add.u32 %0, %0, 0x00010001;" : "+r"(a[i]));
            if (OP == IMAD_LOP3) { asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(m), "r"(c)); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(b[i]) : "r"(m), "r"(c)); }
            if (OP == WIDE_LOP3) { uint64_t w; asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w) : "r"(a[i]), "r"(m)); a[i] = (uint32_t)w ^ (uint32_t)(w >> 32); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(b[i]) : "r"(m), "r"(c)); }
            if (OP == IDP_LOP3) { asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(m), "r"(c)); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(b[i]) : "r"(m), "r"(c)); }
            if (OP == IMAD_PRMT) { asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(m), "r"(c)); asm volatile("prmt.b32 %0, %0, %1, 0x2143;" : "+r"(b[i]) : "r"(m)); }
            if (OP == PHILOX_LIKE) {
                uint64_t w0, w1;
                asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w0) : "r"(a[i]), "r"(0xD2511F53u));
                asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w1) : "r"(b[i]), "r"(0xCD9E8D57u));
                uint32_t x, y;
                asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(x) : "r"((uint32_t)(w1 >> 32)), "r"((uint32_t)w0), "r"(c));
                asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(y) : "r"((uint32_t)(w0 >> 32)), "r"((uint32_t)w1), "r"(m));
                a[i] = x; b[i] = y;
            }
            if (OP == IDP_IMAD) { asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(m), "r"(c)); asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(b[i]) : "r"(m), "r"(c)); }
            if (OP == LDS16) { a[i] = tab[a[i] & 63]; }
            if (OP == LDS16_LOP3) { a[i] = tab[a[i] & 63]; asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(b[i]) : "r"(m), "r"(c)); }
        }
    }
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) r ^= a[i] ^ b[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int OP>
void run(uint32_t* d_out, int sms, double clk_ghz) {
    const int threads = 256, blocks = sms * 4;  // 32 warps per SM = 8 per SMSP
    k<OP><<<blocks, threads>>>(d_out, 12345u);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<OP><<<blocks, threads>>>(d_out, 12345u);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double warp_instr = (double)blocks * threads / 32 * N_ITER * CHAINS * instr_per_step[OP];
    const double per_smsp_clk = warp_instr / (ms * 1e-3 * clk_ghz * 1e9) / (sms * 4);
    printf("%-34s %8.3f ms  %.3f warp-instr/clk/SMSP (assuming %.3f GHz)\n", names[OP], ms, per_smsp_clk, clk_ghz);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const double ghz = clk_khz / 1e6;
    printf("%s, %d SMs, clock attr %.3f GHz\n", p.name, p.multiProcessorCount, ghz);
    uint32_t* d;
    cudaMalloc(&d, 1 << 26);
    const int s = p.multiProcessorCount;
    run<IMAD>(d, s, ghz); run<IMADWIDE>(d, s, ghz); run<IMADHI>(d, s, ghz); run<LOP3>(d, s, ghz); run<PRMT>(d, s, ghz);
    run<IADD3>(d, s, ghz); run<SHF>(d, s, ghz); run<IDP4A>(d, s, ghz); run<ISETP_SEL>(d, s, ghz); run<VIADD>(d, s, ghz);
    run<IMAD_LOP3>(d, s, ghz); run<WIDE_LOP3>(d, s, ghz); run<IDP_LOP3>(d, s, ghz); run<IMAD_PRMT>(d, s, ghz);
    run<PHILOX_LIKE>(d, s, ghz); run<IDP_IMAD>(d, s, ghz); run<LDS16>(d, s, ghz); run<LDS16_LOP3>(d, s, ghz);
    return 0;
}
