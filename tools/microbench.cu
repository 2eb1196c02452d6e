// microbench.cu — issue-rate probes for the integer pipes and shared-memory gathers on B200.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench tools/microbench.cu
// Prints warp-instructions per cycle per SM for each probe (4 SMSPs => 4.0 is the issue limit).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

constexpr int ITERS = 2048;
constexpr int CHAINS = 8;

template <int OP>
__global__ void __launch_bounds__(1024, 1) probe(uint32_t* out, long long* cycles, uint32_t seed) {
    uint32_t a[CHAINS];
    uint32_t y = seed ^ threadIdx.x, z = seed * 3 + 1;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) a[c] = threadIdx.x * 7 + c + seed;
    long long t0 = clock64();
    for (int i = 0; i < ITERS; ++i) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) {
            if (OP == 0) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[c]) : "r"(y), "r"(z));
            if (OP == 1) asm volatile("shf.l.wrap.b32 %0, %0, %1, 5;" : "+r"(a[c]) : "r"(y));
            if (OP == 2) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[c]) : "r"(y));
            if (OP == 3) asm volatile("prmt.b32 %0, %0, %1, 0x2741;" : "+r"(a[c]) : "r"(y));
            if (OP == 4) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[c]) : "r"(y), "r"(z));
            if (OP == 5) asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(a[c]) : "r"(y));
            if (OP == 6) asm volatile("popc.b32 %0, %0;" : "+r"(a[c]));
            if (OP == 7) {  // alternate LOP3 / IMAD: do both pipes dual-issue?
                if (c & 1) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[c]) : "r"(y), "r"(z));
                else asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[c]) : "r"(y), "r"(z));
            }
            if (OP == 8) asm volatile("shl.b32 %0, %0, 3;" : "+r"(a[c]));            // may become IMAD.SHL
            if (OP == 9) asm volatile("shr.u32 %0, %0, 3;" : "+r"(a[c]));
            if (OP == 11) asm volatile("{.reg .pred p; setp.lt.u32 p, %0, %1; selp.u32 %0, %1, %2, p;}" : "+r"(a[c]) : "r"(y), "r"(z));
            if (OP == 12) {  // LOP3 + FFMA mix (fp pipe is separate?)
                if (c & 1) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[c]) : "r"(y), "r"(z));
                else { float f = __uint_as_float(a[c]); asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f) : "f"(1.0001f), "f"(0.5f)); a[c] = __float_as_uint(f); }
            }
            if (OP == 13) asm volatile("vabsdiff4.u32.u32.u32.add %0, %0, %1, %2;" : "+r"(a[c]) : "r"(y), "r"(z));
            if (OP == 14) asm volatile("add.u32 %0, %0, 0x77777777;" : "+r"(a[c]));                       // VIADD?
            if (OP == 15) { if (c & 1) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[c]) : "r"(y), "r"(z));
                            else asm volatile("add.u32 %0, %0, 0x77777777;" : "+r"(a[c])); }             // LOP3 + VIADD
            if (OP == 16) { if (c & 1) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[c]) : "r"(y), "r"(z));
                            else asm volatile("add.u32 %0, %0, %1;" : "+r"(a[c]) : "r"(y)); }            // LOP3 + IADD3
            if (OP == 17) { if (c & 1) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[c]) : "r"(y), "r"(z));
                            else asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(a[c]) : "r"(y)); }         // LOP3 + IMAD.HI
            if (OP == 18) asm volatile("min.u32 %0, %0, %1;" : "+r"(a[c]) : "r"(y));                       // VIMNMX
            if (OP == 19) asm volatile("{.reg .pred p; setp.lt.u32 p, %0, %1; @p add.u32 %0, %0, %2;}" : "+r"(a[c]) : "r"(y), "r"(z));
            if (OP == 20) asm volatile("shf.r.wrap.b32 %0, %0, %1, 7;" : "+r"(a[c]) : "r"(y));
            if (OP == 21) { unsigned long long w; asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w) : "r"(a[c]), "r"(y)); a[c] = (uint32_t)w ^ (uint32_t)(w >> 32); }
            if (OP == 22) { if ((c & 3) == 0) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[c]) : "r"(y), "r"(z));
                            else if ((c & 3) == 1) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[c]) : "r"(y));
                            else asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[c]) : "r"(y), "r"(z)); } // 2 LOP3 : 1 IMAD : 1 IADD
            if (OP == 23) asm volatile("mad.lo.u32 %0, %0, 1, %1;" : "+r"(a[c]) : "r"(y));                 // IMAD.IADD
        }
    }
    long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s ^= a[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// random gathers from a big shared-memory table (u32 entries), 4 independent lookups in flight
template <int WORDS>
__global__ void __launch_bounds__(1024, 1) lds_probe(uint32_t* out, long long* cycles, uint32_t seed) {
    extern __shared__ uint32_t tab[];
    for (int i = threadIdx.x; i < WORDS; i += blockDim.x) tab[i] = i * 2654435761u;
    __syncthreads();
    uint32_t x[4];
    for (int c = 0; c < 4; ++c) x[c] = (threadIdx.x * 2654435761u + c * 40503u + seed);
    long long t0 = clock64();
    for (int i = 0; i < ITERS; ++i) {
#pragma unroll
        for (int c = 0; c < 4; ++c) x[c] = tab[(x[c] >> 8) % WORDS] + x[c] * 5u;
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = x[0] ^ x[1] ^ x[2] ^ x[3];
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP>
int run(const char* name, uint32_t* d_out, long long* d_cyc) {
    probe<OP><<<148, 1024>>>(d_out, d_cyc, 12345u);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    probe<OP><<<148, 1024>>>(d_out, d_cyc, 777u);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long cyc[148];
    CK(cudaMemcpy(cyc, d_cyc, sizeof cyc, cudaMemcpyDeviceToHost));
    double mean = 0; for (int i = 0; i < 148; ++i) mean += cyc[i]; mean /= 148;
    double winstr = (double)ITERS * CHAINS * 32;  // warp instructions per SM (32 warps)
    printf("%-28s %8.3f warp-instr/cycle/SM   (%.0f cycles, %.3f ms, ~%.0f MHz)\n", name, winstr / mean, mean, ms, mean / ms / 1e3);
    return 0;
}

int main() {
    uint32_t* d_out; long long* d_cyc;
    CK(cudaMalloc(&d_out, 148 * 1024 * 4));
    CK(cudaMalloc(&d_cyc, 148 * 8));
    run<0>("LOP3", d_out, d_cyc);
    run<1>("SHF (funnel)", d_out, d_cyc);
    run<2>("IADD", d_out, d_cyc);
    run<3>("PRMT", d_out, d_cyc);
    run<4>("IMAD", d_out, d_cyc);
    run<5>("IMAD.HI", d_out, d_cyc);
    run<6>("POPC", d_out, d_cyc);
    run<7>("LOP3+IMAD alternating", d_out, d_cyc);
    run<8>("SHL imm", d_out, d_cyc);
    run<9>("SHR imm", d_out, d_cyc);
    run<11>("ISETP+SEL pair (x2 instr)", d_out, d_cyc);
    run<12>("LOP3+FFMA alternating", d_out, d_cyc);
    run<13>("vabsdiff4 (emulated?)", d_out, d_cyc);
    run<14>("add imm (VIADD?)", d_out, d_cyc);
    run<15>("LOP3 + add-imm alternating", d_out, d_cyc);
    run<16>("LOP3 + add-reg alternating", d_out, d_cyc);
    run<17>("LOP3 + IMAD.HI alternating", d_out, d_cyc);
    run<18>("min.u32 (VIMNMX)", d_out, d_cyc);
    run<19>("setp + predicated add (x2)", d_out, d_cyc);
    run<20>("SHF.R funnel", d_out, d_cyc);
    run<21>("IMAD.WIDE + xor (x2)", d_out, d_cyc);
    run<22>("2 LOP3 : 1 IMAD : 1 IADD mix", d_out, d_cyc);
    run<23>("mad.lo x,1,y (IMAD.IADD)", d_out, d_cyc);
    {
        constexpr int WORDS = 57344;
        CK(cudaFuncSetAttribute(lds_probe<WORDS>, cudaFuncAttributeMaxDynamicSharedMemorySize, WORDS * 4));
        lds_probe<WORDS><<<148, 1024, WORDS * 4>>>(d_out, d_cyc, 1u);
        CK(cudaDeviceSynchronize());
        lds_probe<WORDS><<<148, 1024, WORDS * 4>>>(d_out, d_cyc, 2u);
        CK(cudaDeviceSynchronize());
        long long cyc[148];
        CK(cudaMemcpy(cyc, d_cyc, sizeof cyc, cudaMemcpyDeviceToHost));
        double mean = 0; for (int i = 0; i < 148; ++i) mean += cyc[i]; mean /= 148;
        printf("%-28s %8.3f cycles per random LDS.32 warp-instr per SM (224 KiB table, 32 warps)\n",
               "LDS random gather", mean / ((double)ITERS * 4 * 32));
    }
    return 0;
}
