// ubench_fp32.cu -- FP32 pipe throughput on sm_100a: scalar FFMA / FADD / FMUL vs the packed f32x2 forms,
// per SM and clock, at the frame kernel's occupancy (16 warps/SM) and at 32 warps/SM.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/ubench_fp32 tools/ubench_fp32.cu
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096
#define CHAINS 8

template <int MODE>
__global__ void k(float *out, float a, float b, long long *cyc)
{
    float x[2 * CHAINS];
    for (int i = 0; i < 2 * CHAINS; i++) x[i] = a + i + threadIdx.x;
    unsigned long long x2[CHAINS];
    for (int i = 0; i < CHAINS; i++) asm("mov.b64 %0, {%1, %2};" : "=l"(x2[i]) : "f"(x[2 * i]), "f"(x[2 * i + 1]));
    unsigned long long ab, bb;
    asm("mov.b64 %0, {%1, %1};" : "=l"(ab) : "f"(a));
    asm("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(b));
    int ii[CHAINS];
    for (int i = 0; i < CHAINS; i++) ii[i] = threadIdx.x + i;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int c = 0; c < CHAINS; c++) {
            if (MODE == 0) { x[2 * c] = fmaf(x[2 * c], a, b); x[2 * c + 1] = fmaf(x[2 * c + 1], a, b); }          // 2 FFMA
            if (MODE == 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x2[c]) : "l"(ab), "l"(bb));      // 1 FFMA2
            if (MODE == 2) { x[2 * c] = x[2 * c] + a; x[2 * c + 1] = x[2 * c + 1] + b; }                          // 2 FADD
            if (MODE == 3) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(x2[c]) : "l"(ab));                   // 1 FADD2
            if (MODE == 4) { x[2 * c] = x[2 * c] * a; x[2 * c + 1] = x[2 * c + 1] * b; }                          // 2 FMUL
            if (MODE == 5) { x[2 * c] = fmaf(x[2 * c], a, b); ii[c] = (ii[c] ^ it) + c; }                          // FFMA + 2 ALU
            if (MODE == 6) { x[2 * c] = fmaf(x[2 * c], x[2 * c + 1], b); x[2 * c + 1] = fmaf(x[2 * c + 1], a, x[2*c]); }   // 2 FFMA 3 distinct regs
        }
    }
    long long t1 = clock64();
    float s = 0;
    for (int i = 0; i < 2 * CHAINS; i++) s += x[i];
    for (int i = 0; i < CHAINS; i++) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(x2[i])); s += lo + hi + ii[i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char *name, int threads, double flop_per_inner)
{
    float *out; long long *cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    k<MODE><<<148, threads>>>(out, 1.0001f, 0.5f, cyc);
    k<MODE><<<148, threads>>>(out, 1.0001f, 0.5f, cyc);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    double c = 0; for (int i = 0; i < 148; i++) c += h[i]; c /= 148;
    // lane-ops per clock per SM
    double ops = (double)ITERS * CHAINS * flop_per_inner * threads / c;
    printf("%-34s threads/SM %4d  cycles %9.0f  lane-ops/clk/SM %7.1f\n", name, threads, c, ops);
    cudaFree(out); cudaFree(cyc);
}

int main()
{
    for (int th : { 512, 1024 }) {
        run<0>("FFMA (reg,imm/const,const) x2", th, 2);
        run<6>("FFMA 3 regs x2", th, 2);
        run<1>("fma.rn.f32x2 (2 lanes-ops each)", th, 2);
        run<2>("FADD x2", th, 2);
        run<3>("add.rn.f32x2", th, 2);
        run<4>("FMUL x2", th, 2);
        run<5>("FFMA + XOR + IADD (3 instr)", th, 3);
    }
    return 0;
}
