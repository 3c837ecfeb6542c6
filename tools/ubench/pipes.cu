// Throughput of IMAD / IDP.4A / PRMT / SHF / LOP3 / F2I on sm_100a (per SM per clock).
#include <cstdio>
#include <cuda_runtime.h>
#define N_ITER 4096
template <int OP>
__global__ void k(int* out, int a, int b) {
    int x0 = threadIdx.x, x1 = a, x2 = b, x3 = a ^ b, x4 = a + 7, x5 = b + 3, x6 = a * 3, x7 = b * 5;
    float f0 = a, f1 = b, f2 = a + 1, f3 = b + 1;
#pragma unroll 1
    for (int i = 0; i < N_ITER; ++i) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (OP == 0) { x0 = x0 * a + b; x1 = x1 * a + b; x2 = x2 * a + b; x3 = x3 * a + b; x4 = x4 * a + b; x5 = x5 * a + b; x6 = x6 * a + b; x7 = x7 * a + b; }
            if (OP == 1) { x0 = __dp4a(x0, a, b); x1 = __dp4a(x1, a, b); x2 = __dp4a(x2, a, b); x3 = __dp4a(x3, a, b); x4 = __dp4a(x4, a, b); x5 = __dp4a(x5, a, b); x6 = __dp4a(x6, a, b); x7 = __dp4a(x7, a, b); }
            if (OP == 2) { x0 = __byte_perm(x0, a, b); x1 = __byte_perm(x1, a, b); x2 = __byte_perm(x2, a, b); x3 = __byte_perm(x3, a, b); x4 = __byte_perm(x4, a, b); x5 = __byte_perm(x5, a, b); x6 = __byte_perm(x6, a, b); x7 = __byte_perm(x7, a, b); }
            if (OP == 3) { x0 = __funnelshift_r(x0, a, b); x1 = __funnelshift_r(x1, a, b); x2 = __funnelshift_r(x2, a, b); x3 = __funnelshift_r(x3, a, b); x4 = __funnelshift_r(x4, a, b); x5 = __funnelshift_r(x5, a, b); x6 = __funnelshift_r(x6, a, b); x7 = __funnelshift_r(x7, a, b); }
            if (OP == 4) { x0 = (x0 & a) ^ b; x1 = (x1 & a) ^ b; x2 = (x2 & a) ^ b; x3 = (x3 & a) ^ b; x4 = (x4 & a) ^ b; x5 = (x5 & a) ^ b; x6 = (x6 & a) ^ b; x7 = (x7 & a) ^ b; }
            if (OP == 5) { x0 += __float2int_rn(f0 + x0); x1 += __float2int_rn(f1 + x1); x2 += __float2int_rn(f2 + x2); x3 += __float2int_rn(f3 + x3); }
            if (OP == 6) { x0 = __dp2a_lo(x0, a, b); x1 = __dp2a_hi(x1, a, b); x2 = __dp2a_lo(x2, a, b); x3 = __dp2a_hi(x3, a, b); x4 = __dp2a_lo(x4, a, b); x5 = __dp2a_hi(x5, a, b); x6 = __dp2a_lo(x6, a, b); x7 = __dp2a_hi(x7, a, b); }
            if (OP == 7) { f0 = f0 * f1 + f2; f1 = f1 * f2 + f3; f2 = f2 * f3 + f0; f3 = f3 * f0 + f1; x0 = x0 * a + b; x1 = x1 * a + b; x2 = x2 * a + b; x3 = x3 * a + b; }
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7 + (int)(f0 + f1 + f2 + f3);
}
template <int OP>
void run(const char* name, int ops_per_unroll) {
    int* out; cudaMalloc(&out, 148 * 8 * 1024 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<OP><<<148 * 8, 256>>>(out, 3, 5);
    cudaEventRecord(e0);
    k<OP><<<148 * 8, 256>>>(out, 3, 5);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double ops = 148.0 * 8 * 256 * N_ITER * 4.0 * ops_per_unroll;
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("%-12s %.3f ms  %.1f thread-ops/clk/SM (at %d MHz nominal)\n", name, ms, ops / (ms * 1e-3) / 148.0 / (clk * 1e3), clk / 1000);
    cudaFree(out);
}
int main() {
    run<0>("IMAD", 8); run<1>("IDP4A", 8); run<2>("PRMT", 8); run<3>("SHF", 8); run<4>("LOP3", 8); run<5>("F2I+FADD+IADD", 4); run<6>("IDP2A", 8); run<7>("FFMA+IMAD", 8);
    return 0;
}
