// Issue-rate probe for the integer / float pipes of sm_100a: thread-instructions per clock per SM for IMAD, IDP.2A, IDP.4A, FFMA, SHF and the
// two candidate inner loops of the LK mismatch (6 IMAD + SHF  vs  2 IDP.2A + 2 IMAD + SHF).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>
template <int OP>
__global__ void __launch_bounds__(256) k(int* out, int a0, int b0, int iters) {
    int a[8]; int b = b0 + threadIdx.x, c = a0 ^ threadIdx.x;
    float f[8]; float fb = (float)b0 * 1e-3f, fc = 1.0001f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = a0 + i * 7 + threadIdx.x; f[i] = (float)(a0 + i); }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (OP == 0) a[i] = a[i] * b + c;                                   // IMAD
                if (OP == 1) a[i] = __dp2a_lo((unsigned)b, (unsigned)a[i], (unsigned)c); // IDP.2A
                if (OP == 2) a[i] = __dp4a((unsigned)b, (unsigned)a[i], (unsigned)c);   // IDP.4A
                if (OP == 3) f[i] = fmaf(f[i], fb, fc);                              // FFMA
                if (OP == 4) a[i] = __funnelshift_r(a[i], b, 9) ;                    // SHF
                if (OP == 5) a[i] = (a[i] + b) ^ c;                                  // IADD3/LOP3
                if (OP == 6) {   // 6 IMAD + SHF + (dependent) : current mismatch pixel
                    int x = a[i] * b + c; x = (x >> 3) * c + x; x = x * b + a[i]; x = (x & 255) * c + x; int d = x >> 9; a[i] = d * b + a[i]; c = d * c + c;
                }
                if (OP == 7) {   // 2 IDP.2A + SHF + 2 IMAD
                    int x = __dp2a_lo((unsigned)b, (unsigned)a[i], (unsigned)c); x = __dp2a_hi((unsigned)c, (unsigned)a[i], (unsigned)x); int d = x >> 9; a[i] = d * b + a[i]; c = d * c + c;
                }
            }
        }
    }
    int s = 0; float fs = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { s += a[i]; fs += f[i]; }
    if (s == 0x12345678 || fs == 1.2345f) out[0] = s + c;
}
template <int OP>
void run(const char* name, int ops_per_inner, int* d, int sms) {
    const int iters = 2048; dim3 g(sms * 8), b(256);
    k<OP><<<g, b>>>(d, 3, 5, 16);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0); k<OP><<<g, b>>>(d, 3, 5, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    double ops = (double)g.x * 256 * iters * 32.0 * ops_per_inner;
    printf("%-28s %8.3f ms  %7.1f thread-instr/clk/SM (at %d MHz nominal)\n", name, ms, ops / (ms * 1e-3) / (khz * 1e3) / sms, khz / 1000);
}
int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int* d; cudaMalloc(&d, 4);
    run<0>("IMAD", 1, d, sms); run<1>("IDP.2A", 1, d, sms); run<2>("IDP.4A", 1, d, sms); run<3>("FFMA", 1, d, sms);
    run<4>("SHF", 1, d, sms); run<5>("IADD+LOP", 2, d, sms); run<6>("pixel: 6 IMAD + 2 SHF + LOP", 9, d, sms); run<7>("pixel: 2 IDP2A + SHF + 2 IMAD", 5, d, sms);
    return 0;
}
