#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
__host__ __device__ __forceinline__ int score_a(int v, const int (&r)[16]) {
    int d[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) d[k] = v - r[k];
    int mn2[16], mx2[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) { mn2[k] = min(d[k], d[(k + 1) & 15]); mx2[k] = max(d[k], d[(k + 1) & 15]); }
    int mn4[16], mx4[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) { mn4[k] = min(mn2[k], mn2[(k + 2) & 15]); mx4[k] = max(mx2[k], mx2[(k + 2) & 15]); }
    int best = -256;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const int mn9 = min(min(mn4[k], mn4[(k + 4) & 15]), d[(k + 8) & 15]);
        const int mx9 = max(max(mx4[k], mx4[(k + 4) & 15]), d[(k + 8) & 15]);
        best = max(best, max(mn9, -mx9));
    }
    return best - 1;
}
__host__ __device__ int score_ref(int v, const int* r) {
    int best = -256;
    for (int k = 0; k < 16; k++) { int mn = 999, mx = -999; for (int j = 0; j < 9; j++) { int d = v - r[(k + j) & 15]; mn = d < mn ? d : mn; mx = d > mx ? d : mx; } int t = mn > -mx ? mn : -mx; best = t > best ? t : best; }
    return best - 1;
}
__global__ void k(const unsigned char* in, int n, int* oa, int* ob) {
    int i = blockIdx.x * blockDim.x + threadIdx.x; if (i >= n) return;
    const unsigned char* p = in + 17 * i; int r[16];
#pragma unroll
    for (int j = 0; j < 16; j++) r[j] = p[1 + j];
    oa[i] = score_a(p[0], r); ob[i] = score_ref(p[0], r);
}
int main() {
    const int n = 100000; unsigned char* h = (unsigned char*)malloc(17 * n); srand(1); for (int i = 0; i < 17 * n; i++) h[i] = rand() & 255;
    // make correlated rings so that real corners occur
    for (int i = 0; i < n; i += 2) { int base = rand() & 255; for (int j = 1; j < 17; j++) h[17*i+j] = (unsigned char)((base + (rand() % 30)) & 255); }
    unsigned char* d; int *oa, *ob; cudaMalloc(&d, 17 * n); cudaMalloc(&oa, 4 * n); cudaMalloc(&ob, 4 * n); cudaMemcpy(d, h, 17 * n, cudaMemcpyHostToDevice);
    k<<<(n + 255) / 256, 256>>>(d, n, oa, ob);
    int* ha = (int*)malloc(4 * n); int* hb = (int*)malloc(4 * n); cudaMemcpy(ha, oa, 4 * n, cudaMemcpyDeviceToHost); cudaMemcpy(hb, ob, 4 * n, cudaMemcpyDeviceToHost);
    int bad_a = 0, bad_b = 0;
    for (int i = 0; i < n; i++) { int r[16]; for (int j = 0; j < 16; j++) r[j] = h[17 * i + 1 + j]; int t = score_ref(h[17 * i], r); if (ha[i] != t) { if (bad_a < 5) printf("A i=%d dev=%d ref=%d\n", i, ha[i], t); bad_a++; } if (hb[i] != t) bad_b++; }
    printf("bad sliding=%d bad naive=%d err=%s\n", bad_a, bad_b, cudaGetErrorString(cudaGetLastError()));
}
