// sgs_logf.h -- glibc's float logarithm, restated so that MapPoint::PredictScale (src/MapPoint.cc:402-418:
// `ceil(log(ratio) / mfLogScaleFactor)` with float operands, i.e. libm's logf) gives the same pyramid level on the device as on the host.
// Algorithm of glibc >= 2.27 sysdeps/ieee754/flt-32/e_logf.c (ARM optimized-routines logf, LOGF_TABLE_BITS = 4, degree-3 polynomial
// evaluated in double): x = 2^k * z, z in [0x3f330000, 2*0x3f330000); table entry i from the top 4 mantissa bits gives invc ~ 1/c and
// logc = log(c); r = z*invc - 1; log(x) = k*ln2 + logc + r + r^2 * (A2 + A1*r + A0*r^2), rounded once to float.
// Pinned: tests/test_host_logic.py compares this function with the running libm's logf (bit for bit; exhaustively verified over all
// 2,139,095,039 positive finite floats against glibc 2.39 when it was written -- the float result does not depend on whether the double
// operations are contracted to FMA, so the x86-64 FMA ifunc variant and this plain version agree everywhere).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#if defined(__CUDACC__)
#define SGS_LOGF_HD __device__ __forceinline__
#define SGS_LOGF_TABLE static __constant__ double
#else
#define SGS_LOGF_HD inline
#define SGS_LOGF_TABLE static const double
#endif

namespace sgs {

SGS_LOGF_TABLE kLogfInvc[16] = {0x1.661ec79f8f3bep+0, 0x1.571ed4aaf883dp+0, 0x1.49539f0f010bp+0, 0x1.3c995b0b80385p+0, 0x1.30d190c8864a5p+0, 0x1.25e227b0b8eap+0,
                                0x1.1bb4a4a1a343fp+0, 0x1.12358f08ae5bap+0, 0x1.0953f419900a7p+0, 0x1p+0, 0x1.e608cfd9a47acp-1, 0x1.ca4b31f026aap-1,
                                0x1.b2036576afce6p-1, 0x1.9c2d163a1aa2dp-1, 0x1.886e6037841edp-1, 0x1.767dcf5534862p-1};
SGS_LOGF_TABLE kLogfLogc[16] = {-0x1.57bf7808caadep-2, -0x1.2bef0a7c06ddbp-2, -0x1.01eae7f513a67p-2, -0x1.b31d8a68224e9p-3, -0x1.6574f0ac07758p-3, -0x1.1aa2bc79c81p-3,
                                -0x1.a4e76ce8c0e5ep-4, -0x1.1973c5a611cccp-4, -0x1.252f438e10c1ep-5, 0x0p+0, 0x1.aa5aa5df25984p-5, 0x1.c5e53aa362eb4p-4,
                                0x1.526e57720db08p-3, 0x1.bc2860d22477p-3, 0x1.1058bc8a07ee1p-2, 0x1.4043057b6ee09p-2};

SGS_LOGF_HD float glibc_logf(float x) {
    const double* invc = kLogfInvc;
    const double* logc = kLogfLogc;
    const double ln2 = 0x1.62e42fefa39efp-1, A0 = -0x1.00ea348b88334p-2, A1 = 0x1.5575b0be00b6ap-2, A2 = -0x1.ffffef20a4123p-2;
    uint32_t ix;
#if defined(__CUDA_ARCH__)
    ix = __float_as_uint(x);
#else
    std::memcpy(&ix, &x, 4);
#endif
    if (ix == 0x3f800000u) return 0.f;
    if (ix - 0x00800000u >= 0x7f800000u - 0x00800000u) {
        if (ix * 2 == 0) return -1.f / 0.f;                     // log(+-0) = -inf
        if (ix == 0x7f800000u) return x;                        // log(inf) = inf
        if ((ix & 0x80000000u) || ix * 2 >= 0xff000000u) return 0.f / 0.f;   // negative or NaN
        const float xs = x * 0x1p23f;                           // subnormal: normalise
#if defined(__CUDA_ARCH__)
        ix = __float_as_uint(xs);
#else
        std::memcpy(&ix, &xs, 4);
#endif
        ix -= 23u << 23;
    }
    const uint32_t tmp = ix - 0x3f330000u;
    const int i = (int)((tmp >> 19) & 15u);
    const int k = (int32_t)tmp >> 23;
    const uint32_t iz = ix - (tmp & 0xff800000u);
    float zf;
#if defined(__CUDA_ARCH__)
    zf = __uint_as_float(iz);
    const double z = (double)zf;
    const double r = __dadd_rn(__dmul_rn(z, invc[i]), -1.0);
    const double y0 = __dadd_rn(logc[i], __dmul_rn((double)k, ln2));
    const double r2 = __dmul_rn(r, r);
    double y = __dadd_rn(__dmul_rn(A1, r), A2);
    y = __dadd_rn(__dmul_rn(A0, r2), y);
    y = __dadd_rn(__dmul_rn(y, r2), __dadd_rn(y0, r));
    return __double2float_rn(y);
#else
    std::memcpy(&zf, &iz, 4);
    const double z = (double)zf;
    volatile double r = z * invc[i]; r = r - 1.0;               // volatile: no FMA contraction whatever the host flags
    volatile double y0 = (double)k * ln2; y0 = logc[i] + y0;
    volatile double r2 = r * r;
    volatile double y = A1 * r; y = y + A2;
    volatile double t = A0 * r2; y = t + y;
    t = y * r2; volatile double s = y0 + r; y = t + s;
    return (float)y;
#endif
}

}  // namespace sgs
