// fp64 exp for the pairwise kernels: 16 FP64-pipe instructions instead of libm's ~23 (measured: CUDA exp()
// = 22.6 DFMA-equivalents on B200, profiles/r01_fp64_peak.json), worst relative error 1.2e-16 on the
// reduced range (tools/exp2_poly.py).  x = k ln2 + r (Cody-Waite, two-part ln2), e^r by a degree-11
// near-minimax polynomial, 2^k applied by an integer add on the exponent field.
// Valid for x in [-708, 708]; callers clamp.  The FP64 pipe is the bound of these kernels, so every
// instruction saved here is throughput (there is no fp64 SFU).
#pragma once

namespace dqgp {

__device__ __forceinline__ double fast_exp(double x) {
    const double LOG2E = 1.4426950408889634;
    const double LN2_HI = 6.93147180369123816490e-01;   // high part: trailing zeros so k*LN2_HI is exact
    const double LN2_LO = 1.90821492927058770002e-10;
    const double MAGIC = 6755399441055744.0;            // 1.5 * 2^52: add/sub rounds to nearest integer
    const double km = fma(x, LOG2E, MAGIC);
    const int k = __double2loint(km);
    const double kf = km - MAGIC;
    double r = fma(kf, -LN2_HI, x);
    r = fma(kf, -LN2_LO, r);
    double p = 0x1.af632a0f7e2cep-26;
    p = fma(p, r, 0x1.28b4101c77212p-22);
    p = fma(p, r, 0x1.71ddf56d8deb5p-19);
    p = fma(p, r, 0x1.a01991a10d9aep-16);
    p = fma(p, r, 0x1.a01a01b1461c5p-13);
    p = fma(p, r, 0x1.6c16c1880029fp-10);
    p = fma(p, r, 0x1.111111110f21ep-7);
    p = fma(p, r, 0x1.555555554f0bap-5);
    p = fma(p, r, 0x1.555555555555ap-3);
    p = fma(p, r, 0x1.0000000000011p-1);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    return __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
}

}  // namespace dqgp
