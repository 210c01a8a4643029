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

// Table variant: x = (32 e + j) ln2/32 + r, |r| <= ln2/64, result = 2^e * T[j] * e^r with a degree-6 polynomial.
// T[j] = 2^(j/32) lives one entry per lane in a register (`tab` = exp_table_entry()) and is fetched with a warp
// shuffle, so the FP64 pipe sees 11 instructions instead of 16.  All lanes of the warp must call it together.
static __device__ __constant__ double DQGP_EXP_T32[32] = {
        0x1.0000000000000p+0, 0x1.059b0d3158574p+0, 0x1.0b5586cf9890fp+0, 0x1.11301d0125b51p+0, 0x1.172b83c7d517bp+0,
        0x1.1d4873168b9aap+0, 0x1.2387a6e756238p+0, 0x1.29e9df51fdee1p+0, 0x1.306fe0a31b715p+0, 0x1.371a7373aa9cbp+0,
        0x1.3dea64c123422p+0, 0x1.44e086061892dp+0, 0x1.4bfdad5362a27p+0, 0x1.5342b569d4f82p+0, 0x1.5ab07dd485429p+0,
        0x1.6247eb03a5585p+0, 0x1.6a09e667f3bcdp+0, 0x1.71f75e8ec5f74p+0, 0x1.7a11473eb0187p+0, 0x1.82589994cce13p+0,
        0x1.8ace5422aa0dbp+0, 0x1.93737b0cdc5e5p+0, 0x1.9c49182a3f090p+0, 0x1.a5503b23e255dp+0, 0x1.ae89f995ad3adp+0,
        0x1.b7f76f2fb5e47p+0, 0x1.c199bdd85529cp+0, 0x1.cb720dcef9069p+0, 0x1.d5818dcfba487p+0, 0x1.dfc97337b9b5fp+0,
        0x1.ea4afa2a490dap+0, 0x1.f50765b6e4540p+0};   // 2^(j/32), correctly rounded
__device__ __forceinline__ double exp_table_entry() { return DQGP_EXP_T32[threadIdx.x & 31]; }

__device__ __forceinline__ double fast_exp_tab(double x, double tab) {
    const double S = 0x1.71547652b82fep+5;                 // 32 / ln 2
    const double C_HI = 0x1.62e42fee00000p-6;     // ln2/32 high part (trailing zeros)
    const double C_LO = 0x1.a39ef35793c76p-38;        // ln2/32 - C_HI
    const double MAGIC = 6755399441055744.0;
    const double km = fma(x, S, MAGIC);
    const int k = __double2loint(km);
    const double kf = km - MAGIC;
    double r = fma(kf, -C_HI, x);
    r = fma(kf, -C_LO, r);
    double p = 1.0 / 720.0;
    p = fma(p, r, 1.0 / 120.0);
    p = fma(p, r, 1.0 / 24.0);
    p = fma(p, r, 1.0 / 6.0);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    const int j = k & 31;
    const int tlo = __shfl_sync(0xffffffffu, __double2loint(tab), j);
    const int thi = __shfl_sync(0xffffffffu, __double2hiint(tab), j);
    p *= __hiloint2double(thi, tlo);
    return __hiloint2double(__double2hiint(p) + ((k >> 5) << 20), __double2loint(p));
}

// Degree-5 variant for the fused gradient (10 FP64-pipe instructions): |r| <= ln2/64 leaves a relative truncation error of
// r^6/720 <= 2.2e-15, far inside the 1e-8 the gradient is held to (it is rounded to 4 decimals), and the gradient kernel is
// bound by exactly this pipe.  The Gram matrices keep the degree-6 fast_exp_tab.
__device__ __forceinline__ double fast_exp_tab5(double x, double tab) {
    const double S = 0x1.71547652b82fep+5;
    const double C_HI = 0x1.62e42fee00000p-6;
    const double C_LO = 0x1.a39ef35793c76p-38;
    const double MAGIC = 6755399441055744.0;
    const double km = fma(x, S, MAGIC);
    const int k = __double2loint(km);
    const double kf = km - MAGIC;
    double r = fma(kf, -C_HI, x);
    r = fma(kf, -C_LO, r);
    double p = 1.0 / 120.0;
    p = fma(p, r, 1.0 / 24.0);
    p = fma(p, r, 1.0 / 6.0);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    const int j = k & 31;
    const int tlo = __shfl_sync(0xffffffffu, __double2loint(tab), j);
    const int thi = __shfl_sync(0xffffffffu, __double2hiint(tab), j);
    p *= __hiloint2double(thi, tlo);
    return __hiloint2double(__double2hiint(p) + ((k >> 5) << 20), __double2loint(p));
}

// The same function of an argument that is ALREADY in units of ln2/32 (xs = x * 32/ln2): the range reduction is then exact in two
// additions (no Cody-Waite pair) and the polynomial's coefficients carry the powers of ln2/32 - 9 FP64-pipe instructions.  The fused
// Gaussian gradient gets the scale for free: it multiplies the row operand and the norms that the DMMA Gram identity starts from.
constexpr double DQGP_EXP_S32 = 0x1.71547652b82fep+5;      // 32 / ln 2
__device__ __forceinline__ double fast_exp_tab5_scaled(double xs, double tab) {
    const double MAGIC = 6755399441055744.0;
    const double L = 0x1.62e42fefa39efp-6;                  // ln2 / 32
    const double km = xs + MAGIC;
    const int k = __double2loint(km);
    const double r = xs - (km - MAGIC);                     // exact: |r| <= 1/2
    double p = L * L * L * L * L / 120.0;
    p = fma(p, r, L * L * L * L / 24.0);
    p = fma(p, r, L * L * L / 6.0);
    p = fma(p, r, L * L * 0.5);
    p = fma(p, r, L);
    p = fma(p, r, 1.0);
    const int j = k & 31;
    const int tlo = __shfl_sync(0xffffffffu, __double2loint(tab), j);
    const int thi = __shfl_sync(0xffffffffu, __double2hiint(tab), j);
    p *= __hiloint2double(thi, tlo);
    return __hiloint2double(__double2hiint(p) + ((k >> 5) << 20), __double2loint(p));
}

}  // namespace dqgp
