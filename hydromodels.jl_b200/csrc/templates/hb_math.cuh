// =============================================================================================
// hb_math — FP64 transcendental helpers for the generated stepper bodies (sm_100a).
//
// Why not libdevice exp()/tanh()/division: on B200 the ExpHydro step is bounded by instruction
// ISSUE, not by the FP64 pipe or HBM (ncu r01: 576 warp instructions per basin-step of which only
// 209 are FP64; 144 are IMAD.MOV constant materialisations for libdevice's polynomial
// coefficients and ~135 are branches/selects of its special-case paths).  These versions are
// branch-free, take their coefficients from the constant bank (DFMA reads c[3][..] directly, no
// register or mov), and are accurate to ~1 ulp over the full double range — far inside the
// 1e-10 parity budget.  The file also compiles as plain C++ (HB_HOST_TEST) so the CPU test suite
// checks the very same code against libm.
// =============================================================================================
#ifdef HB_HOST_TEST
#include <cmath>
#include <cstring>
#define HB_DEV static inline
#define HB_CONST static const
HB_DEV double hb_hiloint2double(int hi, int lo) {
    unsigned long long u = ((unsigned long long)(unsigned)hi << 32) | (unsigned)lo;
    double d;
    memcpy(&d, &u, 8);
    return d;
}
HB_DEV int hb_double2loint(double x) {
    unsigned long long u;
    memcpy(&u, &x, 8);
    return (int)(unsigned)(u & 0xffffffffu);
}
HB_DEV double hb_rcp_seed(double x) {   // emulate MUFU.RCP64H: only the upper 32 bits are meaningful
    double r = 1.0 / x;
    unsigned long long u;
    memcpy(&u, &r, 8);
    u &= 0xffffffff00000000ull;
    memcpy(&r, &u, 8);
    return r;
}
#else
#define HB_DEV __device__ __forceinline__
#define HB_CONST __constant__ const
HB_DEV double hb_hiloint2double(int hi, int lo) { return __hiloint2double(hi, lo); }
HB_DEV int hb_double2loint(double x) { return __double2loint(x); }
HB_DEV double hb_rcp_seed(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    return r;
}
#endif

HB_DEV int hb_double2hiint(double x) {
#ifdef HB_HOST_TEST
    unsigned long long u;
    memcpy(&u, &x, 8);
    return (int)(unsigned)(u >> 32);
#else
    return __double2hiint(x);
#endif
}

// Near-minimax (Chebyshev-interpolated, degree 11) polynomial for exp(r), |r| <= ln2/2; relative
// error 1.7e-17 with these double-rounded coefficients.  The r^1 and r^0 coefficients round to 1.0.
HB_CONST double HB_EXPC[10] = {
    2.5110037605963777712e-08,   // r^11
    2.76326396390410297493e-07,  // r^10
    2.7557240918578969823e-06,   // r^9
    2.4801485482328492419e-05,   // r^8
    1.98412698900471137067e-04,  // r^7
    1.38888889523147746524e-03,  // r^6
    8.33333333331960061091e-03,  // r^5
    4.16666666664880954954e-02,  // r^4
    1.66666666666666808057e-01,  // r^3
    5.00000000000001838553e-01   // r^2
};

// |x| saturated at (just above) a bound given by the HIGH WORD of its double representation (3 integer
// instructions, no FP64-pipe work; the NaN-aware fmin/fmax the compiler emits cost ~6 each).
// A NaN is mapped to +/-bound as well (NaN forcing is not propagated through exp/step_func).
HB_DEV double hb_clamp_abs(double x, int bound_hi) {
    const int hi = hb_double2hiint(x);
    const unsigned ahi = (unsigned)hi & 0x7fffffffu;
    const unsigned chi = (ahi < (unsigned)bound_hi ? ahi : (unsigned)bound_hi) | ((unsigned)hi & 0x80000000u);
    // the low word is kept: a saturated value lies in [bound, bound * (1 + 2^-20)) — still inside the
    // range the callers need (3 instructions: LOP3, VIMNMX.U32, LOP3)
    return hb_hiloint2double((int)chi, hb_double2loint(x));
}
#define HB_HI_700 0x4085e000   // high word of 700.0
#define HB_HI_746 0x40875000   // high word of 746.0

// The 13 constants of the exp kernel.  Inside a long time loop the compiler re-materialises each
// 64-bit immediate with two UMOVs per use (81 of ~490 instructions per ExpHydro step); when
// HB_REGCONST is defined the kernel loads them ONCE into registers (opaque to constant propagation)
// and every helper takes them from `hb_K`.
struct HbExpK { double log2e, nln2hi, nln2lo, c[10]; };
HB_DEV double hb_opaque(double x) {
#ifndef HB_HOST_TEST
    asm volatile("" : "+d"(x));
#endif
    return x;
}
#ifndef HB_HOST_TEST
// mutable module-scope copy: a load from it cannot be constant-folded or re-materialised by ptxas
__device__ double hb_expk_mem[13] = {1.4426950408889634074, -6.93147180369123816490e-01, -1.90821492927058770002e-10,
    2.5110037605963777712e-08, 2.76326396390410297493e-07, 2.7557240918578969823e-06, 2.4801485482328492419e-05,
    1.98412698900471137067e-04, 1.38888889523147746524e-03, 8.33333333331960061091e-03, 4.16666666664880954954e-02,
    1.66666666666666808057e-01, 5.00000000000001838553e-01};
#endif
HB_DEV HbExpK hb_expk_load() {
    HbExpK K;
#ifdef HB_HOST_TEST
    K.log2e = 1.4426950408889634074; K.nln2hi = -6.93147180369123816490e-01; K.nln2lo = -1.90821492927058770002e-10;
    for (int i = 0; i < 10; ++i) K.c[i] = HB_EXPC[i];
#else
    const volatile double *m = hb_expk_mem;
    K.log2e = m[0]; K.nln2hi = m[1]; K.nln2lo = m[2];
#pragma unroll
    for (int i = 0; i < 10; ++i) K.c[i] = m[3 + i];
#endif
    return K;
}

// k = round(x / ln2), p = exp(x - k ln2) ~ [0.70, 1.42]
HB_DEV double hb_exp_reduce_k(double x, int &k, const HbExpK &K) {
    const double MAGIC = 6755399441055744.0;            // 1.5 * 2^52: round-to-nearest-integer trick
    double t = fma(x, K.log2e, MAGIC);                  // x * log2(e)
    k = hb_double2loint(t);
    t -= MAGIC;
    double r = fma(t, K.nln2hi, x);                     // ln2 high part (trailing zeros: t*hi exact)
    r = fma(t, K.nln2lo, r);                            // ln2 low part
    double p = K.c[0];
#ifndef HB_HOST_TEST
#pragma unroll
#endif
    for (int i = 1; i < 10; ++i) p = fma(p, r, K.c[i]);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    return p;
}
HB_DEV double hb_exp_reduce(double x, int &k) {
    const double MAGIC = 6755399441055744.0;
    double t = fma(x, 1.4426950408889634074, MAGIC);
    k = hb_double2loint(t);
    t -= MAGIC;
    double r = fma(t, -6.93147180369123816490e-01, x);
    r = fma(t, -1.90821492927058770002e-10, r);
    double p = HB_EXPC[0];
#ifndef HB_HOST_TEST
#pragma unroll
#endif
    for (int i = 1; i < 10; ++i) p = fma(p, r, HB_EXPC[i]);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    return p;
}

// exp(x) for |x| <= ~700: the result is a normal double, so 2^k is applied by adding k to the
// exponent field of p (one integer instruction).
HB_DEV double hb_exp_narrow(double x) {
    int k;
    const double p = hb_exp_reduce(x, k);
    return hb_hiloint2double(hb_double2hiint(p) + (k << 20), hb_double2loint(p));
}

// Full-range exp, branch-free: |x| saturates at 746 (exp(746) = +inf, exp(-746) = 0) and 2^k is
// applied in two normal factors so overflow / gradual underflow come out of the multiplications.
HB_DEV double hb_exp(double x) {
    int k;
    const double p = hb_exp_reduce(hb_clamp_abs(x, HB_HI_746), k);
    const int k1 = k >> 1, k2 = k - k1;
    const double s1 = hb_hiloint2double((k1 + 1023) << 20, 0);
    const double s2 = hb_hiloint2double((k2 + 1023) << 20, 0);
    return (p * s1) * s2;
}

// 1/x for finite normal x of either sign: MUFU.RCP64H seed (>= 20 bits) + 2 Newton steps (80 bits).
HB_DEV double hb_rcp(double x) {
    double r = hb_rcp_seed(x);
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    return r;
}
HB_DEV double hb_rcp_pos(double x) { return hb_rcp(x); }

// step_func(x) = (tanh(5x)+1)/2 (/root/reference/src/HydroModels.jl:204) == 1/(1+exp(-10x)).
// z = -10x saturates at +/-700: there the logistic is 1 to the last bit / < 1e-304 (the
// reference's tanh form returns exactly 0 from x ~ -3.8 on; both are 0 within 1e-16 absolute).
HB_DEV double hb_step(double x) { return hb_rcp(1.0 + hb_exp_narrow(hb_clamp_abs(-10.0 * x, HB_HI_700))); }
// tanh(x) = 1 - 2/(1+exp(2x)), same saturated logistic (absolute error < 4.5e-16).
HB_DEV double hb_tanh(double x) { return fma(-2.0, hb_rcp(1.0 + hb_exp_narrow(hb_clamp_abs(2.0 * x, HB_HI_700))), 1.0); }
// min / max as one compare + select (the IEEE-NaN-aware fmin/fmax cost twice as much).  NaN
// handling differs from Julia's min/max (which propagate NaN) exactly as fmin/fmax would.
HB_DEV double hb_min(double a, double b) { return a < b ? a : b; }
HB_DEV double hb_max(double a, double b) { return a > b ? a : b; }
HB_DEV double hb_leakyrelu(double x) { return hb_max(0.01 * x, x); }
HB_DEV double hb_clamp(double x, double lo, double hi) { return hb_min(hb_max(x, lo), hi); }

// ---- register-constant variants (HB_REGCONST): same arithmetic, constants from `K` ----------------
HB_DEV double hb_exp_narrow_k(double x, const HbExpK &K) {
    int k;
    const double p = hb_exp_reduce_k(x, k, K);
    return hb_hiloint2double(hb_double2hiint(p) + (k << 20), hb_double2loint(p));
}
HB_DEV double hb_exp_k(double x, const HbExpK &K) {
    int k;
    const double p = hb_exp_reduce_k(hb_clamp_abs(x, HB_HI_746), k, K);
    const int k1 = k >> 1, k2 = k - k1;
    return (p * hb_hiloint2double((k1 + 1023) << 20, 0)) * hb_hiloint2double((k2 + 1023) << 20, 0);
}
HB_DEV double hb_step_k(double x, const HbExpK &K) { return hb_rcp(1.0 + hb_exp_narrow_k(hb_clamp_abs(-10.0 * x, HB_HI_700), K)); }
HB_DEV double hb_tanh_k(double x, const HbExpK &K) { return fma(-2.0, hb_rcp(1.0 + hb_exp_narrow_k(hb_clamp_abs(2.0 * x, HB_HI_700), K)), 1.0); }

// ---- opt-in Float32 mode: same functions on float (libdevice expf is ~1-2 ulp; the FP32 pipe is not the
//      bottleneck, the halved HBM traffic is the point) ---------------------------------------------
#ifndef HB_HOST_TEST
HB_DEV float hb_exp(float x) { return expf(x); }
HB_DEV float hb_rcp(float x) { return __frcp_rn(x); }
HB_DEV float hb_step(float x) { return __frcp_rn(1.0f + expf(fminf(fmaxf(-10.0f * x, -80.0f), 80.0f))); }
HB_DEV float hb_tanh(float x) { return fmaf(-2.0f, __frcp_rn(1.0f + expf(fminf(fmaxf(2.0f * x, -80.0f), 80.0f))), 1.0f); }
HB_DEV float hb_min(float a, float b) { return a < b ? a : b; }
HB_DEV float hb_max(float a, float b) { return a > b ? a : b; }
HB_DEV float hb_leakyrelu(float x) { return hb_max(0.01f * x, x); }
#endif
