// =============================================================================================
// hb_math — FP64 transcendental helpers for the generated stepper bodies (sm_100a).
//
// Why not libdevice exp()/tanh()/division: on B200 the ExpHydro step is bounded by instruction
// ISSUE, not by the FP64 pipe or HBM (ncu r01: 576 warp instructions per basin-step of which only
// 209 are FP64; 144 are IMAD.MOV constant materialisations for libdevice's polynomial
// coefficients and ~135 are branches/selects of its special-case paths).  These versions are
// branch-free and accurate to ~2 ulp over the full double range — far inside the 1e-10 parity budget.  The file also compiles as plain C++ (HB_HOST_TEST) so the CPU test suite
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

// exp(x) = 2^k * 2^(j/64) * exp(r):  t = round(x * 64/ln2), j = t mod 64, k = t div 64, r = x - t*ln2/64,
// |r| <= ln2/128, exp(r) by its degree-5 Taylor polynomial (truncation 3.5e-17 relative).  The 64-entry
// table lives in shared memory (lane-dependent index); 10 FP64-pipe instructions per exp instead of the
// 16 of a table-free degree-11 kernel — the FP64 pipe and the board's power cap are what bound the
// transcendental-heavy bodies (ExpHydro: 6 exp-class evaluations per step, M50: 64 tanh per step).
#define HB_EXP2_TABLE \
    1.0, 1.0108892860517005, 1.0218971486541166, 1.0330248790212284, \
    1.0442737824274138, 1.0556451783605572, 1.0671404006768237, 1.0787607977571199, \
    1.0905077326652577, 1.102382583307841, 1.1143867425958924, 1.1265216186082418, \
    1.1387886347566916, 1.1511892299529827, 1.1637248587775775, 1.1763969916502812, \
    1.189207115002721, 1.202156731452703, 1.215247359980469, 1.22848053610687, \
    1.241857812073484, 1.255380757024691, 1.2690509571917332, 1.2828700160787783, \
    1.2968395546510096, 1.3109612115247644, 1.3252366431597413, 1.339667524053303, \
    1.3542555469368927, 1.3690024229745905, 1.383909881963832, 1.3989796725383112, \
    1.4142135623730951, 1.42961333839197, 1.4451808069770467, 1.460917794180647, \
    1.4768261459394993, 1.4929077282912648, 1.5091644275934228, 1.5255981507445384, \
    1.5422108254079407, 1.559004400237837, 1.5759808451078865, 1.593142151342267, \
    1.6104903319492543, 1.6280274218573478, 1.645755478153965, 1.6636765803267364, \
    1.681792830507429, 1.7001063537185235, 1.718619298122478, 1.7373338352737062, \
    1.7562521603732995, 1.7753764925265212, 1.7947090750031072, 1.8142521755003989, \
    1.8340080864093424, 1.8539791250833855, 1.8741676341103, 1.8945759815869656, \
    1.9152065613971474, 1.9360617934922943, 1.9571441241754002, 1.978456026387951
#ifdef HB_HOST_TEST
static const double hb_exp2t[64] = {HB_EXP2_TABLE};
static inline void hb_math_init() {}
#else
__constant__ const double HB_EXP2T_SRC[64] = {HB_EXP2_TABLE};
__shared__ double hb_exp2t[64];
// every thread of the block calls this once at kernel entry (contains a block barrier)
__device__ __forceinline__ void hb_math_init() {
    for (int i = threadIdx.x; i < 64; i += blockDim.x) hb_exp2t[i] = HB_EXP2T_SRC[i];   // any block size
    __syncthreads();
}
#endif

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

// m = 2^(j/64) * exp(r) in [1, 2), k such that exp(x) = m * 2^k
HB_DEV double hb_exp_reduce(double x, int &k) {
    const double MAGIC = 6755399441055744.0;               // 1.5 * 2^52: round-to-nearest-integer trick
    double t = fma(x, 92.33248261689366, MAGIC);            // x * 64/ln2
    const int ti = hb_double2loint(t);
    t -= MAGIC;
    double r = fma(t, -0.010830424667801708, x);            // ln2/64 high part (24 trailing zero bits: t*hi exact)
    r = fma(t, -2.8447437476627285e-11, r);                 // ln2/64 low part
    k = ti >> 6;
    double q = fma(r, 8.33333333333333333333e-03, 4.16666666666666666667e-02);
    q = fma(q, r, 1.66666666666666666667e-01);
    q = fma(q, r, 0.5);
    q = fma(q, r, 1.0);
    q = fma(q, r, 1.0);                                      // exp(r)
    return hb_exp2t[ti & 63] * q;
}

// exp(x) for |x| <= ~700: the result is a normal double, so 2^k is applied by adding k to the
// exponent field of m (one integer instruction).
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

// 1/x for finite normal x of either sign: MUFU.RCP64H seed (>= 20 bits) + ONE cubic step r (1 + e + e^2), e = 1 - x r:
// residual e^3 <= 2^-60, three FP64 operations instead of the four of two Newton steps.
HB_DEV double hb_rcp(double x) {
    const double r = hb_rcp_seed(x);
    const double e = fma(-x, r, 1.0);
    const double p = fma(e, e, e);
    return fma(r, p, r);
}
HB_DEV double hb_rcp_pos(double x) { return hb_rcp(x); }

// step_func(x) = (tanh(5x)+1)/2 (/root/reference/src/HydroModels.jl:204) == 1/(1+exp(-10x)).
// z = -10x saturates at +/-700: there the logistic is 1 to the last bit / < 1e-304 (the
// reference's tanh form returns exactly 0 from x ~ -3.8 on; both are 0 within 1e-16 absolute).
HB_DEV double hb_step(double x) {
    int k;
    const double MAGIC = 6755399441055744.0;
    const double z = hb_clamp_abs(-10.0 * x, HB_HI_700);
    double t = fma(z, 92.33248261689366, MAGIC);
    const int ti = hb_double2loint(t);
    t -= MAGIC;
    double r = fma(t, -0.010830424667801708, z);
    r = fma(t, -2.8447437476627285e-11, r);
    k = ti >> 6;
    double q = fma(r, 8.33333333333333333333e-03, 4.16666666666666666667e-02);
    q = fma(q, r, 1.66666666666666666667e-01);
    q = fma(q, r, 0.5);
    q = fma(q, r, 1.0);
    q = fma(q, r, 1.0);
    const double tj = hb_exp2t[ti & 63];
    const double tk = hb_hiloint2double(hb_double2hiint(tj) + (k << 20), hb_double2loint(tj));   // 2^k T_j, a normal double for |z| <= 700
    return hb_rcp(fma(tk, q, 1.0));                         // 1 + exp(z) in one fma
}
// tanh(x) = 1 - 2/(1+exp(2x)), same saturated logistic (absolute error < 4.5e-16).  The factor 2 is folded into the
// argument reduction (t = round(x * 128/ln2), r = x - t ln2/128, exp(2x) = 2^(t/64) e^(2r) with the polynomial in r
// carrying the powers of two) and the table product is fused with the `1 +`, so a tanh costs 14 FP64 operations: M50 evaluates 64 of them per basin-step and is bound
// by the FP64 pipe.
#define HB_HI_350 0x4075e000   // high word of 350.0
HB_DEV double hb_tanh(double x) {
    const double MAGIC = 6755399441055744.0;
    const double xc = hb_clamp_abs(x, HB_HI_350);
    double t = fma(xc, 184.66496523378732, MAGIC);          // x * 128/ln2
    const int ti = hb_double2loint(t);
    t -= MAGIC;
    double r = fma(t, -0.005415212333900854, xc);            // ln2/128 high part (half of the ln2/64 split: exact)
    r = fma(t, -1.42237187383136425e-11, r);                 // ln2/128 low part
    double q = fma(r, 2.66666666666666666667e-01, 6.66666666666666666667e-01);   // e^(2r) = sum (2r)^n / n!
    q = fma(q, r, 1.33333333333333333333e+00);
    q = fma(q, r, 2.0);
    q = fma(q, r, 2.0);
    q = fma(q, r, 1.0);
    // 1 + exp(2x) = 1 + (2^k T_j) q in ONE fma: 2^k goes onto the table entry by an integer add to its exponent field
    // (|2x| <= 700: the scaled entry stays a normal double)
    const double tj = hb_exp2t[ti & 63];
    const double tk = hb_hiloint2double(hb_double2hiint(tj) + ((ti >> 6) << 20), hb_double2loint(tj));
    return fma(-2.0, hb_rcp(fma(tk, q, 1.0)), 1.0);
}
// log(x) for positive normal x — the classic argument reduction x = 2^k m, m in [sqrt(1/2), sqrt(2)), f = m - 1, s = f / (2 + f),
// log(1 + f) = f - f^2/2 + s (f^2/2 + R(s^2)) with the degree-7 minimax R of fdlibm's e_log.c (|error| < 2^-58), branch-free:
// ~24 FP64 operations.  x <= 0, subnormal, inf and NaN are the callers' business (hb_pow filters them).
HB_DEV double hb_log_pos(double x) {
    int hi = hb_double2hiint(x);
    const int lo = hb_double2loint(x);
    int k = (hi >> 20) - 1023;
    hi &= 0x000fffff;
    const int up = (hi + 0x95f64) & 0x100000;              // mantissa above sqrt(2): use m/2 and k+1
    k += up >> 20;
    const double m = hb_hiloint2double(hi | (up ^ 0x3ff00000), lo);
    const double f = m - 1.0;
    const double s = f * hb_rcp(2.0 + f);
    const double z = s * s;
    double R = fma(z, 1.479819860511658591e-01, 1.531383769920937332e-01);
    R = fma(R, z, 1.818357216161805012e-01);
    R = fma(R, z, 2.222219843214978396e-01);
    R = fma(R, z, 2.857142874366239149e-01);
    R = fma(R, z, 3.999999999940941908e-01);
    R = fma(R, z, 6.666666666666735130e-01);
    R *= z;
    const double hfsq = 0.5 * f * f;
    const double dk = (double)k;
    return fma(dk, 6.93147180369123816490e-01, f - (hfsq - fma(s, hfsq + R, dk * 1.90821492927058770002e-10)));
}
// pow(x, y) for x >= 0 (the models raise clamped, non-negative ratios to parameter exponents): exp(y log x), 0^y = 0 for y > 0,
// x^0 = 1.  Relative error ~|y log x| * 1.2e-16 (the product is rounded once) — far inside the 1e-10 parity budget — at a
// quarter of the ~200 instructions of libdevice's pow with its special-case branches; negative or NaN x gives NaN like Julia's
// DomainError path would not allow anyway, +inf is not expected here and is returned as exp(y * log(DBL_MAX))-like garbage.
HB_DEV double hb_pow(double x, double y) {
    const double xs = x > 2.2250738585072014e-308 ? x : 1.0;             // keep the reduction on a normal positive argument
    const double r = hb_exp(y * hb_log_pos(xs));
    const double at0 = y == 0.0 ? 1.0 : (y > 0.0 ? 0.0 : 1.0 / 0.0);     // limit for x -> 0+ (subnormals are flushed to it)
    return x > 2.2250738585072014e-308 ? r : (x >= 0.0 ? at0 : x * (0.0 / 0.0) + (0.0 / 0.0));
}
// min / max as one compare + select (the IEEE-NaN-aware fmin/fmax cost twice as much).  NaN
// handling differs from Julia's min/max (which propagate NaN) exactly as fmin/fmax would.
HB_DEV double hb_min(double a, double b) { return a < b ? a : b; }
HB_DEV double hb_max(double a, double b) { return a > b ? a : b; }
HB_DEV double hb_leakyrelu(double x) { return hb_max(0.01 * x, x); }
HB_DEV double hb_clamp(double x, double lo, double hi) { return hb_min(hb_max(x, lo), hi); }

// ---- opt-in Float32 mode: same functions on float (libdevice expf is ~1-2 ulp; the FP32 pipe is not the
//      bottleneck, the halved HBM traffic is the point) ---------------------------------------------
#ifndef HB_HOST_TEST
HB_DEV float hb_exp(float x) { return expf(x); }
HB_DEV float hb_pow(float x, float y) { return powf(x, y); }
HB_DEV float hb_rcp(float x) { return __frcp_rn(x); }
HB_DEV float hb_step(float x) { return __frcp_rn(1.0f + expf(fminf(fmaxf(-10.0f * x, -80.0f), 80.0f))); }
HB_DEV float hb_tanh(float x) { return fmaf(-2.0f, __frcp_rn(1.0f + expf(fminf(fmaxf(2.0f * x, -80.0f), 80.0f))), 1.0f); }
HB_DEV float hb_min(float a, float b) { return a < b ? a : b; }
HB_DEV float hb_max(float a, float b) { return a > b ? a : b; }
HB_DEV float hb_leakyrelu(float x) { return hb_max(0.01f * x, x); }
#endif
