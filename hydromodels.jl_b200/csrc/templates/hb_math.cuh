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
#define HB_MEM inline
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
#define HB_MEM __device__ __forceinline__
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

// exp(x) = 2^k * 2^(j/256) * exp(r):  t = round(x * 256/ln2), j = t mod 256, k = t div 256, r = x - t*ln2/256,
// |r| <= ln2/512, exp(r) by its degree-4 Taylor polynomial (truncation 3.8e-17 relative).  The 256-entry table (2 KB) lives in
// shared memory (lane-dependent index); 9 FP64-pipe instructions per exp instead of the 16 of a table-free degree-11 kernel
// (a 64-entry table needs degree 5: one more) — the FP64 pipe and the board's power cap are what bound the
// transcendental-heavy bodies (ExpHydro: 6 exp-class evaluations per step, M50: 64 tanh per step).
#define HB_EXP2_TABLE \
    1.0, 1.0027112750502025, 1.0054299011128027, 1.0081558981184175, \
    1.0108892860517005, 1.0136300849514894, 1.016378314910953, 1.019133996077738, \
    1.0218971486541166, 1.0246677928971357, 1.0274459491187637, 1.030231637686041, \
    1.0330248790212284, 1.0358256936019572, 1.0386341019613787, 1.041450124688316, \
    1.0442737824274138, 1.0471050958792898, 1.0499440858006872, 1.0527907730046264, \
    1.0556451783605572, 1.0585073227945128, 1.061377227289262, 1.0642549128844645, \
    1.0671404006768237, 1.0700337118202419, 1.0729348675259756, 1.075843889062791, \
    1.0787607977571199, 1.0816856149932152, 1.0846183622133092, 1.0875590609177697, \
    1.0905077326652577, 1.0934643990728858, 1.0964290818163769, 1.099401802630222, \
    1.102382583307841, 1.1053714457017412, 1.1083684117236787, 1.1113735033448175, \
    1.1143867425958924, 1.1174081515673693, 1.1204377524096067, 1.12347556733302, \
    1.1265216186082418, 1.129575928566288, 1.1326385195987192, 1.1357094141578055, \
    1.1387886347566916, 1.1418762039695616, 1.1449721444318042, 1.148076478840179, \
    1.1511892299529827, 1.154310420590216, 1.1574400736337511, 1.1605782120274988, \
    1.1637248587775775, 1.1668800369524817, 1.1700437696832502, 1.1732160801636373, \
    1.1763969916502812, 1.1795865274628758, 1.182784710984341, 1.1859915656609938, \
    1.189207115002721, 1.1924313825831512, 1.1956643920398273, 1.1989061670743806, \
    1.202156731452703, 1.2054161090051239, 1.2086843236265816, 1.2119613992768012, \
    1.215247359980469, 1.2185422298274085, 1.2218460329727576, 1.2251587936371455, \
    1.22848053610687, 1.2318112847340759, 1.2351510639369334, 1.2384998981998165, \
    1.241857812073484, 1.245224830175258, 1.2486009771892048, 1.2519862778663162, \
    1.255380757024691, 1.2587844395497165, 1.2621973503942507, 1.2656195145788063, \
    1.2690509571917332, 1.2724917033894028, 1.275941778396392, 1.2794012075056693, \
    1.2828700160787783, 1.2863482295460256, 1.2898358734066657, 1.2933329732290895, \
    1.2968395546510096, 1.3003556433796506, 1.3038812651919358, 1.3074164459346773, \
    1.3109612115247644, 1.3145155879493546, 1.318079601266064, 1.3216532776031575, \
    1.3252366431597413, 1.3288297242059544, 1.3324325470831615, 1.3360451382041458, \
    1.339667524053303, 1.3432997311868353, 1.3469417862329458, 1.3505937158920345, \
    1.3542555469368927, 1.3579273062129011, 1.3616090206382248, 1.365300717204012, \
    1.3690024229745905, 1.3727141650876684, 1.3764359707545302, 1.380167867260238, \
    1.383909881963832, 1.387662042298529, 1.3914243757719262, 1.3951969099662003, \
    1.3989796725383112, 1.4027726912202048, 1.4065759938190154, 1.4103896082172707, \
    1.4142135623730951, 1.4180478843204152, 1.4218926021691656, 1.4257477441054942, \
    1.42961333839197, 1.433489413367789, 1.4373759974489824, 1.4412731191286257, \
    1.4451808069770467, 1.449099089642035, 1.4530279958490526, 1.4569675544014438, \
    1.460917794180647, 1.4648787441464057, 1.4688504333369818, 1.4728328908693675, \
    1.4768261459394993, 1.4808302278224719, 1.4848451658727524, 1.488870989524397, \
    1.4929077282912648, 1.4969554117672355, 1.5010140696264256, 1.5050837316234065, \
    1.5091644275934228, 1.5132561874526098, 1.5173590411982147, 1.5214730189088146, \
    1.5255981507445384, 1.529734466947287, 1.533881997840956, 1.5380407738316568, \
    1.5422108254079407, 1.5463921831410214, 1.550584877685, 1.5547889397770887, \
    1.559004400237837, 1.5632312899713576, 1.567469639965553, 1.5717194812923414, \
    1.5759808451078865, 1.5802537626528246, 1.5845382652524937, 1.588834384317164, \
    1.593142151342267, 1.597461597908627, 1.6017927556826934, 1.606135656416771, \
    1.6104903319492543, 1.6148568142048607, 1.6192351351948637, 1.6236253270173289, \
    1.6280274218573478, 1.632441451987275, 1.6368674497669644, 1.6413054476440063, \
    1.645755478153965, 1.6502175739206177, 1.6546917676561943, 1.6591780921616162, \
    1.6636765803267364, 1.6681872651305825, 1.6727101796415966, 1.6772453570178785, \
    1.681792830507429, 1.6863526334483934, 1.6909247992693053, 1.6955093614893326, \
    1.7001063537185235, 1.7047158096580513, 1.709337763100463, 1.713972247929926, \
    1.718619298122478, 1.723278947746274, 1.7279512309618377, 1.732636182022311, \
    1.7373338352737062, 1.7420442251551564, 1.746767386199169, 1.7515033530318782, \
    1.7562521603732995, 1.761013843037584, 1.7657884359332727, 1.7705759740635547, \
    1.7753764925265212, 1.7801900265154245, 1.785016611318935, 1.789856282321401, \
    1.7947090750031072, 1.7995750249405351, 1.804454167806624, 1.809346539371032, \
    1.8142521755003989, 1.8191711121586085, 1.8241033854070534, 1.8290490314048973, \
    1.8340080864093424, 1.8389805867758937, 1.843966568958626, 1.8489660695104508, \
    1.8539791250833855, 1.8590057724288205, 1.864046048397789, 1.8690999899412386, \
    1.8741676341103, 1.8792490180565602, 1.8843441790323345, 1.8894531543909392, \
    1.8945759815869656, 1.8997126981765553, 1.9048633418176741, 1.9100279502703899, \
    1.9152065613971474, 1.9203992131630474, 1.925605943636125, 1.930826790987627, \
    1.9360617934922943, 1.9413109895286405, 1.9465744175792332, 1.9518521162309783, \
    1.9571441241754002, 1.9624504802089273, 1.9677712232331759, 1.9731063922552343, \
    1.978456026387951, 1.9838201648502194, 1.9891988469672663, 1.9945921121709402
// Polynomial / reduction constants whose double representation has a non-zero low word cannot be instruction immediates: as
// literals the compiler re-materialises each of them with two UMOV / IMAD.MOV per use INSIDE the time loop (22 + ~10 of the
// 246 instructions of an HBV step, 81 of ExpHydro's 451 — issue slots, and the issue port is what bounds these kernels).  Read
// from a (non-const, so not foldable) __constant__ array they are loaded ONCE per kernel (LDC / LDCU into registers or uniform
// registers — sm_100 FP64 instructions take uniform-register operands — hoisted out of the time loop): no instruction per use.
#ifdef HB_HOST_TEST
#define HB_KBANK static const double
#else
#define HB_KBANK __constant__ double
#endif
HB_KBANK hb_k[] = {
    369.32993046757463,          //  0  256/ln2
    -0.002707606166950427,       //  1  -ln2/256 high part
    -7.111859369156821e-12,      //  2  -ln2/256 low part
    4.16666666666666666667e-02,  //  3  1/24
    1.66666666666666666667e-01,  //  4  1/6
    738.6598609351493,           //  5  512/ln2
    -0.0013538030834752135,      //  6  -ln2/512 high part
    -3.5559296845784106e-12,     //  7  -ln2/512 low part
    6.66666666666666666667e-01,  //  8  2/3
    1.33333333333333333333e+00,  //  9  4/3
    1.479819860511658591e-01,    // 10  log: Lg7
    1.531383769920937332e-01,    // 11  Lg6
    1.818357216161805012e-01,    // 12  Lg5
    2.222219843214978396e-01,    // 13  Lg4
    2.857142874366239149e-01,    // 14  Lg3
    3.999999999940941908e-01,    // 15  Lg2
    6.666666666666735130e-01,    // 16  Lg1
    6.93147180369123816490e-01,  // 17  ln2 high part
    1.90821492927058770002e-10,  // 18  ln2 low part
    -0.0013538030870311431,      // 19  -ln2/512 in one double (HB_TANH_EXACT 0)
    3.33333333333333333333e-01,  // 20  1/3 (table log)
    0.2,                         // 21  1/5
};
// log(x) by table (bodies that raise to a power, `#define HB_LOG_TABLE 1` from the host lowering; host tests always): 128 pairs
// {1/c_j rounded to double, -log(that double)} for the centres c_j = 1 + (j + 1/2)/128 of the mantissa intervals — with the
// ROUNDED reciprocal the identity log m = log1p(m/c - 1) + log c holds exactly, |m/c - 1| <= 2^-8 (2 KB of shared memory).
#ifndef HB_LOG_TABLE
#ifdef HB_HOST_TEST
#define HB_LOG_TABLE 1
#else
#define HB_LOG_TABLE 0
#endif
#endif
#define HB_LOG_TABLE_DATA \
    0.9961089494163424, 0.003898640415657309, 0.9884169884169884, 0.01165061721997525, \
    0.9808429118773946, 0.019342962843130987, 0.973384030418251, 0.026976587698202083, \
    0.9660377358490566, 0.03455238150665973, 0.9588014981273408, 0.042071213920687044, \
    0.9516728624535316, 0.049533935122276676, 0.9446494464944649, 0.05694137640013845, \
    0.9377289377289377, 0.06429435070539725, 0.9309090909090909, 0.07159365318700882, \
    0.924187725631769, 0.078840061707776, 0.9175627240143369, 0.08603433734180316, \
    0.9110320284697508, 0.09317722485418334, 0.9045936395759717, 0.10026945316367517, \
    0.8982456140350877, 0.10731173578908804, 0.89198606271777, 0.11430477128005863, \
    0.8858131487889274, 0.12124924363286965, 0.8797250859106529, 0.12814582269193006, \
    0.8737201365187713, 0.13499516453750482, 0.8677966101694915, 0.1417979118602574, \
    0.8619528619528619, 0.1485546943231372, 0.8561872909698997, 0.15526612891112396, \
    0.8504983388704319, 0.16193282026931324, 0.8448844884488449, 0.16855536102980664, \
    0.839344262295082, 0.17513433212784915, 0.8338762214983714, 0.18167030310763463, \
    0.8284789644012945, 0.18816383241818294, 0.8231511254019293, 0.19461546769967167, \
    0.8178913738019169, 0.2010257460605908, 0.8126984126984127, 0.2073951943460706, \
    0.807570977917981, 0.21372432939771818, 0.8025078369905956, 0.22001365830528213, \
    0.7975077881619937, 0.2262636786504534, 0.7925696594427245, 0.232474878743094, \
    0.7876923076923077, 0.238647737850175, 0.7828746177370031, 0.24478272641769092, \
    0.7781155015197568, 0.25088030628580943, 0.7734138972809668, 0.2569409308975004, \
    0.7687687687687688, 0.26296504550088134, 0.764179104477612, 0.26895308734550394, \
    0.7596439169139466, 0.2749054858727992, 0.7551622418879056, 0.2808226629008878, \
    0.750733137829912, 0.2867050328039543, 0.7463556851311953, 0.29255300268637746, \
    0.7420289855072464, 0.2983669725517973, 0.7377521613832853, 0.3041473354672968, \
    0.7335243553008596, 0.3098944777228647, 0.7293447293447294, 0.3156087789863033, \
    0.7252124645892352, 0.32129061245373425, 0.7211267605633803, 0.3269403449958533, \
    0.7170868347338936, 0.3325583373000766, 0.713091922005571, 0.3381449440087164, \
    0.7091412742382271, 0.34370051385331846, 0.7052341597796143, 0.3492253897852883, \
    0.7013698630136986, 0.354719909102929, 0.6975476839237057, 0.3601844035750078, \
    0.6937669376693767, 0.3656191995609647, 0.6900269541778976, 0.37102461812787263, \
    0.6863270777479893, 0.376400975164253, 0.6826666666666666, 0.3817485814908484, \
    0.6790450928381963, 0.3870677429684483, 0.6754617414248021, 0.3923587606028639, \
    0.6719160104986877, 0.3976219306471385, 0.6684073107049608, 0.4028575447010835, \
    0.6649350649350649, 0.4080658898082217, 0.661498708010336, 0.41324724855021927, \
    0.6580976863753213, 0.41840189913888387, 0.6547314578005116, 0.4235301155058032, \
    0.6513994910941476, 0.42863216738969867, 0.6481012658227848, 0.4337083204215594, \
    0.6448362720403022, 0.43875883620762796, 0.6416040100250626, 0.44378397241030104, \
    0.6384039900249376, 0.4487839828270067, 0.6352357320099256, 0.4537591174671205, \
    0.6320987654320988, 0.4587096226269767, 0.628992628992629, 0.46363574096303256, \
    0.6259168704156479, 0.46853771156323926, 0.6228710462287105, 0.4734157700166721, \
    0.6198547215496368, 0.47827014848147026, 0.6168674698795181, 0.48310107575113576, \
    0.6139088729016786, 0.48790877731923904, 0.6109785202863962, 0.4926934754425752, \
    0.6080760095011877, 0.4974553892028189, 0.6052009456264775, 0.5021947345667155, \
    0.6023529411764705, 0.5069117244448544, 0.5995316159250585, 0.5116065687490621, \
    0.5967365967365967, 0.5162794744484545, 0.5939675174013921, 0.5209306456241853, \
    0.5912240184757506, 0.5255602835229274, 0.5885057471264368, 0.5301685866091216, \
    0.585812356979405, 0.5347557506160276, 0.5831435079726651, 0.5393219685956089, \
    0.5804988662131519, 0.5438674309672835, 0.5778781038374717, 0.5483923255655733, \
    0.5752808988764045, 0.5528968376866776, 0.5727069351230425, 0.5573811501340064, \
    0.5701559020044543, 0.5618454432626918, 0.5676274944567627, 0.5662898950231159, \
    0.565121412803532, 0.5707146810034716, 0.5626373626373626, 0.575119974471388, \
    0.5601750547045952, 0.5795059464146423, 0.5577342047930284, 0.5838727655809826, \
    0.5553145336225597, 0.588220598517086, 0.5529157667386609, 0.5925496096066716, \
    0.5505376344086022, 0.5968599611077938, 0.5481798715203426, 0.6011518131893347, \
    0.5458422174840085, 0.6054253239667169, 0.5435244161358811, 0.6096806495368553, \
    0.5412262156448203, 0.6139179440123704, 0.5389473684210526, 0.6181373595550788, \
    0.5366876310272537, 0.6223390464087787, 0.534446764091858, 0.6265231529313529, \
    0.5322245322245323, 0.6306898256261987, 0.5300207039337475, 0.6348392091730102, \
    0.5278350515463918, 0.6389714464579207, 0.5256673511293635, 0.6430866786030273, \
    0.523517382413088, 0.6471850449953095, 0.5213849287169042, 0.6512666833149582, \
    0.5192697768762677, 0.6553317295631277, 0.5171717171717172, 0.6593803180891278, \
    0.5150905432595574, 0.6634125816170662, 0.5130260521042084, 0.6674286512719563, \
    0.5109780439121756, 0.6714286566053024, 0.5089463220675944, 0.6754127256201768, \
    0.5069306930693069, 0.6793809847957973, 0.504930966469428, 0.6833335591116206, \
    0.5029469548133595, 0.6872705720709603, 0.5009784735812133, 0.691192145724142
#ifdef HB_HOST_TEST
static const double hb_exp2t[256] = {HB_EXP2_TABLE};
static const double hb_logt[256] = {HB_LOG_TABLE_DATA};
static inline void hb_math_init() {}
#else
__constant__ const double HB_EXP2T_SRC[256] = {HB_EXP2_TABLE};
__shared__ double hb_exp2t[256];
#if HB_LOG_TABLE
__constant__ const double HB_LOGT_SRC[256] = {HB_LOG_TABLE_DATA};
__shared__ __align__(16) double hb_logt[256];
#endif
// every thread of the block calls this once at kernel entry (contains a block barrier)
__device__ __forceinline__ void hb_math_init() {
    for (int i = threadIdx.x; i < 256; i += blockDim.x) hb_exp2t[i] = HB_EXP2T_SRC[i];   // any block size
#if HB_LOG_TABLE
    for (int i = threadIdx.x; i < 256; i += blockDim.x) hb_logt[i] = HB_LOGT_SRC[i];
#endif
    __syncthreads();
}
#endif

// |x| saturated at (just above) a bound given by the HIGH WORD of its double representation (3 integer instructions: LOP3,
// VIMNMX.U32, LOP3 — no FP64-pipe work; the NaN-aware fmin/fmax the compiler emits cost ~6 each).
// NaN: by default a NaN is saturated like any value above the bound, i.e. NaN forcing does NOT propagate through exp /
// step_func / tanh (the reference propagates it).  `#define HB_NAN_PROPAGATE 1` (library: HB_EXTRA_DEFINES) lets NaN through
// unchanged — its high word lies above the exponent-all-ones pattern of infinity — at the price of two more integer
// instructions per call (measured: GR4J 100 k x 3650 8.4 -> 9.0 ms, the kernel is close to issue-bound); +/-inf saturates
// either way.
#ifndef HB_NAN_PROPAGATE
#define HB_NAN_PROPAGATE 0
#endif
HB_DEV double hb_clamp_abs(double x, int bound_hi) {
    const int hi = hb_double2hiint(x);
    const unsigned ahi = (unsigned)hi & 0x7fffffffu;
#if HB_NAN_PROPAGATE
    const unsigned lim = ahi > 0x7ff00000u ? ahi : (unsigned)bound_hi;     // NaN: no saturation
#else
    const unsigned lim = (unsigned)bound_hi;
#endif
    const unsigned chi = (ahi < lim ? ahi : lim) | ((unsigned)hi & 0x80000000u);
    // the low word is kept: a saturated value lies in [bound, bound * (1 + 2^-20)) — still inside the
    // range the callers need
    return hb_hiloint2double((int)chi, hb_double2loint(x));
}
#define HB_HI_700 0x4085e000   // high word of 700.0
#define HB_HI_746 0x40875000   // high word of 746.0

// m = 2^(j/256) * exp(r) in [1, 2), k such that exp(x) = m * 2^k
HB_DEV double hb_exp_reduce(double x, int &k) {
    const double MAGIC = 6755399441055744.0;               // 1.5 * 2^52: round-to-nearest-integer trick
    double t = fma(x, hb_k[0], MAGIC);                      // x * 256/ln2
    const int ti = hb_double2loint(t);
    t -= MAGIC;
    double r = fma(t, hb_k[1], x);                          // ln2/256 high part (a quarter of the ln2/64 split: t*hi exact)
    r = fma(t, hb_k[2], r);                                 // ln2/256 low part
    k = ti >> 8;
    double q = fma(r, hb_k[3], hb_k[4]);                    // 1/24, 1/6; |r| <= ln2/512: r^5/120 < 3.8e-17
    q = fma(q, r, 0.5);
    q = fma(q, r, 1.0);
    q = fma(q, r, 1.0);                                      // exp(r)
    return hb_exp2t[ti & 255] * q;
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
    double t = fma(z, hb_k[0], MAGIC);
    const int ti = hb_double2loint(t);
    t -= MAGIC;
    double r = fma(t, hb_k[1], z);
    r = fma(t, hb_k[2], r);
    k = ti >> 8;
    double q = fma(r, hb_k[3], hb_k[4]);
    q = fma(q, r, 0.5);
    q = fma(q, r, 1.0);
    q = fma(q, r, 1.0);
    const double tj = hb_exp2t[ti & 255];
    const double tk = hb_hiloint2double(hb_double2hiint(tj) + (k << 20), hb_double2loint(tj));   // 2^k T_j, a normal double for |z| <= 700
    return hb_rcp(fma(tk, q, 1.0));                         // 1 + exp(z) in one fma
}
// tanh(x) = 1 - 2/(1+exp(2x)), same saturated logistic (absolute error < 4.5e-16).  The factor 2 is folded into the argument
// reduction (t = round(x * 512/ln2), r = x - t ln2/512, exp(2x) = 2^(t/256) e^(2r) with the polynomial in r carrying the powers
// of two) and the table product is fused with the `1 +`, so a tanh costs 13 FP64 operations; M50 evaluates 64 of them per
// basin-step.
// `#define HB_TANH_EXACT 0` (library: HB_EXTRA_DEFINES) selects a form with TEN FP64 operations: r from ONE fma with the full
// double ln2/512 (the product is exact inside the fma; what is lost against the two-constant Cody-Waite form is t * 2^-63
// <= 2e-15 wherever tanh has not saturated), and e^(2r) = (1 + 2r) + [2r^2 + 4/3 r^3 + 2/3 r^4] with the bracket (<= 9.2e-7)
// evaluated in FLOAT (two conversions, four FP32 operations): 8.3e-14 absolute.  It is NOT the default because it buys nothing:
// M50 50 k x 3650 runs 41.65 ms with it and 41.70 ms without — a sub-partition issues one warp instruction per clock and an
// FP64 instruction holds the issue port for two (tools/r2_fp64_mix.cu, profiles/round2_fp64_issue_microbench.txt), so three
// FP64 operations traded for six FP32 / conversion ones leave the issue time where it was.
#define HB_HI_350 0x4075e000   // high word of 350.0
#ifndef HB_TANH_EXACT
#define HB_TANH_EXACT 1
#endif
HB_DEV double hb_tanh(double x) {
    const double MAGIC = 6755399441055744.0;
    const double xc = hb_clamp_abs(x, HB_HI_350);
    double t = fma(xc, hb_k[5], MAGIC);                     // x * 512/ln2
    const int ti = hb_double2loint(t);
    t -= MAGIC;
#if HB_TANH_EXACT
    double r = fma(t, hb_k[6], xc);                          // ln2/512 high part (an eighth of the ln2/64 split: exact)
    r = fma(t, hb_k[7], r);                                  // ln2/512 low part
    double q = fma(r, hb_k[8], hb_k[9]);                     // 2/3, 4/3: e^(2r) = sum (2r)^n / n!, |2r| <= ln2/512
    q = fma(q, r, 2.0);
    q = fma(q, r, 2.0);
    q = fma(q, r, 1.0);
#else
    const double r = fma(t, hb_k[19], xc);                   // |r| <= ln2/1024
    const float rf = (float)r;
    const float tail = (rf * rf) * fmaf(rf, fmaf(rf, 0.66666669f, 1.33333337f), 2.0f);
    const double q = fma(2.0, r, 1.0) + (double)tail;
#endif
    // 1 + exp(2x) = 1 + (2^k T_j) q in ONE fma: 2^k goes onto the table entry by an integer add to its exponent field
    // (|2x| <= 700: the scaled entry stays a normal double)
    const double tj = hb_exp2t[ti & 255];
    const double tk = hb_hiloint2double(hb_double2hiint(tj) + ((ti >> 8) << 20), hb_double2loint(tj));
    return fma(-2.0, hb_rcp(fma(tk, q, 1.0)), 1.0);
}
// log(x) for positive normal x — the classic argument reduction x = 2^k m, m in [sqrt(1/2), sqrt(2)), f = m - 1, s = f / (2 + f),
// log(1 + f) = f - f^2/2 + s (f^2/2 + R(s^2)) with the degree-7 minimax R of fdlibm's e_log.c (|error| < 2^-58), branch-free:
// ~24 FP64 operations.  x <= 0, subnormal, inf and NaN are the callers' business (hb_pow filters them).
#if HB_LOG_TABLE
// log(x) for positive normal x, table form: x = 2^k m, m in [1, 2), j = the top 7 mantissa bits, r = m / c_j - 1 by ONE fma with
// the tabulated reciprocal, log x = k ln2 + log c_j + (r - r^2/2 + r^3/3 - r^4/4 + r^5/5)   (|r| <= 2^-8: r^6/6 < 6e-16).
// 8 FP64 operations instead of the 24 of the table-free form below; absolute error <= 6.2e-16 max(|log x|, 1) (1.1e-15 on
// (1e-6, 1], where the models' ratios live) — exp(y log x) keeps |y| 1e-15 relative, five decades inside the parity bar.
// HBV evaluates one power per step (recharge ~ (soilwater / FC)^(BETA * 5 + 1)): 63.7 -> 57 ms for the calibration ensemble.
HB_DEV double hb_log_pos(double x) {
    const int hi = hb_double2hiint(x);
    const int k = (hi >> 20) - 1023;
    const int j = (hi >> 13) & 127;
    const double m = hb_hiloint2double((hi & 0x000fffff) | 0x3ff00000, hb_double2loint(x));
    const double invc = hb_logt[2 * j], logc = hb_logt[2 * j + 1];
    const double r = fma(m, invc, -1.0);
    double q = fma(r, hb_k[21], -0.25);
    q = fma(q, r, hb_k[20]);
    q = fma(q, r, -0.5);
    const double p = fma(r * r, q, r);
    const double dk = (double)k;
    return fma(dk, hb_k[17], fma(dk, hb_k[18], logc + p));
}
#else
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
    double R = fma(z, hb_k[10], hb_k[11]);
    R = fma(R, z, hb_k[12]);
    R = fma(R, z, hb_k[13]);
    R = fma(R, z, hb_k[14]);
    R = fma(R, z, hb_k[15]);
    R = fma(R, z, hb_k[16]);
    R *= z;
    const double hfsq = 0.5 * f * f;
    const double dk = (double)k;
    return fma(dk, hb_k[17], f - (hfsq - fma(s, hfsq + R, dk * hb_k[18])));
}
#endif
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
HB_DEV double hb_relu(double x) { return hb_double2hiint(x) < 0 ? 0.0 : x; }          // max(0, x) by the sign bit (NaN with a clear sign bit stays NaN)
// NNlib.leakyrelu = max(0.01 x, x): the larger one is x exactly when x's sign bit is clear, so the choice is an integer test of
// the high word (ISETP + 2 SEL) instead of an FP64 compare (DSETP holds the issue port for two cycles) — same bits for every
// non-NaN x, -0.0 included (0.01 * -0.0 = -0.0).
HB_DEV double hb_leakyrelu(double x) { const double m = 0.01 * x; return hb_double2hiint(x) < 0 ? m : x; }
HB_DEV double hb_clamp(double x, double lo, double hi) { return hb_min(hb_max(x, lo), hi); }

HB_DEV double hb_val(double x) { return x; }   // the value of a quantity (gradient runs overload it for hb_dual)

// ---- forward-mode tangents (HB_TANGENTS = D > 0): value + D directional derivatives -------------------------------------------
// The gradient run of the stepper (the reference differentiates the same discrete recurrence with Zygote through its
// ImmutableSolver, /root/reference/README.md:161-187) retypes the generated body's temporaries to `auto` / hb_dual, so the SAME
// statement list propagates d/dp of every state and flux next to the value: exact derivatives of the arithmetic that is
// actually executed (min/max/clamp pass the selected branch's tangent, ties to the second argument like `a > b ? a : b`).
#if defined(HB_TANGENTS) && HB_TANGENTS > 0
#define HB_DK _Pragma("unroll") for (int k = 0; k < HB_TANGENTS; ++k)
struct hb_dual {
    double v;
    double d[HB_TANGENTS];
    HB_MEM hb_dual() {}
    HB_MEM hb_dual(double x) : v(x) { HB_DK d[k] = 0.0; }
};
HB_DEV double hb_val(const hb_dual &x) { return x.v; }
// f(a) with f'(a) = g
HB_DEV hb_dual hb_chain(const hb_dual &a, double f, double g) { hb_dual r; r.v = f; HB_DK r.d[k] = g * a.d[k]; return r; }
HB_DEV hb_dual operator-(const hb_dual &a) { hb_dual r; r.v = -a.v; HB_DK r.d[k] = -a.d[k]; return r; }
HB_DEV hb_dual operator+(const hb_dual &a, const hb_dual &b) { hb_dual r; r.v = a.v + b.v; HB_DK r.d[k] = a.d[k] + b.d[k]; return r; }
HB_DEV hb_dual operator+(const hb_dual &a, double b) { hb_dual r = a; r.v = a.v + b; return r; }
HB_DEV hb_dual operator+(double a, const hb_dual &b) { hb_dual r = b; r.v = a + b.v; return r; }
HB_DEV hb_dual operator-(const hb_dual &a, const hb_dual &b) { hb_dual r; r.v = a.v - b.v; HB_DK r.d[k] = a.d[k] - b.d[k]; return r; }
HB_DEV hb_dual operator-(const hb_dual &a, double b) { hb_dual r = a; r.v = a.v - b; return r; }
HB_DEV hb_dual operator-(double a, const hb_dual &b) { hb_dual r; r.v = a - b.v; HB_DK r.d[k] = -b.d[k]; return r; }
HB_DEV hb_dual operator*(const hb_dual &a, const hb_dual &b) { hb_dual r; r.v = a.v * b.v; HB_DK r.d[k] = fma(a.v, b.d[k], a.d[k] * b.v); return r; }
HB_DEV hb_dual operator*(const hb_dual &a, double b) { hb_dual r; r.v = a.v * b; HB_DK r.d[k] = a.d[k] * b; return r; }
HB_DEV hb_dual operator*(double a, const hb_dual &b) { hb_dual r; r.v = a * b.v; HB_DK r.d[k] = a * b.d[k]; return r; }
HB_DEV hb_dual operator/(const hb_dual &a, const hb_dual &b) {
    hb_dual r; r.v = a.v / b.v; const double ib = 1.0 / b.v; HB_DK r.d[k] = fma(-r.v, b.d[k], a.d[k]) * ib; return r;
}
HB_DEV hb_dual operator/(const hb_dual &a, double b) { hb_dual r; r.v = a.v / b; const double ib = 1.0 / b; HB_DK r.d[k] = a.d[k] * ib; return r; }
HB_DEV hb_dual operator/(double a, const hb_dual &b) { hb_dual r; r.v = a / b.v; const double g = -r.v / b.v; HB_DK r.d[k] = g * b.d[k]; return r; }
HB_DEV hb_dual &operator+=(hb_dual &a, const hb_dual &b) { a.v += b.v; HB_DK a.d[k] += b.d[k]; return a; }
HB_DEV hb_dual &operator*=(hb_dual &a, const hb_dual &b) { a = a * b; return a; }
#define HB_DUAL_CMP(op) \
    HB_DEV bool operator op(const hb_dual &a, const hb_dual &b) { return a.v op b.v; } \
    HB_DEV bool operator op(const hb_dual &a, double b) { return a.v op b; } \
    HB_DEV bool operator op(double a, const hb_dual &b) { return a op b.v; }
HB_DUAL_CMP(<) HB_DUAL_CMP(<=) HB_DUAL_CMP(>) HB_DUAL_CMP(>=)
HB_DEV hb_dual hb_exp(const hb_dual &a) { const double e = hb_exp(a.v); return hb_chain(a, e, e); }
HB_DEV hb_dual hb_step(const hb_dual &a) { const double s = hb_step(a.v); return hb_chain(a, s, 10.0 * s * (1.0 - s)); }   // (tanh(5x)+1)/2
HB_DEV hb_dual hb_tanh(const hb_dual &a) { const double t = hb_tanh(a.v); return hb_chain(a, t, fma(-t, t, 1.0)); }
HB_DEV hb_dual log(const hb_dual &a) { return hb_chain(a, log(a.v), 1.0 / a.v); }
HB_DEV hb_dual sqrt(const hb_dual &a) { const double s = sqrt(a.v); return hb_chain(a, s, 0.5 / s); }
HB_DEV hb_dual fabs(const hb_dual &a) { return a.v < 0.0 ? -a : a; }
HB_DEV hb_dual hb_min(const hb_dual &a, const hb_dual &b) { return a.v < b.v ? a : b; }
HB_DEV hb_dual hb_min(const hb_dual &a, double b) { return a.v < b ? a : hb_dual(b); }
HB_DEV hb_dual hb_min(double a, const hb_dual &b) { return a < b.v ? hb_dual(a) : b; }
HB_DEV hb_dual hb_max(const hb_dual &a, const hb_dual &b) { return a.v > b.v ? a : b; }
HB_DEV hb_dual hb_max(const hb_dual &a, double b) { return a.v > b ? a : hb_dual(b); }
HB_DEV hb_dual hb_max(double a, const hb_dual &b) { return a > b.v ? hb_dual(a) : b; }
HB_DEV hb_dual hb_relu(const hb_dual &a) { return hb_double2hiint(a.v) < 0 ? hb_dual(0.0) : a; }
// x^y: d = y x^(y-1) dx + x^y log(x) dy (the log term is dropped at x <= 0 where the value is 0 / undefined)
HB_DEV hb_dual hb_pow(const hb_dual &x, double y) { return hb_chain(x, hb_pow(x.v, y), y * hb_pow(x.v, y - 1.0)); }
HB_DEV hb_dual hb_pow(double x, const hb_dual &y) {
    const double r = hb_pow(x, y.v);
    return hb_chain(y, r, x > 0.0 ? r * log(x) : 0.0);
}
HB_DEV hb_dual hb_pow(const hb_dual &x, const hb_dual &y) {
    hb_dual r; r.v = hb_pow(x.v, y.v);
    const double gx = y.v * hb_pow(x.v, y.v - 1.0), gy = x.v > 0.0 ? r.v * log(x.v) : 0.0;
    HB_DK r.d[k] = fma(gx, x.d[k], gy * y.d[k]);
    return r;
}
HB_DEV hb_dual fma(const hb_dual &a, const hb_dual &b, const hb_dual &c) {
    hb_dual r; r.v = fma(a.v, b.v, c.v); HB_DK r.d[k] = fma(a.v, b.d[k], fma(a.d[k], b.v, c.d[k])); return r;
}
HB_DEV hb_dual fma(double a, const hb_dual &b, const hb_dual &c) { hb_dual r; r.v = fma(a, b.v, c.v); HB_DK r.d[k] = fma(a, b.d[k], c.d[k]); return r; }
HB_DEV hb_dual fma(const hb_dual &a, double b, const hb_dual &c) { return fma(b, a, c); }
#endif

// ---- opt-in Float32 mode: same functions on float (libdevice expf is ~1-2 ulp; the FP32 pipe is not the
//      bottleneck, the halved HBM traffic is the point) ---------------------------------------------
#ifndef HB_HOST_TEST
HB_DEV double hb_val(float x) { return (double)x; }
HB_DEV float hb_exp(float x) { return expf(x); }
HB_DEV float hb_pow(float x, float y) { return powf(x, y); }
HB_DEV float hb_rcp(float x) { return __frcp_rn(x); }
HB_DEV float hb_step(float x) { return __frcp_rn(1.0f + expf(fminf(fmaxf(-10.0f * x, -80.0f), 80.0f))); }
HB_DEV float hb_tanh(float x) { return fmaf(-2.0f, __frcp_rn(1.0f + expf(fminf(fmaxf(2.0f * x, -80.0f), 80.0f))), 1.0f); }
HB_DEV float hb_min(float a, float b) { return a < b ? a : b; }
HB_DEV float hb_max(float a, float b) { return a > b ? a : b; }
HB_DEV float hb_relu(float x) { return fmaxf(x, 0.0f); }
HB_DEV float hb_leakyrelu(float x) { return hb_max(0.01f * x, x); }
#endif
