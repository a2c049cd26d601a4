// Special functions the reference registers through its SpecialFunctions extension
// (/root/reference/ext/functionlist.jl:6-123, ext/ExaModelsSpecialFunctions.jl) that the CUDA math library does not
// provide: the polygamma family, invdigamma, dawson, erfi and the four Airy functions.  (erf, erfc, erfcx, erfinv,
// erfcinv, tgamma, lgamma, j0/j1/jn, y0/y1/yn come from libdevice.)  The values themselves come from a third-party
// package in the reference (SpecialFunctions.jl / openspecfun, not vendored), so what is restated here are the published
// algorithms: recurrence + Stirling-type asymptotic series for the polygamma functions (reflection for x <= 0), Minka's
// Newton iteration for invdigamma (SpecialFunctions.jl's own method), the positive-term series e^{-x^2} sum x^{2k+1}/(k!(2k+1))
// and the large-x asymptotic series for Dawson's integral, and -- for Airy -- Taylor expansion of y'' = x y about the
// nearest integer (values at the integers tabulated to 20 digits) with the Poincare asymptotic series beyond |x| > 20.5.
//
// This header is prepended to every generated kernel module (see the Makefile: exb_embed.cpp) and is also compiled for the
// host by tests/test_special_functions.py (through tests/special_host.cpp) so that the algorithms can be pinned against scipy
// without a GPU.  Double precision throughout; accuracy ~1e-14 relative away from zeros.
#pragma once
#ifdef __CUDACC__
#define EXB_SF __device__ __noinline__
#define EXB_SFI __device__ __forceinline__
#define EXB_SFCONST __device__ const
#else
#include <cmath>
#define EXB_SF static inline
#define EXB_SFI static inline
#define EXB_SFCONST static const
EXB_SFI double sinpi(double x) { double r = std::fmod(x, 2.0); return std::sin(3.14159265358979323846 * r); }
EXB_SFI double cospi(double x) { double r = std::fmod(x, 2.0); return std::cos(3.14159265358979323846 * r); }
#endif

#define EXB_SF_PI 3.14159265358979323846
#define EXB_SF_SQRTPI 1.77245385090551602730
#define EXB_SF_INVSQRTPI 0.56418958354775628695

// ---- polygamma family -----------------------------------------------------------------------------------------------
// x > 0: psi^(n)(x) = psi^(n)(x + 1) - (-1)^n n! / x^(n+1) up to x >= 12, then the asymptotic series (Bernoulli numbers
// through B_14: the first neglected term is < 1e-16 relative at x = 12).  x <= 0: reflection about 1 - x.
EXB_SFI double exb_sf_digamma_pos(double x) {
  double r = 0.0;
  while (x < 12.0) { r -= 1.0 / x; x += 1.0; }
  const double i = 1.0 / x, i2 = i * i;
  const double s = i2 * (1.0 / 12.0 - i2 * (1.0 / 120.0 - i2 * (1.0 / 252.0 - i2 * (1.0 / 240.0 - i2 * (1.0 / 132.0 - i2 * (691.0 / 32760.0 - i2 * (1.0 / 12.0)))))));
  return r + log(x) - 0.5 * i - s;
}
EXB_SFI double exb_sf_trigamma_pos(double x) {
  double r = 0.0;
  while (x < 12.0) { r += 1.0 / (x * x); x += 1.0; }
  const double i = 1.0 / x, i2 = i * i;
  const double s = i2 * (1.0 / 6.0 - i2 * (1.0 / 30.0 - i2 * (1.0 / 42.0 - i2 * (1.0 / 30.0 - i2 * (5.0 / 66.0 - i2 * (691.0 / 2730.0 - i2 * (7.0 / 6.0)))))));
  return r + i * (1.0 + 0.5 * i + s);
}
EXB_SFI double exb_sf_polygamma2_pos(double x) {   // psi''(x)
  double r = 0.0;
  while (x < 12.0) { r -= 2.0 / (x * x * x); x += 1.0; }
  const double i = 1.0 / x, i2 = i * i;
  const double s = i2 * (0.5 - i2 * (1.0 / 6.0 - i2 * (1.0 / 6.0 - i2 * (3.0 / 10.0 - i2 * (5.0 / 6.0 - i2 * (691.0 / 210.0 - i2 * (35.0 / 2.0)))))));
  return r - i2 * (1.0 + i + s);
}
EXB_SFI double exb_sf_polygamma3_pos(double x) {   // psi'''(x)
  double r = 0.0;
  while (x < 12.0) { const double x2 = x * x; r += 6.0 / (x2 * x2); x += 1.0; }
  const double i = 1.0 / x, i2 = i * i;
  const double s = i2 * (2.0 - i2 * (1.0 - i2 * (4.0 / 3.0 - i2 * (3.0 - i2 * (10.0 - i2 * (691.0 / 15.0 - i2 * 280.0))))));
  return r + i2 * i * (2.0 + 3.0 * i + s);
}
EXB_SF double exb_digamma(double x) {
  if (x > 0.0) return exb_sf_digamma_pos(x);
  const double s = sinpi(x), c = cospi(x);                       // psi(x) = psi(1 - x) - pi cot(pi x)
  return exb_sf_digamma_pos(1.0 - x) - EXB_SF_PI * c / s;
}
EXB_SF double exb_trigamma(double x) {
  if (x > 0.0) return exb_sf_trigamma_pos(x);
  const double s = sinpi(x);                                     // psi1(x) = -psi1(1 - x) + pi^2 / sin^2(pi x)
  return -exb_sf_trigamma_pos(1.0 - x) + (EXB_SF_PI * EXB_SF_PI) / (s * s);
}
EXB_SF double exb_polygamma2(double x) {
  if (x > 0.0) return exb_sf_polygamma2_pos(x);
  const double s = sinpi(x), c = cospi(x);                       // psi2(x) = psi2(1 - x) - 2 pi^3 cot csc^2
  return exb_sf_polygamma2_pos(1.0 - x) - 2.0 * (EXB_SF_PI * EXB_SF_PI * EXB_SF_PI) * c / (s * s * s);
}
EXB_SF double exb_polygamma3(double x) {
  if (x > 0.0) return exb_sf_polygamma3_pos(x);
  const double s = sinpi(x), c = cospi(x), s2 = s * s;           // psi3(x) = -psi3(1 - x) + 2 pi^4 csc^2 (2 cot^2 + csc^2)
  const double p4 = (EXB_SF_PI * EXB_SF_PI) * (EXB_SF_PI * EXB_SF_PI);
  return -exb_sf_polygamma3_pos(1.0 - x) + 2.0 * p4 * (2.0 * c * c + 1.0) / (s2 * s2);
}
// Minka's fixed-point / Newton iteration, as in SpecialFunctions.jl's invdigamma (closed-form start, <= 25 steps, 1e-12)
EXB_SF double exb_invdigamma(double y) {
  double xo = y >= -2.22 ? exp(y) + 0.5 : -1.0 / (y - (-0.57721566490153286061));
  double xn = xo, delta = 1e300;
  for (int it = 0; it < 25 && delta > 1e-12; it++) {
    xn = xo - (exb_digamma(xo) - y) / exb_trigamma(xo);
    delta = fabs(xn - xo);
    xo = xn;
  }
  return xn;
}

// ---- Dawson's integral and erfi ---------------------------------------------------------------------------------------
// sum_{k>=0} x^(2k+1) / (k! (2k+1)) = (sqrt(pi)/2) erfi(x): all terms of one sign, so no cancellation
EXB_SFI double exb_sf_erfi_series(double x) {
  const double x2 = x * x;
  double t = x, s = x;
  for (int k = 1; k < 200; k++) {
    t *= x2 / (double)k;
    const double a = t / (double)(2 * k + 1);
    s += a;
    if (fabs(a) < 1e-17 * fabs(s)) break;
  }
  return s;
}
// D(x) ~ 1/(2x) sum_{k>=0} (2k-1)!! / (2 x^2)^k, truncated at its smallest term (< e^{-x^2}: below 1e-18 for |x| >= 6.5)
EXB_SFI double exb_sf_dawson_asym(double x) {
  const double q = 1.0 / (2.0 * x * x);
  double t = 1.0, s = 1.0;
  for (int k = 1; k < 60; k++) {
    const double tn = t * (double)(2 * k - 1) * q;
    if (fabs(tn) >= fabs(t) || fabs(tn) < 1e-18) break;
    t = tn; s += t;
  }
  return s / (2.0 * x);
}
EXB_SF double exb_dawson(double x) {
  if (fabs(x) < 6.5) return exp(-x * x) * exb_sf_erfi_series(x);
  return exb_sf_dawson_asym(x);
}
EXB_SF double exb_erfi(double x) {
  if (fabs(x) < 6.5) return (2.0 * EXB_SF_INVSQRTPI) * exb_sf_erfi_series(x);
  return (2.0 * EXB_SF_INVSQRTPI) * exp(x * x) * exb_sf_dawson_asym(x);
}

// ---- Airy functions -------------------------------------------------------------------------------------------------
// {Ai, Ai', Bi, Bi'} at x = -20 ... 20 (20 significant digits; generated with mpmath, tests/golden/make_special_golden.py)
EXB_SFCONST double exb_sf_airy_tab[41][4] = {
  {-1.7640612707798468959e-1, 8.928628567364712384e-1, -2.0013930932265134928e-1, -7.9142903383953647936e-1},
  {-1.4166127688042265637e-1, -1.0049611250051395935, 2.3012109009458831467e-1, -6.1447375395607405676e-1},
  {2.7120454080441422158e-1, -1.5903891520496801577e-1, 3.8372488508383998075e-2, 1.1511870941086417987},
  {-1.0526230029095239023e-1, 1.0586845766446600774, -2.5713592100234318214e-1, -4.3780206579098750947e-1},
  {-1.4305793166909969778e-1, -9.7476444162127271796e-1, 2.4312315142822721669e-1, -5.6845560597613537272e-1},
  {2.7821749087082892953e-1, 2.7237420430864202083e-1, -6.9126594531010061186e-2, 1.0764297530843747867},
  {-2.6598348278407779838e-1, 4.4302487700284364117e-1, -1.1966555279762452313e-1, -9.9741181894933352405e-1},
  {1.7151043937053704463e-1, -8.7151967787995336672e-1, 2.4261322909262719933e-1, 6.2309724881928773354e-1},
  {-6.6555175054373129474e-2, 1.0231104533679707299, -2.9571991207807305673e-1, -2.3673219783112331633e-1},
  {-8.75958925570238129e-3, -1.0273278736645794215, 3.0965476742678188633e-1, -2.2022995314464466559e-2},
  {4.0241238486443190689e-2, 9.962650441327900559e-1, -3.1467982964383863316e-1, 1.1941411339990923828e-1},
  {-2.2133721547341403674e-2, -9.7566398092633159471e-1, 3.2494732345524491792e-1, -5.7400513843669254393e-2},
  {-5.2705050356386202622e-2, 9.3556093819830655103e-1, -3.3125158075113785997e-1, -1.5945049781298138935e-1},
  {1.8428083525050563728e-1, -7.7100816841012654773e-1, 2.9376207185441402012e-1, 4.9824459005811348875e-1},
  {-3.2914517362982310523e-1, 3.4593548728134289493e-1, -1.4669837667055703788e-1, -8.1289878510506700042e-1},
  {3.5076100902411431979e-1, 3.2719281855444313679e-1, -1.3836913490160057685e-1, 7.7841177300189924609e-1},
  {-7.0265532949289515099e-2, -7.906285753685813803e-1, 3.9223470570699928955e-1, -1.1667056743834089368e-1},
  {-3.7881429367765807435e-1, 3.1458376921659881365e-1, -1.9828962637492654322e-1, -6.7561122268525853767e-1},
  {2.2740742820168557599e-1, 6.1825902074169104141e-1, -4.1230258795639848808e-1, 2.7879516692116952269e-1},
  {5.355608832923521188e-1, -1.0160567116645209395e-2, 1.0399738949694461189e-1, 5.9237562642279235082e-1},
  {3.5502805388781723926e-1, -2.5881940379280679841e-1, 6.1492662744600073515e-1, 4.4828835735382635791e-1},
  {1.3529241631288141552e-1, -1.5914744129679321279e-1, 1.2074235949528712594, 9.3243593339277563296e-1},
  {3.4924130423274379135e-2, -5.3090384433653631704e-2, 3.2980949999782147103, 4.1006820499328898894},
  {6.5911393574607191443e-3, -1.1912976705951318474e-2, 1.4037328963730232032e+1, 2.2922214966382170185e+1},
  {9.5156385120480187362e-4, -1.9586409502041789001e-3, 8.3847071408468139923e+1, 1.6192668350461340184e+2},
  {1.0834442813607441735e-4, -2.47413890868462476e-4, 6.5779204417117118244e+2, 1.4358190802179825187e+3},
  {9.9476943602528895702e-6, -2.4765200397034954754e-5, 6.5364461048098634538e+3, 1.5725602621930476839e+4},
  {7.4921288639971670808e-7, -2.0081508947387919912e-6, 8.0327790709430247005e+4, 2.0955267087397131951e+5},
  {4.6922076160992316256e-8, -1.3414392979067865743e-7, 1.1995860041244599309e+6, 3.3543423127445388765e+6},
  {2.4711684308724898433e-9, -7.4806413896589464128e-9, 2.1472868891435349093e+7, 6.3807489780908213855e+7},
  {1.1047532552898685934e-10, -3.5206336767389236366e-10, 4.55641153548225141e+8, 1.4292361344828657761e+9},
  {4.2262758649603595913e-12, -1.4111441246628517335e-11, 1.1355782530430476285e+10, 3.7400168196926977015e+10},
  {1.393184688875360839e-13, -4.854736554985308463e-13, 3.2980722582907417618e+11, 1.1355075024433707424e+12},
  {3.981776078833335363e-15, -1.4432080573972626044e-14, 1.1086706719059404747e+13, 3.9757544969908345404e+13},
  {9.9202054911923772663e-17, -3.7293101100179006797e-16, 4.2880536178653414954e+14, 1.5966914115880027886e+15},
  {2.164962520737992299e-18, -8.4205679540177727661e-18, 1.8982099567493589685e+16, 7.3197492034070104962e+16},
  {4.1568888289170243947e-20, -1.6691886768381809559e-19, 9.5721239060491865258e+17, 3.8137435071218626559e+18},
  {7.0501972983886145424e-22, -2.9171482192933137933e-21, 5.4753038113305869824e+19, 2.2494002910657269272e+20},
  {1.0600466825247955656e-23, -4.5120018606819418892e-23, 3.53891825035656867e+21, 1.4964796503287850684e+22},
  {1.4177043777933527189e-25, -6.1981458271300150586e-25, 2.5755355522344585457e+23, 1.1192350063395887811e+24},
  {1.6916728686705403136e-27, -7.5863916257483549605e-27, 2.1037650496511038145e+25, 9.3818393361339643491e+25}};

// y'' = x y expanded about x0: a_{n+2} = (x0 a_n + a_{n-1}) / ((n+2)(n+1)); returns y(x0 + h) and y'(x0 + h)
EXB_SFI void exb_sf_airy_taylor(double x0, double h, double y0, double yp0, double& y, double& yp) {
  double am1 = 0.0, a0 = y0, a1 = yp0;           // a_{n-1}, a_n, a_{n+1}
  double hn = h;                                  // h^(n+1)
  double sy = y0 + yp0 * h, syp = yp0;
  int small = 0;                                  // consecutive negligible terms (every third coefficient can vanish, e.g. x0 = 0)
  for (int n = 0; n < 80; n++) {
    const double a2 = (x0 * a0 + am1) / (double)((n + 2) * (n + 1));
    syp += (double)(n + 2) * a2 * hn;             // (n+2) a_{n+2} h^(n+1)
    hn *= h;
    const double ty = a2 * hn;                    // a_{n+2} h^(n+2)
    sy += ty;
    am1 = a0; a0 = a1; a1 = a2;
    small = (fabs(ty) <= 1e-18 * fabs(sy) && fabs(ty) <= 1e-18 * fabs(syp)) ? small + 1 : 0;
    if (small >= 3) break;
  }
  y = sy; yp = syp;
}
// |x| > 20.5: Poincare series (DLMF 9.7.5-9.7.12), u_k = (6k-5)(6k-3)(6k-1) / ((2k-1) 216 k) u_{k-1}, v_k = (6k+1)/(1-6k) u_k
EXB_SFI void exb_sf_airy_asym(double x, double& ai, double& aip, double& bi, double& bip) {
  const double z = fabs(x), rz = sqrt(z), z14 = sqrt(rz), zeta = (2.0 / 3.0) * z * rz;
  double u[13], v[13];
  u[0] = 1.0; v[0] = 1.0;
  for (int k = 1; k < 13; k++) {
    const double kk = (double)k;
    const double c = (6.0 * kk - 5.0) * (6.0 * kk - 3.0) * (6.0 * kk - 1.0) / ((2.0 * kk - 1.0) * 216.0 * kk);
    u[k] = u[k - 1] * c;
    v[k] = u[k] * (6.0 * kk + 1.0) / (1.0 - 6.0 * kk);
  }
  // u[k], v[k] are the plain coefficients; zeta^-k is folded in while summing
  if (x > 0.0) {
    double sa = 0.0, sap = 0.0, sb = 0.0, sbp = 0.0, p = 1.0, sg = 1.0;
    for (int k = 0; k < 13; k++) { sa += sg * u[k] * p; sap += sg * v[k] * p; sb += u[k] * p; sbp += v[k] * p; p /= zeta; sg = -sg; }
    const double em = exp(-zeta), ep = exp(zeta);
    ai = em / (2.0 * EXB_SF_SQRTPI * z14) * sa;
    aip = -z14 * em / (2.0 * EXB_SF_SQRTPI) * sap;
    bi = ep / (EXB_SF_SQRTPI * z14) * sb;
    bip = z14 * ep / EXB_SF_SQRTPI * sbp;
  } else {
    double pe = 0.0, po = 0.0, qe = 0.0, qo = 0.0, p = 1.0;   // even / odd partial sums with alternating signs
    for (int k = 0; k < 12; k += 2) {
      const double sg = (k & 2) ? -1.0 : 1.0;
      pe += sg * u[k] * p; qe += sg * v[k] * p; p /= zeta;
      po += sg * u[k + 1] * p; qo += sg * v[k + 1] * p; p /= zeta;
    }
    const double th = zeta - 0.25 * EXB_SF_PI, c = cos(th), s = sin(th);
    ai = (c * pe + s * po) / (EXB_SF_SQRTPI * z14);
    aip = z14 / EXB_SF_SQRTPI * (s * qe - c * qo);
    bi = (-s * pe + c * po) / (EXB_SF_SQRTPI * z14);
    bip = z14 / EXB_SF_SQRTPI * (c * qe + s * qo);
  }
}
// which: 0 Ai, 1 Ai', 2 Bi, 3 Bi'
EXB_SF double exb_airy(double x, int which) {
  if (!(fabs(x) < 20.5)) {
    if (x != x) return x;
    double ai, aip, bi, bip;
    exb_sf_airy_asym(x, ai, aip, bi, bip);
    return which == 0 ? ai : which == 1 ? aip : which == 2 ? bi : bip;
  }
  const double x0 = floor(x + 0.5);
  const int k = (int)x0 + 20;
  double y, yp;
  if (which < 2) exb_sf_airy_taylor(x0, x - x0, exb_sf_airy_tab[k][0], exb_sf_airy_tab[k][1], y, yp);
  else exb_sf_airy_taylor(x0, x - x0, exb_sf_airy_tab[k][2], exb_sf_airy_tab[k][3], y, yp);
  return (which & 1) ? yp : y;
}
