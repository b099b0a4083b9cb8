/* TEST INFRASTRUCTURE ONLY (oracle/).  Not part of the product; only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may link or execute this file.
 *
 * Restatement of the part of the GNU Scientific Library that the reference's
 * hot path calls:  MaxCorrelation.c:415
 *     gsl_cdf_hypergeometric_Q(schnitt-1, gr2, cov-gr2, gr1)
 * plus the two symbols its dead code links (MaxCorrelation.c:457-458, 491).
 *
 * Third-party dependency: GSL, version UNPINNED by the reference (README.md:92-95
 * just says "-lgsl -lgslcblas"), not vendored under /root/reference and not
 * installed in this image.  PARITY UNPINNED at this boundary: the reference
 * holds no test or golden vector for it.  What pins this file instead:
 *   - the published GSL 2.x algorithm chain (cdf/hypergeometric.c ->
 *     randist/hypergeometric.c -> specfunc/gamma.c), restated below;
 *   - SURVEY.md Appendix B's known-answer table (exact rational sums);
 *   - tests/test_oracle_shim.py: scipy.stats.hypergeom.sf and mpmath exact sums.
 */
#include <math.h>
#include <float.h>
#include "gsl/gsl_cdf.h"

/* specfunc/gamma.c: exact n! as correctly rounded doubles, n = 0..170
 * (generated from Python's math.factorial, float(n!).hex()). */
static const double rr_fact_table[171] = {
  0x1.0000000000000p+0, /* 0! */
  0x1.0000000000000p+0, /* 1! */
  0x1.0000000000000p+1, /* 2! */
  0x1.8000000000000p+2, /* 3! */
  0x1.8000000000000p+4, /* 4! */
  0x1.e000000000000p+6, /* 5! */
  0x1.6800000000000p+9, /* 6! */
  0x1.3b00000000000p+12, /* 7! */
  0x1.3b00000000000p+15, /* 8! */
  0x1.6260000000000p+18, /* 9! */
  0x1.baf8000000000p+21, /* 10! */
  0x1.308a800000000p+25, /* 11! */
  0x1.c8cfc00000000p+28, /* 12! */
  0x1.7328cc0000000p+32, /* 13! */
  0x1.44c3b28000000p+36, /* 14! */
  0x1.3077775800000p+40, /* 15! */
  0x1.3077775800000p+44, /* 16! */
  0x1.437eeecd80000p+48, /* 17! */
  0x1.6beecca730000p+52, /* 18! */
  0x1.b02b930689000p+56, /* 19! */
  0x1.0e1b3be415a00p+61, /* 20! */
  0x1.6283be9b5c620p+65, /* 21! */
  0x1.e77526159f06cp+69, /* 22! */
  0x1.5e5c335f8a4cep+74, /* 23! */
  0x1.06c52687a7b9ap+79, /* 24! */
  0x1.9a940c33f6121p+83, /* 25! */
  0x1.4d9849ea37eebp+88, /* 26! */
  0x1.19787e5d9f316p+93, /* 27! */
  0x1.ec92dd23d6967p+97, /* 28! */
  0x1.be6518687a785p+102, /* 29! */
  0x1.a27ec6e1f2d0dp+107, /* 30! */
  0x1.956ad0aae33a4p+112, /* 31! */
  0x1.956ad0aae33a4p+117, /* 32! */
  0x1.a21627303a541p+122, /* 33! */
  0x1.bc3789a33df96p+127, /* 34! */
  0x1.e5dcbe8a8bc8cp+132, /* 35! */
  0x1.114c2b2deea0fp+138, /* 36! */
  0x1.3c0011ed1bea1p+143, /* 37! */
  0x1.774015499125fp+148, /* 38! */
  0x1.c95619f1a8e64p+153, /* 39! */
  0x1.1dd5d037098fep+159, /* 40! */
  0x1.6e39f2c684406p+164, /* 41! */
  0x1.e0ac0ea48d948p+169, /* 42! */
  0x1.42f399d68f1fcp+175, /* 43! */
  0x1.bc0ef38704cbbp+180, /* 44! */
  0x1.383a833aef5f3p+186, /* 45! */
  0x1.c0d41ca4b818ep+191, /* 46! */
  0x1.499bc508f7324p+197, /* 47! */
  0x1.ee69a78d72cb6p+202, /* 48! */
  0x1.7a88e4484be3bp+208, /* 49! */
  0x1.27baf2587b49ep+214, /* 50! */
  0x1.d751f23d047dcp+219, /* 51! */
  0x1.7ef294d193a63p+225, /* 52! */
  0x1.3d20e33d8e45ap+231, /* 53! */
  0x1.0b93bfbbf00acp+237, /* 54! */
  0x1.cbe5f18b04928p+242, /* 55! */
  0x1.92693359a4003p+248, /* 56! */
  0x1.6665b1bbd6102p+254, /* 57! */
  0x1.44cc291239feap+260, /* 58! */
  0x1.2b6c35dccd76cp+266, /* 59! */
  0x1.18b5727f009f5p+272, /* 60! */
  0x1.0b8cf1210c97ep+278, /* 61! */
  0x1.0330899804332p+284, /* 62! */
  0x1.fe478ee34844ap+289, /* 63! */
  0x1.fe478ee34844ap+295, /* 64! */
  0x1.0320568f6ab2ep+302, /* 65! */
  0x1.0b395943e6087p+308, /* 66! */
  0x1.17c0097314d0dp+314, /* 67! */
  0x1.293c0a0a461dep+320, /* 68! */
  0x1.4074bad313983p+326, /* 69! */
  0x1.5e7fac56dd6e8p+332, /* 70! */
  0x1.84d5a3305da69p+338, /* 71! */
  0x1.b5705796695b6p+344, /* 72! */
  0x1.f2f423e7902c4p+350, /* 73! */
  0x1.207524c1df599p+357, /* 74! */
  0x1.5209471331bd0p+363, /* 75! */
  0x1.916b0466cb107p+369, /* 76! */
  0x1.e2f4c14bac4fcp+375, /* 77! */
  0x1.264d25ca1d009p+382, /* 78! */
  0x1.6b473aa57bcccp+388, /* 79! */
  0x1.c619094edabffp+394, /* 80! */
  0x1.1f5bd7e3e66d7p+401, /* 81! */
  0x1.702dac9bff3c4p+407, /* 82! */
  0x1.dd7b3bda4f022p+413, /* 83! */
  0x1.3958df4743d96p+420, /* 84! */
  0x1.a02a088aa61cbp+426, /* 85! */
  0x1.179c3dbd279b5p+433, /* 86! */
  0x1.7c1863ed21d72p+439, /* 87! */
  0x1.0550c4b30743ep+446, /* 88! */
  0x1.6b645188f61a6p+452, /* 89! */
  0x1.ff0512a89a152p+458, /* 90! */
  0x1.6b4d9b43dd8b0p+465, /* 91! */
  0x1.051fc798c73bfp+472, /* 92! */
  0x1.7b722e0a01831p+478, /* 93! */
  0x1.16a7d9cf591c4p+485, /* 94! */
  0x1.9da1274fc845fp+491, /* 95! */
  0x1.3638dd7bd6347p+498, /* 96! */
  0x1.d62e2fafb0a78p+504, /* 97! */
  0x1.67fb5c8283404p+511, /* 98! */
  0x1.166c698cf183bp+518, /* 99! */
  0x1.b30964ec395dcp+524, /* 100! */
  0x1.574569a265440p+531, /* 101! */
  0x1.118b502d68b23p+538, /* 102! */
  0x1.b83c3509147ecp+544, /* 103! */
  0x1.65b0eb1760a70p+551, /* 104! */
  0x1.256b20d92d490p+558, /* 105! */
  0x1.e5f96e67b300ep+564, /* 106! */
  0x1.963e824aafa2cp+571, /* 107! */
  0x1.56c4bdef04315p+578, /* 108! */
  0x1.23e389bd89920p+585, /* 109! */
  0x1.f5af14bdc472fp+591, /* 110! */
  0x1.b30dd3fc905bap+598, /* 111! */
  0x1.7cac197cfe503p+605, /* 112! */
  0x1.500fee805882dp+612, /* 113! */
  0x1.2b4e306a4ed48p+619, /* 114! */
  0x1.0ce83f7f82d2fp+626, /* 115! */
  0x1.e764f3171d1e4p+632, /* 116! */
  0x1.bd824633209dbp+639, /* 117! */
  0x1.9ab418b722116p+646, /* 118! */
  0x1.7dd36efa41ac2p+653, /* 119! */
  0x1.65f6380a9d916p+660, /* 120! */
  0x1.5262c0fa08f37p+667, /* 121! */
  0x1.42861fee50880p+674, /* 122! */
  0x1.35ece2af0162bp+681, /* 123! */
  0x1.2c3d7b998957ap+688, /* 124! */
  0x1.25340ab3f01f9p+695, /* 125! */
  0x1.209f3a89205f1p+702, /* 126! */
  0x1.1e5dfc140e1e5p+709, /* 127! */
  0x1.1e5dfc140e1e5p+716, /* 128! */
  0x1.209ab80c363a9p+723, /* 129! */
  0x1.251d22ec67138p+730, /* 130! */
  0x1.2bfbd1bdf17dfp+737, /* 131! */
  0x1.355bb04be109ep+744, /* 132! */
  0x1.4171452ed7d44p+751, /* 133! */
  0x1.5082946d09f23p+758, /* 134! */
  0x1.62e9b88b007d7p+765, /* 135! */
  0x1.79185413b0855p+772, /* 136! */
  0x1.939c09fd12eebp+779, /* 137! */
  0x1.b3243ac4d8695p+786, /* 138! */
  0x1.d88957d1c3026p+793, /* 139! */
  0x1.026b1c06b6a55p+801, /* 140! */
  0x1.1ca9fcdf65321p+808, /* 141! */
  0x1.3bcc9487d4439p+815, /* 142! */
  0x1.60ce8defbf238p+822, /* 143! */
  0x1.8ce85fadb707ep+829, /* 144! */
  0x1.c19f3c62c956fp+836, /* 145! */
  0x1.006cd07056d39p+844, /* 146! */
  0x1.267cf76103b70p+851, /* 147! */
  0x1.54807e082c4b9p+858, /* 148! */
  0x1.8c5d92b583900p+865, /* 149! */
  0x1.d07da7ecb62ccp+872, /* 150! */
  0x1.11fa1e0c9f746p+880, /* 151! */
  0x1.455903aefd5a3p+887, /* 152! */
  0x1.84e466672ad5dp+894, /* 153! */
  0x1.d3e2cb341f894p+901, /* 154! */
  0x1.1b4a51088f182p+909, /* 155! */
  0x1.594292c26e656p+916, /* 156! */
  0x1.a77ba8027b686p+923, /* 157! */
  0x1.055e51b1882a7p+931, /* 158! */
  0x1.44ab297a8724bp+938, /* 159! */
  0x1.95d5f3d928edep+945, /* 160! */
  0x1.fe771cb7257b3p+952, /* 161! */
  0x1.4307602be5b7fp+960, /* 162! */
  0x1.9b5b6477e6884p+967, /* 163! */
  0x1.07868c5ccfaf4p+975, /* 164! */
  0x1.53b370efa3b7fp+982, /* 165! */
  0x1.b88cb676c8529p+989, /* 166! */
  0x1.1f63cb077cadep+997, /* 167! */
  0x1.7932fa79d3a43p+1004, /* 168! */
  0x1.f2054eb4d96ecp+1011, /* 169! */
  0x1.4ab7864418639p+1019, /* 170! */
};

/* specfunc/gamma.c lngamma_lanczos(): Lanczos g=7, 9 coefficients. */
static const double rr_lanczos_7_c[9] = {
    0.99999999999980993227684700473478,
    676.520368121885098567009190444019,
    -1259.13921672240287047156078755283,
    771.3234287776530788486528258894,
    -176.61502916214059906584551354,
    12.507343278686904814458936853,
    -0.13857109526572011689554707,
    9.984369578019570859563e-6,
    1.50563273514931155834e-7};

static double rr_lngamma_lanczos(double x)
{
    int k;
    double Ag, term1, term2;
    x -= 1.0;
    Ag = rr_lanczos_7_c[0];
    for (k = 1; k <= 8; k++) Ag += rr_lanczos_7_c[k] / (x + k);
    term1 = (x + 0.5) * log((x + 7.5) / M_E);
    term2 = 0.9189385332046727418 /* log(sqrt(2 pi)) */ + log(Ag);
    return term1 + (term2 - 7.0);
}

/* gsl_sf_lnfact: table for n <= 170, lngamma(n+1) beyond.  For n+1 >= 172 GSL's
 * lngamma takes the Lanczos branch (x >= 0.5 and not within 0.01 of 1 or 2). */
double rr_oracle_lnfact(unsigned int n)
{
    if (n <= 170) return log(rr_fact_table[n]);
    return rr_lngamma_lanczos((double)n + 1.0);
}

/* gsl_sf_lnchoose */
static double rr_lnchoose(unsigned int n, unsigned int m)
{
    if (m == n || m == 0) return 0.0;
    if (m * 2 > n) m = n - m;
    return rr_oracle_lnfact(n) - rr_oracle_lnfact(m) - rr_oracle_lnfact(n - m);
}

/* gsl_ran_hypergeometric_pdf */
static double rr_hyper_pdf(unsigned int k, unsigned int n1, unsigned int n2, unsigned int t)
{
    if (t > n1 + n2) t = n1 + n2;
    if (k > n1 || k > t) return 0;
    if (t > n2 && k + n2 < t) return 0;
    {
        double c1 = rr_lnchoose(n1, k);
        double c2 = rr_lnchoose(n2, t - k);
        double c3 = rr_lnchoose(n1 + n2, t);
        return exp(c1 + c2 - c3);
    }
}

/* cdf/hypergeometric.c lower_tail / upper_tail: note the mixed unsigned/int/double
 * sub-expressions, kept exactly as GSL writes them. */
static double rr_lower_tail(const unsigned int k, const unsigned int n1,
                            const unsigned int n2, const unsigned int t)
{
    double relerr;
    int i = k;
    double s, P;
    s = rr_hyper_pdf(i, n1, n2, t);
    P = s;
    while (i > 0) {
        double factor = (i / (n1 - i + 1.0)) * ((n2 + i - t) / (t - i + 1.0));
        s *= factor;
        P += s;
        relerr = s / P;
        if (relerr < DBL_EPSILON) break;
        i--;
    }
    return P;
}

static double rr_upper_tail(const unsigned int k, const unsigned int n1,
                            const unsigned int n2, const unsigned int t)
{
    double relerr;
    unsigned int i = k + 1;
    double s, Q;
    s = rr_hyper_pdf(i, n1, n2, t);
    Q = s;
    while (i < t) {
        double factor = ((n1 - i) / (i + 1.0)) * ((t - i) / (n2 + i + 1.0 - t));
        s *= factor;
        Q += s;
        relerr = s / Q;
        if (relerr < DBL_EPSILON) break;
        i++;
    }
    return Q;
}

double gsl_cdf_hypergeometric_P(const unsigned int k, const unsigned int n1,
                                const unsigned int n2, const unsigned int t)
{
    double P;
    if (t > (n1 + n2)) return NAN; /* GSL: domain error */
    if (k >= n1 || k >= t) {
        P = 1.0;
    } else {
        double midpoint = ((double)t * n1) / ((double)n1 + n2);
        if (k >= midpoint) P = 1 - rr_upper_tail(k, n1, n2, t);
        else P = rr_lower_tail(k, n1, n2, t);
    }
    return P;
}

double gsl_cdf_hypergeometric_Q(const unsigned int k, const unsigned int n1,
                                const unsigned int n2, const unsigned int t)
{
    double Q;
    if (t > (n1 + n2)) return NAN; /* GSL: domain error; unreachable from MaxCorrelation.c:415 (gr1 <= cov) */
    if (k >= n1 || k >= t) {
        Q = 0.0;
    } else {
        double midpoint = ((double)t * n1) / ((double)n1 + n2);
        if (k < midpoint) Q = 1 - rr_lower_tail(k, n1, n2, t);
        else Q = rr_upper_tail(k, n1, n2, t);
    }
    return Q;
}

/* Only referenced from dead code (MaxCorrelation.c:491, BestCutoff).  Plain
 * summation of the upper binomial tail; never on the hot path. */
double gsl_cdf_binomial_Q(const unsigned int k, const double p, const unsigned int n)
{
    double Q = 0.0;
    unsigned int x;
    if (p > 1.0 || p < 0.0) return NAN;
    if (k >= n) return 0.0;
    for (x = k + 1; x <= n; x++) {
        double lc = rr_lnchoose(n, x);
        Q += exp(lc + x * log(p) + (n - x) * log1p(-p));
    }
    return Q;
}
