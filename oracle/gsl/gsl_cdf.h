/* TEST INFRASTRUCTURE ONLY (oracle/): stand-in for <gsl/gsl_cdf.h>.
 *
 * The reference includes <gsl/gsl_cdf.h> (MaxCorrelation.c:14) and links -lgsl
 * (README.md:94-95).  GSL is not vendored in the reference and not installed in
 * this image, so the three symbols the file links against are restated in
 * oracle/gsl_shim.c.  Only the prototypes are declared here.
 */
#ifndef RR_ORACLE_GSL_CDF_STUB_H
#define RR_ORACLE_GSL_CDF_STUB_H
#ifdef __cplusplus
extern "C" {
#endif
double gsl_cdf_hypergeometric_P(const unsigned int k, const unsigned int n1,
                                const unsigned int n2, const unsigned int t);
double gsl_cdf_hypergeometric_Q(const unsigned int k, const unsigned int n1,
                                const unsigned int n2, const unsigned int t);
double gsl_cdf_binomial_Q(const unsigned int k, const double p, const unsigned int n);
#ifdef __cplusplus
}
#endif
#endif
