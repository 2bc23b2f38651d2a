/* fftw3.h -- STAND-IN for the FFTW3 header, test infrastructure only.
 *
 * The reference (DmSM-1/C-OFDM) links the system FFTW3 (`-lfftw3`, reference Makefile:3,
 * `#include <fftw3.h>` at OFDM/Frame.hpp:10).  FFTW3 is not installed in this image and there
 * is no network, so the oracle build (oracle/Makefile) puts this directory on the include
 * path instead.  It declares exactly the subset the reference calls (call sites F1-F5,
 * SURVEY.md section 2.1):
 *     fftw_plan_many_dft   (OFDM/Frame.cpp:16-24,108-112,147-150)
 *     fftw_plan_dft_1d     (OFDM/Frame.hpp:289-295)
 *     fftw_execute         (OFDM/Frame.cpp:64,74,152; OFDM/Frame.hpp:118,170,297)
 *     fftw_destroy_plan    (OFDM/Frame.cpp:49-50; OFDM/Frame.hpp:298)
 * Semantics follow the published FFTW3 API: unnormalised complex double DFT,
 * sign -1 = forward (e^{-j}), +1 = backward (e^{+j}); rank-1 "many" layout with
 * howmany/stride/dist; in-place when in == out.  The implementation is
 * fftw3_standin.c (our own mixed-radix Stockham FFT), NOT FFTW.
 */
#ifndef COFDM_ORACLE_FFTW3_STANDIN_H
#define COFDM_ORACLE_FFTW3_STANDIN_H

#ifdef __cplusplus
extern "C" {
#endif

typedef double fftw_complex[2];
typedef struct standin_plan_s *fftw_plan;

#define FFTW_FORWARD (-1)
#define FFTW_BACKWARD (+1)
#define FFTW_MEASURE (0U)
#define FFTW_ESTIMATE (1U << 6)

fftw_plan fftw_plan_many_dft(int rank, const int *n, int howmany,
                             fftw_complex *in, const int *inembed, int istride, int idist,
                             fftw_complex *out, const int *onembed, int ostride, int odist,
                             int sign, unsigned flags);
fftw_plan fftw_plan_dft_1d(int n, fftw_complex *in, fftw_complex *out, int sign, unsigned flags);
void fftw_execute(const fftw_plan p);
void fftw_destroy_plan(fftw_plan p);

#ifdef __cplusplus
}
#endif
#endif
