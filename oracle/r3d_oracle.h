/* r3d_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C, CPU restatement of the Radiative3D phonon-propagate path, operating
 * on the same flattened r3d_model_desc and the same Philox4x32-10 draw stream
 * as the CUDA path, so the two can be compared phonon by phonon.
 *
 * PARITY PIN: this restatement is checked (tests/test_oracle_vs_reference.py)
 * against (a) golden vectors generated from the reference's own compiled
 * objects by oracle/ref_harness.cpp (tests/golden/), including the reference's
 * built-in --rtcoef-test table, and (b) when oracle/_ref is built, against the
 * reference's own Propagate() loop driven by the same Philox stream.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load
 * this library.  The product (radiative3d_b200/) never does.
 */
#ifndef R3D_ORACLE_H_
#define R3D_ORACLE_H_
#include "r3d_gpu.h"
#ifdef __cplusplus
extern "C" {
#endif

/* Trace phonons [first, first+n): accumulate into energies[n_seis][n_bins][5],
 * counts[n_seis][n_bins][2], counters[R3D_NCOUNTERS] (counters[7] = diag OR);
 * finals (may be NULL) receives n end-state records.  nthreads<=1: serial. */
int r3d_oracle_run(const r3d_model_desc *d, uint64_t first, uint64_t n, uint64_t seed,
                   double *energies, uint64_t *counts, uint64_t *counters,
                   r3d_phonon_final *finals, int nthreads);

/* the k-th 31-bit draw of phonon idx (k in [0, 2^31-1] == [0, RAND_MAX]) */
uint32_t r3d_oracle_draw(uint64_t seed, uint64_t idx, uint32_t ordinal);

/* sub-kernels; same in/out layouts as the r3d_test_* hooks in r3d_gpu.h */
void r3d_oracle_cdf_search(const double *cdf, uint32_t n_cdf, const uint32_t *k, uint32_t n, uint32_t *out);
void r3d_oracle_path_to_boundary(const r3d_model_desc *d, const double *in, uint32_t n, double *out);
void r3d_oracle_advance(const r3d_model_desc *d, const double *in, uint32_t n, double *out);
void r3d_oracle_transform(const double *in, uint32_t n, double *out);
void r3d_oracle_rtcoef(const double *in, uint32_t n, double *out);
void r3d_oracle_catch(double bin_dt, uint32_t n_bins, const double *in, uint32_t n, double *out);

/* Scatterer::PopulateProbDists / PopulateWholeProbs / ComputeMFPs (scatterers.cpp:134-220) with ScatterParams::GSATO /
 * XSATO / PSATO (scatparams.cpp:75-194); same outputs as r3d_build_scatterer_tables in r3d_gpu.h */
void r3d_oracle_build_scatterer_tables(const r3d_scatter_params *par, const double *toa_theta, const double *toa_phi,
                                       uint32_t n_toa, double *cdf, double *spol, double *whole_cdf, double *mfp);

#ifdef __cplusplus
}
#endif
#endif
